// host_strict.inl — trig evaluated by the host (cfg.trig_mode == 2), included by batotp_cuda.cu.
//
// batotp's kinematics/dynamics call sin/cos/atan2 of the host libm (robot.cpp:130-136,
// 196-199, 408-419; util.cpp:544-549, 574).  The strict mode of this library (trig_mode 1) runs a
// bit-identical port of that libm's sin/cos on the device (k_trig.cuh).  On a host whose libm is NOT
// the one ported (batotp_cuda_selftest_trig reports mismatches) trig_mode 2 keeps bit parity with
// that host's reference build the slow way: the trig-bearing POINT functions are evaluated by this
// host layer with the host libm between device stages; everything else (splines, sweeps,
// interpolation) stays on the device.  In trig_mode 1 and 2 alike the atan2 of q2aa (util.cpp:574; the
// axis-angle output rows of a UR-type robot) is applied here, to the final rows on their way out.
namespace {

// The serial dynamics evaluated by the host at one point: the caller's point function (cfg.dyn_source = 1,
// batotp_cuda_set_dyn_callback - the plug-in for robots batotp has no model for) or dynRR with the host libm
// (trig_mode 2)
inline void host_dyn_point(const batotp_ctx *h, const double *q, const double *d1, const double *d2, double *a1,
                           double *a2, double *a3, double *a4) {
  if (h->cfg.c.dyn_source == 1)
    h->dynFn(h->dynUser, h->cfg.J, q, d1, d2, a1, a2, a3, a4);
  else
    dyn_rr_point(Trig{0}, q, d1, d2, a1, a2, a3, a4);
}
// runs f(first, last) over [0, n) on the host's cores
template <class F>
void host_parallel_for(int n, F f) {
  const int nth = std::max(1, std::min<int>((int)std::thread::hardware_concurrency(), std::min(64, n)));
  std::vector<std::thread> th;
  for (int t = 1; t < nth; ++t) th.emplace_back([=] { f((int)((long long)n * t / nth), (int)((long long)n * (t + 1) / nth)); });
  f(0, (int)((long long)n / nth));
  for (auto &x : th) x.join();
}

// serial dynamics on the final grid (findDynModel, ba.cpp:905-914): Q/GD/GD2 rows -> A rows (point-major)
void host_dyn_rr_grid(batotp_ctx *h) {
  const DevCfg &c = h->cfg;
  const Ws &w = h->w;
  const int B = h->B, R = c.R;
  h->hst.resize(B);
  g_d2h(h->hst.data(), w.st, (size_t)B * sizeof(TrajState), h->stream);
  g_sync(h->stream);
  int nmax = 0;
  for (int b = 0; b < B; ++b)
    if (!(h->hst[b].status & ST_FATAL_MASK)) nmax = std::max(nmax, h->hst[b].nPts);
  if (nmax <= 0) return;
  const size_t cnt = (size_t)nmax * B * R;
  std::vector<double> q(cnt), d1(cnt), d2(cnt), A((size_t)nmax * B * 4 * w.AD, 0.0);
  g_d2h(q.data(), w.Q, cnt * 8, h->stream);
  g_d2h(d1.data(), w.GD, cnt * 8, h->stream);
  g_d2h(d2.data(), w.GD2, cnt * 8, h->stream);
  g_sync(h->stream);
  const int J = c.J;
  host_parallel_for(B, [&](int b0, int b1) {
    for (int b = b0; b < b1; ++b) {
      const TrajState &s = h->hst[b];
      if (s.status & ST_FATAL_MASK) continue;
      for (int i = 0; i < s.nPts; ++i) {
        const size_t off = ((size_t)i * B + b) * R;
        double a1[MAXD] = {0}, a2[MAXD] = {0}, a3[MAXD] = {0}, a4[MAXD] = {0};
        host_dyn_point(h, &q[off], &d1[off], &d2[off], a1, a2, a3, a4);
        const double *aa[4] = {a1, a2, a3, a4};
        double *Ab = &A[((size_t)i * B + b) * 4 * w.AD];
        for (int k = 0; k < 4; ++k)
          for (int j = 0; j < J; ++j) Ab[k * w.AD + j] = aa[k][j];
      }
    }
  });
  g_h2d(w.A, A.data(), A.size() * 8, h->stream);
  g_sync(h->stream);
}

// dynRR at the output sites (ba.cpp:1815-1825): OA/OD/OD2 rows -> Trq rows (sub-chunk, point-major)
void host_dyn_rr_out(batotp_ctx *h) {
  const DevCfg &c = h->cfg;
  const Ws &w = h->w;
  const int Bo = w.Bo, b0 = w.b0, R = c.R;
  h->hst.resize(h->B);
  g_d2h(h->hst.data(), w.st, (size_t)h->B * sizeof(TrajState), h->stream);
  g_sync(h->stream);
  int nmax = 0;
  for (int bl = 0; bl < Bo; ++bl)
    if (!(h->hst[b0 + bl].status & ST_FATAL_MASK)) nmax = std::max(nmax, h->hst[b0 + bl].nOver);
  if (nmax <= 0) return;
  const size_t cnt = (size_t)nmax * Bo * R;
  std::vector<double> q(cnt), d1(cnt), d2(cnt), T((size_t)nmax * Bo * MAXD, 0.0);
  g_d2h(q.data(), w.OA, cnt * 8, h->stream);
  g_d2h(d1.data(), w.OD, cnt * 8, h->stream);
  g_d2h(d2.data(), w.OD2, cnt * 8, h->stream);
  g_sync(h->stream);
  const int J = c.J;
  host_parallel_for(Bo, [&](int l0, int l1) {
    for (int bl = l0; bl < l1; ++bl) {
      const TrajState &s = h->hst[b0 + bl];
      if (s.status & ST_FATAL_MASK) continue;
      for (int i = 0; i < s.nOver; ++i) {
        const size_t off = ((size_t)i * Bo + bl) * R;
        double a1[MAXD] = {0}, a2[MAXD] = {0}, a3[MAXD] = {0}, a4[MAXD] = {0};
        host_dyn_point(h, &q[off], &d1[off], &d2[off], a1, a2, a3, a4);
        for (int j = 0; j < J; ++j) T[((size_t)i * Bo + bl) * MAXD + j] = a2[j] + a3[j] + a4[j];  // ba.cpp:1823
      }
    }
  });
  g_h2d(w.Trq, T.data(), T.size() * 8, h->stream);
  g_sync(h->stream);
}

// q2aaVect on the final Cartesian rows (ba.cpp:382-403, 1922-1929) and the float cast
void host_q2aa_out(batotp_ctx *h, batotp_batch_out *out, int first) {
  // `first` = global index of the sub-chunk's first trajectory; h->hst holds the chunk's states
  const Ws &w = h->w;
  const int B = w.Bo, OutC = w.OutC, oc = out->out_cap;
  std::vector<double> q((size_t)B * 7 * OutC);
  g_d2h(q.data(), h->d_cartOutD, q.size() * 8, h->stream);
  g_sync(h->stream);
  for (int b = 0; b < B; ++b) {
    const TrajState &s = h->hst[w.b0 + b];
    if (s.status & ST_FATAL_MASK) continue;
    float *o = out->cart_out + (size_t)(first + b) * 6 * oc;
    for (int i = 0; i < s.nCartOut && i < oc; ++i) {
      double cv[7], aa[3];
      for (int r = 0; r < 7; ++r) cv[r] = q[((size_t)b * 7 + r) * OutC + i];
      q2aa_dev(cv + 3, aa);
      for (int r = 0; r < 3; ++r) o[(size_t)r * oc + i] = (float)cv[r];
      for (int r = 0; r < 3; ++r) o[(size_t)(3 + r) * oc + i] = (float)aa[r];
    }
  }
}

template <int J, bool CART, bool TRQ>
void launch_mvc(batotp_ctx *h, double sdotStart, double *d_out, int cap) {
  LAUNCH_TP(h, (k_mvc<J, CART, TRQ>), h->w.Nc, h->B, h->w, sdotStart, d_out, cap);
}

int run_mvc_per_sample(batotp_ctx *h, double sdotStart, double *sdot_out, int cap) {
  const DevCfg &c = h->cfg;
  if (c.trqOn && c.c.is_parallel && !c.c.is_par2ser) {
    h->err = "the per-sample MVC is not instantiated for parallel-mechanism torque limits without isPar2Ser";
    return -1;
  }
  double *d_out = (double *)g_alloc((size_t)h->B * cap * 8);
  g_zero(d_out, (size_t)h->B * cap * 8, h->stream);
  const int key = c.J * 4 + (c.cartOn ? 2 : 0) + (c.trqOn ? 1 : 0);
  int rc = 0;
  switch (key) {
    case 7 * 4 + 0: launch_mvc<7, false, false>(h, sdotStart, d_out, cap); break;
    case 7 * 4 + 2: launch_mvc<7, true, false>(h, sdotStart, d_out, cap); break;
    case 7 * 4 + 3: launch_mvc<7, true, true>(h, sdotStart, d_out, cap); break;
    case 7 * 4 + 1: launch_mvc<7, false, true>(h, sdotStart, d_out, cap); break;
    case 6 * 4 + 3: launch_mvc<6, true, true>(h, sdotStart, d_out, cap); break;
    case 6 * 4 + 2: launch_mvc<6, true, false>(h, sdotStart, d_out, cap); break;
    case 6 * 4 + 0: launch_mvc<6, false, false>(h, sdotStart, d_out, cap); break;
    case 2 * 4 + 3: launch_mvc<2, true, true>(h, sdotStart, d_out, cap); break;
    case 3 * 4 + 3: launch_mvc<3, true, true>(h, sdotStart, d_out, cap); break;
    case 3 * 4 + 2: launch_mvc<3, true, false>(h, sdotStart, d_out, cap); break;
    case 2 * 4 + 2: launch_mvc<2, true, false>(h, sdotStart, d_out, cap); break;
    default: h->err = "no kernel instantiated for this configuration"; rc = -1;
  }
  if (rc == 0) {
    g_d2h(sdot_out, d_out, (size_t)h->B * cap * 8, h->stream);
    g_sync(h->stream);
  }
  g_free(d_out);
  return rc;
}

}  // namespace
