// host_io.cpp — host-side file formats of the BA path, kept byte-compatible with the reference
// (north_star: "reads the same input/*/ robot folders and config.dat options").
//
//   batotp_read_config      BA::readConfigData   batotp/ba.cpp:1942-2087  (+ util.cpp:98-232 readers)
//   batotp_read_traj_bin    BA::trajReadBIN      batotp/ba.cpp:2257-2312
//   batotp_read_traj_csv    BA::trajReadCSV      batotp/ba.cpp:2322-2461
//   batotp_write_traj_bin   BA::trajWriteBIN     batotp/ba.cpp:2582-2651
//   batotp_write_traj_csv   BA::trajWriteCSV     batotp/ba.cpp:2660-2717
//   batotp_write_s_sdot     BA::sdotWrite        batotp/ba.cpp:2726-2759
//
// Plain C ABI (include/batotp_cuda.h); no CUDA here.
#include <clocale>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/batotp_cuda.h"

namespace {

// util.cpp:98-103: consume the rest of the current line
int next_line(FILE *f) {
  int c = 0;
  while (c != '\n' && c != EOF) c = fgetc(f);
  return c;
}

struct Reader {
  FILE *f;
  int count = 0;  // items read, checked against 34+4*nJoints like ba.cpp:2075-2082
  bool str(std::string &out) {
    char tmp[256];
    const int r = fscanf(f, "%255s", tmp);
    if (r == 1) {
      out = tmp;
      count++;
    }
    next_line(f);
    return r == 1;
  }
  bool i32(int &v) {
    const int r = fscanf(f, "%d", &v);
    if (r == 1) count++;
    next_line(f);
    return r == 1;
  }
  bool f64(double &v) {
    const int r = fscanf(f, "%lf", &v);
    if (r == 1) count++;
    next_line(f);
    return r == 1;
  }
  bool vec(double *v, int n) {
    bool ok = true;
    for (int i = 0; i < n; ++i) {
      const int r = fscanf(f, "%lf", &v[i]);
      if (r == 1)
        count++;
      else
        ok = false;
    }
    next_line(f);
    return ok;
  }
  bool flag(int &v) {
    int t = 0;
    const bool ok = i32(t);
    v = (t == 1) ? 1 : 0;
    return ok;
  }
};

struct LocaleGuard {  // ba.cpp:1944-1945: parse with '.' decimals whatever the environment says
  std::string saved;
  LocaleGuard() {
    const char *cur = setlocale(LC_NUMERIC, nullptr);
    saved = cur ? cur : "C";
    if (!setlocale(LC_NUMERIC, "en_US.UTF-8")) setlocale(LC_NUMERIC, "C");
  }
  ~LocaleGuard() { setlocale(LC_NUMERIC, saved.c_str()); }
};

}  // namespace

extern "C" {

int batotp_read_config(const char *path, batotp_cfg *cfg, char *traj_file_name, int name_cap) {
  if (!path || !cfg) return -1;
  LocaleGuard lg;
  FILE *fid = fopen(path, "r");
  if (!fid) {
    printf("\nUnable to open file %s\n", path);
    return -1;
  }
  memset(cfg, 0, sizeof(*cfg));
  Reader rd{fid};
  for (int i = 0; i < 3; ++i) next_line(fid);
  std::string robot, ptype, fname;
  rd.str(robot);
  rd.flag(cfg->is_parallel);
  cfg->robot_type = robot == "KUKA" ? BATOTP_KUKA : robot == "UR" ? BATOTP_UR : robot == "RR" ? BATOTP_RR
                    : robot == "CSPR3DOF" ? BATOTP_CSPR3DOF : robot == "GENJNT" ? BATOTP_GENJNT : 0;
  if (cfg->robot_type == 0) {
    fclose(fid);
    printf("\nreadInputData() error: robotType is %s", robot.c_str());
    printf("It should be 'KUKA', 'UR', 'RR', 'CSPR3DOF', or 'GENJNT'.\n");
    return -1;
  }
  rd.i32(cfg->n_joints);
  rd.i32(cfg->n_cart);
  rd.str(fname);
  if (traj_file_name && name_cap > 0) {
    strncpy(traj_file_name, fname.c_str(), (size_t)name_cap - 1);
    traj_file_name[name_cap - 1] = 0;
  }
  rd.flag(cfg->is_bin_file);
  rd.str(ptype);
  cfg->path_type = ptype == "JOINT" ? BATOTP_JOINT : ptype == "CART" ? BATOTP_CART : ptype == "BOTH" ? BATOTP_BOTH : 0;
  if (cfg->path_type == 0) {
    fclose(fid);
    printf("\nreadInputData() error: pathType is %s", ptype.c_str());
    printf("It should be 'JOINT', 'CART', or 'BOTH'.\n");
    return -1;
  }
  if (cfg->n_joints < 1 || cfg->n_joints > BATOTP_MAX_DOF || cfg->n_cart < 0 || cfg->n_cart > BATOTP_MAX_DOF) {
    fclose(fid);
    printf("\nreadInputData() error: nJoints/nCart outside the supported range 1..%d\n", BATOTP_MAX_DOF);
    return -1;
  }
  next_line(fid);
  next_line(fid);
  const int J = cfg->n_joints;
  rd.flag(cfg->are_jnt_deg);
  rd.flag(cfg->is_jnt_vel_on);
  rd.vec(cfg->jnt_vel_max, J);
  rd.flag(cfg->is_jnt_acc_on);
  rd.vec(cfg->jnt_acc_max, J);
  rd.flag(cfg->is_trq_on);
  rd.vec(cfg->jnt_trq_max, J);
  rd.vec(cfg->jnt_trq_min, J);
  for (int i = 0; i < J; ++i)
    if (std::isnan(cfg->jnt_trq_min[i])) cfg->jnt_trq_min[i] = -cfg->jnt_trq_max[i];  // ba.cpp:2020-2028
  rd.flag(cfg->is_cart_vel_on);
  rd.f64(cfg->cart_vel_max);
  rd.flag(cfg->is_cart_acc_on);
  rd.f64(cfg->cart_acc_max);
  next_line(fid);
  next_line(fid);
  rd.f64(cfg->integ_res);
  rd.f64(cfg->max_integ_time);
  next_line(fid);
  next_line(fid);
  rd.i32(cfg->input_decim_fact);
  rd.i32(cfg->smooth_window);
  rd.flag(cfg->is_sdot_out);
  rd.f64(cfg->jnt_thresh);
  rd.f64(cfg->cart_thresh);
  rd.vec(cfg->s_weights, 3);
  rd.i32(cfg->scale_type);
  rd.f64(cfg->theta_norm_res);
  rd.f64(cfg->theta_norm_res2);
  rd.f64(cfg->cart_norm_res);
  rd.f64(cfg->cart_norm_res2);
  rd.f64(cfg->out_res);
  rd.f64(cfg->out_smooth_fact);
  rd.flag(cfg->is_svd);
  rd.flag(cfg->is_par2ser);
  fclose(fid);
  const double wsum = cfg->s_weights[0] + cfg->s_weights[1] + cfg->s_weights[2];  // ba.cpp:2063-2073
  if (wsum <= 0) {
    printf("Error in readInputData(): sum(sWeights) should be greater than 0.\n");
    return -1;
  }
  for (int i = 0; i < 3; ++i) cfg->s_weights[i] /= wsum;
  const int want = 34 + 4 * J;
  if (rd.count != want) {
    printf("\nfscanf error while reading config.dat file: returned %d; should be %d.\n", rd.count, want);
    return -1;
  }
  cfg->trig_mode = 1;
  return 0;
}

void batotp_free(void *p) { free(p); }

int batotp_read_traj_bin(const char *path, int n_joints, int n_cart, double *tres, int *n0, float **theta,
                         float **cart) {
  if (theta) *theta = nullptr;
  if (cart) *cart = nullptr;
  FILE *fid = fopen(path, "rb");
  if (!fid) {
    printf("\nError! Binary trajectory file '%s' doesn't exist!\n", path);
    return -1;
  }
  float t = 0;
  int n = 0, isTheta = 0, isCart = 0;
  size_t got = fread(&t, 4, 1, fid);
  got += fread(&n, 4, 1, fid);
  got += fread(&isTheta, 4, 1, fid);
  if (got != 3 || n < 0) {
    fclose(fid);
    return -1;
  }
  float *th = nullptr, *ca = nullptr;
  if (isTheta == 1) {
    th = (float *)malloc((size_t)n_joints * n * 4 + 4);
    if (!th) {  // a corrupt header can name any size
      printf("\nError! Cannot allocate %d x %d points for '%s'\n", n_joints, n, path);
      fclose(fid);
      return -1;
    }
    got += fread(th, 4, (size_t)n_joints * n, fid);
  }
  got += fread(&isCart, 4, 1, fid);
  if (isCart == 1) {
    ca = (float *)malloc((size_t)n_cart * n * 4 + 4);
    if (!ca) {
      printf("\nError! Cannot allocate %d x %d points for '%s'\n", n_cart, n, path);
      free(th);
      fclose(fid);
      return -1;
    }
    got += fread(ca, 4, (size_t)n_cart * n, fid);
  }
  fclose(fid);
  const size_t want = (size_t)(isTheta * n_joints + isCart * n_cart) * n + 4;
  if (got != want) {
    printf("\nfread error: %d items read, %d items should have been read.\n", (int)got, (int)want);
    free(th);
    free(ca);
    return -1;
  }
  *tres = (double)t;
  *n0 = n;
  if (theta) *theta = th; else free(th);
  if (cart) *cart = ca; else free(ca);
  return 0;
}

// -> FP64 rows [coord][n0]; header names returned ';'-joined (for trajWriteCSV)
int batotp_read_traj_csv(const char *path, int n_joints, int n_cart, int is_generic, double *tres, int *n0,
                         double **theta, double **cart, double **timestamp, char *header, int header_cap) {
  LocaleGuard lg;
  *theta = *cart = *timestamp = nullptr;
  FILE *fid = fopen(path, "r");
  if (!fid) {
    printf("\nError! File %s doesn't exist", path);
    return -1;
  }
  const size_t nFields = is_generic ? (size_t)n_joints : (size_t)(n_joints + n_cart + 1);
  next_line(fid);
  int n = 0;
  for (;;) {  // count the data lines (ba.cpp:2353-2365)
    double d;
    if (fscanf(fid, "%lf", &d) != 1) break;
    if (next_line(fid) == EOF) break;
    n++;
  }
  if (n == 0) {
    fclose(fid);
    *n0 = 0;
    return 0;
  }
  rewind(fid);
  bool isTs = false, isJ = false, isC = false;
  size_t got = 0;
  std::string hdr;
  for (size_t i = 0; i < nFields; ++i) {
    char tmp[128];
    got += fscanf(fid, " %99[^, \t\n],", tmp);
    const std::string h(tmp);
    if (h == "timestamp") isTs = true;
    if (h == "j1") isJ = true;
    if (h == "x") isC = true;
    hdr += (i ? ";" : "") + h;
  }
  if (header && header_cap > 0) {
    strncpy(header, hdr.c_str(), (size_t)header_cap - 1);
    header[header_cap - 1] = 0;
  }
  double *ts = (double *)calloc((size_t)n, 8);
  double *th = isJ ? (double *)calloc((size_t)n_joints * n, 8) : nullptr;
  double *ca = isC ? (double *)calloc((size_t)n_cart * n, 8) : nullptr;
  for (int i = 0; i < n; ++i) {
    if (isTs) got += fscanf(fid, "%lf,", &ts[i]);
    if (isJ)
      for (int j = 0; j < n_joints; ++j) got += fscanf(fid, "%lf,", &th[(size_t)j * n + i]);
    if (isC)
      for (int j = 0; j < n_cart; ++j) got += fscanf(fid, "%lf,", &ca[(size_t)j * n + i]);
  }
  fclose(fid);
  if (!isTs)
    for (int i = 0; i < n; ++i) ts[i] = 0.2 * (double)i;  // ba.cpp:2440-2444
  *tres = ts[n - 1] / (n - 1);
  *n0 = n;
  *theta = th;
  *cart = ca;
  *timestamp = ts;
  if (nFields * (size_t)(n + 1) != got) {
    printf("trajReadCSV: The number of items read from %s was %d. It should have been %d.\n", path, (int)got,
           (int)(nFields * (n + 1)));
    return -1;
  }
  return 0;
}

// float32 rows in, file out (ba.cpp:2617-2647)
int batotp_write_traj_bin(const char *path, double sres, unsigned n_pts, int n_joints, const float *theta, int n_cart,
                          const float *cart, const float *trq, int pitch) {
  FILE *fid = fopen(path, "wb");
  if (!fid) {
    printf("\nUnable to open file %s", path);
    return -1;
  }
  const float fs = (float)sres;
  const int one = 1, isCart = cart ? 1 : 0, isTrq = trq ? 1 : 0;
  fwrite(&fs, 4, 1, fid);
  fwrite(&n_pts, 4, 1, fid);
  fwrite(&one, 4, 1, fid);
  for (int i = 0; i < n_joints; ++i) fwrite(theta + (size_t)i * pitch, 4, n_pts, fid);
  fwrite(&isCart, 4, 1, fid);
  if (isCart)
    for (int i = 0; i < n_cart; ++i) fwrite(cart + (size_t)i * pitch, 4, n_pts, fid);
  fwrite(&isTrq, 4, 1, fid);
  if (isTrq)
    for (int i = 0; i < n_joints; ++i) fwrite(trq + (size_t)i * pitch, 4, n_pts, fid);
  fclose(fid);
  return 0;
}

// ba.cpp:2735-2749: twice (rev, fwd): f64 sres; i32 n; n f32 s; n f32 sdot
int batotp_write_s_sdot(const char *path, double sres, int n_rev, const float *s_rev, const float *sdot_rev,
                        int n_fwd, const float *s_fwd, const float *sdot_fwd) {
  FILE *fid = fopen(path, "wb");
  if (!fid) {
    printf("\nUnable to open file %s", path);
    return -1;
  }
  const int n[2] = {n_rev, n_fwd};
  const float *s[2] = {s_rev, s_fwd}, *sd[2] = {sdot_rev, sdot_fwd};
  for (int i = 0; i < 2; ++i) {
    if (n[i] <= 0) {
      printf("sdotWrite(): %s was not written because sdot is empty.\n", path);
      fclose(fid);
      return -1;
    }
    fwrite(&sres, 8, 1, fid);
    fwrite(&n[i], 4, 1, fid);
    fwrite(s[i], 4, (size_t)n[i], fid);
    fwrite(sd[i], 4, (size_t)n[i], fid);
  }
  fclose(fid);
  return 0;
}

// ba.cpp:2660-2717 (header ';'-joined as returned by batotp_read_traj_csv; output is always
// "interpolated": time column = i*sres)
int batotp_write_traj_csv(const char *path, const char *header, double sres, int n_pts, int n_joints,
                          const float *theta, int n_cart, const float *cart, int pitch) {
  LocaleGuard lg;
  FILE *fid = fopen(path, "w");
  if (!fid) {
    printf("\nUnable to open file %s", path);
    return -1;
  }
  std::string h(header ? header : "");
  std::vector<std::string> names;
  size_t at = 0;
  while (at <= h.size()) {
    const size_t e = h.find(';', at);
    names.push_back(h.substr(at, e == std::string::npos ? std::string::npos : e - at));
    if (e == std::string::npos) break;
    at = e + 1;
  }
  for (size_t i = 0; i + 1 < names.size(); ++i) fprintf(fid, "%s, ", names[i].c_str());
  if (!names.empty()) fprintf(fid, "%s\n", names.back().c_str());
  for (int i = 0; i < n_pts; ++i) {
    fprintf(fid, "%8.3f", i * sres);
    for (int j = 0; j < n_joints; ++j) fprintf(fid, ", %11.6f", (double)theta[(size_t)j * pitch + i]);
    if (cart)
      for (int j = 0; j < n_cart; ++j) fprintf(fid, ", %9.6f", (double)cart[(size_t)j * pitch + i]);
    fprintf(fid, "\n");
  }
  fclose(fid);
  return 0;
}

}  // extern "C"

// ----------------------------------------------------------------------------- batch writer
// BA::writeOutputData (ba.cpp:2510-2528) for the trajectories of a batch result, on writer threads of its own:
// while they serialise chunk k (traj_out_<index>.dat = trajWriteBIN, s-sdot_<index>.dat = sdotWrite) the
// caller's thread is free to run batotp_cuda_optimize_batch on chunk k+1.
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

struct batotp_writer {
  struct Job {
    batotp_cfg cfg;
    batotp_batch_out out;  // shallow copy: the arrays stay the caller's until batotp_writer_wait returns
    long long base;
    int first, count;
  };
  std::string dir;
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cvWork, cvIdle;
  std::deque<std::pair<Job, int>> queue;  // (job, index within the job)
  int busy = 0;
  bool stop = false;
  std::atomic<long long> written{0}, failed{0};

  static int write_one(const std::string &dir, const Job &j, int b) {
    const batotp_batch_out &o = j.out;
    const int J = j.cfg.n_joints, C = j.cfg.n_cart;
    if (o.status && (o.status[b] & BATOTP_ST_FATAL_MASK)) return 1;  // not optimised: the reference writes nothing
    if (!o.theta_out || !o.n_out || o.out_cap <= 0) return -1;
    const int n = o.n_out[b];
    if (n <= 0 || n > o.out_cap) return -1;
    const double sres = o.out_sres ? o.out_sres[b] : j.cfg.out_res;
    char name[64];
    snprintf(name, sizeof name, "traj_out_%07lld.dat", j.base + b);
    const bool cartFull = o.cart_out && C > 0 && o.n_cart_out && o.n_cart_out[b] == n;  // ba.cpp:2629
    const bool trqFull = j.cfg.is_trq_on && o.trq_out;
    int rc = batotp_write_traj_bin((dir + name).c_str(), sres, (unsigned)n, J,
                                   o.theta_out + (size_t)b * J * o.out_cap, C,
                                   cartFull ? o.cart_out + (size_t)b * C * o.out_cap : nullptr,
                                   trqFull ? o.trq_out + (size_t)b * J * o.out_cap : nullptr, o.out_cap);
    if (rc == 0 && j.cfg.is_sdot_out && !j.cfg.is_interp_only && o.hist && o.hist_cap > 0 && o.n_rev && o.n_fwd) {
      const float *h = o.hist + (size_t)b * 4 * o.hist_cap;
      snprintf(name, sizeof name, "s-sdot_%07lld.dat", j.base + b);
      rc = batotp_write_s_sdot((dir + name).c_str(), sres, o.n_rev[b], h, h + o.hist_cap, o.n_fwd[b],
                               h + 2 * (size_t)o.hist_cap, h + 3 * (size_t)o.hist_cap);
    }
    return rc;
  }
  void run() {
    for (;;) {
      std::pair<Job, int> it;
      {
        std::unique_lock<std::mutex> lk(mu);
        cvWork.wait(lk, [&] { return stop || !queue.empty(); });
        if (queue.empty()) return;
        it = queue.front();
        queue.pop_front();
        busy++;
      }
      // one queue entry = a run of trajectories, so that the lock is taken rarely
      const Job &j = it.first;
      const int lo = it.second, hi = std::min(j.first + j.count, lo + 64);
      for (int b = lo; b < hi; ++b) {
        const int rc = write_one(dir, j, b);
        if (rc == 0)
          written++;
        else if (rc < 0)
          failed++;
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        busy--;
        if (queue.empty() && busy == 0) cvIdle.notify_all();
      }
    }
  }
};

extern "C" {

int batotp_writer_create(const char *dir, int threads, batotp_writer **out) {
  if (!dir || !out || threads < 1) return -1;
  batotp_writer *w = new batotp_writer();
  w->dir = dir;
  if (!w->dir.empty() && w->dir.back() != '/') w->dir += '/';
  for (int t = 0; t < threads; ++t) w->threads.emplace_back([w] { w->run(); });
  *out = w;
  return 0;
}

int batotp_writer_submit(batotp_writer *w, const batotp_cfg *cfg, const batotp_batch_out *out, long long base_index,
                         int first, int count) {
  if (!w || !cfg || !out || first < 0 || count < 0) return -1;
  batotp_writer::Job j{*cfg, *out, base_index, first, count};
  {
    std::lock_guard<std::mutex> lk(w->mu);
    for (int b = first; b < first + count; b += 64) w->queue.emplace_back(j, b);
  }
  w->cvWork.notify_all();
  return 0;
}

int batotp_writer_wait(batotp_writer *w, long long *files_written, long long *files_failed) {
  if (!w) return -1;
  std::unique_lock<std::mutex> lk(w->mu);
  w->cvIdle.wait(lk, [&] { return w->queue.empty() && w->busy == 0; });
  if (files_written) *files_written = w->written.load();
  if (files_failed) *files_failed = w->failed.load();
  return w->failed.load() ? -1 : 0;
}

int batotp_writer_destroy(batotp_writer *w) {
  if (!w) return -1;
  {
    std::lock_guard<std::mutex> lk(w->mu);
    w->stop = true;
  }
  w->cvWork.notify_all();
  for (auto &t : w->threads) t.join();
  delete w;
  return 0;
}

}  // extern "C"
