// ref_harness.cpp — thin C driver around the UNMODIFIED reference library.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/libbatotp_ref.so
// together with /root/reference/batotp/{ba,spline,util,robot}.cpp compiled where they
// lie (no reference source is copied into this repository).  It exposes the reference's
// own BATOTP::BA entry points (ba.h:172-210) and its Traj contents to the tests, and the
// batch runner used as the "reference" CPU baseline.
//
// `private` is re-defined only to READ private state and to call the private per-point
// functions for the instrumented sweep (switching flags); class layout is unchanged.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <thread>
#include <vector>
#include <time.h>
#include <unistd.h>
#include <fcntl.h>

#define private public
#include "ba.h"
#undef private
#include "util.h"
#include "../include/batotp_cfg.h"

using namespace BATOTP;

namespace {
struct Handle {
  BA ba;
  Traj traj = Traj();
  std::vector<unsigned char> flags[2];
  int nRev = 0, nFwd = 0;
  double tRev = 0, tFwd = 0;
};

int g_saved_stdout = -1;

void fill_traj(BA &ba, Traj &traj, int n0, double tres, const float *theta, const float *cart,
               const double *timestamp) {
  // what loadTrajectoryData + trajReadBIN/CSV leave behind (ba.cpp:2206-2245, 2257-2461)
  traj.thetapt.resize(ba._nJoints);
  traj.thetaDpt.resize(ba._nJoints);
  traj.thetaD2pt.resize(ba._nJoints);
  traj.cartpt.resize(ba._nCart);
  traj.cartDpt.resize(ba._nCart);
  traj.cartD2pt.resize(ba._nCart);
  traj.tresInput = tres;
  traj.sres = tres;
  traj.nPts = n0;
  if (theta) {
    traj.theta.assign(ba._nJoints, std::vector<double>(n0));
    for (unsigned j = 0; j < ba._nJoints; ++j)
      for (int i = 0; i < n0; ++i) traj.theta[j][i] = (double)theta[(size_t)j * n0 + i];
  }
  if (cart) {
    traj.cart.assign(ba._nCart, std::vector<double>(n0));
    for (unsigned j = 0; j < ba._nCart; ++j)
      for (int i = 0; i < n0; ++i) traj.cart[j][i] = (double)cart[(size_t)j * n0 + i];
  }
  if (timestamp) traj.timestamp.assign(timestamp, timestamp + n0);
}

int put(const std::vector<double> &v, double *buf, int cap) {
  if (buf)
    for (int i = 0; i < (int)v.size() && i < cap; ++i) buf[i] = v[i];
  return (int)v.size();
}
int put_row(const std::vector<std::vector<double>> &m, int idx, double *buf, int cap) {
  if (idx < 0 || idx >= (int)m.size()) return 0;
  return put(m[idx], buf, cap);
}
}  // namespace

extern "C" {

void ref_silence(int on) {
  fflush(stdout);
  if (on && g_saved_stdout < 0) {
    g_saved_stdout = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
  } else if (!on && g_saved_stdout >= 0) {
    dup2(g_saved_stdout, 1);
    close(g_saved_stdout);
    g_saved_stdout = -1;
  }
}

void *ref_new(const char *config_path, const char *input_folder, const char *output_folder,
              int is_auto_integ_res) {
  Handle *h = new Handle();
  if (input_folder) h->ba.setInputFolder(input_folder);
  if (output_folder) h->ba.setOutputFolder(output_folder);
  h->ba.setIsAutoIntegRes(is_auto_integ_res != 0);  // batest: false (test/main.cpp:53)
  if (h->ba.readConfigData(config_path) == -1) {
    delete h;
    return nullptr;
  }
  return h;
}
void ref_free(void *p) { delete (Handle *)p; }
/* BA::setIsInterpOnly (ba.h public setters): interpInputData then only re-samples the path (ba.cpp:139-159) */
void ref_set_interp_only(void *p, int on) { ((Handle *)p)->ba.setIsInterpOnly(on != 0); }

int ref_load_file(void *p) {
  Handle *h = (Handle *)p;
  return h->ba.loadTrajectoryData(h->traj);
}
int ref_load_raw(void *p, int n0, double tres, const float *theta, const float *cart,
                 const double *timestamp) {
  Handle *h = (Handle *)p;
  fill_traj(h->ba, h->traj, n0, tres, theta, cart, timestamp);
  return 0;
}
int ref_interp_input(void *p) {
  Handle *h = (Handle *)p;
  return h->ba.interpInputData(h->traj);
}
int ref_sweep(void *p, int dir, int last) {
  Handle *h = (Handle *)p;
  h->ba.setIntegDir(dir);
  h->ba.setIsLastSweep(last != 0);
  int r = h->ba.sweep(h->traj);
  if (dir == 1) {
    h->nFwd = h->traj.nPts;
    h->tFwd = h->traj.tTotalTraj;
  } else {
    h->nRev = h->traj.nPts;
    h->tRev = h->traj.tTotalTraj;
  }
  return r;
}
int ref_interp_output(void *p) {
  Handle *h = (Handle *)p;
  return h->ba.interpOutputData(h->traj);
}
int ref_write_output(void *p) {
  Handle *h = (Handle *)p;
  return h->ba.writeOutputData(h->traj);
}

// Instrumented sweep: the step loop is this driver's, every per-point function
// (sdotLim, applyAccelConstraintsBisectionPt and what they call) is the reference's own.
// Used only to obtain the per-step switching flags (SURVEY §8c "limit-active switching
// indices"); the tests check its (s, sdot) history equals BA::sweep's bit for bit.
int ref_sweep_flags(void *p, int dir, double *sOut, double *sdotOut, unsigned char *flags, int cap) {
  Handle *h = (Handle *)p;
  BA &ba = h->ba;
  Traj &traj = h->traj;
  ba._integDir = dir;
  const double absh = ba._integRes;
  const double hh = dir * absh;
  std::array<double, 7> sArr, sdotArr, sddotArr;
  sArr.fill(0);
  sdotArr.fill(0);
  sddotArr.fill(0);
  double sLast;
  int nIter = 0;
  if (dir == 1) {
    traj.curSegC = 0; traj.tauC = 0; sArr[0] = 0;
    traj.curSegMVC = 0; traj.tauMVC = 0; sLast = traj.sC[traj.nPtsC - 1];
  } else {
    traj.curSegC = traj.nPtsC - 2; traj.tauC = 1; sArr[0] = traj.sC[traj.nPtsC - 1];
    traj.curSegMVC = traj.nPts - 2; traj.tauMVC = 1; sLast = 0;
  }
  traj.sCur = sArr[0];
  traj.sdotCur = 0;
  ba.applyAccelConstraintsBisectionPt(traj, sddotArr[0], nIter);
  sdotArr[0] = .1 * hh * sddotArr[0];
  ba._sdotMin = sdotArr[0];
  ba.sdotLim(traj, sdotArr[0], "linear");
  ba._sdotMin = sdotArr[0];
  traj.sdotCur = sdotArr[0];
  ba.applyAccelConstraintsBisectionPt(traj, sddotArr[0], nIter);
  sdotArr[0] = traj.sdotCur;
  ba.sdotLim(traj, sdotArr[0], "linear");
  int n = 0;
  if (n < cap) { sOut[n] = sArr[0]; sdotOut[n] = sdotArr[0]; flags[n] = 0; }
  n++;
  const int maxSteps = (int)std::floor(ba._maxIntegTime / ba._integRes) + 1;
  for (int i = 1;; ++i) {
    double s0 = traj.sCur;
    traj.sdotLimTypeT = false;
    sArr[6] = sArr[0] + hh * sdotArr[0];
    sdotArr[6] = sdotArr[0] + hh * sddotArr[0];
    traj.sCur = sArr[6];
    ba.sdotLim(traj, sdotArr[6], "linear");
    traj.sCur = s0;
    int nLim = 0, nBis = 0;
    for (int j = 0; j < 6; ++j) {
      traj.sdotLimTypeT = false;
      double sdotT = 0, sddotT = 0;
      for (int k = 0; k < j + 1; ++k) {
        sdotT += ba._B[k][j] * sdotArr[k];
        sddotT += ba._B[k][j] * sddotArr[k];
      }
      sArr[j + 1] = sArr[0] + hh * sdotT;
      sdotArr[j + 1] = sdotArr[0] + hh * sddotT;
      sdotArr[j + 1] = std::max(sdotArr[j + 1], 0.0);
      traj.sCur = sArr[j + 1];
      ba.sdotLim(traj, sdotArr[j + 1], "linear");
      traj.sdotCur = sdotArr[j + 1];
      ba.applyAccelConstraintsBisectionPt(traj, sddotArr[j + 1], nIter);
      sdotArr[j + 1] = traj.sdotCur;
      if (traj.sdotLimTypeT) nLim++;
      if (nIter > 0) nBis++;
    }
    sArr[0] = sArr[6];
    sdotArr[0] = sdotArr[6];
    sddotArr[0] = sddotArr[6];
    if (n < cap) {
      sOut[n] = sArr[0];
      sdotOut[n] = sdotArr[0];
      flags[n] = (unsigned char)(nLim | (nBis << 3) | ((traj.isOn_sdot ? 1 : 0) << 6));
    }
    n++;
    if (traj.sCur * dir > sLast) break;
    if (i > maxSteps) return -1;
  }
  return n;
}

// SURVEY §8a A10 (derived product: the per-sample maximum-velocity curve) from the REFERENCE's own private
// per-point functions: evalSplinePartials (ba.cpp:1341), sdotLim (ba.cpp:1204; reverse direction, so no MVC
// term; _sdotMin = 0) and applyAccelConstraintsBisectionPt (ba.cpp:1248) at every knot of the s-grid, starting
// from sdot_start.  This is what pins the restatement's orc_mvc_per_sample (and through it the device kernel
// k_mvc) to the reference.  Call after ref_interp_input.
int ref_mvc_per_sample(void *p, double sdot_start, double *out, int cap) {
  Handle *h = (Handle *)p;
  BA &ba = h->ba;
  Traj &traj = h->traj;
  const int n = traj.nPtsC;
  const int saveDir = ba._integDir;
  const double saveMin = ba._sdotMin, saveSec = traj.sLastSec;
  ba._integDir = -1;
  for (int k = 0; k < n && k < cap; ++k) {
    traj.sCur = traj.sC[k];
    traj.curSegC = std::min(k, n - 2);
    ba._sdotMin = 0.0;
    ba.evalSplinePartials(traj);  // the velocity limits read the partials of the last evaluation (quirk Q2)
    double sd = sdot_start;
    ba.sdotLim(traj, sd, "linear");
    traj.sdotCur = sd;
    double sdd = 0;
    int it = 0;
    traj.sLastSec = 0;  // keeps the bisection from recording it
    ba.applyAccelConstraintsBisectionPt(traj, sdd, it);
    out[k] = traj.sdotCur;
  }
  ba._integDir = saveDir;
  ba._sdotMin = saveMin;
  traj.sLastSec = saveSec;
  return n;
}

int ref_get_vec(void *p, const char *name, int idx, double *buf, int cap) {
  Handle *h = (Handle *)p;
  Traj &t = h->traj;
  std::string n(name);
  if (n == "theta") return put_row(t.theta, idx, buf, cap);
  if (n == "thetaD") return put_row(t.thetaD, idx, buf, cap);
  if (n == "thetaD2") return put_row(t.thetaD2, idx, buf, cap);
  if (n == "cart") return put_row(t.cart, idx, buf, cap);
  if (n == "cartD") return put_row(t.cartD, idx, buf, cap);
  if (n == "cartD2") return put_row(t.cartD2, idx, buf, cap);
  if (n == "trq") return put_row(t.trq, idx, buf, cap);
  if (n == "a1") return put_row(t.a1, idx, buf, cap);
  if (n == "a2") return put_row(t.a2, idx, buf, cap);
  if (n == "a3") return put_row(t.a3, idx, buf, cap);
  if (n == "a4") return put_row(t.a4, idx, buf, cap);
  if (n == "sMVC") return put(t.sMVC, buf, cap);
  if (n == "sdot") return put(t.sdot, buf, cap);
  if (n == "tMVC") return put(t.tMVC, buf, cap);
  if (n == "sC") return put(t.sC, buf, cap);
  if (n == "ptsOrig") return put(t.ptsOrig, buf, cap);
  if (n == "hist_s0") return t.myMVChist.s.size() > 0 ? put(t.myMVChist.s[0], buf, cap) : 0;
  if (n == "hist_sdot0") return t.myMVChist.sdot.size() > 0 ? put(t.myMVChist.sdot[0], buf, cap) : 0;
  if (n == "hist_s1") return t.myMVChist.s.size() > 1 ? put(t.myMVChist.s[1], buf, cap) : 0;
  if (n == "hist_sdot1") return t.myMVChist.sdot.size() > 1 ? put(t.myMVChist.sdot[1], buf, cap) : 0;
  auto coef = [&](std::vector<Spline::splineCoeffs> &c, bool m) -> int {
    if (idx < 0 || idx >= (int)c.size()) return 0;
    if (!m) return put(c[idx].c0, buf, cap);
    // second-derivative solution m_k = 2*c2_k (spline.cpp:206); c2 = sol/2.0 is exact to invert
    std::vector<double> mm(c[idx].c2.size());
    for (size_t k = 0; k < mm.size(); ++k) mm[k] = 2.0 * c[idx].c2[k];
    return put(mm, buf, cap);
  };
  if (n == "thetaC_y") return coef(t.thetaC, false);
  if (n == "thetaC_m") return coef(t.thetaC, true);
  if (n == "cartC_y") return coef(t.cartC, false);
  if (n == "cartC_m") return coef(t.cartC, true);
  if (n == "a1C_m") return coef(t.a1C, true);
  if (n == "a2C_m") return coef(t.a2C, true);
  if (n == "a3C_m") return coef(t.a3C, true);
  if (n == "a4C_m") return coef(t.a4C, true);
  return -1;
}

double ref_get_scalar(void *p, const char *name) {
  Handle *h = (Handle *)p;
  Traj &t = h->traj;
  std::string n(name);
  if (n == "nPts") return t.nPts;
  if (n == "nPtsC") return t.nPtsC;
  if (n == "sres") return t.sres;
  if (n == "sresC") return t.sresC;
  if (n == "vFact") return t.vFact;
  if (n == "aFact") return t.aFact;
  if (n == "tTotalTraj") return t.tTotalTraj;
  if (n == "sLastSec") return t.sLastSec;
  if (n == "nRev") return h->nRev;
  if (n == "nFwd") return h->nFwd;
  if (n == "tRev") return h->tRev;
  if (n == "tFwd") return h->tFwd;
  if (n == "nCart") return h->ba._nCart;
  if (n == "outRes") return h->ba._outRes;
  if (n == "integRes") return h->ba._integRes;
  if (n == "errorOptimization") return (double)h->ba.getErrorOptimization();
  if (n == "cartRows") return (double)t.cart.size();
  if (n == "trqRows") return (double)t.trq.size();
  return NAN;
}

// Read the options the reference parsed into the shared POD (so tests can feed the very
// same values to the oracle restatement and to the CUDA library).
int ref_get_cfg(void *p, batotp_cfg *c) {
  Handle *h = (Handle *)p;
  BA &b = h->ba;
  memset(c, 0, sizeof(*c));
  c->robot_type = b._robotType;
  c->is_parallel = b._isParallelMechOrig;
  c->n_joints = b._nJoints;
  c->n_cart = b._nCart;
  c->is_bin_file = b._isBINfile;
  c->path_type = b._pathType;
  c->are_jnt_deg = b._areJointAnglesDegrees;
  c->is_jnt_vel_on = b._isJntVelConOn;
  c->is_jnt_acc_on = b._isJntAccConOn;
  c->is_trq_on = b._isTrqConOn;
  c->is_cart_vel_on = b._isCartVelConOn;
  c->is_cart_acc_on = b._isCartAccConOn;
  c->input_decim_fact = b._inputDecimFact;
  c->smooth_window = b._smoothWindow;
  c->is_sdot_out = b.is_sdotOut;
  c->scale_type = b._scaleType;
  c->is_svd = b._isSVD;
  c->is_par2ser = b._isPar2Ser;
  c->is_interp_only = b._isInterpOnly;
  c->is_auto_integ_res = b._isAutoIntegRes;
  c->trig_mode = 1;
  for (unsigned i = 0; i < b._nJoints && i < BATOTP_MAX_DOF; ++i) {
    c->jnt_vel_max[i] = b._JntVelMax[i];
    c->jnt_acc_max[i] = b._JntAccMax[i];
    c->jnt_trq_max[i] = b._JntTrqMax[i];
    c->jnt_trq_min[i] = b._JntTrqMin[i];
  }
  c->cart_vel_max = b._CartVelMax;
  c->cart_acc_max = b._CartAccMax;
  c->integ_res = b._integRes;
  c->max_integ_time = b._maxIntegTime;
  c->jnt_thresh = b._jntThresh;
  c->cart_thresh = b._cartThresh;
  for (int i = 0; i < 3; ++i) c->s_weights[i] = b._sWeights[i];
  c->theta_norm_res = b._thetaNormRes;
  c->theta_norm_res2 = b._thetaNormRes2;
  c->cart_norm_res = b._cartNormRes;
  c->cart_norm_res2 = b._cartNormRes2;
  c->out_res = b._outRes;
  c->out_smooth_fact = b._outSmoothFact;
  return 0;
}

// Batch CPU baseline ("reference"): BASELINE.md §3 — T worker threads, a fresh BA (copy of the
// parsed template) and a fresh Traj per trajectory, timing interpInputData + 2 sweeps +
// interpOutputData only.  Same argument meaning as orc_batch_run.
double ref_batch_run(const char *config_path, int B, int n0, double tres, const float *theta,
                     const float *cart, int n_threads, double *t_total, int *n_rev, int *n_fwd,
                     int *n_out, int *status, float *theta_out, int out_cap) {
  BA tmpl;
  tmpl.setIsAutoIntegRes(false);
  ref_silence(1);
  if (tmpl.readConfigData(config_path) == -1) {
    ref_silence(0);
    return -1.0;
  }
  if (n_threads < 1) n_threads = 1;
  const int J = tmpl._nJoints, C = tmpl._nCart;
  auto worker = [&](int tid) {
    for (int b = tid; b < B; b += n_threads) {
      BA ba = tmpl;
      Traj traj = Traj();
      fill_traj(ba, traj, n0, tres, theta ? theta + (size_t)b * J * n0 : nullptr,
                cart ? cart + (size_t)b * C * n0 : nullptr, nullptr);
      int r = ba.interpInputData(traj);
      int nr = 0, nf = 0;
      if (r == 0) {
        ba.setIntegDir(-1);
        ba.setIsLastSweep(false);
        r = ba.sweep(traj);
        nr = traj.nPts;
      }
      if (r == 0) {
        ba.setIntegDir(1);
        ba.setIsLastSweep(true);
        r = ba.sweep(traj);
        nf = traj.nPts;
      }
      if (r == 0) ba.interpOutputData(traj);
      if (status) status[b] = r;
      if (t_total) t_total[b] = traj.tTotalTraj;
      if (n_rev) n_rev[b] = nr;
      if (n_fwd) n_fwd[b] = nf;
      if (n_out) n_out[b] = (r == 0) ? (int)traj.theta[0].size() : 0;
      if (theta_out && r == 0)
        for (int j = 0; j < J; ++j)
          for (int k = 0; k < (int)traj.theta[j].size() && k < out_cap; ++k)
            theta_out[((size_t)b * J + j) * out_cap + k] = (float)traj.theta[j][k];
    }
  };
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) th.emplace_back(worker, i);
  for (auto &x : th) x.join();
  clock_gettime(CLOCK_MONOTONIC, &t1);
  ref_silence(0);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

}  // extern "C"
