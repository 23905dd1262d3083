// ba.cpp — BATOTP::BA facade over the CUDA library (see ba.h).  Host code only: option and file
// handling stay on the host (host_io.cpp, byte-compatible with the reference), the numerical
// path runs on the device through include/batotp_cuda.h.
#include "ba.h"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "util.h"

namespace BATOTP {

BA::BA(void) {
  memset(&_cfg, 0, sizeof(_cfg));
  _cfg.is_auto_integ_res = 1;  // ba.h:309 default; batest switches it off (test/main.cpp:53)
  _cfg.trig_mode = 1;
  _HomeFolder = "../";  // ba.cpp:67-74
  _InputFolder = _HomeFolder + "input/";
  _OutputFolder = _HomeFolder + "output/";
}

BA::~BA(void) {
  if (_h) batotp_cuda_destroy(_h);
}

void BA::setJointMaximalVelocity(const std::vector<double> &v) {
  for (size_t i = 0; i < v.size() && i < BATOTP_MAX_DOF; ++i) _cfg.jnt_vel_max[i] = v[i];
}
void BA::setJointMaximalAcceleration(const std::vector<double> &a) {
  for (size_t i = 0; i < a.size() && i < BATOTP_MAX_DOF; ++i) _cfg.jnt_acc_max[i] = a[i];
}
std::vector<double> BA::getJointMaximalVelocity() const {
  return std::vector<double>(_cfg.jnt_vel_max, _cfg.jnt_vel_max + _cfg.n_joints);
}
std::vector<double> BA::getJointMaximalAcceleration() const {
  return std::vector<double>(_cfg.jnt_acc_max, _cfg.jnt_acc_max + _cfg.n_joints);
}

int BA::ensureDevice() {
  if (_h) return 0;
  if (batotp_cuda_create(_device, &_h) != 0) {
    printf("BA: no CUDA device available (this build has no CPU path).\n");
    return -1;
  }
  batotp_cuda_set_keep_f64(_h, 1);
  return 0;
}

int BA::readConfigData(const char *filename) {
  const int autoRes = _cfg.is_auto_integ_res, interpOnly = _cfg.is_interp_only, trig = _cfg.trig_mode;
  char name[1024];
  printf("\nConfiguration file: '%s'\n", filename);
  if (batotp_read_config(filename, &_cfg, name, (int)sizeof(name)) != 0) return -1;
  _cfg.is_auto_integ_res = autoRes;
  _cfg.is_interp_only = interpOnly;
  _cfg.trig_mode = trig;
  _trajFileName = _InputFolder + name;
  static const char *names[] = {"", "KUKA", "UR", "RR", "CSPR3DOF", "GENJNT"};
  _robotTypeStr = names[_cfg.robot_type];
  return 0;
}

int BA::loadConfigData(const Config &c) {
  const int autoRes = _cfg.is_auto_integ_res, interpOnly = _cfg.is_interp_only, trig = _cfg.trig_mode;
  memset(&_cfg, 0, sizeof(_cfg));
  _cfg.is_auto_integ_res = autoRes;
  _cfg.is_interp_only = interpOnly;
  _cfg.trig_mode = trig;
  _robotTypeStr = c.robotTypeStr;
  _cfg.robot_type = c.robotTypeStr == "KUKA" ? BATOTP_KUKA : c.robotTypeStr == "UR" ? BATOTP_UR
                    : c.robotTypeStr == "RR" ? BATOTP_RR : c.robotTypeStr == "CSPR3DOF" ? BATOTP_CSPR3DOF
                    : c.robotTypeStr == "GENJNT" ? BATOTP_GENJNT : 0;
  if (_cfg.robot_type == 0) {
    printf("\nreadInputData() error: robotType is %s", c.robotTypeStr.c_str());
    printf("It should be 'KUKA', 'UR', 'RR', 'CSPR3DOF', or 'GENJNT'.\n");
    return -1;
  }
  _cfg.is_parallel = c.isParallelMech;
  _cfg.n_joints = c.nJoints;
  _cfg.n_cart = c.nCart;
  _trajFileName = c.trajFileName;
  _cfg.is_bin_file = c.isBinFile;
  _cfg.path_type = c.pathType == "JOINT" ? BATOTP_JOINT : c.pathType == "CART" ? BATOTP_CART
                   : c.pathType == "BOTH" ? BATOTP_BOTH : 0;
  if (_cfg.path_type == 0) {
    printf("\nreadInputData() error: pathType is %s", c.pathType.c_str());
    printf("It should be 'JOINT', 'CART', or 'BOTH'.\n");
    return -1;
  }
  if (c.nJoints < 1 || c.nJoints > BATOTP_MAX_DOF || c.nCart < 0 || c.nCart > BATOTP_MAX_DOF) return -1;
  _cfg.is_jnt_vel_on = c.isJntVelConon;
  _cfg.is_jnt_acc_on = c.isJntAccConOn;
  _cfg.is_trq_on = c.isTrqConOn;
  for (int i = 0; i < c.nJoints; ++i) {
    _cfg.jnt_vel_max[i] = i < (int)c.jntVelLims.size() ? c.jntVelLims[i] : 0;
    _cfg.jnt_acc_max[i] = i < (int)c.jntAccLims.size() ? c.jntAccLims[i] : 0;
    _cfg.jnt_trq_max[i] = i < (int)c.jntTrqMax.size() ? c.jntTrqMax[i] : 0;
    const double mn = i < (int)c.jntTrqMin.size() ? c.jntTrqMin[i] : 0;
    _cfg.jnt_trq_min[i] = std::isnan(mn) ? -_cfg.jnt_trq_max[i] : mn;  // ba.cpp:2150-2155
  }
  _cfg.is_cart_vel_on = c.isCartVelConOn;
  _cfg.cart_vel_max = c.cartVelMax;
  _cfg.is_cart_acc_on = c.isCarAccConOn;
  _cfg.cart_acc_max = c.cartAccMax;
  _cfg.integ_res = c.integRes;
  _cfg.max_integ_time = c.maxIntegTime;
  _cfg.input_decim_fact = c.inputDecimFact;
  _cfg.smooth_window = c.smoothWindow;
  _cfg.is_sdot_out = c.is_sdotOut;
  _cfg.jnt_thresh = c.jntThresh;
  _cfg.cart_thresh = c.cartThresh;
  double w[3] = {0, 0, 0};
  for (int i = 0; i < 3 && i < (int)c.sWeights.size(); ++i) w[i] = c.sWeights[i];
  const double ws = w[0] + w[1] + w[2];
  if (ws <= 0) {
    printf("Error in readInputData(): sum(sWeights) should be greater than 0.\n");
    return -1;
  }
  for (int i = 0; i < 3; ++i) _cfg.s_weights[i] = w[i] / ws;
  _cfg.scale_type = c.scaleType;
  _cfg.theta_norm_res = c.thetaNormRes;
  _cfg.theta_norm_res2 = c.thetaNormRes2;
  _cfg.cart_norm_res = c.cartNormRes;
  _cfg.cart_norm_res2 = c.cartNormRes2;
  _cfg.out_res = c.outRes;
  _cfg.out_smooth_fact = c.outSmoothFact;
  _cfg.is_svd = c.isSVD;
  _cfg.is_par2ser = c.isPar2Ser;
  return 0;
}

int BA::loadTrajectoryData(Traj &traj) {
  traj.trajFileName = _trajFileName;
  const char *fn = _trajFileName.c_str();
  if (doesFileExist(fn) != 0) {
    printf("Error: The file '%s' does not exist.\n", fn);
    return -1;
  }
  const int J = _cfg.n_joints, C = _cfg.n_cart;
  traj.rawTheta32.clear();
  traj.rawCart32.clear();
  traj.rawTheta64.clear();
  traj.rawCart64.clear();
  traj.timestamp.clear();
  int n0 = 0;
  double tres = 0;
  if (_cfg.is_bin_file) {
    float *th = nullptr, *ca = nullptr;
    if (batotp_read_traj_bin(fn, J, C, &tres, &n0, &th, &ca) != 0) return -1;
    if (th) traj.rawTheta32.assign(th, th + (size_t)J * n0);
    if (ca) traj.rawCart32.assign(ca, ca + (size_t)C * n0);
    batotp_free(th);
    batotp_free(ca);
  } else {
    double *th = nullptr, *ca = nullptr, *ts = nullptr;
    char header[2048];
    const int rc = batotp_read_traj_csv(fn, J, C, _cfg.robot_type == BATOTP_GENJNT, &tres, &n0, &th, &ca, &ts,
                                        header, (int)sizeof(header));
    if (rc == 0 && n0 > 0) {
      if (th) traj.rawTheta64.assign(th, th + (size_t)J * n0);
      if (ca) traj.rawCart64.assign(ca, ca + (size_t)C * n0);
      if (ts) traj.timestamp.assign(ts, ts + n0);
      _csvHeader = header;
      traj.trajFileHeader.clear();
      size_t at = 0;
      const std::string h(header);
      while (at <= h.size()) {
        const size_t e = h.find(';', at);
        traj.trajFileHeader.push_back(h.substr(at, e == std::string::npos ? std::string::npos : e - at));
        if (e == std::string::npos) break;
        at = e + 1;
      }
    }
    batotp_free(th);
    batotp_free(ca);
    batotp_free(ts);
    if (rc != 0) return -1;
  }
  traj.tresInput = tres;
  traj.sres = tres;
  traj.nPts = (unsigned)n0;
  traj.nRaw = n0;
  // printInputData (ba.cpp:2470-2501)
  printf("\nRobot: %s \n", _robotTypeStr.c_str());
  printf("Number of robot joints: %u \n", (unsigned)J);
  printf("Input  traj. file : %s\n", traj.trajFileName.c_str());
  printf("Input resolution  :  %.4f s\n", traj.tresInput);
  printf("Number of traj pts: %d\n", traj.nPts);
  printf("Joint velocity limits : ");
  for (int i = 0; i < J; ++i) printf("%.1f ", _cfg.jnt_vel_max[i]);
  printf("\nJoint accel.   limits : ");
  for (int i = 0; i < J; ++i) printf("%.1f ", _cfg.jnt_acc_max[i]);
  printf("\nCartesian speed  limit: %.4f\n", _cfg.cart_vel_max);
  printf("Integration resolution: %.4f s\n", _cfg.integ_res);
  printf("Output      resolution: %.4f s\n", _cfg.out_res);
  printf("Max. integration time : %.0f s\n\n", _cfg.max_integ_time);
  return 0;
}

int BA::pull(const char *name, int row, std::vector<double> &v) {
  const int n = batotp_cuda_get_f64(_h, name, 0, row, nullptr, 0);
  if (n < 0) return -1;
  v.resize((size_t)n);
  if (n > 0) batotp_cuda_get_f64(_h, name, 0, row, v.data(), n);
  return n;
}

int BA::interpInputData(Traj &traj) {
  if (ensureDevice() != 0) return -1;
  batotp_batch_in in;
  memset(&in, 0, sizeof(in));
  in.B = 1;
  in.n0_max = traj.nRaw;
  in.tres_all = traj.tresInput;
  in.theta_f32 = traj.rawTheta32.empty() ? nullptr : traj.rawTheta32.data();
  in.cart_f32 = traj.rawCart32.empty() ? nullptr : traj.rawCart32.data();
  in.theta_f64 = traj.rawTheta64.empty() ? nullptr : traj.rawTheta64.data();
  in.cart_f64 = traj.rawCart64.empty() ? nullptr : traj.rawCart64.data();
  in.timestamp = traj.timestamp.empty() ? nullptr : traj.timestamp.data();
  _sweepsDone = false;
  _errorOptimization = NO_ERROR;
  if (batotp_cuda_load(_h, &_cfg, &in) != 0 || batotp_cuda_interp_input(_h) != 0) {
    printf("interpInputData(): %s\n", batotp_cuda_last_error(_h));
    return -1;
  }
  if (_cfg.is_interp_only) {  // ba.cpp:139-159: the re-sampled path is the result; the reference returns -1 here
    _sweepsDone = true;
    const bool was = _cfg.is_trq_on != 0;
    _cfg.is_trq_on = 0;  // no torque rows in this mode
    fillOutput(traj);
    _cfg.is_trq_on = was;
    _sweepsDone = false;
    return -1;
  }
  // status and grid size come back with the fetch after the sweeps; the grid size is available now
  std::vector<double> y;
  const int n = pull("thetaC_y", 0, y);
  if (n <= 0) {
    printf("Input trajectory could not be interpolated; no optimization will be performed.\n");
    return -1;
  }
  traj.nGrid = n;
  traj.nPts = (unsigned)n;
  traj.myMVChist.s.assign(4, std::vector<double>());
  traj.myMVChist.sdot.assign(4, std::vector<double>());
  printf("Number of points on MVC, theta, and cart arrays after splineFact: %d\n", traj.nPts);
  return 0;
}

int BA::sweep(Traj &traj) {
  if (!_h) return -1;
  if (_integDir == -1) {
    if (batotp_cuda_sweeps(_h) != 0) {
      printf("sweep(): %s\n", batotp_cuda_last_error(_h));
      return -1;
    }
    _sweepsDone = true;
  } else if (!_sweepsDone) {
    printf("sweep(): the reverse sweep (integDir = -1) must be requested before the forward sweep.\n");
    return -1;
  }
  std::vector<double> s, sd;
  const bool fwd = (_integDir == 1);
  if (pull(fwd ? "s_fwd" : "s_rev", 0, s) <= 0 || pull(fwd ? "sdot_fwd" : "sdot_rev", 0, sd) <= 0) {
    printf("Error in sweep(): integration did not complete (maxIntegTime of %.1f s exceeded or numerical failure).\n",
           _cfg.max_integ_time);
    _errorOptimization = MAX_INTEGRATION_TIME;
    return -1;
  }
  const int nPts = (int)s.size();
  // the step the device integrated with: _integRes, or the automatically chosen one (ba.cpp:493-556, 631)
  double absh = _cfg.integ_res;
  std::vector<double> one;
  if (pull("integ_res", 0, one) == 1) absh = one[0];
  const double tElapsed = absh * (nPts - 1);
  printf("%s integ.: %4d steps; %5d ODE evals; %3d failed steps; traj time. %.3f sec.; avg. step size %f sec.\n",
         fwd ? "fwd." : "rev.", nPts, 4 * (nPts - 1), 0, tElapsed, tElapsed / nPts);
  traj.sMVC = s;
  traj.sdot = sd;
  traj.nPts = (unsigned)nPts;
  traj.tTotalTraj = tElapsed;
  if (fwd) {
    traj.nFwd = nPts;
    traj.tMVC.resize((size_t)nPts);
    for (int i = 0; i < nPts; ++i) traj.tMVC[i] = absh * (double)i;
  } else
    traj.nRev = nPts;
  if (_cfg.is_sdot_out) {
    traj.myMVChist.s[fwd ? 1 : 0] = s;
    traj.myMVChist.sdot[fwd ? 1 : 0] = sd;
  }
  return 0;
}

int BA::interpOutputData(Traj &traj) {
  if (!_h || !_sweepsDone) return -1;
  if (batotp_cuda_interp_output(_h) != 0) {
    printf("interpOutputData(): %s\n", batotp_cuda_last_error(_h));
    return -1;
  }
  return fillOutput(traj);
}

// copies the packed result of the resident trajectory into Traj (FP64 rows before the float cast)
int BA::fillOutput(Traj &traj) {
  batotp_batch_out o;
  memset(&o, 0, sizeof(o));
  int status = 0, nOut = 0, nCart = 0;
  double sres = 0, sLastSec = 0;
  o.status = &status;
  o.n_out = &nOut;
  o.n_cart_out = &nCart;
  o.out_sres = &sres;
  o.s_last_sec = &sLastSec;
  if (batotp_cuda_fetch(_h, &o) != 0) return -1;
  traj.status = status;
  traj.sLastSec = sLastSec;
  const int J = _cfg.n_joints, C = _cfg.n_cart;
  traj.theta.assign((size_t)J, std::vector<double>());
  for (int j = 0; j < J; ++j) pull("theta_out", j, traj.theta[j]);
  traj.cart.assign((size_t)C, std::vector<double>());
  for (int j = 0; j < C; ++j) pull("cart_out", j, traj.cart[j]);
  traj.trq.clear();
  if (_cfg.is_trq_on) {
    traj.trq.assign((size_t)J, std::vector<double>());
    for (int j = 0; j < J; ++j) pull("trq_out", j, traj.trq[j]);
  }
  traj.nCartOut = nCart;
  traj.nPts = (unsigned)nOut;
  traj.sres = sres;
  return 0;
}

int BA::writeOutputData(Traj &traj) {
  if (traj.theta.empty()) {
    printf("trajWrite(): myTraj is empty; no file was written.\n");
    return -1;
  }
  const int J = _cfg.n_joints, C = _cfg.n_cart;
  const int n = (int)traj.theta[0].size();
  const bool cartFull = (int)traj.cart.size() == C && C > 0 && (int)traj.cart[0].size() == n;
  const bool trqFull = _cfg.is_trq_on && !traj.trq.empty() && !traj.trq[0].empty();
  std::vector<float> th((size_t)J * n), ca, tq;
  for (int j = 0; j < J; ++j)
    for (int i = 0; i < n; ++i) th[(size_t)j * n + i] = (float)traj.theta[j][i];
  if (cartFull) {
    ca.resize((size_t)C * n);
    for (int j = 0; j < C; ++j)
      for (int i = 0; i < n; ++i) ca[(size_t)j * n + i] = (float)traj.cart[j][i];
  }
  if (trqFull) {
    tq.resize((size_t)J * n);
    for (int j = 0; j < J; ++j)
      for (int i = 0; i < n; ++i) tq[(size_t)j * n + i] = (float)traj.trq[j][i];
  }
  std::string fn = _OutputFolder + "traj_out.dat";
  batotp_write_traj_bin(fn.c_str(), traj.sres, traj.nPts, J, th.data(), C, cartFull ? ca.data() : nullptr,
                        trqFull ? tq.data() : nullptr, n);
  if (!_cfg.is_bin_file) {  // ba.cpp:2514-2518 (written from the FP64 values, like the reference)
    fn = _OutputFolder + "traj_out.csv";
    FILE *fid = fopen(fn.c_str(), "w");
    if (fid) {
      for (size_t i = 0; i + 1 < traj.trajFileHeader.size(); ++i) fprintf(fid, "%s, ", traj.trajFileHeader[i].c_str());
      if (!traj.trajFileHeader.empty()) fprintf(fid, "%s\n", traj.trajFileHeader.back().c_str());
      for (int i = 0; i < n; ++i) {
        fprintf(fid, "%8.3f", i * traj.sres);
        for (int j = 0; j < J; ++j) fprintf(fid, ", %11.6f", traj.theta[j][i]);
        if (cartFull)
          for (int j = 0; j < C; ++j) fprintf(fid, ", %9.6f", traj.cart[j][i]);
        fprintf(fid, "\n");
      }
      fclose(fid);
    }
  }
  if (_cfg.is_sdot_out && !_cfg.is_interp_only && traj.myMVChist.s.size() >= 2) {
    fn = _OutputFolder + "s-sdot.dat";
    std::vector<float> f[4];
    for (int k = 0; k < 2; ++k) {
      f[2 * k].assign(traj.myMVChist.s[k].begin(), traj.myMVChist.s[k].end());
      f[2 * k + 1].assign(traj.myMVChist.sdot[k].begin(), traj.myMVChist.sdot[k].end());
    }
    batotp_write_s_sdot(fn.c_str(), traj.sres, (int)f[0].size(), f[0].data(), f[1].data(), (int)f[2].size(),
                        f[2].data(), f[3].data());
  }
  printf("\nOutput trajectory is %.3f sec.\n", (traj.nPts - 1) * traj.sres);
  return 0;
}

int BA::optimize(Traj &traj) {  // ba.cpp:2538-2573
  _errorOptimization = NO_ERROR;
  if (interpInputData(traj) == -1) return -1;
  if (traj.nPts < 4) return -1;
  setIntegDir(-1);
  setIsLastSweep(false);
  if (sweep(traj) == -1) return -1;
  setIntegDir(1);
  setIsLastSweep(true);
  if (sweep(traj) == -1) return -1;
  interpOutputData(traj);
  return 0;
}

}  // namespace BATOTP
