// ba.h — drop-in host facade: the BATOTP::BA entry points batest calls (reference
// batotp/ba.h:168-255, test/main.cpp:30-116), implemented on top of the extern "C" CUDA
// library (include/batotp_cuda.h).  Same method names, argument meaning and 0 / -1 return
// convention; Traj is the same in/out carrier with the fields a caller reads.
//
// Differences a maintainer should know (all documented in INTEGRATION.md):
//  * the two sweeps run in ONE device launch: sweep() with integDir = -1 runs both and hands
//    back the reverse curve, the following sweep() with integDir = +1 hands back the forward
//    curve (calling them in the other order returns -1);
//  * Traj holds results only (no spline-coefficient members); thetaD/thetaD2/cartD/cartD2 are
//    not filled after interpOutputData (the reference leaves oddly scaled values there).
#ifndef BATOTP_B200_BA_H
#define BATOTP_B200_BA_H

#include <array>
#include <string>
#include <vector>

#include "../../include/batotp_cuda.h"

namespace BATOTP {

struct Traj {  // reference ba.h:59-153 (result-bearing members)
  double tresInput = 0;
  double sres = 0;
  unsigned int nPts = 0;
  double tTotalTraj = 0;
  std::string trajFileName;
  std::vector<std::string> trajFileHeader;
  std::vector<double> timestamp;
  std::vector<std::vector<double>> theta, cart, trq;  // [coordinate][point]
  std::vector<double> sMVC, tMVC, sdot;                // s-sdot curve of the last sweep
  struct MVChist {
    std::vector<std::vector<double>> s, sdot;  // [0] after reverse, [1] after forward integration
  } myMVChist;
  double sLastSec = 0;
  // raw file payloads as read (float32 for BIN files, float64 for CSV files)
  std::vector<float> rawTheta32, rawCart32;
  std::vector<double> rawTheta64, rawCart64;
  int nRaw = 0;
  // counts reported by the device
  int nGrid = 0, nRev = 0, nFwd = 0, nCartOut = 0, status = 0;
};

class BA {
 public:
  BA(void);
  ~BA(void);
  BA(const BA &) = delete;
  BA &operator=(const BA &) = delete;

  enum ErrorOptimization { NO_ERROR, MAX_INTEGRATION_TIME };

  struct Config {  // reference ba.h:213-255, same names and defaults
    std::string robotTypeStr = "UR";
    bool isParallelMech = false;
    int nJoints = 6;
    int nCart = 6;
    std::string trajFileName = "urtraj.csv";
    bool isBinFile = false;
    std::string pathType = "BOTH";
    bool isJntVelConon = true;
    std::vector<double> jntVelLims = std::vector<double>(6, 190);
    bool isJntAccConOn = true;
    std::vector<double> jntAccLims = std::vector<double>(6, 500);
    bool isTrqConOn = false;
    std::vector<double> jntTrqMax = std::vector<double>(6, 0);
    std::vector<double> jntTrqMin = std::vector<double>(6, 0);
    bool isCartVelConOn = true;
    double cartVelMax = 0.4;
    bool isCarAccConOn = true;
    double cartAccMax = 5.0;
    double integRes = 0.016;
    double maxIntegTime = 60000;
    int inputDecimFact = 1;
    int smoothWindow = 1;
    bool is_sdotOut = false;
    double jntThresh = 1e-6;
    double cartThresh = 1e-6;
    std::vector<double> sWeights = {0, 0.1, 1};
    int scaleType = 2;
    double thetaNormRes = 0.01;
    double thetaNormRes2 = 0.01;
    double cartNormRes = 0.002;
    double cartNormRes2 = 0.002;
    double outRes = 0.008;
    int outSmoothFact = 1;
    bool isSVD = false;
    bool isPar2Ser = false;
  };

  int readConfigData(const char *filename);  // ba.cpp:1942
  int loadConfigData(const Config &conf);    // ba.cpp:2100
  int loadTrajectoryData(Traj &traj);        // ba.cpp:2206
  int interpInputData(Traj &traj);           // ba.cpp:95
  int sweep(Traj &traj);                     // ba.cpp:979
  int interpOutputData(Traj &traj);          // ba.cpp:1661
  int writeOutputData(Traj &traj);           // ba.cpp:2510
  int optimize(Traj &traj);                  // ba.cpp:2538

  inline void setIsLastSweep(bool v) { _isLastSweep = v; }
  inline void setIntegDir(int d) { _integDir = d; }
  inline void setIsInterpOnly(bool v) { _cfg.is_interp_only = v ? 1 : 0; }
  inline void setCartesianMaximalVelocity(const double &v) { _cfg.cart_vel_max = v; }
  inline void setCartesianMaximalAcceleration(const double &a) { _cfg.cart_acc_max = a; }
  void setJointMaximalVelocity(const std::vector<double> &v);
  void setJointMaximalAcceleration(const std::vector<double> &a);
  inline void setIsAutoIntegRes(const bool v) { _cfg.is_auto_integ_res = v ? 1 : 0; }
  inline void setHomeFolder(const std::string &f) { _HomeFolder = f; }
  inline void setInputFolder(const std::string &f) { _InputFolder = f; }
  inline void setOutputFolder(const std::string &f) { _OutputFolder = f; }
  // not in the reference: 1 = host-libm trig (bit parity), 0 = device trig (throughput)
  inline void setTrigMode(int m) { _cfg.trig_mode = m; }
  inline void setDevice(int d) { _device = d; }

  inline double getCartesianMaximalVelocity() const { return _cfg.cart_vel_max; }
  inline double getCartesianMaximalAcceleration() const { return _cfg.cart_acc_max; }
  std::vector<double> getJointMaximalVelocity() const;
  std::vector<double> getJointMaximalAcceleration() const;
  inline ErrorOptimization getErrorOptimization() const { return _errorOptimization; }
  inline double getOutTimeRes() const { return _cfg.out_res; }
  inline std::string getHomeFolder() const { return _HomeFolder; }
  inline std::string getInputFolder() const { return _InputFolder; }
  inline std::string getOutputFolder() const { return _OutputFolder; }

 private:
  batotp_cfg _cfg;
  batotp_handle _h = nullptr;
  int _device = 0;
  std::string _robotTypeStr, _trajFileName, _csvHeader;
  std::string _HomeFolder, _InputFolder, _OutputFolder;
  bool _isLastSweep = false;
  int _integDir = -1;
  bool _sweepsDone = false;
  ErrorOptimization _errorOptimization = NO_ERROR;

  int ensureDevice();
  int pull(const char *name, int row, std::vector<double> &v);
  int fillOutput(Traj &traj);
};

}  // namespace BATOTP
#endif
