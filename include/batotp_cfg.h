/* batotp_cfg.h — plain-data mirror of the options BATOTP::BA keeps as private
 * members after readConfigData()/loadConfigData().
 *
 * Reference: batotp/ba.h:261-302 (the members), batotp/ba.cpp:1958-2073 (how
 * config.dat fills them, including sWeights normalisation and NAN JntTrqMin),
 * batotp/ba.h:306-311 (_isInterpOnly, _isAutoIntegRes), batotp/robot.h:33-42
 * (robot-type and path-type codes).
 *
 * The struct crosses the extern "C" boundary by pointer, has no padding
 * surprises (ints first, doubles after, both 8-byte aligned blocks) and is the
 * single description of a run shared by the CUDA library, the C++ BA facade,
 * the Python ctypes mirror and the test oracle.
 */
#ifndef BATOTP_CFG_H
#define BATOTP_CFG_H

#ifdef __cplusplus
extern "C" {
#endif

#define BATOTP_MAX_DOF 7 /* joints (<=7) and Cartesian coordinates (<=7: xyz+quaternion) */

/* robot.h:33-37 */
enum { BATOTP_KUKA = 1, BATOTP_UR = 2, BATOTP_RR = 3, BATOTP_CSPR3DOF = 4, BATOTP_GENJNT = 5 };
/* robot.h:40-42 */
enum { BATOTP_JOINT = 1, BATOTP_CART = 2, BATOTP_BOTH = 3 };

typedef struct batotp_cfg {
  /* ---- integers / flags (32 x int32) ---- */
  int robot_type;        /* _robotType           ba.h:313 */
  int is_parallel;       /* _isParallelMech      ba.h:262 */
  int n_joints;          /* _nJoints             ba.h:263 */
  int n_cart;            /* _nCart               ba.h:264 (6 for UR on input; 7 internally after aa->quat) */
  int is_bin_file;       /* _isBINfile           ba.h:266 (host I/O only) */
  int path_type;         /* _pathType            ba.h:267 */
  int are_jnt_deg;       /* _areJointAnglesDegrees ba.h:270 */
  int is_jnt_vel_on;     /* _isJntVelConOn       ba.h:271 (NB: sdotLim applies the limits regardless, ba.cpp:1219) */
  int is_jnt_acc_on;     /* _isJntAccConOn       ba.h:273 */
  int is_trq_on;         /* _isTrqConOn          ba.h:275 */
  int is_cart_vel_on;    /* _isCartVelConOn      ba.h:278 */
  int is_cart_acc_on;    /* _isCartAccConOn      ba.h:280 */
  int input_decim_fact;  /* _inputDecimFact      ba.h:288 */
  int smooth_window;     /* _smoothWindow        ba.h:289 */
  int is_sdot_out;       /* is_sdotOut           ba.h:290 */
  int scale_type;        /* _scaleType           ba.h:294 */
  int is_svd;            /* _isSVD               ba.h:301 (1 is not supported on the device: returns -1) */
  int is_par2ser;        /* _isPar2Ser           ba.h:302 */
  int is_interp_only;    /* _isInterpOnly        ba.h:306 (re-sample the path at outRes only, ba.cpp:139-159; needs joint rows) */
  int is_auto_integ_res; /* _isAutoIntegRes      ba.h:309 (batest forces 0, test/main.cpp:53) */
  int trig_mode;         /* sin/cos of the kinematics / dynamics point functions (robot.cpp:130-136, 196-199, 408-419,
                            util.cpp:544-549), all evaluated on the device unless 2:
                            0: CUDA's sin/cos (1-2 ulp: results within the bisection tolerance, not bit-identical);
                            1: strict parity — a bit-identical device port of the host libm's sin/cos
                               (glibc 2.39 x86-64; batotp_cuda_selftest_trig verifies it against the running host);
                            2: strict parity the slow way — the host evaluates those point functions with its own
                               libm between device stages (for a host whose libm is not the one ported).
                            In modes 1 and 2 the atan2 of the axis-angle output rows (util.cpp:574) is applied by
                            the host to the final rows */
  int dyn_source;        /* torque limits for a robot batotp has no dynamic model for (the reference's
                            README asks a user to add the model to robot.cpp): 0 = the built-in models (dynRR robot.cpp:377,
                            dynCSPR3DOF robot.cpp:487); 1 = a1..a4 come from the caller's point function
                            (batotp_cuda_set_dyn_callback, same contract as Robot::call_dynSerial robot.cpp:349-360);
                            serial mechanisms only */
  int reserved_i[10];
  /* ---- doubles ---- */
  double jnt_vel_max[BATOTP_MAX_DOF]; /* _JntVelMax */
  double jnt_acc_max[BATOTP_MAX_DOF]; /* _JntAccMax */
  double jnt_trq_max[BATOTP_MAX_DOF]; /* _JntTrqMax */
  double jnt_trq_min[BATOTP_MAX_DOF]; /* _JntTrqMin (NAN already replaced by -max) */
  double cart_vel_max;                /* _CartVelMax */
  double cart_acc_max;                /* _CartAccMax */
  double integ_res;                   /* _integRes */
  double max_integ_time;              /* _maxIntegTime */
  double jnt_thresh;                  /* _jntThresh */
  double cart_thresh;                 /* _cartThresh (quadraticRadThresh = cart_thresh^2) */
  double s_weights[3];                /* _sWeights, already normalised to sum 1 */
  double theta_norm_res;              /* _thetaNormRes */
  double theta_norm_res2;             /* _thetaNormRes2 */
  double cart_norm_res;               /* _cartNormRes */
  double cart_norm_res2;              /* _cartNormRes2 */
  double out_res;                     /* _outRes */
  double out_smooth_fact;             /* _outSmoothFact (double in the reference, ba.h:300) */
  double reserved_d[8];
} batotp_cfg;

#ifdef __cplusplus
}
#endif
#endif /* BATOTP_CFG_H */
