/* ba_oracle.h — CPU restatement of batotp's Bisection Algorithm path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under batotp_b200/ (the product) may
 * include, link or call this; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do (as the checker).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks this
 * restatement bit-for-bit against (a) the reference sources compiled in place
 * into oracle/_ref (oracle/Makefile) on the five stock input folders and on
 * seeded synthetic paths, and (b) the sha256 fingerprints of the prebuilt
 * /root/reference/bin/batest outputs recorded in SURVEY.md §8c
 * (tests/golden/stock_fingerprints.json).
 *
 * Every function cites the reference lines it follows.
 */
#ifndef BA_ORACLE_H
#define BA_ORACLE_H

#include <stdint.h>
#include "../include/batotp_cfg.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_traj orc_traj;

/* status bits shared with the product's per-trajectory status word */
#define ORC_OK 0
#define ORC_ERR (-1)

orc_traj *orc_new(const batotp_cfg *cfg);
void orc_free(orc_traj *t);

/* what trajReadBIN/trajReadCSV leave in Traj (ba.cpp:2257-2461): float32 payload
 * widened to double, coordinate-major [coord][n0]; NULL = that block absent;
 * timestamp NULL for BIN files. */
int orc_load_raw(orc_traj *t, int n0, double tres, const float *theta, const float *cart,
                 const double *timestamp);
int orc_load_raw_f64(orc_traj *t, int n0, double tres, const double *theta, const double *cart,
                     const double *timestamp);

int orc_interp_input(orc_traj *t);                        /* ba.cpp:95-316   */
int orc_sweep(orc_traj *t, int integ_dir, int is_last);   /* ba.cpp:979-1195 */
int orc_interp_output(orc_traj *t);                       /* ba.cpp:1661-1931 */
int orc_optimize(orc_traj *t);                            /* ba.cpp:2538-2573 */

/* derived per-sample max-velocity curve: for every grid point k
 * evaluate the spline partials at sC[k], apply the velocity limits without the
 * MVC term starting from sdot_start, then bisect on the acceleration constraints.
 * Must be called after orc_interp_input and before the sweeps. */
int orc_mvc_per_sample(orc_traj *t, double sdot_start, double *sdot_out, int cap);

/* cfg.dyn_source = 1: the generalized forces tau = a1*sddot + a2*sdot^2 + a3*sdot + a4 of a serial robot come from
 * the caller's point function instead of Robot::call_dynSerial (robot.cpp:349-360; same arguments: the joint values
 * and their first / second derivatives at one point in, a1..a4 out).  orc_demo_dyn_rr restates dynRR (robot.cpp:
 * 377-431) behind that signature, so that the plug-in path can be pinned to the reference on the RR folder;
 * orc_demo_dyn_serial is a decoupled n-joint model (inertia, viscous friction, a gravity-like term) for robots the
 * reference has no model for (the KUKA torque variant of SURVEY 8d C3). */
typedef void (*orc_dyn_fn)(void *user, int n_joints, const double *theta, const double *thetaD, const double *thetaD2,
                           double *a1, double *a2, double *a3, double *a4);
void orc_set_dyn_callback(orc_traj *t, orc_dyn_fn fn, void *user);
void orc_demo_dyn_rr(void *user, int n_joints, const double *theta, const double *thetaD, const double *thetaD2,
                     double *a1, double *a2, double *a3, double *a4);
void orc_demo_dyn_serial(void *user, int n_joints, const double *theta, const double *thetaD, const double *thetaD2,
                         double *a1, double *a2, double *a3, double *a4);

/* vectors: returns length (copies min(len,cap)); -1 unknown name.
 * names: theta thetaD thetaD2 cart cartD cartD2 trq a1 a2 a3 a4 (idx = coordinate)
 *        sMVC sdot tMVC sC ptsOrig hist_s0 hist_sdot0 hist_s1 hist_sdot1 (idx ignored)
 *        thetaC_y thetaC_m cartC_y cartC_m a1C_y a1C_m ... (knot values / 2nd-derivative solution)
 *        flags0 flags1 (per-step switching flags of the rev / fwd sweep, as doubles) */
int orc_get_vec(const orc_traj *t, const char *name, int idx, double *buf, int cap);
/* scalars: nPts nPtsC sres sresC vFact aFact tTotalTraj sLastSec nRev nFwd tRev tFwd
 *          nCart outRes integRes nA5 nA4 nA2 nSteps errorOptimization */
double orc_get_scalar(const orc_traj *t, const char *name);

/* pack exactly what trajWriteBIN / sdotWrite put in traj_out.dat / s-sdot.dat
 * (ba.cpp:2582-2651, 2726-2759); returns byte count (or needed size if cap too small) */
long orc_pack_traj_out(const orc_traj *t, unsigned char *buf, long cap);
long orc_pack_s_sdot(const orc_traj *t, unsigned char *buf, long cap);

/* batch CPU baseline ("port"): n_threads workers, one trajectory at a time each,
 * timing interp_input + 2 sweeps + interp_output only.  theta/cart: [B][coord][n0] f32.
 * outputs (any may be NULL): t_total[B], n_rev[B], n_fwd[B], n_out[B], status[B],
 * theta_out [B][J][out_cap] f32.  returns elapsed seconds. */
double orc_batch_run(const batotp_cfg *cfg, int B, int n0, double tres, const float *theta,
                     const float *cart, int n_threads, double *t_total, int *n_rev, int *n_fwd,
                     int *n_out, int *status, float *theta_out, int out_cap);

#ifdef __cplusplus
}
#endif
#endif
