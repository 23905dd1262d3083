/* ba_oracle.c — CPU restatement (C99) of the batotp Bisection Algorithm path.
 *
 * TEST INFRASTRUCTURE ONLY — see ba_oracle.h.  Parity status: PINNED against the
 * reference sources built into oracle/_ref and the prebuilt-binary fingerprints.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off -fno-fast-math (oracle/Makefile).
 * The reference is built by compile.sh with plain x86-64 SSE2 arithmetic (no FMA),
 * so every expression below keeps the reference's operand order and association;
 * min/max follow std::min/std::max ((b<a)?b:a and (a<b)?b:a).
 *
 * Layout of this file (reference file:line in each function header):
 *   1. small vectors        2. Spline (spline.cpp)      3. util.cpp numerics
 *   4. Robot (robot.cpp)    5. BA input interpolation   6. BA sweeps
 *   7. BA output interp.    8. accessors / packers / batch runner
 */
#include "ba_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAXD BATOTP_MAX_DOF

/* ------------------------------------------------------------------ 1. vectors */
typedef struct {
  double *p;
  int n, cap;
} dvec;

static void dv_reserve(dvec *v, int cap) {
  if (cap > v->cap) {
    int nc = v->cap ? v->cap : 16;
    while (nc < cap) nc *= 2;
    v->p = (double *)realloc(v->p, (size_t)nc * sizeof(double));
    v->cap = nc;
  }
}
/* std::vector<double>::resize(n): keeps the prefix, new elements are 0 */
static void dv_resize(dvec *v, int n) {
  dv_reserve(v, n);
  for (int i = v->n; i < n; ++i) v->p[i] = 0.0;
  v->n = n;
}
static void dv_assign(dvec *v, int n, double x) {
  dv_reserve(v, n);
  for (int i = 0; i < n; ++i) v->p[i] = x;
  v->n = n;
}
static void dv_copy(dvec *d, const dvec *s) {
  dv_reserve(d, s->n);
  if (s->n) memcpy(d->p, s->p, (size_t)s->n * sizeof(double));
  d->n = s->n;
}
static void dv_iota(dvec *v, double start) {
  for (int i = 0; i < v->n; ++i) v->p[i] = start + (double)i;
}
/* util.h:101-107: C*a multiplies a in place (quirk Q1) */
static void dv_scale(dvec *v, double C) {
  for (int i = 0; i < v->n; ++i) v->p[i] = C * v->p[i];
}
static void dv_free(dvec *v) {
  free(v->p);
  v->p = NULL;
  v->n = v->cap = 0;
}

static inline double dmin(double a, double b) { return (b < a) ? b : a; } /* std::min */
static inline double dmax(double a, double b) { return (a < b) ? b : a; } /* std::max */
static inline int imax(int a, int b) { return (a < b) ? b : a; }
static inline int imin(int a, int b) { return (b < a) ? b : a; }
static inline int sgn_d(double v) { return (0.0 < v) - (v < 0.0); } /* util.h:94-96 */

/* ------------------------------------------------------------------ 2. Spline */
typedef struct {
  dvec c0, c1, c2, c3;
} splc;
typedef struct {
  int *seg;
  double *tau;
  int n, cap;
} segs_t;

static void splc_free(splc *c) {
  dv_free(&c->c0);
  dv_free(&c->c1);
  dv_free(&c->c2);
  dv_free(&c->c3);
}
static void segs_resize(segs_t *s, int n) {
  if (n > s->cap) {
    s->seg = (int *)realloc(s->seg, (size_t)n * sizeof(int));
    s->tau = (double *)realloc(s->tau, (size_t)n * sizeof(double));
    s->cap = n;
  }
  s->n = n;
}
static void segs_free(segs_t *s) {
  free(s->seg);
  free(s->tau);
  memset(s, 0, sizeof(*s));
}

/* spline.cpp:252-276 — "natural": m0 = 0 but the last unknown is an ordinary row (Q4) */
static void tridiag_natural(double *d, int npts) {
  int n = npts - 1;
  double a = 1.0, b = 4.0;
  double *c = (double *)malloc((size_t)(n > 2 ? n : 2) * sizeof(double));
  for (int i = 0; i < n; ++i) c[i] = 1.0;
  c[1] /= b;
  d[1] /= b;
  for (int i = 2; i < n; ++i) {
    c[i] /= b - a * c[i - 1];
    d[i] = (d[i] - a * d[i - 1]) / (b - a * c[i - 1]);
  }
  d[n] = (d[n] - a * d[n - 1]) / (b - a * c[n - 1]);
  for (int i = n; i > 1; --i) d[i - 1] -= c[i - 1] * d[i];
  free(c);
}

/* spline.cpp:225-243 — "clamped": back-substitution starts at n-3 (Q4) */
static void tridiag_clamped(double *d, int n) {
  double a = 1.0;
  double *c = (double *)malloc((size_t)n * sizeof(double));
  double *b = (double *)malloc((size_t)n * sizeof(double));
  for (int i = 0; i < n; ++i) {
    c[i] = 1.0;
    b[i] = 4.0;
  }
  b[0] = 2.0;
  b[n - 1] = 2.0;
  c[0] /= b[0];
  d[0] /= b[0];
  for (int i = 1; i < n; ++i) {
    c[i] /= b[i] - a * c[i - 1];
    d[i] = (d[i] - a * d[i - 1]) / (b[i] - a * c[i - 1]);
  }
  for (int i = n - 2; i-- > 0;) d[i] -= c[i] * d[i + 1];
  free(c);
  free(b);
}

/* spline.cpp:168-211.  `sol_out` (optional) receives the second-derivative solution. */
static int spl_coeffs(const dvec *y, splc *yC, int clamped, dvec *sol_out) {
  int npts = y->n;
  if (npts != yC->c0.n) {
    dv_resize(&yC->c0, npts);
    dv_resize(&yC->c1, npts);
    dv_resize(&yC->c2, npts);
    dv_resize(&yC->c3, npts);
  }
  double *sol = (double *)calloc((size_t)npts, sizeof(double));
  for (int i = 1; i < npts - 1; ++i) sol[i] = 6 * (y->p[i - 1] - 2 * y->p[i] + y->p[i + 1]);
  if (clamped)
    tridiag_clamped(sol, npts);
  else
    tridiag_natural(sol, npts);
  for (int i = 0; i < npts - 1; ++i) {
    yC->c3.p[i] = (sol[i + 1] - sol[i]) / 6.0;
    yC->c2.p[i] = sol[i] / 2.0;
    yC->c1.p[i] = y->p[i + 1] - y->p[i] - (sol[i + 1] + 2 * sol[i]) / 6.0;
    yC->c0.p[i] = y->p[i];
  }
  if (sol_out) {
    dv_resize(sol_out, npts);
    memcpy(sol_out->p, sol, (size_t)npts * sizeof(double));
  }
  free(sol);
  return 0;
}

/* spline.cpp:56-99 */
static int spl_find_segs(const dvec *aIn, const dvec *aOut, segs_t *sg) {
  int nIn = aIn->n, nOut = aOut->n;
  double *den = (double *)calloc((size_t)nIn, sizeof(double));
  segs_resize(sg, nOut);
  int cur = 0;
  for (int i = 0; i < nOut; ++i) {
    double a = aOut->p[i];
    for (;;) {
      if (a < aIn->p[cur + 1] || cur == nIn - 2) {
        sg->seg[i] = cur;
        break;
      }
      cur++;
    }
  }
  for (int i = 0; i < nIn - 1; ++i) {
    den[i] = aIn->p[i + 1] - aIn->p[i];
    if (den[i] < 1e-20) {
      free(den);
      return -1;
    }
  }
  for (int i = 0; i < nOut; ++i) {
    int s = sg->seg[i];
    sg->tau[i] = (aOut->p[i] - aIn->p[s]) / den[s];
  }
  free(den);
  return 0;
}

/* spline.cpp:108-120 */
static void spl_interp_linear(dvec *b, const segs_t *sg) {
  int n = sg->n;
  double *o = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
  for (int i = 0; i < n; ++i) {
    int s = sg->seg[i];
    o[i] = b->p[s] + (b->p[s + 1] - b->p[s]) * sg->tau[i];
  }
  dv_resize(b, n);
  memcpy(b->p, o, (size_t)n * sizeof(double));
  free(o);
}

/* spline.cpp:129-155 (bD / bD2 may be NULL for the reference's dummy outputs) */
static void spl_interp_spline(dvec *b, dvec *bD, dvec *bD2, const splc *bC, const segs_t *sg,
                              double tfact) {
  int n = sg->n;
  dv_resize(b, n);
  if (bD) dv_resize(bD, n);
  if (bD2) dv_resize(bD2, n);
  double vfact = 1.0 / tfact;
  double afact = vfact * vfact;
  for (int i = 0; i < n; ++i) {
    int j = sg->seg[i];
    double tau = sg->tau[i];
    double tau2 = tau * tau, tau3 = tau2 * tau;
    double c3 = bC->c3.p[j], c2 = bC->c2.p[j], c1 = bC->c1.p[j], c0 = bC->c0.p[j];
    b->p[i] = c3 * tau3 + c2 * tau2 + c1 * tau + c0;
    if (bD) bD->p[i] = (3 * c3 * tau2 + 2 * c2 * tau + c1) * vfact;
    if (bD2) bD2->p[i] = (6 * c3 * tau + 2 * c2) * afact;
  }
}

/* ------------------------------------------------------------------ 3. util.cpp */
/* util.cpp:254-288 */
static void u_smooth(dvec *x, int w) {
  int n = x->n;
  w = imin(w, n);
  int wMid = w / 2 + w % 2 - 1;
  w = 2 * wMid + 1;
  double *x2 = (double *)calloc((size_t)n, sizeof(double));
  x2[0] = x->p[0];
  x2[n - 1] = x->p[n - 1];
  for (int i = 1; i < wMid; ++i) {
    double xt = 0, xte = 0;
    int nT = 2 * i + 1;
    for (int j = 0; j < nT; ++j) {
      xt += x->p[j];
      xte += x->p[n - j - 1];
    }
    x2[i] = xt / nT;
    x2[n - i - 1] = xte / nT;
  }
  for (int i = wMid; i < n - wMid; ++i) {
    double xt = 0;
    for (int j = i - wMid; j < i + wMid + 1; ++j) xt += x->p[j];
    x2[i] = xt / w;
  }
  memcpy(x->p, x2, (size_t)n * sizeof(double));
  free(x2);
}

/* util.cpp:343-352 */
static void u_decimate(dvec *x, int w) {
  int nIn = x->n;
  int nOut = (nIn - 1) / w + 1;
  for (int i = 0; i < nOut; ++i) x->p[i] = x->p[w * i];
  if (w * (nOut - 1) + 1 != nIn) x->p[nOut - 1] = x->p[nIn - 1];
  dv_resize(x, nOut);
}

/* util.cpp:361-383 */
static int u_solve_quadratic(double A, double B, double C, double *s1, double *s2) {
  if (fabs(A) < 1e-308) {
    if (fabs(B) < 1e-308) return -2;
    *s1 = -C / B;
    *s2 = *s1;
    return 0;
  }
  double rad = B * B - 4 * A * C;
  if (rad < 0) return -1;
  double den = 2 * A;
  double F1 = -B / den, F2 = sqrt(rad) / den;
  *s1 = F1 + F2;
  *s2 = F1 - F2;
  return 0;
}

/* util.cpp:413-442 with isSVD=0: x = A.lu().solve(b).  Eigen is not in the reference
 * tree (README.md:32-50, un-vendored, "3.3.4"); its PartialPivLU for small dynamic
 * matrices is restated here exactly as in oracle/eigen_standin/Eigen/Dense:
 * first-max partial pivoting, sub-column divided by the pivot, rank-1 trailing
 * update, column-oriented forward/back substitution.  Pinned by the prebuilt
 * bin/batest CSPR3DOF fingerprints (the only shipped config reaching it). */
static int u_solve_lin_sys(int n, double A[MAXD][MAXD], const double *b, double *x, int isSVD) {
  if (isSVD) return -1; /* not restated: no shipped config uses it */
  double lu[MAXD][MAXD];
  int perm[MAXD];
  for (int i = 0; i < n; ++i) {
    perm[i] = i;
    for (int j = 0; j < n; ++j) lu[i][j] = A[i][j];
  }
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(lu[k][k]);
    for (int i = k + 1; i < n; ++i) {
      double v = fabs(lu[i][k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    if (p != k) {
      for (int j = 0; j < n; ++j) {
        double t = lu[k][j];
        lu[k][j] = lu[p][j];
        lu[p][j] = t;
      }
      int t = perm[k];
      perm[k] = perm[p];
      perm[p] = t;
    }
    if (best != 0.0) {
      double piv = lu[k][k];
      for (int i = k + 1; i < n; ++i) lu[i][k] /= piv;
    }
    for (int i = k + 1; i < n; ++i)
      for (int j = k + 1; j < n; ++j) lu[i][j] -= lu[i][k] * lu[k][j];
  }
  for (int i = 0; i < n; ++i) x[i] = b[perm[i]];
  for (int i = 0; i < n; ++i)
    for (int r = i + 1; r < n; ++r) x[r] -= x[i] * lu[r][i];
  for (int i = n - 1; i >= 0; --i) {
    x[i] /= lu[i][i];
    for (int r = 0; r < i; ++r) x[r] -= x[i] * lu[r][i];
  }
  return 0;
}

/* util.cpp:452-524: x drives, y follows (ny may be 0) */
static void u_rem_close_pts(dvec *x, int nx, dvec *y, int ny, double thresh) {
  double thrSQ = thresh * thresh;
  int nPts = x[0].n;
  double *isRem = (double *)calloc((size_t)(nPts > 0 ? nPts : 1), sizeof(double));
  for (;;) {
    int anyRem = 0;
    for (int i = 1; i < nPts; ++i) {
      double sum = 0;
      for (int j = 0; j < nx; ++j) {
        double d = x[j].p[i] - x[j].p[i - 1];
        sum += d * d;
      }
      if (sum < thrSQ && !isRem[i - 1]) {
        isRem[i] = 1;
        anyRem = 1;
      }
    }
    if (isRem[nPts - 1] && nPts > 2) {
      isRem[nPts - 1] = 0;
      isRem[nPts - 2] = 1;
      isRem[nPts - 3] = 0;
    }
    if (!anyRem) break;
    int cur = 0;
    for (int i = 0; i < nPts; ++i) {
      if (!isRem[i]) {
        for (int j = 0; j < nx; ++j) x[j].p[cur] = x[j].p[i];
        for (int j = 0; j < ny; ++j) y[j].p[cur] = y[j].p[i];
        cur++;
      }
    }
    nPts = cur;
    for (int j = 0; j < nx; ++j) dv_resize(&x[j], cur);
    for (int j = 0; j < ny; ++j) dv_resize(&y[j], cur);
    for (int i = 0; i < nPts; ++i) isRem[i] = 0;
  }
  free(isRem);
}

/* util.cpp:534-554 */
static void u_aa2q(const double aa[3], double q[4]) {
  double theta = sqrt(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
  if (theta < 1e-6) {
    q[0] = 1.0;
    q[1] = q[2] = q[3] = 0.0;
  } else {
    double sh = sin(0.5 * theta);
    q[0] = cos(0.5 * theta);
    for (int i = 0; i < 3; ++i) q[i + 1] = aa[i] * sh / theta;
  }
}
/* util.cpp:563-580 */
static void u_q2aa(const double q[4], double aa[3]) {
  double nrm = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (nrm < 1e-6) {
    aa[0] = aa[1] = aa[2] = 0.0;
  } else {
    double theta = 2.0 * atan2(nrm, q[0]) / nrm;
    for (int i = 0; i < 3; ++i) aa[i] = theta * q[i + 1];
  }
}

/* ------------------------------------------------------------------ state */
static const double C_PI = 3.14159265358979323846; /* config.h:27-30 */
#define C_DEG2RAD (C_PI / 180.0)
#define C_RAD2DEG (180.0 / C_PI)
static const double C_G = 9.81;

typedef struct {
  unsigned char *p;
  int n, cap;
} bvec;
static void bv_push(bvec *v, unsigned char x) {
  if (v->n == v->cap) {
    v->cap = v->cap ? 2 * v->cap : 1024;
    v->p = (unsigned char *)realloc(v->p, (size_t)v->cap);
  }
  v->p[v->n++] = x;
}

struct orc_traj {
  orc_dyn_fn dynFn; /* cfg.dyn_source = 1: the caller's point function for a1..a4 */
  void *dynUser;
  batotp_cfg cfg; /* option values as loaded; the mutable ones are copied below (Q9) */
  /* BA private state (ba.h:261-331) that changes during a run */
  int nJoints, nCart;
  int isParallelMech, isParallelMechOrig, isGenericRobot;
  int isInterpolated, isLastSweep, integDir, dynDim;
  int scaleType;
  double sWeights[3];
  double integRes, outRes, outSmoothFact;
  double quadThresh; /* _quadraticRadThresh = cartThresh^2 (ba.cpp:2048) */
  double sdotMin;    /* _sdotMin */
  double B[6][6];    /* _B[k][j] (ba.cpp:58-63) */
  double dA[6];      /* _dA, zeroed by the first sweep (Q1) */
  int errorOptimization;
  double pmat[3][3];
  int havePmat;
  /* Traj (ba.h:59-153) */
  double tresInput, sres;
  int nPts;
  double tTotalTraj;
  dvec timestamp, ptsOrig;
  splc ptsOrigC;
  dvec theta[MAXD], thetaD[MAXD], thetaD2[MAXD];
  dvec cart[MAXD], cartD[MAXD], cartD2[MAXD];
  int thetaRows, cartRows;
  dvec trq[MAXD], a1[MAXD], a2[MAXD], a3[MAXD], a4[MAXD];
  int trqRows;
  int curSegMVC;
  double tauMVC, sCur, sdotCur;
  int sdotLimTypeT;
  double sddotH, sddotL;
  dvec sMVC, tMVC, sdot;
  dvec histS[2], histSdot[2];
  double thetapt[MAXD], thetaDpt[MAXD], thetaD2pt[MAXD];
  double a1pt[MAXD], a2pt[MAXD], a3pt[MAXD], a4pt[MAXD];
  double cartpt[MAXD], cartDpt[MAXD], cartD2pt[MAXD];
  double CartAccCoeffs[3];
  double Apt[MAXD][MAXD];
  double sLastSec;
  int isOn_sdot;
  int nPtsC;
  double sresC, vFact, aFact;
  int curSegC;
  double tauC;
  dvec sC;
  splc thetaC[MAXD], cartC[MAXD], a1C[MAXD], a2C[MAXD], a3C[MAXD], a4C[MAXD];
  dvec thetaM[MAXD], cartM[MAXD], a1M[MAXD], a2M[MAXD], a3M[MAXD], a4M[MAXD]; /* 2nd-deriv solutions */
  int aCRows; /* a1C.size() */
  /* instrumentation (not in the reference) */
  bvec flags[2];
  long nA5, nA4, nA2, nSteps;
  int nRev, nFwd;
  double tRev, tFwd;
  int numericFail;
};

orc_traj *orc_new(const batotp_cfg *cfg) {
  orc_traj *t = (orc_traj *)calloc(1, sizeof(orc_traj));
  t->cfg = *cfg;
  t->nJoints = cfg->n_joints;
  t->nCart = cfg->n_cart;
  t->isParallelMech = cfg->is_parallel;
  t->isParallelMechOrig = cfg->is_parallel;
  t->isGenericRobot = (cfg->robot_type == BATOTP_GENJNT);
  t->scaleType = cfg->scale_type;
  for (int i = 0; i < 3; ++i) t->sWeights[i] = cfg->s_weights[i];
  t->integRes = cfg->integ_res;
  t->outRes = cfg->out_res;
  t->outSmoothFact = cfg->out_smooth_fact;
  t->quadThresh = cfg->cart_thresh * cfg->cart_thresh;
  t->sdotMin = DBL_MAX; /* ba.h:319 */
  /* ba.cpp:48-63 — the literals are kept as written there */
  double A[6] = {1. / 5, 3. / 10, 4. / 5, 8. / 9, 1.0, 1.0};
  for (int i = 0; i < 6; ++i) t->dA[i] = A[i];
  for (int i = 1; i < 6; ++i) t->dA[i] = A[i] - A[i - 1];
  t->dA[5] = 1.0e-6;
  double Bt[6][6] = {{1. / 5, 3. / 40, 44. / 45, 19372. / 6561, 9017. / 3168, 35. / 384},
                     {0, 9. / 40, -56. / 15, -25360. / 2187, -355. / 33, 0},
                     {0, 0, 32. / 9, 64448. / 6561, 46732. / 5247, 500. / 1113},
                     {0, 0, 0, -212. / 729, 49. / 176, 125. / 192},
                     {0, 0, 0, 0, -5103. / 18656, -2187. / 6784},
                     {0, 0, 0, 0, 0, 11. / 84}};
  memcpy(t->B, Bt, sizeof(Bt));
  t->sLastSec = 0;
  return t;
}

void orc_free(orc_traj *t) {
  if (!t) return;
  dv_free(&t->timestamp);
  dv_free(&t->ptsOrig);
  splc_free(&t->ptsOrigC);
  for (int i = 0; i < MAXD; ++i) {
    dv_free(&t->theta[i]); dv_free(&t->thetaD[i]); dv_free(&t->thetaD2[i]);
    dv_free(&t->cart[i]); dv_free(&t->cartD[i]); dv_free(&t->cartD2[i]);
    dv_free(&t->trq[i]); dv_free(&t->a1[i]); dv_free(&t->a2[i]); dv_free(&t->a3[i]); dv_free(&t->a4[i]);
    splc_free(&t->thetaC[i]); splc_free(&t->cartC[i]);
    splc_free(&t->a1C[i]); splc_free(&t->a2C[i]); splc_free(&t->a3C[i]); splc_free(&t->a4C[i]);
    dv_free(&t->thetaM[i]); dv_free(&t->cartM[i]);
    dv_free(&t->a1M[i]); dv_free(&t->a2M[i]); dv_free(&t->a3M[i]); dv_free(&t->a4M[i]);
  }
  dv_free(&t->sMVC); dv_free(&t->tMVC); dv_free(&t->sdot); dv_free(&t->sC);
  for (int i = 0; i < 2; ++i) {
    dv_free(&t->histS[i]);
    dv_free(&t->histSdot[i]);
    free(t->flags[i].p);
  }
  free(t);
}

/* ba.cpp:2257-2312 (BIN) / 2322-2461 (CSV): what the readers leave in Traj */
int orc_load_raw_f64(orc_traj *t, int n0, double tres, const double *theta, const double *cart,
                     const double *timestamp) {
  t->tresInput = tres;
  t->sres = tres;
  t->nPts = n0;
  t->thetaRows = 0;
  t->cartRows = 0;
  if (theta) {
    t->thetaRows = t->nJoints;
    for (int j = 0; j < t->nJoints; ++j) {
      dv_resize(&t->theta[j], n0);
      memcpy(t->theta[j].p, theta + (size_t)j * n0, (size_t)n0 * sizeof(double));
    }
  }
  if (cart) {
    t->cartRows = t->nCart;
    for (int j = 0; j < t->nCart; ++j) {
      dv_resize(&t->cart[j], n0);
      memcpy(t->cart[j].p, cart + (size_t)j * n0, (size_t)n0 * sizeof(double));
    }
  }
  t->timestamp.n = 0;
  if (timestamp) {
    dv_resize(&t->timestamp, n0);
    memcpy(t->timestamp.p, timestamp, (size_t)n0 * sizeof(double));
  }
  return 0;
}

int orc_load_raw(orc_traj *t, int n0, double tres, const float *theta, const float *cart,
                 const double *timestamp) {
  double *th = NULL, *ca = NULL;
  if (theta) {
    th = (double *)malloc((size_t)t->nJoints * n0 * sizeof(double));
    for (size_t i = 0; i < (size_t)t->nJoints * n0; ++i) th[i] = (double)theta[i];
  }
  if (cart) {
    ca = (double *)malloc((size_t)t->nCart * n0 * sizeof(double));
    for (size_t i = 0; i < (size_t)t->nCart * n0; ++i) ca[i] = (double)cart[i];
  }
  int r = orc_load_raw_f64(t, n0, tres, th, ca, timestamp);
  free(th);
  free(ca);
  return r;
}

/* ------------------------------------------------------------------ 4. Robot */
/* robot.cpp:291-322 */
static void rb_find_pmat(orc_traj *t) {
  double cible1[3] = {1.0941, -4.9074, 2.5542};
  double delta1[3] = {-0.765, 0.112, 3.74};
  double cible3[3] = {0.2098, 5.3409, 2.6236};
  double delta2[3] = {0.43, 0.125, 3.615};
  double p1[3], p2[3], p3[3] = {-5.9751, 0.1399, 6.1543};
  for (int i = 0; i < 3; ++i) {
    p1[i] = cible1[i] + delta1[i];
    p2[i] = cible3[i] + delta2[i];
  }
  int ind[3] = {1, 0, 2};
  for (int i = 0; i < 3; ++i) {
    int it = ind[i];
    t->pmat[i][0] = -p1[it];
    t->pmat[i][1] = -p2[it];
    t->pmat[i][2] = -p3[it];
  }
  double cen[3];
  for (int i = 0; i < 3; ++i) cen[i] = 1 / 3.0 * (t->pmat[i][0] + t->pmat[i][1] + t->pmat[i][2]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t->pmat[i][j] -= cen[i];
  t->havePmat = 1;
}

static void mat3_mul(double R[3][3], double A[3][3], double Bm[3][3]) {
  /* Eigen 3x3 product as restated in oracle/eigen_standin/Eigen/Core: left-to-right inner sum */
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = A[i][0] * Bm[0][j] + A[i][1] * Bm[1][j] + A[i][2] * Bm[2][j];
}

/* robot.cpp:105-176 */
static void rb_fwd_kin_kuka(orc_traj *t) {
  int nPts = t->theta[0].n;
  double tool[3] = {0, -.08, .545};
  double a0 = .3105, a1 = .4, a2 = .39;
  t->cartRows = 3;
  for (int i = 0; i < 3; ++i) dv_resize(&t->cart[i], nPts);
  for (int i = 0; i < nPts; ++i) {
    double c[7], s[7];
    for (int k = 0; k < 7; ++k) {
      double tk = C_DEG2RAD * t->theta[k].p[i];
      c[k] = cos(tk);
      s[k] = sin(tk);
    }
    double c1 = c[0], c2 = c[1], c3 = c[2], c4 = c[3], c5 = c[4], c6 = c[5], c7 = c[6];
    double s1 = s[0], s2 = s[1], s3 = s[2], s4 = s[3], s5 = s[4], s6 = s[5], s7 = s[6];
    double Q12[3][3] = {{c1 * c2, -s1, -c1 * s2}, {c2 * s1, c1, -s1 * s2}, {s2, 0, c2}};
    double Q34[3][3] = {{c3 * c4, -s3, c3 * s4}, {c4 * s3, c3, s3 * s4}, {-s4, 0, c4}};
    double Q567[3][3] = {{c5 * c6 * c7 - s5 * s7, -c7 * s5 - c5 * c6 * s7, -c5 * s6},
                         {c5 * s7 + c6 * c7 * s5, c5 * c7 - c6 * s5 * s7, -s5 * s6},
                         {c7 * s6, -s6 * s7, c6}};
    double Q1234[3][3], Q[3][3];
    mat3_mul(Q1234, Q12, Q34);
    mat3_mul(Q, Q1234, Q567);
    double x1 = a1 * Q12[0][2], y1 = a1 * Q12[1][2], z1 = a1 * Q12[2][2] + a0;
    double x2 = x1 + a2 * Q1234[0][2], y2 = y1 + a2 * Q1234[1][2], z2 = z1 + a2 * Q1234[2][2];
    double x3 = x2 + (Q[0][0] * tool[0] + Q[0][1] * tool[1] + Q[0][2] * tool[2]);
    double y3 = y2 + (Q[1][0] * tool[0] + Q[1][1] * tool[1] + Q[1][2] * tool[2]);
    double z3 = z2 + (Q[2][0] * tool[0] + Q[2][1] * tool[1] + Q[2][2] * tool[2]);
    t->cart[0].p[i] = x3;
    t->cart[1].p[i] = y3;
    t->cart[2].p[i] = z3;
  }
}

/* robot.cpp:185-202 */
static void rb_fwd_kin_rr(orc_traj *t) {
  int nPts = t->theta[0].n;
  double a1 = .4, a2 = .6;
  t->cartRows = 3;
  for (int i = 0; i < 3; ++i) dv_resize(&t->cart[i], nPts);
  for (int i = 0; i < nPts; ++i) {
    double th1 = C_DEG2RAD * t->theta[0].p[i];
    double th2 = C_DEG2RAD * t->theta[1].p[i];
    t->cart[0].p[i] = a1 * cos(th1) + a2 * cos(th1 + th2);
    t->cart[1].p[i] = a1 * sin(th1) + a2 * sin(th1 + th2);
  }
}

/* robot.cpp:74-94: returns -1 when the robot has no model */
static int rb_fwd_kin(orc_traj *t) {
  switch (t->cfg.robot_type) {
    case BATOTP_KUKA: rb_fwd_kin_kuka(t); return 0;
    case BATOTP_RR: rb_fwd_kin_rr(t); return 0;
    default: return -1;
  }
}

/* robot.cpp:243-278 (+ util.h:140-146 norm) */
static int rb_inv_kin(orc_traj *t) {
  if (t->cfg.robot_type != BATOTP_CSPR3DOF) return -1;
  int nDim = t->cartRows;
  int nPts = t->cart[0].n;
  t->thetaRows = 3;
  for (int i = 0; i < nDim && i < 3; ++i) dv_resize(&t->theta[i], nPts);
  if (!t->havePmat) rb_find_pmat(t);
  for (int i = 0; i < nPts; ++i) {
    double x = t->cart[0].p[i], y = t->cart[1].p[i], z = t->cart[2].p[i];
    for (int k = 0; k < 3; ++k) {
      double rv[3] = {x - t->pmat[0][k], y - t->pmat[1][k], z - t->pmat[2][k]};
      double sumSQ = 0.0;
      for (int q = 0; q < 3; ++q) sumSQ += rv[q] * rv[q];
      t->theta[k].p[i] = sqrt(sumSQ);
    }
  }
  return 0;
}

/* robot.cpp:377-431 */
static void rb_dyn_rr(orc_traj *t) {
  int nPts = t->theta[0].n;
  double A1 = .4, A2 = .6, m1 = 4, m2 = 8;
  for (int i = 0; i < 2; ++i) {
    dv_resize(&t->a1[i], nPts);
    dv_resize(&t->a2[i], nPts);
    dv_resize(&t->a3[i], nPts);
    dv_resize(&t->a4[i], nPts);
  }
  for (int i = 0; i < nPts; ++i) {
    double th1 = C_DEG2RAD * t->theta[0].p[i];
    double th2 = C_DEG2RAD * t->theta[1].p[i];
    double dth1 = C_DEG2RAD * t->thetaD[0].p[i];
    double dth2 = C_DEG2RAD * t->thetaD[1].p[i];
    double ddth1 = C_DEG2RAD * t->thetaD2[0].p[i];
    double ddth2 = C_DEG2RAD * t->thetaD2[1].p[i];
    double c1 = cos(th1), c2 = cos(th2), c12 = cos(th1 + th2);
    double A11 = .25 * m1 * A1 * A1 + m2 * (A1 * A1 + .25 * A2 * A2 + A1 * A2 * c2);
    double A12 = .5 * m2 * (.5 * A2 * A2 + A1 * A2 * c2);
    double A22 = .25 * m2 * A2 * A2;
    t->a1[0].p[i] = A11 * dth1 + A12 * dth2;
    t->a1[1].p[i] = A12 * dth1 + A22 * dth2;
    double ccFact = m2 * A1 * A2 * sin(th2);
    t->a2[0].p[i] = A11 * ddth1 + A12 * ddth2 - ccFact * dth2 * (dth1 + .5 * dth2);
    t->a2[1].p[i] = A12 * ddth1 + A22 * ddth2 - .5 * ccFact * dth1 * dth1;
    t->a3[0].p[i] = 10 * dth1;
    t->a3[1].p[i] = 10 * dth2;
    t->a4[0].p[i] = .5 * C_G * (m1 * A1 * c1 + m2 * (2.0 * A1 * c1 + A2 * c12));
    t->a4[1].p[i] = .5 * C_G * m2 * A2 * c12;
  }
}

/* cfg.dyn_source = 1: Robot::call_dynSerial (robot.cpp:349-360) replaced by the caller's point function */
static void rb_dyn_callback(orc_traj *t) {
  int nPts = t->theta[0].n, J = t->nJoints;
  for (int i = 0; i < J; ++i) {
    dv_resize(&t->a1[i], nPts);
    dv_resize(&t->a2[i], nPts);
    dv_resize(&t->a3[i], nPts);
    dv_resize(&t->a4[i], nPts);
  }
  for (int i = 0; i < nPts; ++i) {
    double th[MAXD], d1[MAXD], d2[MAXD], a1[MAXD] = {0}, a2[MAXD] = {0}, a3[MAXD] = {0}, a4[MAXD] = {0};
    for (int j = 0; j < J; ++j) {
      th[j] = t->theta[j].p[i];
      d1[j] = t->thetaD[j].p[i];
      d2[j] = t->thetaD2[j].p[i];
    }
    t->dynFn(t->dynUser, J, th, d1, d2, a1, a2, a3, a4);
    for (int j = 0; j < J; ++j) {
      t->a1[j].p[i] = a1[j];
      t->a2[j].p[i] = a2[j];
      t->a3[j].p[i] = a3[j];
      t->a4[j].p[i] = a4[j];
    }
  }
}
void orc_set_dyn_callback(orc_traj *t, orc_dyn_fn fn, void *user) {
  t->dynFn = fn;
  t->dynUser = user;
}
/* dynRR (robot.cpp:377-431) behind the plug-in signature */
void orc_demo_dyn_rr(void *user, int n_joints, const double *theta, const double *thetaD, const double *thetaD2,
                     double *a1, double *a2, double *a3, double *a4) {
  (void)user;
  (void)n_joints;
  double A1 = .4, A2 = .6, m1 = 4, m2 = 8;
  double th1 = C_DEG2RAD * theta[0], th2 = C_DEG2RAD * theta[1];
  double dth1 = C_DEG2RAD * thetaD[0], dth2 = C_DEG2RAD * thetaD[1];
  double ddth1 = C_DEG2RAD * thetaD2[0], ddth2 = C_DEG2RAD * thetaD2[1];
  double c1 = cos(th1), c2 = cos(th2), c12 = cos(th1 + th2);
  double A11 = .25 * m1 * A1 * A1 + m2 * (A1 * A1 + .25 * A2 * A2 + A1 * A2 * c2);
  double A12 = .5 * m2 * (.5 * A2 * A2 + A1 * A2 * c2);
  double A22 = .25 * m2 * A2 * A2;
  a1[0] = A11 * dth1 + A12 * dth2;
  a1[1] = A12 * dth1 + A22 * dth2;
  double ccFact = m2 * A1 * A2 * sin(th2);
  a2[0] = A11 * ddth1 + A12 * ddth2 - ccFact * dth2 * (dth1 + .5 * dth2);
  a2[1] = A12 * ddth1 + A22 * ddth2 - .5 * ccFact * dth1 * dth1;
  a3[0] = 10 * dth1;
  a3[1] = 10 * dth2;
  a4[0] = .5 * C_G * (m1 * A1 * c1 + m2 * (2.0 * A1 * c1 + A2 * c12));
  a4[1] = .5 * C_G * m2 * A2 * c12;
}
/* a decoupled n-joint model (angles in degrees): tau_j = I_j thetaddot_j + b_j thetadot_j + g_j cos(theta_j), i.e.
 * a1 = I theta', a2 = I theta'', a3 = b theta', a4 = g cos(theta) in the path parameter; inertias and gravity loads
 * falling from the base to the wrist.  TEST MODEL for robots the reference has no dynamics for. */
void orc_demo_dyn_serial(void *user, int n_joints, const double *theta, const double *thetaD, const double *thetaD2,
                         double *a1, double *a2, double *a3, double *a4) {
  (void)user;
  for (int j = 0; j < n_joints; ++j) {
    double I = 2.4 / (1.0 + j), bf = 1.5 / (1.0 + j), g = (j % 2 ? 18.0 : 3.0) / (1.0 + 0.5 * j);
    double d1 = C_DEG2RAD * thetaD[j], d2 = C_DEG2RAD * thetaD2[j];
    a1[j] = I * d1;
    a2[j] = I * d2;
    a3[j] = bf * d1;
    a4[j] = g * cos(C_DEG2RAD * theta[j]);
  }
}

/* robot.cpp:487-517 */
static void rb_dyn_cspr(orc_traj *t) {
  int nPts = t->cartD[0].n;
  for (int i = 0; i < 3; ++i) {
    dv_resize(&t->a1[i], nPts);
    dv_resize(&t->a2[i], nPts);
    dv_assign(&t->a3[i], nPts, 0.0);
    dv_assign(&t->a4[i], nPts, 0.0);
  }
  for (int i = 0; i < nPts; ++i) {
    for (int j = 0; j < 3; ++j) {
      t->a1[j].p[i] = -t->cartD[j].p[i];
      t->a2[j].p[i] = -t->cartD2[j].p[i];
    }
    t->a4[2].p[i] = C_G;
  }
}

/* robot.cpp:534-558 */
static void rb_set_A(orc_traj *t, const double *theta, const double *cart, double A[MAXD][MAXD]) {
  if (!t->havePmat) rb_find_pmat(t);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = (cart[i] - t->pmat[i][j]) / theta[j];
}

/* ------------------------------------------------------------------ 5. input interpolation */
static void ba_interp_traj_linear(orc_traj *t, int nNew);
static int ba_eval_spline_partials(orc_traj *t);

/* ba.cpp:327-368 */
static void ba_aa2q_vect(orc_traj *t) {
  int nPts = t->cart[0].n;
  t->nCart = 7;
  t->cartRows = 7;
  dv_resize(&t->cart[6], nPts);
  double aa[3] = {t->cart[3].p[0], t->cart[4].p[0], t->cart[5].p[0]};
  double q[4], qprev[4];
  u_aa2q(aa, qprev);
  for (int i = 0; i < nPts; ++i) {
    aa[0] = t->cart[3].p[i];
    aa[1] = t->cart[4].p[i];
    aa[2] = t->cart[5].p[i];
    u_aa2q(aa, q);
    double qdir = 0;
    for (int j = 0; j < 4; ++j) qdir += q[j] * qprev[j];
    if (qdir < 0.0)
      for (int j = 0; j < 4; ++j) q[j] = -q[j];
    for (int j = 0; j < 4; ++j) {
      qprev[j] = q[j];
      t->cart[3 + j].p[i] = q[j];
    }
  }
}

/* ba.cpp:382-403 */
static void ba_q2aa_vect(orc_traj *t) {
  int nPts = t->cart[0].n;
  for (int i = 0; i < nPts; ++i) {
    double q[4] = {t->cart[3].p[i], t->cart[4].p[i], t->cart[5].p[i], t->cart[6].p[i]};
    double aa[3];
    u_q2aa(q, aa);
    for (int j = 0; j < 3; ++j) t->cart[3 + j].p[i] = aa[j];
  }
  t->nCart = 6;
  t->cartRows = 6;
  t->cart[6].n = 0;
}

/* ba.cpp:790-863 */
static int ba_eval_spline_full_traj(orc_traj *t, double oldRes, double newRes) {
  int nOld = t->nPts;
  t->nPtsC = nOld;
  int nNew = (int)ceil(oldRes / newRes * (nOld - 1)) + 1;
  nNew = imax(nNew, 4);
  newRes = oldRes * (nOld - 1) / (nNew - 1);
  if (t->sC.n != nOld) {
    dv_resize(&t->sC, nOld);
    dv_iota(&t->sC, 0);
    dv_scale(&t->sC, t->sres);
  }
  dv_resize(&t->sMVC, nNew);
  dv_iota(&t->sMVC, 0);
  double sScale = t->sC.p[nOld - 1] / t->sMVC.p[nNew - 1];
  dv_scale(&t->sMVC, sScale);
  t->sresC = t->sres;
  t->vFact = 1 / t->sresC;
  t->aFact = t->vFact * t->vFact;
  t->sres = newRes;
  t->nPts = nNew;
  for (int i = 0; i < t->nJoints; ++i) spl_coeffs(&t->theta[i], &t->thetaC[i], 0, &t->thetaM[i]);
  for (int i = 0; i < t->nCart; ++i) spl_coeffs(&t->cart[i], &t->cartC[i], 0, &t->cartM[i]);
  spl_coeffs(&t->ptsOrig, &t->ptsOrigC, 0, NULL);
  segs_t sg = {0};
  if (spl_find_segs(&t->sC, &t->sMVC, &sg) == -1) {
    segs_free(&sg);
    return -1;
  }
  for (int i = 0; i < t->nJoints; ++i)
    spl_interp_spline(&t->theta[i], &t->thetaD[i], &t->thetaD2[i], &t->thetaC[i], &sg, oldRes);
  for (int i = 0; i < t->nCart; ++i)
    spl_interp_spline(&t->cart[i], &t->cartD[i], &t->cartD2[i], &t->cartC[i], &sg, oldRes);
  spl_interp_spline(&t->ptsOrig, NULL, NULL, &t->ptsOrigC, &sg, oldRes);
  t->isInterpolated = 1;
  segs_free(&sg);
  return 0;
}

typedef struct {
  double tTeachFact, thetaNormFact, cartPosNormFact, sLast, sResNew, sResi;
} interp_vars;

/* ba.cpp:651-781 */
static int ba_interp_special(orc_traj *t, const interp_vars *iv) {
  int J = t->nJoints, C = t->nCart;
  for (int i = 0; i < J; ++i) spl_coeffs(&t->theta[i], &t->thetaC[i], 0, &t->thetaM[i]);
  for (int i = 0; i < C; ++i) spl_coeffs(&t->cart[i], &t->cartC[i], 0, &t->cartM[i]);
  t->curSegC = 0;
  t->tauC = 0;
  int nPts2 = (int)ceil(iv->sLast / iv->sResNew) + 1;
  nPts2 = imax(nPts2, 4);
  dvec sC2 = {0}, th[MAXD], ca[MAXD];
  memset(th, 0, sizeof(th));
  memset(ca, 0, sizeof(ca));
  dv_resize(&sC2, nPts2);
  for (int i = 0; i < J; ++i) {
    dv_resize(&th[i], nPts2);
    th[i].p[0] = t->theta[i].p[0];
  }
  for (int i = 0; i < C; ++i) {
    dv_resize(&ca[i], nPts2);
    ca[i].p[0] = t->cart[i].p[0];
  }
  double sPrv = 0, prv_ds = 0;
  int CurNewPt = 1, CurOldPt = 1;
  int isDone = 0;
  while (!isDone) {
    double dthetaSQ = 0;
    for (int j = 0; j < J; ++j) {
      double d = t->theta[j].p[CurOldPt] - th[j].p[CurNewPt - 1];
      dthetaSQ += d * d;
    }
    double dcartSQ = 0;
    for (int j = 0; j < 3; ++j) {
      double d = t->cart[j].p[CurOldPt] - ca[j].p[CurNewPt - 1];
      dcartSQ += d * d;
    }
    double cur_ds = iv->tTeachFact * iv->sResi * t->ptsOrig.p[CurOldPt] +
                    iv->thetaNormFact * sqrt(dthetaSQ) + iv->cartPosNormFact * sqrt(dcartSQ);
    if (cur_ds > iv->sResNew) {
      sC2.p[CurNewPt] = sPrv + iv->sResNew - prv_ds;
      prv_ds = 0;
      sPrv = sC2.p[CurNewPt];
      t->sCur = sPrv;
      if (t->sCur > t->sC.p[t->nPts - 1]) isDone = 1;
      if (!isDone) {
        ba_eval_spline_partials(t);
        for (int i = 0; i < J; ++i) th[i].p[CurNewPt] = t->thetapt[i];
        for (int i = 0; i < C; ++i) ca[i].p[CurNewPt] = t->cartpt[i];
        CurOldPt = t->curSegC + 1;
        CurNewPt++;
        if (CurNewPt == th[0].n) {
          dv_resize(&sC2, CurNewPt + nPts2);
          for (int i = 0; i < J; ++i) dv_resize(&th[i], CurNewPt + nPts2);
          for (int i = 0; i < C; ++i) dv_resize(&ca[i], CurNewPt + nPts2);
        }
      }
    } else {
      if (CurOldPt == t->nPts - 1) {
        isDone = 1;
      } else {
        prv_ds = cur_ds;
        sPrv = t->sC.p[CurOldPt];
        CurOldPt++;
      }
    }
  }
  for (int i = 0; i < J; ++i) {
    th[i].p[CurNewPt] = t->theta[i].p[t->nPts - 1];
    dv_resize(&th[i], CurNewPt + 1);
  }
  for (int i = 0; i < C; ++i) {
    ca[i].p[CurNewPt] = t->cart[i].p[t->nPts - 1];
    dv_resize(&ca[i], CurNewPt + 1);
  }
  t->nPts = CurNewPt + 1;
  t->sres = iv->sResNew;
  for (int i = 0; i < J; ++i) {
    dv_copy(&t->theta[i], &th[i]);
    dv_free(&th[i]);
  }
  for (int i = 0; i < C; ++i) {
    dv_copy(&t->cart[i], &ca[i]);
    dv_free(&ca[i]);
  }
  t->thetaRows = J;
  t->cartRows = C;
  dv_free(&sC2);
  if (t->nPts < 4) ba_interp_traj_linear(t, 4);
  dv_resize(&t->ptsOrig, t->nPts);
  dv_iota(&t->ptsOrig, 0);
  return 0;
}

/* ba.cpp:412-638 */
static int ba_adjust_s(orc_traj *t, int special) {
  if (t->sWeights[1] + t->sWeights[2] < 1e-8) return 0;
  double cartNormRes = special ? t->cfg.cart_norm_res : t->cfg.cart_norm_res2;
  double thetaNormRes = special ? t->cfg.theta_norm_res : t->cfg.theta_norm_res2;
  int nPts = t->nPts;
  int J = t->nJoints;
  double *thetaNorm = (double *)calloc((size_t)nPts, sizeof(double));
  double *cartPosNorm = (double *)calloc((size_t)nPts, sizeof(double));
  dv_resize(&t->sC, nPts);
  dv_resize(&t->ptsOrig, nPts);
  double sResi = t->sres;
  double MinRatio = 1.0 / t->quadThresh;
  double thetaWindow = 5;
  double thetaNormLast = 0, cartPosNormLast = 0;
  if (!t->cfg.are_jnt_deg) thetaWindow *= C_DEG2RAD;
  for (int i = 0; i < nPts - 1; ++i) {
    double dthetaSQ = 0;
    for (int j = 0; j < J; ++j) {
      double d = t->theta[j].p[i + 1] - t->theta[j].p[i];
      dthetaSQ += d * d;
    }
    thetaNorm[i + 1] = thetaNorm[i] + sqrt(dthetaSQ);
    double dcartSQ = 0;
    for (int j = 0; j < 3; ++j) {
      double d = t->cart[j].p[i + 1] - t->cart[j].p[i];
      dcartSQ += d * d;
    }
    cartPosNorm[i + 1] = cartPosNorm[i] + sqrt(dcartSQ);
    if (t->cfg.is_auto_integ_res) {
      double thetaChange = thetaNorm[i + 1] - thetaNormLast;
      double cartChange = cartPosNorm[i + 1] - cartPosNormLast;
      if (thetaChange > thetaWindow) {
        MinRatio = dmin(MinRatio, 3.0 * cartChange / thetaChange);
        thetaNormLast = thetaNorm[i + 1];
        cartPosNormLast = cartPosNorm[i + 1];
      }
    }
  }
  if (thetaNorm[nPts - 1] < thetaNormRes) {
    free(thetaNorm);
    free(cartPosNorm);
    return -1;
  }
  double sLast = 0, sResNew = 0;
  if (t->cfg.is_auto_integ_res) { /* ba.cpp:493-556 */
    if ((cartPosNorm[nPts - 1] < cartNormRes) && t->scaleType == 2) {
      t->sWeights[1] = t->sWeights[1] + t->sWeights[2];
      t->sWeights[2] = 0;
      t->scaleType = 1;
    }
    double sW12in = t->sWeights[1] + t->sWeights[2];
    double cartRat = 500.0 * cartPosNorm[nPts - 1];
    double thetaRat = thetaNorm[nPts - 1];
    if (!t->cfg.are_jnt_deg) thetaRat *= C_RAD2DEG;
    double minIntegRes = 0.004, maxIntegRes = 0.2, K = 0.0003;
    double newIntegRes = K * t->cfg.cart_acc_max / t->cfg.cart_vel_max;
    for (int i = 0; i < J; ++i)
      newIntegRes = dmax(newIntegRes, K * t->cfg.jnt_acc_max[i] / t->cfg.jnt_vel_max[i]);
    newIntegRes = dmin(newIntegRes, maxIntegRes);
    double changeRat = cartRat / thetaRat;
    double jointIntegRes = maxIntegRes * changeRat * changeRat;
    double jointWin = maxIntegRes * MinRatio * MinRatio;
    jointWin = dmax(jointWin, 0.016);
    jointIntegRes = dmin(jointIntegRes, jointWin);
    if (jointIntegRes < newIntegRes) newIntegRes = jointIntegRes;
    newIntegRes = dmax(newIntegRes, minIntegRes);
    t->integRes = newIntegRes;
    double sW12out = cartRat + thetaRat;
    double outScale = sW12in / sW12out;
    cartRat *= outScale;
    thetaRat *= outScale;
    if (thetaRat > t->sWeights[1]) {
      t->sWeights[1] = thetaRat;
      t->sWeights[2] = cartRat;
    }
    if (t->sWeights[2] > 0) cartNormRes = dmin(cartNormRes, cartNormRes * t->sWeights[2] / t->sWeights[1]);
  }
  switch (t->scaleType) {
    case 0: sLast = sResi * t->ptsOrig.p[nPts - 1]; sResNew = sResi; break;
    case 1: sLast = thetaNorm[nPts - 1]; sResNew = thetaNormRes; break;
    case 2: sLast = cartPosNorm[nPts - 1]; sResNew = cartNormRes; break;
  }
  double cartPosNormFact, tTeachFact, thetaNormFact;
  if (cartPosNorm[nPts - 1] >= cartNormRes)
    cartPosNormFact = t->sWeights[2] * sLast / cartPosNorm[nPts - 1];
  else
    cartPosNormFact = 0;
  tTeachFact = t->sWeights[0] * sLast / (sResi * t->ptsOrig.p[nPts - 1]);
  thetaNormFact = t->sWeights[1] * sLast / thetaNorm[nPts - 1];
  t->sres = sLast / (nPts - 1);
  for (int i = 0; i < nPts; ++i)
    t->sC.p[i] = tTeachFact * sResi * t->ptsOrig.p[i] + thetaNormFact * thetaNorm[i] +
                 cartPosNormFact * cartPosNorm[i];
  free(thetaNorm);
  free(cartPosNorm);
  if (special) {
    interp_vars iv = {tTeachFact, thetaNormFact, cartPosNormFact, sLast, sResNew, sResi};
    ba_interp_special(t, &iv);
  } else {
    for (int i = 1; i < nPts; ++i)
      if (t->sC.p[i] - t->sC.p[i - 1] < 1e-12 * t->sres) return -1;
    ba_eval_spline_full_traj(t, t->sres, sResNew);
  }
  if (t->cfg.path_type == BATOTP_JOINT) {
    if (t->cfg.robot_type == BATOTP_GENJNT) {
      t->cartRows = t->nCart;
      for (int i = 0; i < t->nCart; ++i) dv_resize(&t->cart[i], t->nPts);
    } else {
      rb_fwd_kin(t);
    }
  }
  if (t->cfg.path_type == BATOTP_CART) rb_inv_kin(t);
  return 0;
}

/* ba.cpp:2768-2794 */
static void ba_interp_traj_linear(orc_traj *t, int nNew) {
  int nOld = t->nPts;
  dvec pOld = {0}, pNew = {0};
  dv_resize(&pOld, nOld);
  dv_iota(&pOld, 0);
  dv_scale(&pOld, 1.0 / (nOld - 1));
  dv_resize(&pNew, nNew);
  dv_iota(&pNew, 0);
  dv_scale(&pNew, 1.0 / (nNew - 1));
  segs_t sg = {0};
  spl_find_segs(&pOld, &pNew, &sg);
  /* rows that do not exist yet (e.g. the generic robot's Cartesian block before ba.cpp:257) are out-of-bounds
   * reads in the reference (undefined behaviour); here they are skipped and stay empty */
  for (int i = 0; i < t->nJoints; ++i)
    if (t->theta[i].n == nOld) spl_interp_linear(&t->theta[i], &sg);
  for (int i = 0; i < t->nCart; ++i)
    if (t->cart[i].n == nOld) spl_interp_linear(&t->cart[i], &sg);
  t->sres = t->sres * (nOld - 1) / (nNew - 1);
  t->nPts = nNew;
  segs_free(&sg);
  dv_free(&pOld);
  dv_free(&pNew);
}

/* ba.cpp:958-967 */
static void ba_dyn_coeffs2ser(orc_traj *t, dvec *b, int pt) {
  int n = t->dynDim;
  double bs[MAXD], xs[MAXD];
  for (int j = 0; j < n; ++j) bs[j] = b[j].p[pt];
  u_solve_lin_sys(n, t->Apt, bs, xs, t->cfg.is_svd);
  for (int j = 0; j < n; ++j) b[j].p[pt] = xs[j];
}

/* ba.cpp:873-949 */
static int ba_find_dyn_model(orc_traj *t) {
  t->dynDim = t->isParallelMech ? t->nCart : t->nJoints;
  if (t->isParallelMechOrig) {
    if (t->cfg.robot_type != BATOTP_CSPR3DOF) return -1;
    rb_dyn_cspr(t);
  } else if (t->cfg.dyn_source == 1) {
    if (!t->dynFn) return -1;
    rb_dyn_callback(t);
  } else {
    if (t->cfg.robot_type != BATOTP_RR) return -1;
    rb_dyn_rr(t);
  }
  if (t->isParallelMech && t->cfg.is_par2ser) {
    double cartpt[MAXD], thetapt[MAXD];
    for (int i = 0; i < t->nPts; ++i) {
      for (int j = 0; j < t->nCart; ++j) cartpt[j] = t->cart[j].p[i];
      for (int j = 0; j < t->nJoints; ++j) thetapt[j] = t->theta[j].p[i];
      rb_set_A(t, thetapt, cartpt, t->Apt);
      ba_dyn_coeffs2ser(t, t->a1, i);
      ba_dyn_coeffs2ser(t, t->a2, i);
      ba_dyn_coeffs2ser(t, t->a3, i);
      ba_dyn_coeffs2ser(t, t->a4, i);
    }
    t->isParallelMech = 0;
  }
  for (int i = 0; i < t->nJoints; ++i) { /* Q7: over nJoints */
    spl_coeffs(&t->a1[i], &t->a1C[i], 0, &t->a1M[i]);
    spl_coeffs(&t->a2[i], &t->a2C[i], 0, &t->a2M[i]);
    spl_coeffs(&t->a3[i], &t->a3C[i], 0, &t->a3M[i]);
    spl_coeffs(&t->a4[i], &t->a4C[i], 0, &t->a4M[i]);
  }
  t->aCRows = t->dynDim;
  return 0;
}

/* ba.cpp:95-316 */
int orc_interp_input(orc_traj *t) {
  if (t->timestamp.n > 0) { /* ba.cpp:98-127; isRem is a vector<uint8_t> of INDICES (wraps at 256) */
    int cap = t->nPts > 0 ? t->nPts : 1;
    unsigned char *isRem = (unsigned char *)malloc((size_t)cap);
    int nRem = 0;
    for (int i = 1; i < t->nPts; ++i)
      if (t->timestamp.p[i] == t->timestamp.p[i - 1]) isRem[nRem++] = (unsigned char)i;
    for (int r = nRem - 1; r >= 0; --r) {
      int k = isRem[r];
      memmove(t->timestamp.p + k, t->timestamp.p + k + 1, (size_t)(t->timestamp.n - k - 1) * sizeof(double));
      t->timestamp.n--;
      for (int j = 0; j < t->nJoints; ++j) {
        memmove(t->theta[j].p + k, t->theta[j].p + k + 1, (size_t)(t->theta[j].n - k - 1) * sizeof(double));
        t->theta[j].n--;
      }
      for (int j = 0; j < t->nCart; ++j) {
        memmove(t->cart[j].p + k, t->cart[j].p + k + 1, (size_t)(t->cart[j].n - k - 1) * sizeof(double));
        t->cart[j].n--;
      }
    }
    free(isRem);
    t->nPts = t->timestamp.n;
    t->tresInput = t->timestamp.p[t->timestamp.n - 1] / (t->nPts - 1);
    t->sres = t->tresInput;
    dv_copy(&t->sC, &t->timestamp);
  }
  if (t->nPts == 1) return -1;
  if (t->nPts < 4) ba_interp_traj_linear(t, 4);
  int pt = t->cfg.path_type;
  if (t->cfg.is_interp_only) { /* ba.cpp:139-159 */
    t->nPts = t->theta[0].n;
    double oldRes = t->sres;
    dv_resize(&t->ptsOrig, t->nPts);
    dv_iota(&t->ptsOrig, 0);
    if ((pt == BATOTP_CART || pt == BATOTP_BOTH) && t->nCart == 6) ba_aa2q_vect(t);
    ba_eval_spline_full_traj(t, oldRes, t->outRes);
    if (t->nCart == 7) ba_q2aa_vect(t);
    t->sres = t->outRes;
    return -1;
  }
  t->sLastSec = -1;
  if (pt == BATOTP_CART) {
    u_rem_close_pts(t->cart, t->cartRows, t->theta, t->thetaRows, t->cfg.cart_thresh);
    t->nPts = t->cart[0].n;
  } else {
    u_rem_close_pts(t->theta, t->thetaRows, t->cart, t->cartRows, t->cfg.jnt_thresh);
    t->nPts = t->theta[0].n;
  }
  if (t->nPts == 1) return -1;
  if (t->nPts < 4) ba_interp_traj_linear(t, 4);
  if ((pt == BATOTP_CART || pt == BATOTP_BOTH) && t->nCart == 6) ba_aa2q_vect(t);
  int df = t->cfg.input_decim_fact;
  if (df > 1) {
    if (pt == BATOTP_JOINT || pt == BATOTP_BOTH) {
      for (int i = 0; i < t->nJoints; ++i) u_smooth(&t->theta[i], df);
      for (int i = 0; i < t->nJoints; ++i) u_decimate(&t->theta[i], df);
      t->nPts = t->theta[0].n;
    }
    if (pt == BATOTP_CART || pt == BATOTP_BOTH) {
      for (int i = 0; i < t->nCart; ++i) u_smooth(&t->cart[i], df);
      for (int i = 0; i < t->nCart; ++i) u_decimate(&t->cart[i], df);
      t->nPts = t->cart[0].n;
    }
    t->tresInput *= df;
    t->sres *= df;
    t->isInterpolated = 1;
  }
  if (t->cfg.smooth_window > 1) { /* Q6: window is inputDecimFact */
    if (pt == BATOTP_JOINT || pt == BATOTP_BOTH)
      for (int i = 0; i < t->nJoints; ++i) u_smooth(&t->theta[i], df);
    if (pt == BATOTP_CART || pt == BATOTP_BOTH)
      for (int i = 0; i < t->nCart; ++i) u_smooth(&t->cart[i], df);
  }
  if (pt == BATOTP_JOINT) {
    if (t->cfg.is_cart_vel_on || t->cfg.is_cart_acc_on) {
      if (rb_fwd_kin(t) == -1) return -1;
    } else {
      t->cartRows = t->nCart;
      for (int i = 0; i < t->nCart; ++i) dv_resize(&t->cart[i], t->nPts);
    }
  }
  if (pt == BATOTP_CART) {
    if (t->cfg.is_jnt_vel_on || t->cfg.is_jnt_acc_on || t->cfg.is_trq_on) {
      rb_inv_kin(t);
    } else {
      t->thetaRows = 3;
      for (int i = 0; i < t->nJoints; ++i) dv_resize(&t->theta[i], t->nPts);
    }
  }
  dv_resize(&t->ptsOrig, t->nPts);
  dv_iota(&t->ptsOrig, 0);
  if (ba_adjust_s(t, 1) == -1) return -1;
  if (ba_adjust_s(t, 0) == -1) return -1;
  t->sC.n = 0;
  ba_eval_spline_full_traj(t, t->sres, t->sres);
  dv_assign(&t->sdot, 0, 0.0);
  /* traj.sdot.resize(nPts, DBL_MAX): sdot is empty on a fresh Traj */
  dv_assign(&t->sdot, t->nPts, DBL_MAX);
  if (t->cfg.is_trq_on) ba_find_dyn_model(t);
  return 0;
}

/* ------------------------------------------------------------------ 6. sweeps */
/* ba.cpp:1617-1652 (+ guard: a NaN sCur would spin forever in the reference) */
static void ba_update_cur_seg(orc_traj *t, const dvec *s, double sCur, int *curSeg, double *tau) {
  double sSeg;
  const int lastSeg = s->n - 2;
  long guard = 0;
  for (;;) {
    sSeg = s->p[*curSeg];
    if (sCur >= sSeg && sCur <= s->p[*curSeg + 1]) break;
    if (sCur > sSeg) {
      if (*curSeg >= lastSeg) {
        *curSeg = lastSeg;
        break;
      }
      (*curSeg)++;
    }
    if (sCur < sSeg) {
      if (*curSeg <= 0) {
        *curSeg = 0;
        break;
      }
      (*curSeg)--;
    }
    if (++guard > 4L * s->n + 16) {
      t->numericFail = 1;
      break;
    }
  }
  *tau = (sCur - sSeg) / (s->p[*curSeg + 1] - sSeg);
}

/* ba.cpp:1423-1439 */
static void ba_eval_cart_quad(orc_traj *t) {
  double vx = t->cartDpt[0], vy = t->cartDpt[1], vz = t->cartDpt[2];
  double ax = t->cartD2pt[0], ay = t->cartD2pt[1], az = t->cartD2pt[2];
  t->CartAccCoeffs[0] = vx * vx + vy * vy + vz * vz;
  t->CartAccCoeffs[1] = 2 * (vx * ax + vy * ay + vz * az);
  t->CartAccCoeffs[2] = ax * ax + ay * ay + az * az;
}

/* ba.cpp:1341-1413 */
static int ba_eval_spline_partials(orc_traj *t) {
  t->nA5++;
  ba_update_cur_seg(t, &t->sC, t->sCur, &t->curSegC, &t->tauC);
  int k = t->curSegC;
  double tau = t->tauC, tau2 = tau * tau, tau3 = tau2 * tau;
  for (int i = 0; i < t->nJoints; ++i) {
    double c0 = t->thetaC[i].c0.p[k], c1 = t->thetaC[i].c1.p[k];
    double c2 = t->thetaC[i].c2.p[k], c3 = t->thetaC[i].c3.p[k];
    t->thetapt[i] = c3 * tau3 + c2 * tau2 + c1 * tau + c0;
    t->thetaDpt[i] = (3 * c3 * tau2 + 2 * c2 * tau + c1) * t->vFact;
    t->thetaD2pt[i] = (6 * c3 * tau + 2 * c2) * t->aFact;
  }
  if (t->cfg.is_cart_vel_on || t->cfg.is_cart_acc_on) {
    for (int i = 0; i < t->nCart; ++i) {
      double c0 = t->cartC[i].c0.p[k], c1 = t->cartC[i].c1.p[k];
      double c2 = t->cartC[i].c2.p[k], c3 = t->cartC[i].c3.p[k];
      t->cartpt[i] = c3 * tau3 + c2 * tau2 + c1 * tau + c0;
      t->cartDpt[i] = (3 * c3 * tau2 + 2 * c2 * tau + c1) * t->vFact;
      t->cartD2pt[i] = (6 * c3 * tau + 2 * c2) * t->aFact;
    }
    ba_eval_cart_quad(t);
  }
  if (t->cfg.is_trq_on && t->aCRows > 0) {
    for (int i = 0; i < t->nJoints; ++i) {
      t->a1pt[i] = t->a1C[i].c3.p[k] * tau3 + t->a1C[i].c2.p[k] * tau2 + t->a1C[i].c1.p[k] * tau + t->a1C[i].c0.p[k];
      t->a2pt[i] = t->a2C[i].c3.p[k] * tau3 + t->a2C[i].c2.p[k] * tau2 + t->a2C[i].c1.p[k] * tau + t->a2C[i].c0.p[k];
      t->a3pt[i] = t->a3C[i].c3.p[k] * tau3 + t->a3C[i].c2.p[k] * tau2 + t->a3C[i].c1.p[k] * tau + t->a3C[i].c0.p[k];
      t->a4pt[i] = t->a4C[i].c3.p[k] * tau3 + t->a4C[i].c2.p[k] * tau2 + t->a4C[i].c1.p[k] * tau + t->a4C[i].c0.p[k];
    }
    if (t->isParallelMech) rb_set_A(t, t->thetapt, t->cartpt, t->Apt);
  }
  return 0;
}

/* ba.cpp:1590-1607 ("linear" is the only type any caller passes; Q5) */
static double ba_eval_sdot(orc_traj *t) {
  ba_update_cur_seg(t, &t->sMVC, t->sCur, &t->curSegMVC, &t->tauMVC);
  int seg = t->curSegMVC;
  double v = t->sdot.p[seg] + t->tauMVC * (t->sdot.p[seg + 1] - t->sdot.p[seg]);
  return dmax(v, t->sdotMin);
}

/* ba.cpp:1204-1236 */
static void ba_sdot_lim(orc_traj *t, double *sdot) {
  t->nA2++;
  double sdoti = *sdot;
  if (t->integDir == 1) {
    double m = ba_eval_sdot(t);
    if (*sdot > m) {
      t->isOn_sdot = 1;
      *sdot = m;
    } else
      t->isOn_sdot = 0;
  }
  *sdot = dmin(*sdot, t->sC.p[t->sC.n - 1] / t->integRes);
  *sdot = dmax(*sdot, t->sdotMin);
  for (int i = 0; i < t->nJoints; ++i)
    if (fabs(t->thetaDpt[i]) > t->cfg.jnt_thresh * t->vFact)
      *sdot = dmin(*sdot, fabs(t->cfg.jnt_vel_max[i] / t->thetaDpt[i]));
  if (t->cfg.is_cart_vel_on && t->CartAccCoeffs[0] > t->quadThresh * t->aFact)
    *sdot = dmin(*sdot, t->cfg.cart_vel_max / sqrt(t->CartAccCoeffs[0]));
  if (*sdot < sdoti) t->sdotLimTypeT = 1;
}

/* ba.cpp:1449-1581: returns 1 when the constraints are violated */
static int ba_verify(orc_traj *t, double sdotCur, double sddotMax) {
  t->nA4++;
  t->sddotL = -sddotMax;
  t->sddotH = sddotMax;
  double sdotSQ = sdotCur * sdotCur;
  double CartAccMaxSQ = t->cfg.cart_acc_max * t->cfg.cart_acc_max;
  int J = t->nJoints, C = t->nCart;
  if (t->cfg.is_trq_on) {
    if (t->isParallelMech) { /* ba.cpp:1463-1491 */
      double cStar1[MAXD], bStar[MAXD], xStar[MAXD], Astar[MAXD][MAXD], sol[2], lim[2];
      for (int i = 0; i < C; ++i) cStar1[i] = sdotSQ * t->a2pt[i] + sdotCur * t->a3pt[i] + t->a4pt[i];
      for (int j = 0; j < J; ++j) {
        lim[0] = t->cfg.jnt_trq_min[j];
        lim[1] = t->cfg.jnt_trq_max[j];
        for (int ii = 0; ii < 2; ++ii) {
          memcpy(Astar, t->Apt, sizeof(Astar));
          for (int k = 0; k < C; ++k) {
            bStar[k] = cStar1[k] - t->Apt[k][j] * lim[ii];
            Astar[k][j] = -t->a1pt[k];
          }
          u_solve_lin_sys(C, Astar, bStar, xStar, t->cfg.is_svd);
          sol[ii] = xStar[j];
        }
        t->sddotH = dmin(t->sddotH, dmax(sol[0], sol[1]));
        t->sddotL = dmax(t->sddotL, dmin(sol[0], sol[1]));
        if (t->sddotL > t->sddotH) return 1;
      }
    } else { /* ba.cpp:1495-1509 */
      for (int j = 0; j < J; ++j) {
        double a1 = t->a1pt[j];
        double tmp1 = t->a3pt[j] * sdotCur + t->a4pt[j];
        if (fabs(a1) < t->cfg.jnt_thresh * t->vFact) continue;
        double tmp2 = t->a2pt[j] * sdotSQ + tmp1;
        double s0 = (t->cfg.jnt_trq_max[j] - tmp2) / a1;
        double s1 = (t->cfg.jnt_trq_min[j] - tmp2) / a1;
        t->sddotH = dmin(t->sddotH, dmax(s0, s1));
        t->sddotL = dmax(t->sddotL, dmin(s0, s1));
        if (t->sddotL > t->sddotH) return 1;
      }
    }
  }
  if (t->cfg.is_jnt_acc_on) { /* ba.cpp:1514-1533 */
    for (int j = 0; j < J; ++j) {
      double vpt = t->thetaDpt[j];
      if (fabs(vpt) < t->cfg.jnt_thresh * t->vFact) {
        if (fabs(t->thetaD2pt[j]) < t->cfg.jnt_thresh * t->aFact) continue;
        if (sdotSQ > t->cfg.jnt_acc_max[j] / fabs(t->thetaD2pt[j])) return 1;
        continue;
      }
      int sv = sgn_d(vpt);
      double vTerm = t->thetaD2pt[j] * sdotSQ;
      t->sddotH = dmin(t->sddotH, (sv * t->cfg.jnt_acc_max[j] - vTerm) / vpt);
      t->sddotL = dmax(t->sddotL, (-sv * t->cfg.jnt_acc_max[j] - vTerm) / vpt);
      if (t->sddotL > t->sddotH) return 1;
    }
  }
  if (t->cfg.is_cart_acc_on) { /* ba.cpp:1535-1578 */
    double A = t->CartAccCoeffs[0];
    if (A > t->quadThresh * t->aFact) {
      double Bq = t->CartAccCoeffs[1] * sdotSQ;
      double Cq = t->CartAccCoeffs[2] * sdotSQ * sdotSQ - CartAccMaxSQ;
      double s1 = 0, s2 = 0; /* on the -2 path the reference reads them uninitialised */
      int e = u_solve_quadratic(A, Bq, Cq, &s1, &s2);
      if (e == -1) return 1;
      t->sddotH = dmin(t->sddotH, dmax(s1, s2));
      t->sddotL = dmax(t->sddotL, dmin(s1, s2));
      if (t->sddotL > t->sddotH) return 1;
    } else {
      double Cq = t->CartAccCoeffs[2];
      if (Cq < t->quadThresh * t->quadThresh * t->aFact * t->aFact) return 0;
      if (sdotSQ * sdotSQ > CartAccMaxSQ / Cq) return 1;
      return 0;
    }
  }
  return 0;
}

/* ba.cpp:1248-1332 */
static int ba_bisect_pt(orc_traj *t, double *sddot, int *nIter) {
  const double errThresh = .001;
  double lowFact = .01;
  double sdotMin = 0;
  double sdotGood = sdotMin, sdotGoodLast, sdotErr;
  int anyGood = 0;
  double sddotmax = 2 * t->sC.p[t->sC.n - 1] / (t->integRes * t->integRes);
  double sdotL = sdotGood;
  double sdotH = t->sdotCur;
  double sdotCur = sdotH;
  *nIter = 0;
  ba_eval_spline_partials(t);
  for (;;) {
    int viol = ba_verify(t, sdotCur, sddotmax);
    if (viol) {
      if (t->integDir == -1 && t->sLastSec < 0) t->sLastSec = t->sCur;
      sdotH = sdotCur;
      if (!anyGood) {
        lowFact *= 2.0;
        sdotL = dmax(.999 * sdotMin, (1.0 - lowFact) * sdotH);
      }
    } else {
      if (*nIter == 0) break;
      anyGood = 1;
      sdotGoodLast = sdotGood;
      sdotGood = sdotCur;
      sdotErr = fabs(sdotGood - sdotGoodLast) / sdotGood;
      if (sdotErr < errThresh || sdotCur < sdotMin) {
        t->sdotCur = sdotCur;
        break;
      }
      sdotL = sdotCur;
    }
    (*nIter)++;
    if (*nIter > 100) return -1;
    if (sdotCur < 0 || ((sdotH - sdotL) / sdotH < 1e-20 && !anyGood)) return -1;
    sdotCur = .5 * (sdotH + sdotL);
  }
  *sddot = (t->integDir == 1) ? t->sddotH : t->sddotL;
  return 0;
}

/* ba.cpp:979-1195 */
int orc_sweep(orc_traj *t, int integ_dir, int is_last) {
  t->integDir = integ_dir;
  t->isLastSweep = is_last;
  int nIter = 0;
  const int maxIntegSteps = (int)floor(t->cfg.max_integ_time / t->integRes) + 1;
  const int nChunk = 10000;
  int nIntegPts = nChunk;
  double h, sLast;
  double absh = t->integRes;
  double tElapsed = 0;
  dvec sInteg = {0}, tInteg = {0}, sdotInteg = {0};
  dv_resize(&sInteg, nIntegPts);
  dv_resize(&sdotInteg, nIntegPts);
  double sArr[7] = {0}, sdotArr[7] = {0}, sddotArr[7] = {0};
  bvec *fl = &t->flags[integ_dir == 1 ? 1 : 0];
  fl->n = 0;
  /* ba.cpp:998 fits a spline through traj.sdot that no caller reads (only "linear"
   * MVC lookups are ever requested) — omitted, it has no observable effect. */
  if (integ_dir == 1) {
    t->curSegC = 0;
    t->tauC = 0;
    sArr[0] = 0;
    t->curSegMVC = 0;
    t->tauMVC = 0;
    sLast = t->sC.p[t->nPtsC - 1];
  } else {
    t->curSegC = t->nPtsC - 2;
    t->tauC = 1;
    sArr[0] = t->sC.p[t->nPtsC - 1];
    t->curSegMVC = t->nPts - 2;
    t->tauMVC = 1;
    sLast = 0;
  }
  t->sCur = sArr[0];
  t->sdotCur = 0;
  ba_bisect_pt(t, &sddotArr[0], &nIter);
  h = integ_dir * absh;
  sdotArr[0] = .1 * h * sddotArr[0];
  t->sdotMin = sdotArr[0];
  ba_sdot_lim(t, &sdotArr[0]);
  t->sdotMin = sdotArr[0];
  t->sdotCur = sdotArr[0];
  sInteg.p[0] = sArr[0];
  ba_bisect_pt(t, &sddotArr[0], &nIter);
  sdotArr[0] = t->sdotCur;
  ba_sdot_lim(t, &sdotArr[0]);
  sdotInteg.p[0] = sdotArr[0];
  bv_push(fl, 0);
  double sdotT, sddotT;
  int nPts = 0;
  /* ba.cpp:1050-1051 (Q1): dsMin = 1e-6*sArr.back()/7 with sArr.back()==0 -> dsMinV == 0, and the
   * in-place operator* zeroes _dA for the lifetime of the BA object. */
  double dsMin = 1.0e-6 * sArr[6] / 7;
  double dsMinV[6];
  for (int q = 0; q < 6; ++q) {
    t->dA[q] = dsMin * t->dA[q];
    dsMinV[q] = t->dA[q];
  }
  int i;
  for (i = 1; i < nIntegPts; ++i) {
    double s0 = t->sCur;
    t->sdotLimTypeT = 0;
    sdotT = sdotArr[0];
    sddotT = sddotArr[0];
    sArr[6] = sArr[0] + h * sdotT;
    sdotArr[6] = sdotArr[0] + h * sddotT;
    t->sCur = sArr[6];
    ba_sdot_lim(t, &sdotArr[6]);
    t->sCur = s0;
    int nLim = 0, nBis = 0;
    for (int j = 0; j < 6; ++j) {
      t->sdotLimTypeT = 0;
      sdotT = 0;
      sddotT = 0;
      for (int k = 0; k < j + 1; ++k) {
        sdotT += t->B[k][j] * sdotArr[k];
        sddotT += t->B[k][j] * sddotArr[k];
      }
      sArr[j + 1] = sArr[0] + h * sdotT;
      sdotArr[j + 1] = sdotArr[0] + h * sddotT;
      sdotArr[j + 1] = dmax(sdotArr[j + 1], dsMinV[j] / absh);
      t->sCur = sArr[j + 1];
      ba_sdot_lim(t, &sdotArr[j + 1]);
      t->sdotCur = sdotArr[j + 1];
      ba_bisect_pt(t, &sddotArr[j + 1], &nIter);
      sdotArr[j + 1] = t->sdotCur;
      if (t->sdotLimTypeT) nLim++;
      if (nIter > 0) nBis++;
    }
    sArr[0] = sArr[6];
    sdotArr[0] = sdotArr[6];
    sddotArr[0] = sddotArr[6];
    sInteg.p[i] = sArr[0];
    sdotInteg.p[i] = sdotArr[0];
    bv_push(fl, (unsigned char)(nLim | (nBis << 3) | ((t->isOn_sdot ? 1 : 0) << 6)));
    t->nSteps++;
    if (i == nIntegPts - 1) {
      nIntegPts += nChunk;
      dv_resize(&sInteg, nIntegPts);
      dv_resize(&sdotInteg, nIntegPts);
    }
    if (t->sCur * integ_dir > sLast) {
      tElapsed = absh * i;
      nPts = i + 1;
      break;
    }
    if (i > maxIntegSteps || t->numericFail) {
      t->errorOptimization = 1;
      dv_free(&sInteg);
      dv_free(&sdotInteg);
      return -1;
    }
  }
  dv_resize(&sInteg, nPts);
  dv_resize(&sdotInteg, nPts);
  dv_resize(&tInteg, nPts);
  double sRat = (sLast - sInteg.p[nPts - 2]) / (sInteg.p[nPts - 1] - sInteg.p[nPts - 2]);
  sdotInteg.p[nPts - 1] = sdotInteg.p[nPts - 2] + sRat * (sdotInteg.p[nPts - 1] - sdotInteg.p[nPts - 2]);
  sInteg.p[nPts - 1] = sLast;
  if (integ_dir == 1) {
    sdotInteg.p[nPts - 1] = t->sdot.p[t->nPts - 1];
    t->nFwd = nPts;
    t->tFwd = tElapsed;
  } else {
    for (int a = 0, b = nPts - 1; a < b; ++a, --b) {
      double x = sInteg.p[a]; sInteg.p[a] = sInteg.p[b]; sInteg.p[b] = x;
      x = sdotInteg.p[a]; sdotInteg.p[a] = sdotInteg.p[b]; sdotInteg.p[b] = x;
    }
    t->nRev = nPts;
    t->tRev = tElapsed;
  }
  dv_iota(&tInteg, 0);
  dv_scale(&tInteg, absh);
  t->tTotalTraj = tElapsed;
  if (t->cfg.is_sdot_out) {
    int w = is_last ? 1 : 0;
    dv_copy(&t->histS[w], &sInteg);
    dv_copy(&t->histSdot[w], &sdotInteg);
  }
  if (nPts < 4) { /* ba.cpp:1171-1184 */
    double tResNew = tInteg.p[nPts - 1] / 3.;
    nPts = 4;
    dvec tNew = {0};
    dv_resize(&tNew, nPts);
    dv_iota(&tNew, 0);
    dv_scale(&tNew, tResNew);
    segs_t sg = {0};
    spl_find_segs(&tInteg, &tNew, &sg);
    spl_interp_linear(&sInteg, &sg);
    spl_interp_linear(&sdotInteg, &sg);
    dv_copy(&tInteg, &tNew);
    dv_free(&tNew);
    segs_free(&sg);
  }
  if (is_last) dv_copy(&t->tMVC, &tInteg);
  dv_copy(&t->sMVC, &sInteg);
  dv_copy(&t->sdot, &sdotInteg);
  t->nPts = nPts;
  dv_free(&sInteg);
  dv_free(&sdotInteg);
  dv_free(&tInteg);
  return 0;
}

/* SURVEY §8a A10 — derived product, defined here and mirrored by the device kernel */
int orc_mvc_per_sample(orc_traj *t, double sdot_start, double *out, int cap) {
  int n = t->nPtsC;
  int saveDir = t->integDir;
  double saveMin = t->sdotMin, saveSec = t->sLastSec;
  t->integDir = -1;
  for (int k = 0; k < n && k < cap; ++k) {
    t->sCur = t->sC.p[k];
    t->curSegC = imin(k, n - 2);
    t->sdotMin = 0.0;
    ba_eval_spline_partials(t); /* refresh the per-point partials used by the velocity limits */
    double sd = sdot_start;
    ba_sdot_lim(t, &sd);
    t->sdotCur = sd;
    double sdd = 0;
    int it = 0;
    t->sLastSec = 0; /* keep the bisection from recording it */
    ba_bisect_pt(t, &sdd, &it);
    out[k] = t->sdotCur;
  }
  t->integDir = saveDir;
  t->sdotMin = saveMin;
  t->sLastSec = saveSec;
  return n;
}

/* ------------------------------------------------------------------ 7. output interpolation */
/* ba.cpp:1661-1931 */
int orc_interp_output(orc_traj *t) {
  int J = t->nJoints;
  int pt = t->cfg.path_type;
  int isReinterp = 0;
  double outResT = t->outRes;
  if (t->outRes < t->integRes) {
    isReinterp = 1;
    t->outRes = t->integRes;
    t->outSmoothFact *= dmax(outResT / t->outRes, 1.);
  }
  dvec sMVCout = {0}, tMVCout = {0};
  segs_t sg = {0};
  double tLast = t->tMVC.p[t->tMVC.n - 1];
  dv_copy(&sMVCout, &t->sMVC);
  int nOutPts = (int)(t->outSmoothFact * ceil(t->tMVC.p[t->tMVC.n - 1] / t->outRes + 1.));
  nOutPts = imax(nOutPts, 4);
  dv_resize(&tMVCout, nOutPts);
  dv_iota(&tMVCout, -1);
  tMVCout.p[0] = 0;
  tMVCout.p[1] = 1.0 / 3.0;
  tMVCout.p[nOutPts - 1] = tMVCout.p[nOutPts - 2];
  tMVCout.p[nOutPts - 2] = tMVCout.p[nOutPts - 2] - 1.0 / 3.0;
  dv_scale(&tMVCout, t->tMVC.p[t->tMVC.n - 1] / tMVCout.p[nOutPts - 1]);
  splc sCo;
  memset(&sCo, 0, sizeof(sCo));
  spl_find_segs(&t->tMVC, &tMVCout, &sg);
  spl_coeffs(&t->sMVC, &sCo, 0, NULL);
  spl_interp_spline(&sMVCout, NULL, NULL, &sCo, &sg, t->sres / t->outSmoothFact);
  spl_find_segs(&t->sC, &sMVCout, &sg);
  t->nPts = nOutPts;
  t->sres = t->outRes;
  if (pt == BATOTP_JOINT || pt == BATOTP_BOTH) {
    for (int i = 0; i < J; ++i)
      spl_interp_spline(&t->theta[i], &t->thetaD[i], &t->thetaD2[i], &t->thetaC[i], &sg, t->sres);
    if (pt == BATOTP_JOINT && t->cfg.robot_type != BATOTP_GENJNT) rb_fwd_kin(t);
  }
  if (pt == BATOTP_CART || pt == BATOTP_BOTH) {
    for (int i = 0; i < t->nCart; ++i)
      spl_interp_spline(&t->cart[i], &t->cartD[i], &t->cartD2[i], &t->cartC[i], &sg, t->sres);
    if (pt == BATOTP_CART) rb_inv_kin(t);
  }
  if (t->cfg.is_trq_on) { /* ba.cpp:1744-1827 */
    segs_resize(&sg, t->nPts);
    for (int i = 0; i < t->nPts; ++i) {
      sg.seg[i] = i - 1;
      sg.tau[i] = 1;
    }
    sg.seg[0] = 0;
    sg.tau[0] = 0;
    if (t->isParallelMechOrig) {
      for (int i = 0; i < J; ++i) {
        spl_coeffs(&t->theta[i], &t->thetaC[i], 0, NULL);
        spl_interp_spline(&t->theta[i], &t->thetaD[i], &t->thetaD2[i], &t->thetaC[i], &sg, t->sres / t->outSmoothFact);
      }
      for (int i = 0; i < t->nCart; ++i) {
        spl_coeffs(&t->cart[i], &t->cartC[i], 0, NULL);
        spl_interp_spline(&t->cart[i], &t->cartD[i], &t->cartD2[i], &t->cartC[i], &sg, t->sres / t->outSmoothFact);
      }
      t->nPts = t->theta[0].n;
      rb_dyn_cspr(t);
      double bStar[MAXD], xStar[MAXD], cartpt[MAXD], thetapt[MAXD];
      t->trqRows = J;
      for (int i = 0; i < J; ++i) dv_resize(&t->trq[i], t->nPts);
      for (int i = 0; i < t->nPts; ++i) {
        for (int j = 0; j < t->nCart; ++j) bStar[j] = t->a2[j].p[i] + t->a3[j].p[i] + t->a4[j].p[i];
        for (int j = 0; j < t->nCart; ++j) cartpt[j] = t->cart[j].p[i];
        for (int j = 0; j < J; ++j) thetapt[j] = t->theta[j].p[i];
        rb_set_A(t, thetapt, cartpt, t->Apt);
        u_solve_lin_sys(t->nCart, t->Apt, bStar, xStar, t->cfg.is_svd);
        for (int j = 0; j < J; ++j) t->trq[j].p[i] = xStar[j];
      }
    } else {
      for (int i = 0; i < J; ++i) {
        spl_coeffs(&t->theta[i], &t->thetaC[i], 1, NULL);
        spl_interp_spline(&t->theta[i], &t->thetaD[i], &t->thetaD2[i], &t->thetaC[i], &sg, t->sres / t->outSmoothFact);
      }
      t->nPts = t->theta[0].n;
      if (t->cfg.dyn_source == 1)
        rb_dyn_callback(t);
      else
        rb_dyn_rr(t);
      t->trqRows = J;
      for (int i = 0; i < J; ++i) {
        dv_resize(&t->trq[i], t->nPts);
        for (int j = 0; j < t->nPts; ++j) t->trq[i].p[j] = t->a2[i].p[j] + t->a3[i].p[j] + t->a4[i].p[j];
      }
    }
  }
  if (t->cart[0].n != t->theta[0].n)
    for (int i = 0; i < 3; ++i) dv_resize(&t->cart[i], t->nPts);
  if (t->outSmoothFact > 1.5) { /* ba.cpp:1838-1871 */
    int nIn = t->nPts;
    int nOut = imax((int)((nIn - 1) / t->outSmoothFact) + 1, 4);
    dvec inS = {0}, outS = {0};
    dv_resize(&inS, nIn);
    dv_iota(&inS, 0);
    dv_resize(&outS, nOut);
    dv_iota(&outS, 0);
    dv_scale(&outS, inS.p[nIn - 1] / outS.p[nOut - 1]);
    spl_find_segs(&inS, &outS, &sg);
    int w = (int)t->outSmoothFact;
    t->nPts = nOut;
    for (int i = 0; i < J; ++i) {
      u_smooth(&t->theta[i], w);
      spl_interp_linear(&t->theta[i], &sg);
    }
    if (t->cfg.is_trq_on)
      for (int i = 0; i < J; ++i) {
        u_smooth(&t->trq[i], w);
        spl_interp_linear(&t->trq[i], &sg);
      }
    for (int i = 0; i < t->nCart; ++i) {
      u_smooth(&t->cart[i], w);
      spl_interp_linear(&t->cart[i], &sg);
    }
    dv_free(&inS);
    dv_free(&outS);
  }
  if (isReinterp) { /* ba.cpp:1873-1919 */
    int nPtsOut = imax((int)(ceil(tLast / outResT)), 4);
    dvec s1 = {0}, s2 = {0};
    dv_resize(&s1, t->nPts);
    dv_iota(&s1, 0);
    dv_resize(&s2, nPtsOut);
    dv_iota(&s2, 0);
    dv_scale(&s1, 1. / s1.p[t->nPts - 1]);
    dv_scale(&s2, 1. / s2.p[nPtsOut - 1]);
    spl_find_segs(&s1, &s2, &sg);
    for (int i = 0; i < J; ++i) {
      spl_coeffs(&t->theta[i], &t->thetaC[i], 0, NULL);
      spl_interp_spline(&t->theta[i], &t->thetaD[i], &t->thetaD2[i], &t->thetaC[i], &sg, outResT);
    }
    if (!t->isGenericRobot)
      for (int i = 0; i < t->nCart; ++i) {
        spl_coeffs(&t->cart[i], &t->cartC[i], 0, NULL);
        spl_interp_spline(&t->cart[i], &t->cartD[i], &t->cartD2[i], &t->cartC[i], &sg, outResT);
      }
    if (t->cfg.is_trq_on)
      for (int i = 0; i < J; ++i) {
        splc dC;
        memset(&dC, 0, sizeof(dC));
        spl_coeffs(&t->trq[i], &dC, 0, NULL);
        spl_interp_spline(&t->trq[i], NULL, NULL, &dC, &sg, outResT);
        splc_free(&dC);
      }
    t->nPts = nPtsOut;
    t->outRes = outResT;
    dv_free(&s1);
    dv_free(&s2);
  }
  t->sres = t->outRes;
  t->nPts = t->theta[0].n;
  if (t->nCart == 7) ba_q2aa_vect(t);
  dv_free(&sMVCout);
  dv_free(&tMVCout);
  splc_free(&sCo);
  segs_free(&sg);
  return 0;
}

/* ba.cpp:2538-2573 */
int orc_optimize(orc_traj *t) {
  t->errorOptimization = 0;
  if (orc_interp_input(t) == -1) return -1;
  if (t->nPts < 4) return -1;
  if (orc_sweep(t, -1, 0) == -1) return -1;
  if (orc_sweep(t, 1, 1) == -1) return -1;
  orc_interp_output(t);
  return 0;
}

/* ------------------------------------------------------------------ 8. accessors */
static int put_vec(const dvec *v, double *buf, int cap) {
  int n = v->n;
  if (buf)
    for (int i = 0; i < n && i < cap; ++i) buf[i] = v->p[i];
  return n;
}

int orc_get_vec(const orc_traj *t, const char *name, int idx, double *buf, int cap) {
#define ROW(nm, arr) \
  if (!strcmp(name, nm)) return (idx >= 0 && idx < MAXD) ? put_vec(&t->arr[idx], buf, cap) : -1;
#define ONE(nm, v) \
  if (!strcmp(name, nm)) return put_vec(&t->v, buf, cap);
  ROW("theta", theta) ROW("thetaD", thetaD) ROW("thetaD2", thetaD2)
  ROW("cart", cart) ROW("cartD", cartD) ROW("cartD2", cartD2)
  ROW("trq", trq) ROW("a1", a1) ROW("a2", a2) ROW("a3", a3) ROW("a4", a4)
  ROW("thetaC_m", thetaM) ROW("cartC_m", cartM)
  ROW("a1C_m", a1M) ROW("a2C_m", a2M) ROW("a3C_m", a3M) ROW("a4C_m", a4M)
  ONE("sMVC", sMVC) ONE("sdot", sdot) ONE("tMVC", tMVC) ONE("sC", sC) ONE("ptsOrig", ptsOrig)
  ONE("hist_s0", histS[0]) ONE("hist_sdot0", histSdot[0])
  ONE("hist_s1", histS[1]) ONE("hist_sdot1", histSdot[1])
  if (!strcmp(name, "thetaC_y")) return (idx >= 0 && idx < MAXD) ? put_vec(&t->thetaC[idx].c0, buf, cap) : -1;
  if (!strcmp(name, "cartC_y")) return (idx >= 0 && idx < MAXD) ? put_vec(&t->cartC[idx].c0, buf, cap) : -1;
  if (!strcmp(name, "flags0") || !strcmp(name, "flags1")) {
    const bvec *f = &t->flags[name[5] - '0'];
    if (buf)
      for (int i = 0; i < f->n && i < cap; ++i) buf[i] = (double)f->p[i];
    return f->n;
  }
#undef ROW
#undef ONE
  return -1;
}

double orc_get_scalar(const orc_traj *t, const char *name) {
#define S(nm, v) \
  if (!strcmp(name, nm)) return (double)(t->v);
  S("nPts", nPts) S("nPtsC", nPtsC) S("sres", sres) S("sresC", sresC) S("vFact", vFact) S("aFact", aFact)
  S("tTotalTraj", tTotalTraj) S("sLastSec", sLastSec) S("nRev", nRev) S("nFwd", nFwd) S("tRev", tRev)
  S("tFwd", tFwd) S("nCart", nCart) S("outRes", outRes) S("integRes", integRes) S("nA5", nA5) S("nA4", nA4)
  S("nA2", nA2) S("nSteps", nSteps) S("errorOptimization", errorOptimization) S("cartRows", cartRows)
  S("trqRows", trqRows)
#undef S
  return NAN;
}

/* ba.cpp:2582-2651 */
long orc_pack_traj_out(const orc_traj *t, unsigned char *buf, long cap) {
  int n = t->theta[0].n;
  int isCart = (t->cartRows == t->nCart && t->cart[0].n == n) ? 1 : 0;
  int isTrq = (t->cfg.is_trq_on && t->trqRows > 0 && t->trq[0].n > 0) ? 1 : 0;
  long need = 4 + 4 + 4 + (long)t->nJoints * n * 4 + 4 + (isCart ? (long)t->nCart * n * 4 : 0) + 4 +
              (isTrq ? (long)t->nJoints * n * 4 : 0);
  if (!buf || cap < need) return need;
  unsigned char *w = buf;
  float f = (float)t->sres;
  unsigned int un = (unsigned int)t->nPts;
  int one = 1;
  memcpy(w, &f, 4); w += 4;
  memcpy(w, &un, 4); w += 4;
  memcpy(w, &one, 4); w += 4;
  for (int i = 0; i < t->nJoints; ++i)
    for (int k = 0; k < n; ++k) { f = (float)t->theta[i].p[k]; memcpy(w, &f, 4); w += 4; }
  memcpy(w, &isCart, 4); w += 4;
  if (isCart)
    for (int i = 0; i < t->nCart; ++i)
      for (int k = 0; k < n; ++k) { f = (float)t->cart[i].p[k]; memcpy(w, &f, 4); w += 4; }
  memcpy(w, &isTrq, 4); w += 4;
  if (isTrq)
    for (int i = 0; i < t->nJoints; ++i)
      for (int k = 0; k < n; ++k) { f = (float)t->trq[i].p[k]; memcpy(w, &f, 4); w += 4; }
  return need;
}

/* ba.cpp:2726-2759 */
long orc_pack_s_sdot(const orc_traj *t, unsigned char *buf, long cap) {
  long need = 0;
  for (int i = 0; i < 2; ++i) need += 8 + 4 + 8L * t->histS[i].n;
  if (!buf || cap < need) return need;
  unsigned char *w = buf;
  for (int i = 0; i < 2; ++i) {
    int n = t->histS[i].n;
    memcpy(w, &t->sres, 8); w += 8;
    memcpy(w, &n, 4); w += 4;
    for (int k = 0; k < n; ++k) { float f = (float)t->histS[i].p[k]; memcpy(w, &f, 4); w += 4; }
    for (int k = 0; k < n; ++k) { float f = (float)t->histSdot[i].p[k]; memcpy(w, &f, 4); w += 4; }
  }
  return need;
}

/* ------------------------------------------------------------------ batch runner (cpu_baseline "port") */
typedef struct {
  const batotp_cfg *cfg;
  int B, n0, tid, nth;
  double tres;
  const float *theta, *cart;
  double *t_total;
  int *n_rev, *n_fwd, *n_out, *status;
  float *theta_out;
  int out_cap;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *jb = (batch_job *)arg;
  int J = jb->cfg->n_joints, C = jb->cfg->n_cart;
  for (int b = jb->tid; b < jb->B; b += jb->nth) {
    orc_traj *t = orc_new(jb->cfg);
    orc_load_raw(t, jb->n0, jb->tres, jb->theta ? jb->theta + (size_t)b * J * jb->n0 : NULL,
                 jb->cart ? jb->cart + (size_t)b * C * jb->n0 : NULL, NULL);
    int r = orc_optimize(t);
    if (jb->status) jb->status[b] = r;
    if (jb->t_total) jb->t_total[b] = t->tTotalTraj;
    if (jb->n_rev) jb->n_rev[b] = t->nRev;
    if (jb->n_fwd) jb->n_fwd[b] = t->nFwd;
    if (jb->n_out) jb->n_out[b] = (r == 0) ? t->theta[0].n : 0;
    if (jb->theta_out && r == 0)
      for (int j = 0; j < J; ++j)
        for (int k = 0; k < t->theta[j].n && k < jb->out_cap; ++k)
          jb->theta_out[((size_t)b * J + j) * jb->out_cap + k] = (float)t->theta[j].p[k];
    orc_free(t);
  }
  return NULL;
}

double orc_batch_run(const batotp_cfg *cfg, int B, int n0, double tres, const float *theta,
                     const float *cart, int n_threads, double *t_total, int *n_rev, int *n_fwd,
                     int *n_out, int *status, float *theta_out, int out_cap) {
  if (n_threads < 1) n_threads = 1;
  pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
  batch_job *jb = (batch_job *)malloc((size_t)n_threads * sizeof(batch_job));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int i = 0; i < n_threads; ++i) {
    batch_job j = {cfg, B, n0, i, n_threads, tres, theta, cart, t_total, n_rev, n_fwd, n_out, status, theta_out, out_cap};
    jb[i] = j;
    pthread_create(&th[i], NULL, batch_worker, &jb[i]);
  }
  for (int i = 0; i < n_threads; ++i) pthread_join(th[i], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th);
  free(jb);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
