// k_mvc.cuh — derived product: the per-sample maximum-velocity curve.
// One thread per (trajectory, grid knot): the sweep's own per-point functions
// (evalSplinePartials ba.cpp:1341, the velocity caps of sdotLim ba.cpp:1216-1229 without the
// MVC term, applyAccelConstraintsBisectionPt ba.cpp:1248) evaluated independently at every
// knot s_k, starting from `sdotStart`.  The reference has no such stage (its MVC is the
// reverse sweep); the oracle's orc_mvc_per_sample defines the expected values.
#pragma once
#include "k_sweep.cuh"

template <int J, bool CART, bool TRQ>
__global__ void k_mvc(Ws w, double sdotStart, double *out, int cap, int nblk) {
  TP_DECOMP(nblk);
  if (b >= w.B) return;
  const TrajState &s = w.st[b];
  if (s.status & ST_FATAL_MASK) return;
  if (i >= s.nPtsC || i >= cap) return;
  SweepLane<J, CART, TRQ> L;
  L.b = b;
  L.status = 0;
  L.dir = -1;
  L.absh = s.integRes;
  L.h = -L.absh;
  L.sresC = s.sresC;
  L.vFact = s.vFact;
  L.aFact = s.aFact;
  L.nPtsC = s.nPtsC;
  L.lastSeg = s.nPtsC - 2;
  L.sBack = s.sresC * (double)(s.nPtsC - 1);
  L.sdotCap = L.sBack / L.absh;
  L.sddotmax = 2 * L.sBack / (L.absh * L.absh);
  L.thrV = CFG.c.jnt_thresh * L.vFact;
  L.thrA = CFG.c.jnt_thresh * L.aFact;
  L.thrQ = CFG.quadThresh * L.aFact;
  L.thrQ2 = CFG.quadThresh * CFG.quadThresh * L.aFact * L.aFact;
  L.amaxSQ = CFG.c.cart_acc_max * CFG.c.cart_acc_max;
  L.tab = w.tab + (size_t)b * w.Nc * (size_t)w.RT * 4;
  L.nM = 0;
  L.sM = L.sdM = nullptr;
  L.segM = 0;
  L.seg = imin_(i, s.nPtsC - 2);
  L.segLoaded = -1;
  L.sCur = s.sresC * (double)i;
  L.sdotMin = 0.0;
  L.limT = 0;
  L.isOn = 0;
  L.sLastSec = 0;
  L.eval_partials(w.Nc);
  const double sd = L.sdot_lim(sdotStart);
  L.bisect_begin(sd);
  for (;;) {  // ba.cpp:1270-1321
    const bool viol = L.verify(L.sdotCur);
    if (viol) {
      L.sdotH = L.sdotCur;
      if (!L.anyGood) {
        L.lowFact *= 2.0;
        L.sdotL = dmax_(.999 * 0.0, (1.0 - L.lowFact) * L.sdotH);
      }
    } else {
      if (L.nIter == 0) break;
      L.anyGood = 1;
      const double last = L.sdotGood;
      L.sdotGood = L.sdotCur;
      const double err = fabs(L.sdotGood - last) / L.sdotGood;
      if (err < .001 || L.sdotCur < 0.0) {
        L.sdotIn = L.sdotCur;
        break;
      }
      L.sdotL = L.sdotCur;
    }
    L.nIter++;
    if (L.nIter > 100) break;
    if (L.sdotCur < 0 || ((L.sdotH - L.sdotL) / L.sdotH < 1e-20 && !L.anyGood)) break;
    L.sdotCur = .5 * (L.sdotH + L.sdotL);
  }
  out[(size_t)b * cap + i] = L.sdotIn;
}
