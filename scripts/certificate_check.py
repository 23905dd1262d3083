#!/usr/bin/env python
"""Evidence run for the sweep kernel's float certificates (CPU, host build of the kernel sources).

The host build cross-checks every shortcut of k_sweep.cuh against the full exact computation while it runs
(TEST-ONLY blocks: float decisions vs verify_acc_exact, certified binding quotient vs the full intersection,
velocity-cap shortcut vs the all-joint cap).  This script runs GEN7DOF paths through eight limit / step regimes and
prints how many certificates were issued and how many disagreed (must be 0).

    python scripts/certificate_check.py [--paths 40]
"""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--paths", type=int, default=40)
    a = ap.parse_args()
    import __graft_entry__ as g
    import _parity as P
    from batotp_b200 import native
    ctx = native.Context(0, g.build_emu())
    out = (C.c_longlong * 16)()
    tot = [0] * 16
    t0 = time.time()
    for acc, vel, integ in [(1, 1, 1), (0.05, 1, 1), (20, 0.2, 1), (1, 5, 1), (1, 1, 0.25), (300, 30, 0.5),
                            (0.3, 0.5, 1), (3, 2, 2)]:
        cfg, tres, th, _ = P.load_synth("GEN7DOF", 300000, a.paths)
        for i in range(cfg.n_joints):
            cfg.jnt_acc_max[i] *= acc
            cfg.jnt_vel_max[i] *= vel
        cfg.integ_res *= integ
        ctx.L.batotp_emu_filter_stats(out, 16, 1)
        r = P.run_device(ctx, cfg, tres, th, None, out_cap=65536, hist_cap=65536)
        ctx.L.batotp_emu_filter_stats(out, 16, 1)
        v = list(out)
        tot = [x + y for x, y in zip(tot, v)]
        print("acc x%-5g vel x%-4g integRes x%-4g optimised %d/%d  float decisions %9d  exact fallbacks %6d  "
              "disagreements (decision, bound, velocity cap) %s  [%.0f s]"
              % (acc, vel, integ, int((r.status & native.ST_FATAL_MASK == 0).sum()), a.paths, v[0], v[1], v[8:11],
                 time.time() - t0), flush=True)
    print("TOTAL float decisions %d, certified bounds %d, velocity caps %d: disagreements %s"
          % (tot[0], tot[2], tot[4] + tot[5] + tot[6], tot[8:11]))
    return 0 if tot[8:11] == [0, 0, 0] else 1


if __name__ == "__main__":
    sys.exit(main())
