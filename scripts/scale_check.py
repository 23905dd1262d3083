#!/usr/bin/env python
"""Scale check of BASELINE configs[2] (KUKA x4096) and configs[3] (CSPR3DOF x65536, here --cspr paths) on one
GPU: runs the synthetic batches through the C-ABI, reports status counts and throughput, and compares a random
sample against the oracle restatement (bit-exact switching counts and total time)."""
import argparse
import functools
import os
import sys
import time

print = functools.partial(print, flush=True)

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kuka", type=int, default=4096)
    ap.add_argument("--cspr", type=int, default=16384)
    ap.add_argument("--check", type=int, default=12)
    a = ap.parse_args()
    import _parity as P
    from batotp_b200 import native
    ctx = native.Context(0)
    for name, count, trig in (("KUKA", a.kuka, 1), ("KUKA", a.kuka, 0), ("CSPR3DOF", a.cspr, 1)):
        if count <= 0:
            continue
        cfg, tres, th, ca = P.load_synth(name, 0, count)
        cfg = cfg.copy()
        cfg.trig_mode = trig
        ref = th if th is not None else ca
        J = cfg.n_joints
        res = native.BatchResult(count, J, cfg.n_cart, 0, 0, bool(cfg.is_trq_on), want_rows=False, want_hist=False)
        bi = ctx.make_in(theta=th, cart=ca, tres=tres)
        ctx.optimize_batch(cfg, bi, res)
        t0 = time.perf_counter()
        ctx.optimize_batch(cfg, bi, res)
        dt = time.perf_counter() - t0
        ok = int((res.status & native.ST_FATAL_MASK == 0).sum())
        print("   status histogram:", {int(k): int(v) for k, v in zip(*np.unique(res.status, return_counts=True))})
        print("%s x%d trig_mode=%d: %.0f paths/s (%.1f ms), optimised %d, bisect-fail flag %d, mean steps rev %.0f fwd %.0f"
              % (name, count, trig, count / dt, dt * 1e3, ok, int(((res.status & native.ST_BISECT_FAIL) != 0).sum()),
                 res.n_rev[res.n_rev > 0].mean(), res.n_fwd[res.n_fwd > 0].mean()))
        if trig == 1 and a.check > 0:
            rng = np.random.RandomState(7)
            bad = 0
            for b in rng.choice(count, min(a.check, count), replace=False):
                orc = P.OracleRun(cfg, tres, None if th is None else th[b], None if ca is None else ca[b])
                if (res.status[b] & native.ST_FATAL_MASK) == 0:
                    if (orc.n_rev, orc.n_fwd) != (int(res.n_rev[b]), int(res.n_fwd[b])) or orc.t_total != res.t_total[b]:
                        bad += 1
            print("   oracle sample of %d: %d mismatches" % (min(a.check, count), bad))
    ctx.close()


if __name__ == "__main__":
    main()
