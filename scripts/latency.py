#!/usr/bin/env python
"""Single-trajectory latency (BASELINE configs[1]: UR5 at fine discretisation; also the five stock folders).

A trajectory is sequential in its Runge-Kutta steps, so one path keeps one lane of one warp busy: this number
is LATENCY, stated beside the reference's CPU time for the same path (oracle/_ref when present, else the oracle
port), and the result is checked bit for bit against the oracle.  "UR5-fine" = input/UR5 with integRes 0.008 ->
0.001, thetaNormRes(2) 0.3 -> 0.03, cartNormRes(2) 0.002 -> 0.0002, outRes 0.008 -> 0.001."""
import functools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
print = functools.partial(print, flush=True)


def main():
    import _parity as P
    from _oracle import Oracle
    from batotp_b200 import native
    ctx = native.Context(0)
    out = {}
    cases = [(n, n, None) for n in P.STOCK] + [("UR5-fine", "UR5", dict(integ_res=0.001, theta_norm_res=0.03,
                                                                         theta_norm_res2=0.03, cart_norm_res=0.0002,
                                                                         cart_norm_res2=0.0002, out_res=0.001))]
    for label, name, mod in cases:
        cfg, tres, th, ca, ts = P.load_stock(name)
        if mod:
            cfg = cfg.copy()
            for k, v in mod.items():
                setattr(cfg, k, v)
        cap = 200000 if mod else 32768
        res = P.run_device(ctx, cfg, tres, th, ca, ts, out_cap=cap, hist_cap=cap)  # warm-up (capacities, tables)
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            res = P.run_device(ctx, cfg, tres, th, ca, ts, out_cap=cap, hist_cap=cap)
            t.append(time.perf_counter() - t0)
        o = Oracle(cfg)
        n0 = (th if th is not None else ca).shape[2]
        o.load_raw(n0, tres, None if th is None else th[0], None if ca is None else ca[0], None if ts is None else ts[0])
        c0 = time.perf_counter()
        rc = o.optimize()
        cpu = time.perf_counter() - c0
        same = (rc == 0 and int(o.scalar("nRev")) == int(res.n_rev[0]) and int(o.scalar("nFwd")) == int(res.n_fwd[0])
                and o.scalar("tTotalTraj") == res.t_total[0]
                and np.array_equal(o.rows("theta", cfg.n_joints).astype(np.float32),
                                   res.theta_out[0, :, :int(res.n_out[0])]))
        out[label] = dict(gpu_ms=round(min(t) * 1e3, 2), cpu_oracle_port_ms=round(cpu * 1e3, 2), steps_rev=int(res.n_rev[0]),
                          steps_fwd=int(res.n_fwd[0]), grid=int(res.n_grid[0]), n_out=int(res.n_out[0]),
                          bit_exact_vs_oracle=bool(same))
        print(label, out[label])
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
