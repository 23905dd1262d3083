#!/usr/bin/env python
"""Tuning aid (not a benchmark): per-kernel device time of one step of the GEN7DOF batch path, measured
with the library's own event brackets (batotp_cuda_set_profile serialises the launches), plus an
unprofiled resident step and an unprofiled host-buffer step for comparison.

    python scripts/kprof.py --batch 32768 [--chunk N] [--out-chunk M] [--lib path]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--out-chunk", type=int, default=8192)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--workload", default="GEN7DOF", choices=["GEN7DOF", "KUKA", "CSPR3DOF"])
    ap.add_argument("--trig", type=int, default=1)
    ap.add_argument("--max-steps", type=int, default=0)
    ap.add_argument("--skip-host", action="store_true")
    ap.add_argument("--sweep-kernel", type=int, default=0, help="0 automatic, 1 lane per trajectory, 2 group per trajectory")
    ap.add_argument("--walker-kernel", type=int, default=0, help="0 automatic, 1 thread per trajectory, 2 point-parallel + group march")
    a = ap.parse_args()
    import torch
    from batotp_b200 import native, synth
    from batotp_b200.config import read_config
    cfg, _ = read_config(os.path.join(ROOT, "tests", "golden", "synthetic", a.workload.replace("KUKA", "KUKA") + "_config.dat"))
    cfg.trig_mode = a.trig
    B = a.batch
    gen = {"GEN7DOF": synth.gen7dof_paths, "KUKA": synth.kuka_paths, "CSPR3DOF": synth.cspr_paths}[a.workload]
    parts = []
    for at in range(0, B, 4096):
        tres, p = gen(at, min(4096, B - at))
        parts.append(p)
    theta = np.concatenate(parts, axis=0)
    is_cart = a.workload == "CSPR3DOF"
    h_theta = torch.from_numpy(theta).pin_memory()
    d_theta = h_theta.cuda()
    ctx = native.Context(0, a.lib)
    if a.chunk:
        ctx.set_chunk(a.chunk)
    if a.max_steps:
        ctx.set_max_steps(a.max_steps)
    ctx.set_out_chunk(a.out_chunk)
    ctx.set_sweep_kernel(a.sweep_kernel)
    ctx.set_walker_kernel(a.walker_kernel)
    J = cfg.n_joints
    bi_dev = ctx.make_in(tres=tres, device_ptrs=dict(theta=None if is_cart else d_theta.data_ptr(),
                                                     cart=d_theta.data_ptr() if is_cart else None, B=B, n0_max=theta.shape[2]))
    res = native.BatchResult(B, J, cfg.n_cart, 0, 0, bool(cfg.is_trq_on), want_rows=False, want_hist=False)
    for _ in range(2):
        ctx.optimize_batch(cfg, bi_dev, res)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        ctx.timer_start()
        ctx.optimize_batch(cfg, bi_dev, res)
        ms = ctx.timer_stop_ms()
        ts.append((ms, (time.perf_counter() - t0) * 1e3))
    print("resident step: device ms %s  wall ms %s  -> %.0f traj/s" % (
        [round(x[0], 1) for x in ts], [round(x[1], 1) for x in ts], B / (min(x[0] for x in ts) * 1e-3)))
    ref_t = res.t_total.copy()
    print("sum t_total %.6f  ok %d  mean nfwd %.1f  max nfwd %d  mean ngrid %.0f max ngrid %d" % (
        ref_t.sum(), int((res.status & native.ST_FATAL_MASK == 0).sum()), res.n_fwd.mean(), res.n_fwd.max(),
        res.n_grid.mean(), res.n_grid.max()))
    print("status histogram:", {int(k): int(v) for k, v in zip(*np.unique(res.status, return_counts=True))})
    # host-buffer leg
    if not a.skip_host:
        out_cap = int(res.n_out.max()) + 64
        res_e = native.BatchResult(B, J, cfg.n_cart, out_cap, 0, bool(cfg.is_trq_on), want_rows=True, want_hist=False, pinned=True)
        bi = ctx.make_in(theta=None if is_cart else h_theta.numpy(), cart=h_theta.numpy() if is_cart else None, tres=tres)
        ctx.optimize_batch(cfg, bi, res_e)
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            ctx.optimize_batch(cfg, bi, res_e)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        print("host-buffer step: wall ms %s -> %.0f traj/s; identical t_total: %s; rows nonzero: %s" % (
            [round(x, 1) for x in ts], B / (min(ts) * 1e-3), bool(np.array_equal(ref_t, res_e.t_total)),
            bool(np.abs(res_e.theta_out).sum() > 0)))
    # profiled resident step
    ctx.set_profile(True)
    ctx.optimize_batch(cfg, bi_dev, res)
    pr = ctx.profile()
    ctx.set_profile(False)
    tot = sum(v[0] for v in pr.values())
    print("%-28s %10s %6s %7s" % ("kernel", "ms", "n", "share"))
    for k, v in sorted(pr.items(), key=lambda kv: -kv[1][0]):
        print("%-28s %10.3f %6d %7.3f" % (k, v[0], v[1], v[0] / tot))
    print("%-28s %10.3f" % ("TOTAL", tot))
    st = ctx.stats()
    print("stats", st)
    ctx.close()


if __name__ == "__main__":
    main()
