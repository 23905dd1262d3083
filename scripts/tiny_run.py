#!/usr/bin/env python
"""Small batches through the whole path (for compute-sanitizer runs): GEN7DOF with both sweep kernels (pitched and
ragged rows, several chunks and output sub-chunks, the tail helper), the CSPR3DOF with and without Par2Ser, the RR
robot, and - unless TINY_KUKA=0 - one KUKA path (strict trig on the device)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _parity as P  # noqa: E402
from batotp_b200 import native  # noqa: E402

ctx = native.Context(0)
n = int(os.environ.get("TINY_N", "48"))
cfg, tres, th, ca = P.load_synth("GEN7DOF", 0, n)
for kernel in (1, 2):
    ctx.set_sweep_kernel(kernel)
    ctx.set_walker_kernel(kernel)  # 1: one thread per trajectory, 2: point-parallel increments + group march
    ctx.set_chunk(max(8, n // 3))
    ctx.set_out_chunk(7)
    res = P.run_device(ctx, cfg, tres, th, ca, out_cap=8192, hist_cap=8192)
    rag = native.BatchResult(n, cfg.n_joints, 0, 0, 0, False, want_hist=False, ragged_cap=int(res.n_out.sum()) + 8)
    ctx.optimize_batch(cfg, ctx.make_in(th, None, tres), rag)
    print("GEN7DOF kernel", kernel, "ok", int((res.status & native.ST_FATAL_MASK == 0).sum()), "of", n,
          "t_total sum", float(res.t_total.sum()), "ragged equal", bool((rag.rows(1) == res.theta_out[1, :, :res.n_out[1]]).all()))
ctx.set_sweep_kernel(0)
ctx.set_walker_kernel(0)
ctx.set_chunk(0)
ctx.set_out_chunk(40)
cfg, tres, th, ca = P.load_synth("CSPR3DOF", 0, 3)
for p2s in (1, 0):
    c = cfg.copy()
    c.is_par2ser = p2s
    res = P.run_device(ctx, c, tres, th, ca, out_cap=8192, hist_cap=8192)
    print("CSPR3DOF isPar2Ser", p2s, "ok", int((res.status & native.ST_FATAL_MASK == 0).sum()), "t_total", res.t_total.tolist())
cfg, tres, th, ca, ts = P.load_stock("RR")
res = P.run_device(ctx, cfg, tres, th, ca, ts)
print("RR ok", int(res.status[0]), float(res.t_total[0]))
if os.environ.get("TINY_KUKA", "1") != "0":
    cfg, tres, th, ca, ts = P.load_stock("KUKA-LWR-IV")
    res = P.run_device(ctx, cfg, tres, th, ca, ts)
    print("KUKA ok", int(res.status[0]), float(res.t_total[0]))
ctx.close()
