#!/usr/bin/env python
"""A small batch through the whole path (for compute-sanitizer runs): 96 GEN7DOF paths, rows + histories."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _parity as P  # noqa: E402
from batotp_b200 import native  # noqa: E402

ctx = native.Context(0)
for name, n in (("GEN7DOF", int(os.environ.get("TINY_N", "96"))), ("CSPR3DOF", 4)):
    cfg, tres, th, ca = P.load_synth(name, 0, n)
    ctx.set_out_chunk(40)
    res = P.run_device(ctx, cfg, tres, th, ca, out_cap=8192, hist_cap=8192)
    print(name, "ok", int((res.status & native.ST_FATAL_MASK == 0).sum()), "of", n, "t_total sum", float(res.t_total.sum()))
ctx.close()
