#!/usr/bin/env python
"""Tuning aid: throughput of K contexts (own stream + workspace each) driven by K host threads on one GPU,
every context working through its own share of the batch.  Shows how much the latency-bound kernels of
different chunks overlap."""
import argparse
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=113664)
    ap.add_argument("--ctxs", type=int, default=2)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from batotp_b200 import native, synth
    from batotp_b200.config import read_config
    cfg, _ = read_config(os.path.join(ROOT, "tests", "golden", "synthetic", "GEN7DOF_config.dat"))
    K = a.ctxs
    per = a.batch // K
    parts = []
    for at in range(0, per * K, 16384):
        tres, p = synth.gen7dof_paths(at, min(16384, per * K - at))
        parts.append(p)
    theta = torch.from_numpy(np.concatenate(parts, axis=0)).cuda()
    J = cfg.n_joints
    ctxs, ins, outs = [], [], []
    for k in range(K):
        c = native.Context(0)
        c.set_chunk(a.chunk or per)
        ctxs.append(c)
        ins.append(c.make_in(tres=tres, device_ptrs=dict(theta=theta[k * per:(k + 1) * per].data_ptr(), cart=None, B=per,
                                                         n0_max=theta.shape[2])))
        outs.append(native.BatchResult(per, J, 0, 0, 0, False, want_rows=False, want_hist=False))

    def work(k):
        ctxs[k].optimize_batch(cfg, ins[k], outs[k])

    def step():
        th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()

    step()
    step()
    ts = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    print("ctxs %d  batch %d (%d each, chunk %d): wall ms %s -> %.0f traj/s" % (
        K, per * K, per, a.chunk or per, [round(x * 1e3, 1) for x in ts], per * K / min(ts)))
    print("sum t_total", sum(float(o.t_total.sum()) for o in outs))


if __name__ == "__main__":
    main()
