"""Shared parity helpers: run a case through a batotp_cuda context (the product library on
the GPU, or the host-emulation build of the same kernels on CPU) and compare it with the
oracle restatement / the golden files.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import json
import os

import numpy as np

from _oracle import Oracle
from batotp_b200 import native, synth
from batotp_b200.config import (GENJNT, pack_s_sdot, pack_traj_out, read_config, read_traj_bin,
                                read_traj_csv)

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
STOCK = ["GEN7DOF", "RR", "UR5", "KUKA-LWR-IV", "CSPR3DOF"]


def golden_json():
    return json.load(open(os.path.join(GOLD, "golden.json")))


def synthetic_json():
    return json.load(open(os.path.join(GOLD, "synthetic.json")))


def load_stock(name):
    """-> (cfg, tres, theta[1,J,n] or None, cart[1,C,n] or None, timestamp[1,n] or None)"""
    d = os.path.join(GOLD, "stock", name)
    cfg, fname = read_config(os.path.join(d, "config.dat"))
    ts = None
    if cfg.is_bin_file:
        tres, n0, th, ca = read_traj_bin(os.path.join(d, fname), cfg.n_joints, cfg.n_cart)
    else:
        tres, n0, th, ca, ts, _ = read_traj_csv(os.path.join(d, fname), cfg.n_joints, cfg.n_cart,
                                                cfg.robot_type == GENJNT)
        ts = np.ascontiguousarray(ts[None])
    th = None if th is None else np.ascontiguousarray(th[None])
    ca = None if ca is None else np.ascontiguousarray(ca[None])
    return cfg, tres, th, ca, ts


def variant_config(name, dst_dir, **lines):
    """Writes a copy of a stock config.dat with some option lines replaced (the option name as it appears in the
    file's comment, e.g. inputDecimFact=3) into dst_dir and returns its path; the path file stays in the stock
    folder (pass that as the input folder)."""
    src = os.path.join(GOLD, "stock", name, "config.dat")
    out = []
    left = dict(lines)
    for ln in open(src).read().splitlines():
        for k in list(left):
            if "// " + k in ln:
                ln = "%s // %s" % (left.pop(k), k)
        out.append(ln)
    assert not left, "options not found: %s" % list(left)
    dst = os.path.join(str(dst_dir), "config_%s.dat" % name)
    open(dst, "w").write("\n".join(out) + "\n")
    return dst


def load_stock_variant(name, dst_dir, **lines):
    """load_stock with some options of the stock config replaced."""
    cfg, tres, th, ca, ts = load_stock(name)
    cfg2, _ = read_config(variant_config(name, dst_dir, **lines))
    return cfg2, tres, th, ca, ts


def load_synth(name, first, count):
    """-> (cfg, tres, theta or None, cart or None)"""
    cfg, _ = read_config(os.path.join(GOLD, "synthetic", name + "_config.dat"))
    if name == "GEN7DOF":
        tres, th = synth.gen7dof_paths(first, count)
        return cfg, tres, th, None
    if name == "KUKA":
        tres, th = synth.kuka_paths(first, count)
        return cfg, tres, th, None
    tres, ca = synth.cspr_paths(first, count)
    return cfg, tres, None, ca


def run_device(ctx, cfg, tres, th, ca, ts=None, n0=None, out_cap=32768, hist_cap=32768):
    ref = th if th is not None else ca
    B = ref.shape[0]
    bi = ctx.make_in(th, ca, tres, n0=n0, timestamp=ts)
    res = native.BatchResult(B, cfg.n_joints, cfg.n_cart, out_cap, hist_cap, bool(cfg.is_trq_on))
    ctx.optimize_batch(cfg, bi, res)
    return res


class OracleRun:
    def __init__(self, cfg, tres, th, ca, ts=None, n0=None, dyn_fn=None):
        self.o = Oracle(cfg)
        if dyn_fn is not None:
            self.o.set_dyn_callback(dyn_fn)
        ref = th if th is not None else ca
        n = ref.shape[1] if n0 is None else n0
        self.o.load_raw(n, tres, None if th is None else np.ascontiguousarray(th[:, :n]),
                        None if ca is None else np.ascontiguousarray(ca[:, :n]),
                        None if ts is None else ts[:n])
        self.rc = self.o.optimize()
        o = self.o
        self.J = cfg.n_joints
        self.n_rev, self.n_fwd = int(o.scalar("nRev")), int(o.scalar("nFwd"))
        self.t_total = o.scalar("tTotalTraj")
        self.s_last_sec = o.scalar("sLastSec")
        self.ok = self.rc == 0
        if self.ok:
            self.n_out = int(o.scalar("nPts"))
            self.theta = o.rows("theta", self.J).astype(np.float32)
            cr = int(o.scalar("cartRows"))
            self.cart = o.rows("cart", cr).astype(np.float32) if cr > 0 else None
            self.trq = o.rows("trq", self.J).astype(np.float32) if cfg.is_trq_on else None
            self.flags = [o.vec("flags0").astype(np.uint8), o.vec("flags1").astype(np.uint8)]
            self.hist = [o.vec("hist_s0"), o.vec("hist_sdot0"), o.vec("hist_s1"), o.vec("hist_sdot1")]


def device_traj_out_bytes(cfg, res, b):
    """trajWriteBIN (ba.cpp:2582-2651) from a batch result."""
    n = int(res.n_out[b])
    nc = int(res.n_cart_out[b])
    theta = res.theta_out[b, :, :n]
    cart = res.cart_out[b, :cfg.n_cart, :n] if (res.cart_out is not None and cfg.n_cart > 0 and nc == n) else None
    trq = res.trq_out[b, :, :n] if (cfg.is_trq_on and res.trq_out is not None) else None
    return pack_traj_out(res.out_sres[b], n, theta, cart, trq)


def device_s_sdot_bytes(res, b):
    nr, nf = int(res.n_rev[b]), int(res.n_fwd[b])
    return pack_s_sdot(res.out_sres[b], [(res.hist[b, 0, :nr], res.hist[b, 1, :nr]),
                                         (res.hist[b, 2, :nf], res.hist[b, 3, :nf])])


def compare(cfg, res, b, orc: OracleRun, check_hist=True):
    """Bit-exact comparison of one trajectory of a device batch result with the oracle."""
    msgs = []
    fatal = bool(res.status[b] & native.ST_FATAL_MASK)
    if fatal != (not orc.ok):
        msgs.append("status %d vs oracle rc %d" % (res.status[b], orc.rc))
        return msgs
    if fatal:
        return msgs
    if res.n_rev[b] != orc.n_rev or res.n_fwd[b] != orc.n_fwd:
        msgs.append("steps rev %d/%d fwd %d/%d" % (res.n_rev[b], orc.n_rev, res.n_fwd[b], orc.n_fwd))
    if res.t_total[b] != orc.t_total:
        msgs.append("tTotalTraj %r vs %r" % (res.t_total[b], orc.t_total))
    if res.s_last_sec[b] != orc.s_last_sec:
        msgs.append("sLastSec %r vs %r" % (res.s_last_sec[b], orc.s_last_sec))
    if res.n_out[b] != orc.n_out:
        msgs.append("n_out %d vs %d" % (res.n_out[b], orc.n_out))
        return msgs
    n = orc.n_out
    if res.theta_out is not None and not np.array_equal(res.theta_out[b, :, :n], orc.theta):
        msgs.append("theta_out differs (max %g)" % np.abs(res.theta_out[b, :, :n] - orc.theta).max())
    if orc.cart is not None and res.cart_out is not None:
        nc = orc.cart.shape[1]
        if res.n_cart_out[b] != nc:
            msgs.append("n_cart_out %d vs %d" % (res.n_cart_out[b], nc))
        elif not np.array_equal(res.cart_out[b, :orc.cart.shape[0], :nc], orc.cart):
            msgs.append("cart_out differs")
    if orc.trq is not None and res.trq_out is not None:
        if not np.array_equal(res.trq_out[b, :, :n], orc.trq):
            msgs.append("trq_out differs")
    if check_hist and res.hist is not None:
        for k, (nn, nm) in enumerate(((orc.n_rev, "s_rev"), (orc.n_rev, "sdot_rev"), (orc.n_fwd, "s_fwd"),
                                      (orc.n_fwd, "sdot_fwd"))):
            if len(orc.hist[k]) and not np.array_equal(res.hist[b, k, :nn], orc.hist[k].astype(np.float32)):
                msgs.append(nm + " differs")
        for k in range(2):
            f = orc.flags[k]
            if not np.array_equal(res.flags[b, k, :len(f)], f):
                msgs.append("switching flags differ (sweep %d)" % k)
    return msgs
