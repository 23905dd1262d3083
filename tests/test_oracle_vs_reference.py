"""CPU: the oracle restatement against the UNMODIFIED reference sources compiled into
oracle/_ref (skipped when that library is absent, i.e. where /root/reference never existed)."""
import numpy as np
import pytest

import _parity as P
from _oracle import Oracle, Ref, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref/libbatotp_ref.so not built")


def _pair(name, tmp_path, auto=False):
    d = P.GOLD + "/stock/" + name + "/"
    r = Ref(d + "config.dat", d, str(tmp_path) + "/", auto_integ_res=auto)
    assert r.load_file() == 0
    cfg, tres, th, ca, ts = P.load_stock(name)
    if auto:
        cfg = cfg.copy()
        cfg.is_auto_integ_res = 1
    rc = r.cfg()
    import ctypes as C
    assert bytes(C.string_at(C.byref(cfg), C.sizeof(cfg))) == bytes(C.string_at(C.byref(rc), C.sizeof(rc))), \
        "Python config parser disagrees with BA::readConfigData"
    o = Oracle(cfg)
    n0 = (th if th is not None else ca).shape[2]
    o.load_raw(n0, tres, None if th is None else th[0], None if ca is None else ca[0], None if ts is None else ts[0])
    return cfg, r, o


@pytest.mark.parametrize("name", P.STOCK)
def test_every_stage_matches_bit_for_bit(name, tmp_path):
    cfg, r, o = _pair(name, tmp_path)
    J = cfg.n_joints
    assert r.interp_input() == 0 and o.interp_input() == 0
    assert r.scalar("nPts") == o.scalar("nPts")
    for nm in ("theta", "thetaD", "thetaD2", "thetaC_y", "thetaC_m"):
        for j in range(J):
            x, y = r.vec(nm, j), o.vec(nm, j)
            if nm.startswith("thetaC"):
                x, y = x[:-1], y[:-1]  # the last coefficient slot is never written (spline.cpp:203)
            assert np.array_equal(x, y), (nm, j)
    if cfg.is_trq_on:
        for nm in ("a1", "a2", "a3", "a4", "a1C_m", "a4C_m"):
            for j in range(J):
                x, y = r.vec(nm, j), o.vec(nm, j)
                if nm.endswith("_m"):
                    x, y = x[:-1], y[:-1]
                assert np.array_equal(x, y), (nm, j)
    for d, last in ((-1, 0), (1, 1)):
        assert r.sweep(d, last) == 0 and o.sweep(d, last) == 0
        assert np.array_equal(r.vec("sMVC"), o.vec("sMVC")) and np.array_equal(r.vec("sdot"), o.vec("sdot"))
    assert r.scalar("tTotalTraj") == o.scalar("tTotalTraj") and r.scalar("sLastSec") == o.scalar("sLastSec")
    r.interp_output()
    o.interp_output()
    for nm, n in (("theta", J), ("cart", int(r.scalar("cartRows"))), ("trq", int(r.scalar("trqRows")))):
        assert n == int(o.scalar({"theta": "nCart", "cart": "cartRows", "trq": "trqRows"}[nm])) or nm == "theta"
        for j in range(n):
            assert np.array_equal(r.vec(nm, j), o.vec(nm, j)), (nm, j)


@pytest.mark.parametrize("name", P.STOCK)
def test_switching_flags_match_instrumented_reference(name, tmp_path):
    """The per-step switching flags: the reference's own per-point functions driven by the harness
    loop (ref_sweep_flags) must (a) reproduce BA::sweep's history bit for bit and (b) give the
    flags the oracle records."""
    cfg, r, o = _pair(name, tmp_path)
    assert r.interp_input() == 0 and o.interp_input() == 0
    s, sd, fl = r.sweep_flags(-1)
    assert o.sweep(-1, 0) == 0
    hs, hsd = o.vec("hist_s0"), o.vec("hist_sdot0")
    assert len(s) == len(hs)
    # integration order vs ascending-s storage; the last integrated point is snapped afterwards
    assert np.array_equal(s[:-1][::-1], hs[1:]) and np.array_equal(sd[:-1][::-1], hsd[1:])
    assert np.array_equal(fl, o.vec("flags0").astype(np.uint8))
    assert r.sweep(-1, 0) == 0  # the real reverse sweep, to set up the forward pass identically
    s, sd, fl = r.sweep_flags(1)
    assert o.sweep(1, 1) == 0
    assert np.array_equal(s[:-1], o.vec("hist_s1")[:-1]) and np.array_equal(sd[:-1], o.vec("hist_sdot1")[:-1])
    assert np.array_equal(fl, o.vec("flags1").astype(np.uint8))


@pytest.mark.parametrize("name", ["RR", "UR5", "KUKA-LWR-IV", "CSPR3DOF"])
def test_automatic_integration_resolution_matches(name, tmp_path):
    """_isAutoIntegRes = true (the library default, ba.h:309; batest switches it off): adjust_s rewrites
    sWeights / scaleType / integRes per trajectory (ba.cpp:471-556).  The restatement must follow the unmodified
    reference through all three phases bit for bit.  (The generic robot has no Cartesian path to derive the
    resolution from: both sides reject GEN7DOF in this mode.)"""
    cfg, r, o = _pair(name, tmp_path, auto=True)
    J = cfg.n_joints
    assert r.interp_input() == 0 and o.interp_input() == 0
    assert r.scalar("nPts") == o.scalar("nPts") and r.scalar("integRes") == o.scalar("integRes")
    for d, last in ((-1, 0), (1, 1)):
        assert r.sweep(d, last) == 0 and o.sweep(d, last) == 0
        assert np.array_equal(r.vec("sMVC"), o.vec("sMVC")) and np.array_equal(r.vec("sdot"), o.vec("sdot"))
    assert r.scalar("tTotalTraj") == o.scalar("tTotalTraj")
    r.interp_output()
    o.interp_output()
    assert r.scalar("outRes") == o.scalar("outRes")
    for j in range(J):
        assert np.array_equal(r.vec("theta", j), o.vec("theta", j)), j
    for j in range(int(r.scalar("cartRows"))):
        assert np.array_equal(r.vec("cart", j), o.vec("cart", j)), j


def test_interpolation_only_mode_matches(tmp_path):
    """_isInterpOnly (ba.cpp:139-159): interpInputData only re-samples the path at outRes and returns -1.  UR5 is
    the stock folder that carries both joint and Cartesian rows (the reference reads traj.cart[i] of an empty
    vector for the joint-only files in this mode, so those have no defined result)."""
    d = P.GOLD + "/stock/UR5/"
    r = Ref(d + "config.dat", d, str(tmp_path) + "/")
    r.set_interp_only(True)
    assert r.load_file() == 0
    cfg, tres, th, ca, ts = P.load_stock("UR5")
    cfg = cfg.copy()
    cfg.is_interp_only = 1
    o = Oracle(cfg)
    o.load_raw(th.shape[2], tres, th[0], ca[0], None if ts is None else ts[0])
    assert r.interp_input() == -1 and o.interp_input() == -1
    assert r.scalar("nPts") == o.scalar("nPts") and r.scalar("sres") == o.scalar("sres") == cfg.out_res
    for j in range(cfg.n_joints):
        assert np.array_equal(r.vec("theta", j), o.vec("theta", j)), j
    assert int(r.scalar("cartRows")) == int(o.scalar("cartRows")) == 6
    for j in range(6):
        assert np.array_equal(r.vec("cart", j), o.vec("cart", j)), j


def test_batch_runner_agrees(tmp_path):
    import ctypes as C
    from _oracle import orc_lib, ref_lib
    B = 12
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 500, B)
    cfgp = (P.GOLD + "/synthetic/GEN7DOF_config.dat").encode()
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    outs = []
    for fn, first in ((ref_lib().ref_batch_run, cfgp), (orc_lib().orc_batch_run, C.byref(cfg))):
        tt = np.zeros(B); nr = np.zeros(B, np.int32); nf = np.zeros(B, np.int32); no = np.zeros(B, np.int32)
        st = np.zeros(B, np.int32); out = np.zeros((B, 7, 4096), np.float32)
        el = fn(first, B, th.shape[2], tres, th.ctypes.data_as(fp), None, 2, tt.ctypes.data_as(dp),
                nr.ctypes.data_as(ip), nf.ctypes.data_as(ip), no.ctypes.data_as(ip), st.ctypes.data_as(ip),
                out.ctypes.data_as(fp), 4096)
        assert el > 0
        outs.append((tt, nr, nf, no, st, out))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["GEN7DOF", "RR", "UR5", "CSPR3DOF", "KUKA-LWR-IV"])
@pytest.mark.parametrize("decim,window", [(3, 1), (1, 3), (3, 3), (2, 4)])
def test_input_decimation_and_smoothing_match(name, decim, window, tmp_path):
    """inputDecimFact > 1 / smoothWindow > 1 (ba.cpp:195-242 with util.cpp:254-288 smooth and 343-352 decimate;
    quirk Q6: the smoothWindow branch smooths with inputDecimFact as the window): no shipped config switches them
    on, so the restatement is pinned here against the unmodified reference, all three phases bit for bit."""
    import ctypes as C
    d = P.GOLD + "/stock/" + name + "/"
    cfgp = P.variant_config(name, tmp_path, inputDecimFact=decim, smoothWindow=window)
    r = Ref(cfgp, d, str(tmp_path) + "/")
    assert r.load_file() == 0
    cfg, tres, th, ca, ts = P.load_stock_variant(name, tmp_path, inputDecimFact=decim, smoothWindow=window)
    rc = r.cfg()
    assert bytes(C.string_at(C.byref(cfg), C.sizeof(cfg))) == bytes(C.string_at(C.byref(rc), C.sizeof(rc)))
    assert cfg.input_decim_fact == decim and cfg.smooth_window == window
    o = Oracle(cfg)
    n0 = (th if th is not None else ca).shape[2]
    o.load_raw(n0, tres, None if th is None else th[0], None if ca is None else ca[0], None if ts is None else ts[0])
    J = cfg.n_joints
    ri, oi = r.interp_input(), o.interp_input()
    assert ri == oi
    if ri != 0:
        return
    assert r.scalar("nPts") == o.scalar("nPts")
    for nm in ("theta", "thetaD", "thetaD2"):
        for j in range(J):
            assert np.array_equal(r.vec(nm, j), o.vec(nm, j)), (nm, j)
    for dd, last in ((-1, 0), (1, 1)):
        assert r.sweep(dd, last) == o.sweep(dd, last) == 0
        assert np.array_equal(r.vec("sMVC"), o.vec("sMVC")) and np.array_equal(r.vec("sdot"), o.vec("sdot"))
    assert r.scalar("tTotalTraj") == o.scalar("tTotalTraj")
    r.interp_output()
    o.interp_output()
    for j in range(J):
        assert np.array_equal(r.vec("theta", j), o.vec("theta", j)), j
    for j in range(int(r.scalar("cartRows"))):
        assert np.array_equal(r.vec("cart", j), o.vec("cart", j)), j


@pytest.mark.parametrize("name", P.STOCK)
@pytest.mark.parametrize("start", [1.0e3, 0.5])
def test_per_sample_mvc_is_the_reference_functions(name, start, tmp_path):
    """the per-sample maximum-velocity curve has no reference output of its own; it is DEFINED as
    the reference's private per-point functions (evalSplinePartials, sdotLim, applyAccelConstraintsBisectionPt)
    evaluated at every knot.  ref_mvc_per_sample drives exactly those (oracle/ref_harness.cpp); the restatement
    must agree bit for bit on every robot: joint limits (GEN7DOF), Cartesian (UR5, KUKA), serial torque (RR),
    Par2Ser torque (CSPR3DOF)."""
    cfg, r, o = _pair(name, tmp_path)
    assert r.interp_input() == 0 and o.interp_input() == 0
    a, b = r.mvc_per_sample(start), o.mvc_per_sample(start)
    assert len(a) == len(b) == int(o.scalar("nPtsC")) and len(a) > 100
    assert np.array_equal(a, b)
    assert np.isfinite(b).all() and b.max() <= start


def test_caller_supplied_dynamics_reproduce_the_reference_on_rr(tmp_path):
    """cfg.dyn_source = 1: a1..a4 come from the caller's point function instead of Robot::call_dynSerial.  With
    dynRR restated behind that signature (orc_demo_dyn_rr) the plug-in path of the restatement must give what the
    unmodified reference gives with its built-in model: grid tables, both sweeps, torque rows - bit for bit."""
    from _oracle import dyn_fn_address
    cfg, r, o = _pair("RR", tmp_path)
    cfg2 = cfg.copy()
    cfg2.dyn_source = 1
    o = Oracle(cfg2)
    _, tres, th, ca, ts = P.load_stock("RR")
    o.load_raw(th.shape[2], tres, th[0], None, None)
    o.set_dyn_callback(dyn_fn_address("orc_demo_dyn_rr"))
    assert r.interp_input() == 0 and o.interp_input() == 0
    for nm in ("a1", "a2", "a3", "a4", "a1C_m", "a4C_m"):
        for j in range(2):
            x, y = r.vec(nm, j), o.vec(nm, j)
            if nm.endswith("_m"):
                x, y = x[:-1], y[:-1]
            assert np.array_equal(x, y), (nm, j)
    for d, last in ((-1, 0), (1, 1)):
        assert r.sweep(d, last) == 0 and o.sweep(d, last) == 0
        assert np.array_equal(r.vec("sMVC"), o.vec("sMVC")) and np.array_equal(r.vec("sdot"), o.vec("sdot"))
    r.interp_output()
    o.interp_output()
    for nm in ("theta", "trq"):
        for j in range(2):
            assert np.array_equal(r.vec(nm, j), o.vec(nm, j)), (nm, j)


def test_parallel_torque_without_par2ser_matches(tmp_path):
    """isPar2Ser = 0 on the CSPR3DOF folder (ba.cpp:1463-1491; no shipped config uses it): restatement against the
    unmodified reference, every phase bit for bit."""
    d = P.GOLD + "/stock/CSPR3DOF/"
    r = Ref(P.variant_config("CSPR3DOF", tmp_path, isPar2Ser=0), d, str(tmp_path) + "/")
    assert r.load_file() == 0
    cfg, tres, th, ca, ts = P.load_stock_variant("CSPR3DOF", tmp_path, isPar2Ser=0)
    o = Oracle(cfg)
    o.load_raw(ca.shape[2], tres, None, ca[0], None)
    assert r.interp_input() == 0 and o.interp_input() == 0
    for dd, last in ((-1, 0), (1, 1)):
        assert r.sweep(dd, last) == 0 and o.sweep(dd, last) == 0
        assert np.array_equal(r.vec("sMVC"), o.vec("sMVC")) and np.array_equal(r.vec("sdot"), o.vec("sdot"))
    assert r.scalar("tTotalTraj") == o.scalar("tTotalTraj")
    r.interp_output()
    o.interp_output()
    for nm in ("theta", "trq"):
        for j in range(3):
            assert np.array_equal(r.vec(nm, j), o.vec(nm, j)), (nm, j)
