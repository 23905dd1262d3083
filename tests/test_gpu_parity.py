"""GPU (-m gpu): the product library batotp_b200/lib/libbatotp_cuda.so on a B200, through the C-ABI,
against the oracle restatement, the golden reference files and size-independent properties."""
import hashlib

import numpy as np
import pytest

import _parity as P
from batotp_b200 import native

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = native.Context(0)  # product library; raises without a CUDA device
    yield c
    c.close()


@pytest.fixture(params=[1, 2], ids=["lane_kernel", "group_kernel"])
def sweep_kernel(request, ctx):
    """Both sweep kernels on the same cases: 1 = one trajectory per lane (k_sweep.cuh), 2 = a group of lanes per
    trajectory (k_sweep_group.cuh).  Left alone (0) the library picks by chunk size."""
    ctx.set_sweep_kernel(request.param)
    yield request.param
    ctx.set_sweep_kernel(0)


@pytest.mark.parametrize("name", P.STOCK)
def test_stock_folders_reproduce_reference_files(ctx, name, sweep_kernel):
    cfg, tres, th, ca, ts = P.load_stock(name)
    res = P.run_device(ctx, cfg, tres, th, ca, ts)
    assert res.status[0] & native.ST_FATAL_MASK == 0
    g = P.golden_json()[name]
    assert (res.n_grid[0], res.n_rev[0], res.n_fwd[0], res.n_out[0]) == (g["n_grid"], g["n_rev"], g["n_fwd"], g["n_out"])
    assert res.t_total[0] == g["t_total"]  # total trajectory time: exact, hence within 1e-9 relative
    d = P.GOLD + "/stock/" + name
    assert P.device_traj_out_bytes(cfg, res, 0) == open(d + "/ref_traj_out.dat", "rb").read()
    assert P.device_s_sdot_bytes(res, 0) == open(d + "/ref_s-sdot.dat", "rb").read()
    orc = P.OracleRun(cfg, tres, None if th is None else th[0], None if ca is None else ca[0],
                      None if ts is None else ts[0])
    assert P.compare(cfg, res, 0, orc) == []  # includes the limit-active switching flags


@pytest.mark.parametrize("name,count,check", [("GEN7DOF", 512, 48), ("KUKA", 8, 4), ("CSPR3DOF", 32, 8)])
def test_synthetic_batches_match_oracle_and_golden(ctx, name, count, check, sweep_kernel):
    g = P.synthetic_json()[name]
    cfg, tres, th, ca = P.load_synth(name, 0, count)
    res = P.run_device(ctx, cfg, tres, th, ca)
    for b in range(min(count, g["B"])):
        assert (res.n_rev[b], res.n_fwd[b], res.n_out[b]) == (g["n_rev"][b], g["n_fwd"][b], g["n_out"][b]), b
        assert res.t_total[b] == g["t_total"][b]
        assert hashlib.sha256(res.theta_out[b, :, :res.n_out[b]].tobytes()).hexdigest() == g["theta_out_sha256"][b]
    rng = np.random.RandomState(1)
    for b in rng.choice(count, check, replace=False):
        orc = P.OracleRun(cfg, tres, None if th is None else th[b], None if ca is None else ca[b])
        assert P.compare(cfg, res, b, orc) == [], b


def test_chunking_does_not_change_results(ctx):
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 1000, 300)
    ctx.set_chunk(128)
    a = P.run_device(ctx, cfg, tres, th, None, out_cap=4096, hist_cap=4096)
    ctx.set_chunk(16384)
    ctx.set_out_chunk(100)
    b = P.run_device(ctx, cfg, tres, th, None, out_cap=4096, hist_cap=4096)
    ctx.set_out_chunk(8192)
    for nm in ("status", "n_rev", "n_fwd", "n_out", "t_total", "theta_out", "hist", "flags"):
        assert np.array_equal(getattr(a, nm), getattr(b, nm)), nm


def test_tail_overlap_does_not_change_results(ctx):
    """A last chunk that fits one sweep CTA per SM runs on the library's second context beside the full chunks
    (batotp_cuda_set_tail_overlap); every output and the work counters are the same as with one context."""
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 7000, 2 * 1024 + 300)
    ctx.set_chunk(1024)
    try:
        runs = []
        for on in (True, False, True):
            ctx.set_tail_overlap(on)
            ctx.stats_reset()
            r = P.run_device(ctx, cfg, tres, th, None, out_cap=4096, hist_cap=4096)
            runs.append((r, ctx.stats()))
    finally:
        ctx.set_tail_overlap(True)
        ctx.set_chunk(16384)
    (a, sa), (b, sb), (c, sc) = runs
    assert (a.status & native.ST_FATAL_MASK == 0).all()
    for nm in ("status", "n_rev", "n_fwd", "n_out", "t_total", "theta_out", "hist", "flags"):
        assert np.array_equal(getattr(a, nm), getattr(b, nm)), nm
        assert np.array_equal(getattr(a, nm), getattr(c, nm)), nm
    for k in ("verifies", "steps", "trajectories", "sweep_launches"):
        assert sa[k] == sb[k] == sc[k], k


def test_ragged_short_and_degenerate_inputs(ctx):
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 7, 6)
    th = th.copy()
    n0 = np.array([400, 57, 3, 2, 1, 400], dtype=np.int32)
    th[5, :, :] = th[5, :, :1]
    th[1, :, 20:25] = th[1, :, 19:20]
    res = P.run_device(ctx, cfg, tres, th, None, n0=n0)
    for b in range(6):
        orc = P.OracleRun(cfg, tres, th[b], None, n0=int(n0[b]))
        assert P.compare(cfg, res, b, orc) == [], (b, res.status[b])
    assert res.status[4] & 1 and res.status[5] & (1 | 2)


def test_large_batch_properties(ctx):
    """BASELINE-size behaviour through properties that need no oracle run per path: every path is
    optimised, total time == integRes * steps, results do not depend on the position in the batch or on
    the run (idempotence), and the output respects the joint limits the config imposes."""
    B = 16384
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 50000, B)
    res = native.BatchResult(B, 7, 0, 3200, 0, False, want_hist=False)
    ctx.optimize_batch(cfg, ctx.make_in(th, None, tres), res)
    assert np.all(res.status & native.ST_FATAL_MASK == 0)
    assert np.array_equal(res.t_total, cfg.integ_res * (res.n_fwd - 1))
    assert res.n_fwd.min() > 1000 and res.n_fwd.max() < 3200 and np.all(res.n_fwd > res.n_rev)
    perm = np.random.RandomState(3).permutation(B)
    res2 = native.BatchResult(B, 7, 0, 3200, 0, False, want_hist=False)
    ctx.optimize_batch(cfg, ctx.make_in(np.ascontiguousarray(th[perm]), None, tres), res2)
    assert np.array_equal(res2.n_fwd, res.n_fwd[perm]) and np.array_equal(res2.n_rev, res.n_rev[perm])
    assert np.array_equal(res2.theta_out, res.theta_out[perm])
    # joint velocity of the output trajectory stays within the 5 rad/s limit (+ bisection/interp tolerance)
    for b in range(64):
        n = int(res.n_out[b])
        v = np.diff(res.theta_out[b, :, :n].astype(np.float64), axis=1) / res.out_sres[b]
        assert np.abs(v).max() < 5.0 * 1.05, b
    # spot-check 8 random paths against the oracle
    for b in np.random.RandomState(5).choice(B, 8, replace=False):
        orc = P.OracleRun(cfg, tres, th[b], None)
        assert P.compare(cfg, res, b, orc, check_hist=False) == [], b


def test_per_sample_mvc_matches_oracle(ctx):
    from _oracle import Oracle
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 3, 4)
    ctx.load(cfg, ctx.make_in(th, None, tres))
    ctx.interp_input()
    got = ctx.mvc_per_sample(4, 2048, 1.0e3)
    for b in range(4):
        o = Oracle(cfg)
        o.load_raw(400, tres, th[b])
        assert o.interp_input() == 0
        want = o.mvc_per_sample(1.0e3)
        assert np.array_equal(got[b, :len(want)], want)


def test_device_trig_mode_is_within_tolerance(ctx):
    """trig_mode 0 (CUDA sincos in the KUKA forward kinematics) is not bit-identical to the host libm;
    it must stay within the bisection tolerance (1e-3 relative) on the s-sdot profile and within
    a handful of RK steps on the switching counts."""
    cfg, tres, th, ca, ts = P.load_stock("KUKA-LWR-IV")
    strict = P.run_device(ctx, cfg, tres, th, ca, ts)
    cfg2 = cfg.copy()
    cfg2.trig_mode = 0
    fast = P.run_device(ctx, cfg2, tres, th, ca, ts)
    assert fast.status[0] & native.ST_FATAL_MASK == 0
    assert abs(int(fast.n_fwd[0]) - int(strict.n_fwd[0])) <= 8
    assert abs(fast.t_total[0] - strict.t_total[0]) / strict.t_total[0] < 2e-3
    n = min(int(fast.n_out[0]), int(strict.n_out[0]))
    # degrees over a 20 s move: a timing shift of a few integration steps (allowed above) at joint speeds of
    # some tens of deg/s moves theta(t) by up to ~1 degree at a fixed output index
    assert np.abs(fast.theta_out[0, :, :n] - strict.theta_out[0, :, :n]).max() < 2.0


def test_shared_reciprocal_division_is_ieee(ctx):
    """The sweep kernel divides several numerators by one denominator through a shared refined
    reciprocal (k_sweep.cuh sdiv::).  It must return the compiler's correctly rounded '/' bit for bit."""
    total_fast = 0
    for seed in (1, 2, 3):
        bad, fast = ctx.selftest_div(seed, 600_000_000)
        assert bad == 0
        total_fast += fast
    assert total_fast > 1_000_000_000  # the fast path really is what gets exercised


def test_large_batch_matches_oracle_bit_for_bit(ctx, sweep_kernel):
    """4096 GEN7DOF paths (BASELINE configs[2] size) through the C-ABI against the oracle restatement run on
    the host cores: switching counts, total time and every float32 output sample.  The sweep kernel takes its
    bisection decisions from certified float models of the bounds and forms only the binding quotients exactly; a wrong
    certificate anywhere would change a step count here (the algorithm is chaotic at the last bit)."""
    import ctypes as C
    import os
    from _oracle import orc_lib
    from batotp_b200 import synth
    from batotp_b200.config import read_config
    cfg, _ = read_config(os.path.join(P.GOLD, "synthetic", "GEN7DOF_config.dat"))
    B = 4096
    tres, th = synth.gen7dof_paths(500000, B)
    out_cap = 3200
    res = native.BatchResult(B, cfg.n_joints, 0, out_cap, 0, False, want_rows=True, want_hist=False)
    ctx.set_chunk(4096)
    ctx.optimize_batch(cfg, ctx.make_in(theta=th, tres=tres), res)
    ctx.set_chunk(16384)
    L = orc_lib()
    tt = np.zeros(B)
    nr, nf, no, st = (np.zeros(B, np.int32) for _ in range(4))
    rows = np.zeros((B, cfg.n_joints, out_cap), np.float32)
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.orc_batch_run(C.byref(cfg), B, th.shape[2], tres, th.ctypes.data_as(fp), None, os.cpu_count() or 1,
                    tt.ctypes.data_as(dp), nr.ctypes.data_as(ip), nf.ctypes.data_as(ip), no.ctypes.data_as(ip),
                    st.ctypes.data_as(ip), rows.ctypes.data_as(fp), out_cap)
    assert (res.status & native.ST_FATAL_MASK == 0).all() and (st == 0).all()
    assert np.array_equal(res.n_rev, nr) and np.array_equal(res.n_fwd, nf) and np.array_equal(res.n_out, no)
    assert np.array_equal(res.t_total, tt)
    assert int(no.max()) <= out_cap
    assert np.array_equal(res.theta_out, rows)


@pytest.mark.parametrize("name", ["RR", "UR5", "KUKA-LWR-IV", "CSPR3DOF"])
def test_automatic_integration_resolution(ctx, name, sweep_kernel):
    """_isAutoIntegRes = true (per-trajectory integRes / weights, ba.cpp:471-556)."""
    cfg, tres, th, ca, ts = P.load_stock(name)
    cfg = cfg.copy()
    cfg.is_auto_integ_res = 1
    res = P.run_device(ctx, cfg, tres, th, ca, ts)
    assert res.status[0] & native.ST_FATAL_MASK == 0
    orc = P.OracleRun(cfg, tres, None if th is None else th[0], None if ca is None else ca[0],
                      None if ts is None else ts[0])
    assert orc.ok and P.compare(cfg, res, 0, orc) == []


def test_interpolation_only_mode(ctx):
    """_isInterpOnly (ba.cpp:139-159) — the path is only re-sampled at outRes.  UR5 carries
    joint and Cartesian rows (axis-angle -> quaternion -> axis-angle around the resample)."""
    from _oracle import Oracle
    cfg, tres, th, ca, ts = P.load_stock("UR5")
    cfg = cfg.copy()
    cfg.is_interp_only = 1
    res = P.run_device(ctx, cfg, tres, th, ca, ts, out_cap=8192)
    o = Oracle(cfg)
    o.load_raw(th.shape[2], tres, th[0], ca[0], None if ts is None else ts[0])
    assert o.interp_input() == -1
    n = int(o.scalar("nPts"))
    assert res.status[0] == 0 and res.n_out[0] == n and res.n_cart_out[0] == n and res.out_sres[0] == cfg.out_res
    assert np.array_equal(res.theta_out[0, :, :n], o.rows("theta", cfg.n_joints).astype(np.float32))
    assert np.array_equal(res.cart_out[0, :6, :n], o.rows("cart", 6).astype(np.float32))
    assert res.n_rev[0] == 0 and res.n_fwd[0] == 0 and not res.hist.any()
    # a batch of joint-only paths: absent Cartesian rows are zeros (the reference has no defined result there)
    cfg2, tres2, th2, _ = P.load_synth("GEN7DOF", 0, 5)
    cfg2 = cfg2.copy()
    cfg2.is_interp_only = 1
    r2 = P.run_device(ctx, cfg2, tres2, th2, None, out_cap=1024)
    assert (r2.status == 0).all() and (r2.n_out == r2.n_out[0]).all() and r2.n_out[0] > 400
    assert np.abs(r2.theta_out[:, :, 0] - th2[:, :, 0]).max() == 0  # a spline interpolates its first knot exactly


@pytest.mark.parametrize("acc,vel,integ", [(0.05, 1.0, 1.0), (20.0, 0.2, 1.0), (1.0, 5.0, 1.0), (1.0, 1.0, 0.25),
                                           (300.0, 30.0, 0.5)])
def test_limit_regimes_match_oracle(ctx, acc, vel, integ, sweep_kernel):
    """GEN7DOF paths with the limits and the step scaled far away from the stock values (acceleration-starved,
    velocity-bound, nearly unconstrained, fine step): the float certificates of the sweep kernel are relative to
    the bounds they see, so every regime must still equal the oracle bit for bit (rcp.approx and FFMA on the
    device, against the plain divisions of the host emulation in tests/test_emu_parity.py)."""
    B = 48
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 20000, B)
    for i in range(cfg.n_joints):
        cfg.jnt_acc_max[i] *= acc
        cfg.jnt_vel_max[i] *= vel
    cfg.integ_res *= integ
    res = P.run_device(ctx, cfg, tres, th, None, out_cap=65536, hist_cap=65536)
    assert (res.status & native.ST_FATAL_MASK == 0).all()
    for b in range(0, B, 3):
        orc = P.OracleRun(cfg, tres, th[b], None)
        assert P.compare(cfg, res, b, orc) == [], b


# ----------------------------------------------------------------------------- round 2
def _oracle_batch(cfg, tres, th, ca, out_cap):
    """orc_batch_run on all host cores -> (t_total, n_rev, n_fwd, n_out, status, theta_out[B,J,out_cap])."""
    import ctypes as C
    import os
    from _oracle import orc_lib
    ref = th if th is not None else ca
    B = ref.shape[0]
    tt = np.zeros(B)
    nr, nf, no, st = (np.zeros(B, np.int32) for _ in range(4))
    rows = np.zeros((B, cfg.n_joints, out_cap), np.float32)
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    orc_lib().orc_batch_run(C.byref(cfg), B, ref.shape[2], tres, None if th is None else th.ctypes.data_as(fp),
                            None if ca is None else ca.ctypes.data_as(fp), os.cpu_count() or 1,
                            tt.ctypes.data_as(dp), nr.ctypes.data_as(ip), nf.ctypes.data_as(ip),
                            no.ctypes.data_as(ip), st.ctypes.data_as(ip), rows.ctypes.data_as(fp), out_cap)
    return tt, nr, nf, no, st, rows


def test_strict_trig_port_reproduces_the_host_libm_on_1e9_arguments(ctx):
    """cfg.trig_mode 1: the device's sin / cos (k_trig.cuh, a port of the host libm's algorithm in the arithmetic
    glibc selected on this machine) against this host's sin / cos: 10^9 arguments = 2*10^9 values, bit for bit."""
    bad, variant = ctx.selftest_trig(20261017, 1_000_000_000)
    assert variant in (1, 3)
    assert bad == 0


def test_branch_free_bracket_update_on_the_device(ctx):
    assert ctx.selftest_bisect(99, 50_000_000) == 0


def test_kuka_4096_matches_oracle_bit_for_bit(ctx, sweep_kernel):
    """BASELINE configs[2] at its stated size: 4096 synthetic KUKA-LWR-IV paths (Cartesian velocity / acceleration
    limits through the forward kinematics, ~18 600 RK steps per sweep), strict trigonometry ON THE DEVICE, against
    the oracle restatement on the host cores: switching counts, total time and every float32 output sample."""
    B, slab = 4096, 1024
    cfg, tres, th, _ = P.load_synth("KUKA", 0, B)
    assert cfg.trig_mode == 1
    scal = native.BatchResult(B, cfg.n_joints, cfg.n_cart, 0, 0, False, want_rows=False, want_hist=False)
    ctx.optimize_batch(cfg, ctx.make_in(theta=th, tres=tres), scal)
    assert (scal.status & native.ST_FATAL_MASK == 0).all()
    out_cap = int(scal.n_out.max()) + 8
    for at in range(0, B, slab):
        res = native.BatchResult(slab, cfg.n_joints, cfg.n_cart, out_cap, 0, False, want_rows=True, want_hist=False)
        ctx.optimize_batch(cfg, ctx.make_in(theta=th[at:at + slab], tres=tres), res)
        tt, nr, nf, no, st, rows = _oracle_batch(cfg, tres, th[at:at + slab], None, out_cap)
        assert (st == 0).all()
        assert np.array_equal(res.n_rev, nr) and np.array_equal(res.n_fwd, nf) and np.array_equal(res.n_out, no)
        assert np.array_equal(res.t_total, tt) and np.array_equal(res.t_total, scal.t_total[at:at + slab])
        assert np.array_equal(res.theta_out, rows)


def test_cspr_8192_matches_oracle_bit_for_bit(ctx, sweep_kernel):
    """BASELINE configs[3] (an eighth of its stated size; bench.py --workload cspr runs all 65 536): 8192 synthetic
    CSPR3DOF paths inside the static workspace (cable-tension limits through dynCSPR3DOF + setA + Par2Ser LU,
    Cartesian-driven, ~4300-knot grids) against the oracle on the host cores, bit for bit; plus the raw candidate
    family, where some paths leave the workspace (failed bisections, crawling sweeps): same statuses / counts for
    everything the step ceiling lets finish."""
    B = 8192
    cfg, tres, _, ca = P.load_synth("CSPR3DOF", 100000, B)
    scal = native.BatchResult(B, cfg.n_joints, cfg.n_cart, 0, 0, True, want_rows=False, want_hist=False)
    ctx.optimize_batch(cfg, ctx.make_in(cart=ca, tres=tres), scal)
    assert (scal.status & native.ST_FATAL_MASK == 0).all()
    assert (scal.status & native.ST_BISECT_FAIL == 0).all()  # inside the static workspace no bisection fails
    out_cap = int(scal.n_out.max()) + 8
    res = native.BatchResult(B, cfg.n_joints, cfg.n_cart, out_cap, 0, True, want_rows=True, want_hist=False)
    ctx.optimize_batch(cfg, ctx.make_in(cart=ca, tres=tres), res)
    tt, nr, nf, no, st, rows = _oracle_batch(cfg, tres, None, ca, out_cap)
    assert (st == 0).all()
    assert np.array_equal(res.n_rev, nr) and np.array_equal(res.n_fwd, nf) and np.array_equal(res.n_out, no)
    assert np.array_equal(res.t_total, tt)
    assert np.array_equal(res.theta_out, rows)
    # cable-tension rows, Cartesian rows and sLastSec of a sample (the batch oracle returns the joint rows only)
    rng = np.random.RandomState(3)
    for b in rng.choice(B, 24, replace=False):
        assert P.compare(cfg, res, b, P.OracleRun(cfg, tres, None, ca[b]), check_hist=False) == [], b


def test_ur5_fine_discretisation(ctx, sweep_kernel):
    """BASELINE configs[1]: the UR5 path at fine discretisation (integRes 0.008 -> 0.001, norm resolutions / 10,
    outRes 0.001: ten times the grid and the steps), joint + Cartesian limits, axis-angle rows: bit for bit."""
    cfg, tres, th, ca, ts = P.load_stock("UR5")
    cfg = cfg.copy()
    cfg.integ_res = 0.001
    cfg.theta_norm_res, cfg.theta_norm_res2 = cfg.theta_norm_res / 10, cfg.theta_norm_res2 / 10
    cfg.cart_norm_res, cfg.cart_norm_res2 = cfg.cart_norm_res / 10, cfg.cart_norm_res2 / 10
    cfg.out_res = 0.001
    res = P.run_device(ctx, cfg, tres, th, ca, ts, out_cap=65536, hist_cap=65536)
    assert res.status[0] & native.ST_FATAL_MASK == 0 and res.n_fwd[0] > 5000
    orc = P.OracleRun(cfg, tres, th[0], ca[0], ts[0])
    assert orc.ok and P.compare(cfg, res, 0, orc) == []


@pytest.mark.parametrize("name,max_step_diff,max_deg", [("RR", 8, 2.0), ("UR5", 8, 2.0)])
def test_cuda_trig_mode_on_rr_and_ur5(ctx, name, max_step_diff, max_deg):
    """trig_mode 0: CUDA's own sin / cos in fwdKinRR + dynRR (RR) and aa2q / q2aa (UR5).  Not bit-identical to
    the reference; stated tolerance: switching counts within 8 RK steps, total time within 2e-3 relative,
    theta(t) within 2 degrees at a fixed output index (a shift of a few steps at some tens of deg/s)."""
    cfg, tres, th, ca, ts = P.load_stock(name)
    strict = P.run_device(ctx, cfg, tres, th, ca, ts)
    c0 = cfg.copy()
    c0.trig_mode = 0
    fast = P.run_device(ctx, c0, tres, th, ca, ts)
    assert fast.status[0] & native.ST_FATAL_MASK == 0
    assert abs(int(fast.n_fwd[0]) - int(strict.n_fwd[0])) <= max_step_diff
    assert abs(int(fast.n_rev[0]) - int(strict.n_rev[0])) <= max_step_diff
    assert abs(fast.t_total[0] - strict.t_total[0]) / strict.t_total[0] < 2e-3
    n = min(int(fast.n_out[0]), int(strict.n_out[0]))
    assert np.abs(fast.theta_out[0, :, :n] - strict.theta_out[0, :, :n]).max() < max_deg
    if fast.cart_out is not None:
        nc = min(int(fast.n_cart_out[0]), int(strict.n_cart_out[0]))
        assert np.abs(fast.cart_out[0, :3, :nc] - strict.cart_out[0, :3, :nc]).max() < 0.02  # metres


@pytest.mark.parametrize("name", ["RR", "KUKA-LWR-IV", "UR5"])
def test_host_evaluated_trig_mode_gives_the_same_bytes(ctx, name):
    cfg, tres, th, ca, ts = P.load_stock(name)
    d = P.GOLD + "/stock/" + name
    c2 = cfg.copy()
    c2.trig_mode = 2
    res = P.run_device(ctx, c2, tres, th, ca, ts)
    assert P.device_traj_out_bytes(c2, res, 0) == open(d + "/ref_traj_out.dat", "rb").read()
    assert P.device_s_sdot_bytes(res, 0) == open(d + "/ref_s-sdot.dat", "rb").read()


@pytest.mark.parametrize("name,decim,window", [("GEN7DOF", 3, 1), ("GEN7DOF", 2, 4), ("RR", 3, 3), ("UR5", 1, 3),
                                               ("UR5", 3, 3), ("CSPR3DOF", 3, 3), ("KUKA-LWR-IV", 2, 4)])
def test_input_decimation_and_smoothing(ctx, name, decim, window, tmp_path):
    cfg, tres, th, ca, ts = P.load_stock_variant(name, tmp_path, inputDecimFact=decim, smoothWindow=window)
    res = P.run_device(ctx, cfg, tres, th, ca, ts)
    orc = P.OracleRun(cfg, tres, None if th is None else th[0], None if ca is None else ca[0],
                      None if ts is None else ts[0])
    assert orc.ok and res.status[0] & native.ST_FATAL_MASK == 0
    assert P.compare(cfg, res, 0, orc) == []


def test_results_into_device_resident_buffers(ctx):
    """batotp_batch_out.on_device: scalars, rows, histories and flags land in caller-owned HBM."""
    import torch
    B, cap = 700, 4096
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 7000, B)
    ctx.set_chunk(256)
    ctx.set_out_chunk(100)
    try:
        a = P.run_device(ctx, cfg, tres, th, None, out_cap=cap, hist_cap=cap)
        dev = {}
        b = native.BatchResult(B, cfg.n_joints, cfg.n_cart, cap, cap, False)
        for nm in ("status", "n_rev", "n_fwd", "n_out", "n_cart_out", "n_grid", "t_total", "t_rev", "s_last_sec",
                   "out_sres", "theta_out", "hist", "flags"):
            host = getattr(b, nm)
            dev[nm] = torch.zeros(host.shape, dtype={"int32": torch.int32, "float64": torch.float64,
                                                     "float32": torch.float32, "uint8": torch.uint8}[str(host.dtype)],
                                  device="cuda:0")
            ptr = dev[nm].data_ptr()
            if nm in ("theta_out", "hist", "flags"):
                setattr(b.c, nm, ptr)
            else:
                import ctypes as C
                setattr(b.c, nm, C.cast(ptr, type(getattr(b.c, nm))))
        b.c.cart_out = None
        b.c.on_device = 1
        ctx.optimize_batch(cfg, ctx.make_in(th, None, tres), b)
        torch.cuda.synchronize()
    finally:
        ctx.set_chunk(16384)
        ctx.set_out_chunk(8192)
    for nm, t in dev.items():
        assert np.array_equal(getattr(a, nm), t.cpu().numpy()), nm


def test_stragglers_and_two_contexts(ctx):
    """Stragglers (a few trajectories outgrowing the step capacity are re-run after the batch) and two contexts
    with different configurations driven in an interleaved order, on the device."""
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 100, 300)
    ctx.set_step_hint(0)
    a = P.run_device(ctx, cfg, tres, th, None, out_cap=4096, hist_cap=4096)
    steps = np.sort(np.maximum(a.n_rev, a.n_fwd))
    try:
        ctx.set_step_hint(int(steps[-4]))
        ctx.stats_reset()
        b = P.run_device(ctx, cfg, tres, th, None, out_cap=4096, hist_cap=4096)
        st = ctx.stats()
    finally:
        ctx.set_step_hint(0)
    assert st["sweep_launches"] == 2 and st["trajectories"] == 300
    for nm in ("status", "n_rev", "n_fwd", "n_out", "t_total", "s_last_sec", "theta_out", "hist", "flags"):
        assert np.array_equal(getattr(a, nm), getattr(b, nm)), nm
    other = native.Context(0)
    try:
        cfgB, tresB, thB, _ = P.load_synth("KUKA", 0, 2)
        ctx.load(cfg, ctx.make_in(th[:4], None, tres))
        ctx.interp_input()
        other.load(cfgB, other.make_in(thB, None, tresB))
        other.interp_input()
        ctx.sweeps()
        other.sweeps()
        ctx.interp_output()
        other.interp_output()
        ra = ctx.fetch(native.BatchResult(4, cfg.n_joints, cfg.n_cart, 8192, 8192, False))
        rb = other.fetch(native.BatchResult(2, cfgB.n_joints, cfgB.n_cart, 32768, 32768, False))
        for k in range(4):
            assert P.compare(cfg, ra, k, P.OracleRun(cfg, tres, th[k], None)) == []
        for k in range(2):
            assert P.compare(cfgB, rb, k, P.OracleRun(cfgB, tresB, thB[k], None)) == []
    finally:
        other.close()


@pytest.mark.parametrize("name", ["RR", "UR5", "CSPR3DOF", "KUKA-LWR-IV"])
def test_per_sample_mvc_on_cartesian_and_torque_robots(ctx, name):
    """SURVEY 8a A10 on robots with Cartesian limits (UR5, KUKA), serial torque (RR) and Par2Ser torque (CSPR3DOF):
    k_mvc against the oracle, which tests/test_oracle_vs_reference.py pins to the reference's private per-point
    functions on the same folders."""
    from _oracle import Oracle
    cfg, tres, th, ca, ts = P.load_stock(name)
    ctx.load(cfg, ctx.make_in(th, ca, tres, timestamp=ts))
    ctx.interp_input()
    o = Oracle(cfg)
    n0 = (th if th is not None else ca).shape[2]
    o.load_raw(n0, tres, None if th is None else th[0], None if ca is None else ca[0], None if ts is None else ts[0])
    assert o.interp_input() == 0
    for start in (1.0e3, 0.5):
        want = o.mvc_per_sample(start)
        got = ctx.mvc_per_sample(1, len(want) + 8, start)
        assert np.array_equal(got[0, :len(want)], want), start


def _kuka_torque_cfg(cfg):
    """The KUKA torque variant of SURVEY 8d C3: the stock KUKA options plus torque limits on a caller-supplied model."""
    c = cfg.copy()
    c.is_trq_on = 1
    c.dyn_source = 1
    for j, v in enumerate((60.0, 60.0, 30.0, 30.0, 15.0, 15.0, 8.0)):
        c.jnt_trq_max[j] = v
        c.jnt_trq_min[j] = -v
    return c


def test_caller_supplied_dynamics(ctx):
    """cfg.dyn_source = 1 (batotp_cuda_set_dyn_callback): (a) dynRR behind the plug-in signature gives the reference's
    RR files byte for byte; (b) the KUKA-LWR-IV stock path with torque limits on a 7-joint model the reference does
    not have (k_sweep<7,true,true>): same host point function on both sides, device against the oracle bit for bit
    - switching flags, torque rows and all - with the limits actually binding."""
    from _oracle import dyn_fn_address
    cfg, tres, th, ca, ts = P.load_stock("RR")
    c1 = cfg.copy()
    c1.dyn_source = 1
    try:
        ctx.set_dyn_callback(dyn_fn_address("orc_demo_dyn_rr"))
        res = P.run_device(ctx, c1, tres, th, ca, ts)
        d = P.GOLD + "/stock/RR"
        assert P.device_traj_out_bytes(c1, res, 0) == open(d + "/ref_traj_out.dat", "rb").read()
        assert P.device_s_sdot_bytes(res, 0) == open(d + "/ref_s-sdot.dat", "rb").read()
        cfg, tres, th, ca, ts = P.load_stock("KUKA-LWR-IV")
        plain = P.run_device(ctx, cfg, tres, th, ca, ts)
        c2 = _kuka_torque_cfg(cfg)
        fn = dyn_fn_address("orc_demo_dyn_serial")
        ctx.set_dyn_callback(fn)
        res = P.run_device(ctx, c2, tres, th, ca, ts)
        orc = P.OracleRun(c2, tres, th[0], None, dyn_fn=fn)
        assert orc.ok and res.status[0] & native.ST_FATAL_MASK == 0
        assert P.compare(c2, res, 0, orc) == []
        assert res.t_total[0] > plain.t_total[0]  # the torque limits bind: the move takes longer
    finally:
        ctx.set_dyn_callback(None)
    c3 = _kuka_torque_cfg(cfg)
    with pytest.raises(native.NativeError, match="dyn_source = 1 needs a point function"):
        P.run_device(ctx, c3, tres, th, ca, ts)


def test_kuka_torque_variant_batch(ctx, sweep_kernel):
    """SURVEY 8d C3, torque variant: synthetic KUKA paths with torque limits on a caller-supplied 7-joint model, both
    sweep kernels (k_sweep<7,true,true> / k_sweep_group<7,true,true>), against the oracle fed the same point function."""
    from _oracle import dyn_fn_address
    B = 40
    cfg, tres, th, _ = P.load_synth("KUKA", 300, B)
    c2 = _kuka_torque_cfg(cfg)
    fn = dyn_fn_address("orc_demo_dyn_serial")
    try:
        ctx.set_dyn_callback(fn)
        res = P.run_device(ctx, c2, tres, th, None, out_cap=49152, hist_cap=49152)
    finally:
        ctx.set_dyn_callback(None)
    assert (res.status & native.ST_FATAL_MASK == 0).all()
    for b in range(0, B, 5):
        orc = P.OracleRun(c2, tres, th[b], None, dyn_fn=fn)
        assert orc.ok and P.compare(c2, res, b, orc) == [], b


def test_ragged_rows_hold_the_same_samples(ctx):
    """batotp_batch_out.row_offset: joint rows packed at their own length ([J][n_out] blocks, the payload of
    trajWriteBIN) instead of the pitch of the longest trajectory - same float32 samples as the pitched layout, blocks
    disjoint and inside the capacity, also across chunks, output sub-chunks, the tail helper and re-run stragglers."""
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 400, 21)
    a = P.run_device(ctx, cfg, tres, th, None, out_cap=4096, hist_cap=64)
    total = int(a.n_out.sum())
    steps = np.sort(np.maximum(a.n_rev, a.n_fwd))
    ctx.set_chunk(8)
    ctx.set_out_chunk(3)
    try:
        for hint in (0, int(steps[-3])):
            ctx.set_step_hint(hint)
            b = native.BatchResult(21, cfg.n_joints, cfg.n_cart, 0, 0, False, want_hist=False, ragged_cap=total + 5)
            ctx.optimize_batch(cfg, ctx.make_in(th, None, tres), b)
            assert np.array_equal(a.n_out, b.n_out) and np.array_equal(a.t_total, b.t_total)
            spans = sorted((int(b.row_offset[k]), int(b.row_offset[k]) + int(b.n_out[k])) for k in range(21))
            assert spans[0][0] >= 0 and spans[-1][1] <= total and all(x[1] <= y[0] for x, y in zip(spans, spans[1:]))
            for k in range(21):
                assert np.array_equal(b.rows(k), a.theta_out[k, :, :a.n_out[k]]), k
        small = native.BatchResult(21, cfg.n_joints, cfg.n_cart, 0, 0, False, want_hist=False, ragged_cap=total // 2)
        with pytest.raises(native.NativeError, match="ragged_cap"):
            ctx.optimize_batch(cfg, ctx.make_in(th, None, tres), small)
    finally:
        ctx.set_step_hint(0)
        ctx.set_chunk(16384)
        ctx.set_out_chunk(8192)
    # torque rows travel the same way (RR: serial torque)
    cfg, tres, th, ca, ts = P.load_stock("RR")
    p = P.run_device(ctx, cfg, tres, th, ca, ts)
    r = native.BatchResult(1, cfg.n_joints, cfg.n_cart, 0, 0, True, want_hist=False, ragged_cap=4096)
    ctx.optimize_batch(cfg, ctx.make_in(th, ca, tres, timestamp=ts), r)
    n = int(p.n_out[0])
    assert np.array_equal(r.rows(0), p.theta_out[0, :, :n]) and np.array_equal(r.rows(0, "trq_out"), p.trq_out[0, :, :n])


def test_parallel_torque_without_par2ser(ctx, tmp_path):
    """cable-tension limits of the CSPR3DOF with isPar2Ser = 0 (ba.cpp:1463-1491: the structure
    matrix is rebuilt at every point and every verification solves 2 x 3 three-by-three systems with a replaced
    column) - k_sweep<3,true,true,PAR> against the oracle, which tests/test_oracle_vs_reference.py pins to the
    unmodified reference for this option."""
    cfg, tres, th, ca, ts = P.load_stock_variant("CSPR3DOF", tmp_path, isPar2Ser=0)
    assert cfg.is_par2ser == 0
    res = P.run_device(ctx, cfg, tres, th, ca, ts)
    orc = P.OracleRun(cfg, tres, None, ca[0])
    assert orc.ok and res.status[0] & native.ST_FATAL_MASK == 0
    assert P.compare(cfg, res, 0, orc) == []
    cfg2, tres2, _, ca2 = P.load_synth("CSPR3DOF", 50, 3)
    cfg2 = cfg2.copy()
    cfg2.is_par2ser = 0
    r2 = P.run_device(ctx, cfg2, tres2, None, ca2)
    for b in range(3):
        assert P.compare(cfg2, r2, b, P.OracleRun(cfg2, tres2, None, ca2[b])) == [], b


def test_max_integration_time_is_reported_like_the_reference(ctx):
    """maxIntegTime (ba.cpp:1117-1122): a sweep that does not reach the end of the path within maxIntegTime/integRes
    steps ends with BA::MAX_INTEGRATION_TIME in the reference (sweep returns -1); here the trajectory gets
    BATOTP_ST_MAX_INTEG_TIME (not the internal step ceiling) and no output, the rest of the batch is untouched."""
    cfg, tres, th, _ = P.load_synth("GEN7DOF", 900, 4)
    ref = P.run_device(ctx, cfg, tres, th, None)
    c2 = cfg.copy()
    c2.max_integ_time = 0.6 * float(ref.t_total.min())  # every path needs longer than this
    res = P.run_device(ctx, c2, tres, th, None)
    assert ((res.status & 16) != 0).all() and ((res.status & 32) == 0).all() and (res.n_out == 0).all()
    for b in range(2):
        orc = P.OracleRun(c2, tres, th[b], None)
        assert not orc.ok
    c3 = cfg.copy()
    c3.max_integ_time = 0.5 * (float(np.sort(ref.t_total)[1]) + float(np.sort(ref.t_total)[2]))  # two of the four finish
    mix = P.run_device(ctx, c3, tres, th, None)
    done = (mix.status & native.ST_FATAL_MASK) == 0
    assert 1 <= done.sum() <= 3
    for b in range(4):
        orc = P.OracleRun(c3, tres, th[b], None)
        assert orc.ok == bool(done[b]), b
        if done[b]:
            assert P.compare(c3, mix, b, orc) == [], b


@pytest.mark.gpu
@pytest.mark.parametrize("name", P.STOCK)
def test_walker_kernels_give_the_same_bytes(ctx, name):
    """Both shapes of the sequential walkers of interpInputData (one thread per trajectory; point-parallel norm
    increments + k_march_group, 16 lanes per trajectory) reproduce the reference's files on every stock robot."""
    cfg, tres, th, ca, ts = P.load_stock(name)
    d = P.GOLD + "/stock/" + name
    for mode in (1, 2):
        ctx.set_walker_kernel(mode)
        try:
            res = P.run_device(ctx, cfg, tres, th, ca, ts)
        finally:
            ctx.set_walker_kernel(0)
        assert P.device_traj_out_bytes(cfg, res, 0) == open(d + "/ref_traj_out.dat", "rb").read(), mode
        assert P.device_s_sdot_bytes(res, 0) == open(d + "/ref_s-sdot.dat", "rb").read(), mode


@pytest.mark.gpu
@pytest.mark.parametrize("name,count", [("GEN7DOF", 3000), ("KUKA", 96), ("CSPR3DOF", 300)])
def test_walker_kernels_on_batches(ctx, name, count):
    """The two walker shapes on batches (many groups per CTA, ragged lengths): every result array identical, and a
    sample of the paths bit for bit against the oracle."""
    cfg, tres, th, ca = P.load_synth(name, 1000, count)
    n0 = np.full(count, (th if th is not None else ca).shape[2], np.int32)
    n0[::7] = np.maximum(n0[::7] // 3, 5)
    runs = []
    for mode in (1, 2):
        ctx.set_walker_kernel(mode)
        try:
            runs.append(P.run_device(ctx, cfg, tres, th, ca, n0=n0, out_cap=65536, hist_cap=65536))
        finally:
            ctx.set_walker_kernel(0)
    a, b = runs
    for nm in ("status", "n_rev", "n_fwd", "n_out", "n_grid", "t_total", "s_last_sec", "theta_out", "hist", "flags"):
        assert np.array_equal(getattr(a, nm), getattr(b, nm)), nm
    for k in (0, 7, count - 1):
        orc = P.OracleRun(cfg, tres, None if th is None else th[k], None if ca is None else ca[k], n0=int(n0[k]))
        assert P.compare(cfg, b, k, orc) == [], k
