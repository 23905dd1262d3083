"""CPU: host-side logic — option parsing, file formats, the synthetic generators, and that the
C-ABI library loads, exports every symbol include/*.h declares, and refuses to run without a GPU."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import _parity as P
from batotp_b200 import native, synth
from batotp_b200.config import BatotpCfg, pack_traj_out, read_config, read_traj_bin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg_struct_layout_matches_header():
    # 32 ints + (4*7 + 6 + 3 + 6 + 8) doubles, no padding surprises
    assert C.sizeof(BatotpCfg) == 32 * 4 + (4 * 7 + 6 + 3 + 6 + 8) * 8


def test_read_config_stock_values():
    cfg, name = read_config(P.GOLD + "/stock/RR/config.dat")
    assert name == "RRlemniscate.dat" and cfg.n_joints == 2 and cfg.is_trq_on == 1
    assert list(cfg.jnt_trq_min)[:2] == [-60.0, -40.0]  # NAN -> -JntTrqMax (ba.cpp:2020-2028)
    assert list(cfg.s_weights) == [0.0, 1.0, 0.0]
    cfg, _ = read_config(P.GOLD + "/stock/KUKA-LWR-IV/config.dat")
    assert abs(sum(cfg.s_weights) - 1.0) < 1e-15 and cfg.s_weights[1] == 0.1 / 1.1


def test_bin_reader_and_writer_roundtrip(tmp_path):
    tres, n0, th, ca = read_traj_bin(P.GOLD + "/stock/RR/RRlemniscate.dat", 2, 3)
    assert n0 == 3601 and th.shape == (2, 3601) and ca is None
    b = pack_traj_out(0.008, 5, np.arange(10.0).reshape(2, 5), None, None)
    assert len(b) == 4 + 4 + 4 + 40 + 4 + 4


def test_synthetic_generator_is_deterministic():
    g = P.synthetic_json()
    for nm, gen in (("GEN7DOF", synth.gen7dof_paths), ("KUKA", synth.kuka_paths), ("CSPR3DOF", synth.cspr_paths)):
        tres, pay = gen(0, g[nm]["B"])
        assert pay.flags["C_CONTIGUOUS"] and pay.dtype == np.float32
        assert hashlib.sha256(pay.tobytes()).hexdigest() == g[nm]["payload_sha256"]
    a = synth.gen7dof_paths(5, 3)[1]
    b = synth.gen7dof_paths(0, 8)[1][5:8]
    assert np.array_equal(a, b), "paths must depend on (config, index) only"


def _declared_symbols():
    syms = []
    for f in os.listdir(os.path.join(ROOT, "include")):
        txt = open(os.path.join(ROOT, "include", f)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms += re.findall(r"\b(batotp_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(syms))


def test_c_abi_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build_cuda()
    L = native.load()
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), "libbatotp_cuda.so does not export " + s


def test_product_library_refuses_to_run_without_gpu():
    L = native.load()
    if L.batotp_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(native.NativeError):
        native.Context(0)


def test_cspr_paths_are_redrawn_into_the_static_workspace():
    """a CSPR candidate is kept only if the cable tensions that hold the platform at rest stay within
    [1.05, 11.5] N along it (inside the static workspace the reference's bisection cannot fail); a rejected index
    draws its next candidate, so every path depends on (index) only."""
    raw = synth.cspr_paths(0, 64, redraw=False)[1]
    ok = synth.cspr_accept(raw)
    assert 0.5 < ok.mean() < 0.95  # a good part of the raw family leaves the workspace
    tres, pay = synth.cspr_paths(0, 64)
    assert synth.cspr_accept(pay).all() and synth.cspr_accept(pay, stride=1).all()
    assert np.array_equal(pay[ok], raw[ok])  # accepted first draws are kept as they are
    assert np.array_equal(synth.cspr_paths(40, 8)[1], pay[40:48])
    tau = synth.cspr_static_tensions(pay[:4])
    assert tau.shape == (4, pay.shape[2], 3) and tau.min() >= 1.05 and tau.max() <= 11.5
    # the static tensions solve A tau = (0, 0, g) with A's columns the unit vectors along the cables
    x = pay[0, :, 100].astype(np.float64)
    d = x[:, None] - synth.cspr_pmat()
    A = d / np.sqrt((d * d).sum(axis=0, keepdims=True))
    assert np.allclose(A @ tau[0, 100], [0.0, 0.0, 9.81], atol=1e-12)
