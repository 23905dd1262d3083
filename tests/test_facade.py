"""The BATOTP::BA-compatible host facade and its batest driver (batotp_b200/host): same call
sequence and files as the reference's test/main.cpp.  CPU: linked against the host-emulation
build of the kernels; GPU (-m gpu): the shipped batotp_b200/lib/batest."""
import os
import shutil
import subprocess

import pytest

import _parity as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_batest(exe_files, name, tmp_path):
    w = tmp_path / name
    for sub in ("bin", "input", "output"):
        (w / sub).mkdir(parents=True)
    src = os.path.join(P.GOLD, "stock", name)
    for f in os.listdir(src):
        if not f.startswith("ref_"):
            shutil.copy(os.path.join(src, f), w / "input" / f)
    for f in exe_files:
        shutil.copy(f, w / "bin" / os.path.basename(f))
    exe = "./" + os.path.basename(exe_files[0])
    out = subprocess.run([exe], cwd=w / "bin", capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    g = P.golden_json()[name]
    assert "rev. integ.: %4d steps" % g["n_rev"] in out.stdout
    assert "fwd. integ.: %4d steps" % g["n_fwd"] in out.stdout
    for f in ("traj_out.dat", "s-sdot.dat"):
        assert open(w / "output" / f, "rb").read() == open(os.path.join(src, "ref_" + f), "rb").read(), f
    assert os.path.getsize(w / "output" / "compTimes.dat") == 12


@pytest.mark.parametrize("name", P.STOCK)
def test_batest_facade_emulated(name, tmp_path):
    import __graft_entry__ as g
    g.build_emu()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "batotp_b200", "host"), "emu"], check=True)
    emu = os.path.join(ROOT, "tests", "_emu")
    _run_batest([os.path.join(emu, "batest_emu"), os.path.join(emu, "libbatotp_emu.so")], name, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", P.STOCK)
def test_batest_facade_on_gpu(name, tmp_path):
    lib = os.path.join(ROOT, "batotp_b200", "lib")
    _run_batest([os.path.join(lib, "batest"), os.path.join(lib, "libbatotp_cuda.so")], name, tmp_path)
