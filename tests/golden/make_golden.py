"""Regenerates tests/golden/ (run in the build container, where /root/reference exists).

Inputs : the reference's own example folders  /root/reference/input/{RR,UR5,KUKA-LWR-IV,CSPR3DOF,GEN7DOF}
         (config.dat + path file; DATA fixtures only — no reference source is copied).
Outputs: for each folder the files the UNMODIFIED reference writes (traj_out.dat, s-sdot.dat),
         produced by oracle/_ref/libbatotp_ref.so (reference sources + Eigen stand-in, oracle/Makefile)
         and, for cross-checking, by the prebuilt /root/reference/bin/batest;
         golden.json with step counts, sizes and sha256 fingerprints of both;
         synthetic.json with per-path results of the reference on seeded synthetic paths
         (batotp_b200/synth.py) for the three batch configurations.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _oracle import Ref, ref_lib  # noqa: E402
from batotp_b200 import synth  # noqa: E402
from batotp_b200.config import read_config  # noqa: E402

REF = "/root/reference"
FOLDERS = ["RR", "UR5", "KUKA-LWR-IV", "CSPR3DOF", "GEN7DOF"]


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    gold = {}
    for d in FOLDERS:
        src = os.path.join(REF, "input", d)
        dst = os.path.join(HERE, "stock", d)
        os.makedirs(dst, exist_ok=True)
        cfg, name = read_config(os.path.join(src, "config.dat"))
        for f in ("config.dat", name):
            shutil.copy(os.path.join(src, f), os.path.join(dst, f))
        out = tempfile.mkdtemp() + "/"
        r = Ref(os.path.join(dst, "config.dat"), dst + "/", out)
        assert r.load_file() == 0
        assert r.interp_input() == 0
        n_grid = int(r.scalar("nPts"))
        assert r.sweep(-1, 0) == 0
        assert r.sweep(1, 1) == 0
        r.interp_output()
        r.write_output()
        entry = dict(n_grid=n_grid, n_rev=int(r.scalar("nRev")), n_fwd=int(r.scalar("nFwd")),
                     t_rev=r.scalar("tRev"), t_total=r.scalar("tFwd"), n_out=int(r.scalar("nPts")),
                     out_sres=r.scalar("sres"))
        for f in ("traj_out.dat", "s-sdot.dat"):
            b = open(out + f, "rb").read()
            open(os.path.join(dst, "ref_" + f), "wb").write(b)
            entry["sha256_" + f] = sha(b)
        # prebuilt binary (glibc dynamic) on the same folder
        w = tempfile.mkdtemp()
        for sub in ("bin", "input", "output"):
            os.makedirs(os.path.join(w, sub))
        for f in ("config.dat", name):
            shutil.copy(os.path.join(src, f), os.path.join(w, "input", f))
        shutil.copy(os.path.join(REF, "bin", "batest"), os.path.join(w, "bin", "batest"))
        os.chmod(os.path.join(w, "bin", "batest"), 0o755)
        subprocess.run(["./batest"], cwd=os.path.join(w, "bin"), stdout=subprocess.DEVNULL, check=True)
        for f in ("traj_out.dat", "s-sdot.dat"):
            entry["prebuilt_sha256_" + f] = sha(open(os.path.join(w, "output", f), "rb").read())
        gold[d] = entry
        print(d, entry)
    json.dump(gold, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)

    # synthetic paths through the reference's batch runner
    import ctypes as C
    syn = {}
    L = ref_lib()
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    cfgdir = os.path.join(HERE, "synthetic")
    os.makedirs(cfgdir, exist_ok=True)
    gen_cfg = open(os.path.join(REF, "input/GEN7DOF/config.dat")).read().replace(
        "0        // isBINfile", "1        // isBINfile")
    open(os.path.join(cfgdir, "GEN7DOF_config.dat"), "w").write(gen_cfg)
    shutil.copy(os.path.join(REF, "input/KUKA-LWR-IV/config.dat"), os.path.join(cfgdir, "KUKA_config.dat"))
    shutil.copy(os.path.join(REF, "input/CSPR3DOF/config.dat"), os.path.join(cfgdir, "CSPR3DOF_config.dat"))
    for nm, B, gen in (("GEN7DOF", 64, synth.gen7dof_paths), ("KUKA", 4, synth.kuka_paths),
                       ("CSPR3DOF", 8, synth.cspr_paths)):
        cfgp = os.path.join(cfgdir, nm + "_config.dat")
        cfg, _ = read_config(cfgp)
        tres, pay = gen(0, B)
        th = pay if nm != "CSPR3DOF" else None
        ca = pay if nm == "CSPR3DOF" else None
        J = cfg.n_joints
        cap = 32768
        tt = np.zeros(B); nr = np.zeros(B, np.int32); nf = np.zeros(B, np.int32)
        no = np.zeros(B, np.int32); st = np.zeros(B, np.int32)
        out = np.zeros((B, J, cap), np.float32)
        L.ref_batch_run(cfgp.encode(), B, pay.shape[2], tres,
                        th.ctypes.data_as(fp) if th is not None else None,
                        ca.ctypes.data_as(fp) if ca is not None else None, 4,
                        tt.ctypes.data_as(dp), nr.ctypes.data_as(ip), nf.ctypes.data_as(ip),
                        no.ctypes.data_as(ip), st.ctypes.data_as(ip), out.ctypes.data_as(fp), cap)
        syn[nm] = dict(B=B, payload_sha256=sha(pay.tobytes()), t_total=tt.tolist(), n_rev=nr.tolist(),
                       n_fwd=nf.tolist(), n_out=no.tolist(), status=st.tolist(),
                       theta_out_sha256=[sha(out[b, :, :no[b]].tobytes()) for b in range(B)])
        print(nm, "n_rev", nr[:4], "n_fwd", nf[:4])
    json.dump(syn, open(os.path.join(HERE, "synthetic.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
