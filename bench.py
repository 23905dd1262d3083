#!/usr/bin/env python
"""bench.py — trajectories/s of the BA time-optimisation path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU implementation on the host cores

Workload: GEN7DOF synthetic spline paths (input/GEN7DOF/generateGEN7DOFpath.m recipe, seeded;
batotp_b200/synth.py), `--batch` paths per GPU (default 131072, i.e. the 2^20-path configuration
at 8 GPUs; weak scaling: every rank owns its own contiguous slice, no collective on the data path).
A "step" is one pass interpInputData -> sweep(rev) -> sweep(fwd) -> interpOutputData over that batch.

value : inputs already resident in HBM, per-trajectory scalars read back, device time by CUDA events
        on the library's stream, max over ranks.
e2e   : the same metric through the C-ABI call with pinned HOST buffers, host->device copy of the
        float32 paths and device->host copy of the float32 output trajectories inside the timed region.
roofline : FP64 (the path is scalar FP64, no tensor cores): algorithmic flops of the sweep kernel
        (SURVEY §8d counting, from the kernel's own call counters) / its CUDA-event time, against the
        FP64 peak measured on this device by a DFMA micro-kernel (MEASURED_PEAKS.json has none).
cpu_baseline : oracle/_ref (the unmodified reference, "reference") or the oracle port ("port") on the
        host cores, one trajectory per thread, on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "trajectories/sec (GEN7DOF batch)"
UNIT = "trajectories/s"
CFG_PATH = os.path.join(ROOT, "tests", "golden", "synthetic", "GEN7DOF_config.dat")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=131072, help="paths per GPU per step")
    ap.add_argument("--chunk", type=int, default=0,
                    help="paths resident per device pass (input interp + sweeps); 0 = the library's automatic split")
    ap.add_argument("--out-chunk", type=int, default=8192, help="paths per output-interpolation pass")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--lib", default=None, help="experiments only: alternative build of libbatotp_cuda.so")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 6:
                continue
            try:
                sm.append(float(t[0]))
                mx.append(float(t[1]))
            except ValueError:
                continue
            for nm, v in zip(names, t[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        return out


# ----------------------------------------------------------------------------- CPU arm
def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_runner():
    """-> (kind, callable(theta[B,7,400] f32, tres, threads) -> seconds)  (the checker libraries, used here
    only as the timed CPU baseline)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    from batotp_b200.config import read_config
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    if O.ref_available():
        L = O.ref_lib()

        def run(th, tres, nth):
            B = th.shape[0]
            st = np.zeros(B, np.int32)
            return L.ref_batch_run(CFG_PATH.encode(), B, th.shape[2], tres, th.ctypes.data_as(fp), None, nth,
                                   None, None, None, None, st.ctypes.data_as(ip), None, 0)
        return "reference", run
    L = O.orc_lib()
    cfg, _ = read_config(CFG_PATH)

    def run(th, tres, nth):
        B = th.shape[0]
        st = np.zeros(B, np.int32)
        return L.orc_batch_run(C.byref(cfg), B, th.shape[2], tres, th.ctypes.data_as(fp), None, nth,
                               None, None, None, None, st.ctypes.data_as(ip), None, 0)
    return "port", run


def cpu_sample_rate(theta, tres, seconds):
    """Times the CPU implementation on a bounded prefix of the workload. -> dict"""
    kind, run = cpu_runner()
    nth = cpu_threads()
    pilot = min(theta.shape[0], 8 * nth)
    t = run(theta[:pilot], tres, nth)
    rate = pilot / max(t, 1e-9)
    n = int(min(theta.shape[0], max(pilot, rate * seconds)))
    t = run(theta[:n], tres, nth)
    return dict(value=n / t, unit=UNIT, cores=nth, kind=kind,
                sample="first %d of the step's paths, %.1f s, interpInputData+2 sweeps+interpOutputData, "
                       "one trajectory per thread" % (n, t))


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from batotp_b200 import synth
    kind, run = cpu_runner()
    nth = cpu_threads()
    # bounded sample per step: sized from a pilot so that (steps+warmup) steps stay within minutes
    tres, pilot = synth.gen7dof_paths(0, 8 * nth)
    t = run(pilot, tres, nth)
    rate = pilot.shape[0] / max(t, 1e-9)
    per_step = int(max(8 * nth, min(args.batch, rate * max(2.0, 90.0 / (args.steps + args.warmup)))))
    tres, theta = synth.gen7dof_paths(0, per_step)
    for _ in range(args.warmup):
        run(theta, tres, nth)
    t0 = time.time()
    tot = 0.0
    for _ in range(args.steps):
        tot += run(theta, tres, nth)
    wall = time.time() - t0
    value = per_step * args.steps / tot
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * tot / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference",
                config=dict(workload="GEN7DOF synthetic spline paths (20 knots U[0,5]^7 -> 400 pts), "
                                     "stock GEN7DOF config.dat", paths_per_step=per_step,
                            note="bounded sample of the b200 arm's workload; CPU only, rank 0"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=nth, kind=kind,
                                  sample="%d paths per step, %d timed steps, wall %.1f s" % (per_step, args.steps, wall)),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    emit_line(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- B200 arm
def main_b200(args):
    import torch
    import torch.distributed as dist
    from batotp_b200 import native, synth
    from batotp_b200.config import read_config

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg, _ = read_config(CFG_PATH)
    B = args.batch
    J = cfg.n_joints
    # every rank owns the contiguous slice [rank*B, (rank+1)*B) of the seeded path family
    tres, theta = None, None
    parts = []
    for at in range(0, B, 16384):
        tres, p = synth.gen7dof_paths(rank * B + at, min(16384, B - at))
        parts.append(p)
    theta = np.concatenate(parts, axis=0)
    del parts
    n0 = theta.shape[2]
    h_theta = torch.from_numpy(theta).pin_memory()
    d_theta = h_theta.cuda()
    ctx = native.Context(local, args.lib)
    ctx.set_chunk(args.chunk)
    ctx.set_out_chunk(args.out_chunk)
    peak_fma, peak_nofma = ctx.fp64_peak()

    # ---- value: inputs resident in HBM, scalars back --------------------------------
    bi_dev = ctx.make_in(tres=tres, device_ptrs=dict(theta=d_theta.data_ptr(), cart=None, B=B, n0_max=n0))
    res_s = native.BatchResult(B, J, cfg.n_cart, 0, 0, False, want_rows=False, want_hist=False)

    def step_resident():
        ctx.optimize_batch(cfg, bi_dev, res_s)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    ctx.stats_reset()
    clk = ClockSampler(local)
    ctx.timer_start()
    for _ in range(args.steps):
        step_resident()
    ms = ctx.timer_stop_ms()
    barrier()
    clocks = clk.stop()
    st = ctx.stats()
    ms = max_over_ranks(ms)
    ok = int((res_s.status & native.ST_FATAL_MASK == 0).sum())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (the sweeps) ---------------------------------
    steps_rk, ver, ntraj = st["steps"], st["verifies"], max(st["trajectories"], 1)
    n_eval = 6 * steps_rk + 4 * ntraj           # evalSplinePartials calls (A5): 6/step + 2 per sweep prologue
    n_lim = 7 * steps_rk + 4 * ntraj            # sdotLim calls (A2)
    flops = n_eval * (5 + 18 * J) + ver * (2 + 7 * J) + n_lim * (1 + 2 * J + 3) + 110 * steps_rk
    sweep_s = st["sweep_ms"] * 1e-3
    achieved = flops / max(sweep_s, 1e-12) * 1e-12
    # DRAM traffic of the sweep kernel: from the committed ncu --set full capture (profiles/), scaled from the
    # trajectories of the captured launch to the trajectories per launch of this run
    traffic, traffic_src = None, None
    try:
        import csv
        prof = os.path.join(ROOT, "profiles", "r1f_sweep_ncu_full_summary.csv")
        vals = {r[0]: (r[1], float(r[2])) for r in csv.reader(open(prof)) if len(r) == 3 and not r[0].startswith("#")
                and r[0] != "metric"}
        gb = sum(v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                 for k, (u, v) in vals.items() if k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        per_launch_traj = ntraj / max(st["sweep_launches"], 1)
        traffic = gb / 56832.0 * per_launch_traj
        traffic_src = "profiles/r1f_sweep_ncu_full_summary.csv (dram__bytes_read+write of a 56832-path launch, per path)"
    except Exception:
        pass
    roof = dict(bound="fp64", achieved=achieved, peak=peak_fma, unit="TFLOP/s", frac=achieved / max(peak_fma, 1e-12),
                traffic=traffic, traffic_unit="bytes per launch", traffic_source=traffic_src,
                kernel="k_sweep<7,false,false>", launches=st["sweep_launches"],
                avg_launch_ms=st["sweep_ms"] / max(st["sweep_launches"], 1),
                peak_source="measured in this run (DFMA chains); MEASURED_PEAKS.json has no FP64 entry",
                peak_nofma=peak_nofma, frac_of_nofma_peak=achieved / max(peak_nofma, 1e-12),
                flops_per_trajectory=flops / ntraj, share_of_step=sweep_s / max(ms * 1e-3, 1e-12),
                share_note="sum of the sweep launches' own durations / step time; the launch of a tail chunk runs "
                           "beside other kernels (tail overlap), so with a tail this exceeds the sweep's share of "
                           "the timeline (49 % on one full wave: profiles/r1f_launch_list_step.csv)",
                counting="SURVEY 8d: 1 flop per FP64 add/sub/mul/div/sqrt; A5=5+18J, A4=2+7J, A2=4+2J, RK=110/step")
    launches_value = st["launches"]

    # ---- e2e: pinned host inputs in, float32 trajectories out --------------------------------------
    # One C-ABI call per step over the rank's whole batch: inside it the library moves chunk k+1's rows to the
    # device and chunk k's results to the host while it computes.  (If the full-size pinned result buffer cannot
    # be had, one call per resident chunk into a reusable buffer.)
    if args.chunk > 0:
        lanes_chunk = min(args.chunk, B)
    else:  # the library's automatic split (batotp_cuda.cu auto_chunk): full waves of SMs*3*128 resident lanes
        lanes = torch.cuda.get_device_properties(local).multi_processor_count * 3 * 128
        lanes_chunk = max(128, min(lanes, -(-B // 128) * 128))
    out_cap = int(res_s.n_out.max()) + 64 if ok else 4096
    h_np = h_theta.numpy()
    sl = lanes_chunk if args.skip_e2e else B
    try:
        res_e = native.BatchResult(sl, J, 0, out_cap, 0, False, want_rows=True, want_hist=False, pinned=True)
    except Exception:
        sl = lanes_chunk
        res_e = native.BatchResult(sl, J, 0, out_cap, 0, False, want_rows=True, want_hist=False, pinned=True)

    def step_e2e():
        chk = 0.0
        for at in range(0, B, sl):
            n = min(sl, B - at)
            bi = ctx.make_in(theta=h_np[at:at + n], tres=tres)
            ctx.optimize_batch(cfg, bi, res_e)
            chk += float(res_e.t_total[:n].sum())
        return chk

    for _ in range(0 if args.skip_e2e else max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(1 if args.skip_e2e else args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = world * B * args.steps / e2e_s
    nsl = (B + sl - 1) // sl
    e2e = dict(value=e2e_val, unit=UNIT, h2d_bytes_per_step=int(theta.nbytes),
               d2h_bytes_per_step=int(nsl * res_e.d2h_bytes()),
               note="C-ABI batotp_cuda_optimize_batch on pinned host buffers, %d call(s) of %d paths per step "
                    "(resident chunks of %d); float32 theta(t) rows + per-trajectory scalars copied back"
                    % (nsl, sl, lanes_chunk))

    line = None
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload="GEN7DOF synthetic spline paths (generateGEN7DOFpath.m recipe: 20 knots "
                                         "U[0,5]^7 -> not-a-knot spline -> 400 pts, seeded), stock GEN7DOF config.dat "
                                         "(joint vel 5 / acc 10 limits, integRes 0.01, outRes 0.008, outSmoothFact 5)",
                                paths_per_gpu=B, paths_total=world * B, chunk=lanes_chunk,
                                parallelism="independent slices per GPU, no collective",
                                l2="inputs (%.0f MB/GPU) and per-chunk tables exceed the 126 MB L2" % (theta.nbytes / 1e6),
                                optimised=ok, mean_rk_steps=steps_rk / ntraj, mean_verifies_per_stage=ver / max(6 * steps_rk, 1)),
                    clocks=clocks, e2e=e2e, gpu_launches=int(launches_value), roofline=roof)
        if not args.no_cpu_baseline and world >= 1:
            try:
                line["cpu_baseline"] = cpu_sample_rate(theta, tres, args.cpu_seconds)
            except Exception as e:  # the checker libraries are optional at bench time
                line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=cpu_threads(), kind="unavailable", sample=str(e))
        emit_line(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


_REAL_STDOUT = None


def emit_line(text):
    """The one JSON line of this process, on the real stdout (see main)."""
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(text, flush=True)
        return
    os.write(_REAL_STDOUT, (text + "\n").encode())


def main():
    global _REAL_STDOUT
    args = parse()
    # Rank 0's stdout carries exactly one JSON line: whatever libraries print there meanwhile (NCCL's version
    # banner, for one) goes to stderr; emit_line writes to the saved descriptor.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return main_reference(args)
    return main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
