"""CPU, world_size 2 over gloo: the multi-GPU logic of bench.py that does not need a GPU —
contiguous per-rank slices of the seeded path family and the max-over-ranks reduction."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_slices_are_disjoint_and_reduction_is_max(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import os, sys, hashlib
        sys.path.insert(0, %r)
        import numpy as np, torch, torch.distributed as dist
        from batotp_b200 import synth
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        B = 8
        tres, th = synth.gen7dof_paths(r * B, B)
        full = synth.gen7dof_paths(0, w * B)[1]
        assert np.array_equal(th, full[r * B:(r + 1) * B])          # contiguous slice of one family
        h = torch.tensor([int(hashlib.sha256(th.tobytes()).hexdigest()[:12], 16)], dtype=torch.int64)
        hs = [torch.zeros_like(h) for _ in range(w)]
        dist.all_gather(hs, h)
        assert len({int(x) for x in hs}) == w                        # ranks hold different paths
        t = torch.tensor([10.0 + r], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == 10.0 + (w - 1)                            # timing is the max over ranks
        dist.barrier()
        print("rank", r, "ok")
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
