"""world_size-2 (gloo, CPU) check of the N>1 host logic: slab assignment and the max-over-ranks /
sum-over-ranks aggregation bench.py uses for the whole-job events/s."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from mulskips_b200.decomp import slab_grid, rank_block, aggregate
    dist.init_process_group("gloo")
    r, w = dist.get_rank(), dist.get_world_size()
    gx, gy = slab_grid(w)
    assert gx * gy == w
    x0, y0, lx, ly = rank_block(r, w, 4608, 4608)
    assert lx %% 48 == 0 and ly %% 48 == 0
    tot = aggregate(dist, [1.0 + r, 10.0 * (r + 1)], ["max", "sum"], device="cpu")
    assert tot == [float(w), 10.0 * w * (w + 1) / 2], tot
    if r == 0: print("OK", gx, gy, tot)
    dist.destroy_process_group()
""") % ROOT


def test_two_ranks(tmp_path):
    p = tmp_path / "w.py"
    p.write_text(SCRIPT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29543", str(p)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
