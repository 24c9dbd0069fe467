"""world_size-2 (gloo, CPU) check of the N>1 host logic: slab assignment, neighbour ranks, the seam index
shift, the unique-id broadcast bench.py uses to connect the ranks, and the max-over-ranks / sum-over-ranks
aggregation behind the whole-job events/s."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from mulskips_b200.decomp import slab_block, neighbours, to_local, seam_shift, aggregate
    dist.init_process_group("gloo")
    r, w = dist.get_rank(), dist.get_world_size()
    G = 4608 * w
    x0, owned, lx = slab_block(r, w, G)
    assert owned == 4608 and lx == 4608 + 48 and x0 == r * 4608
    left, right = neighbours(r, w)
    assert left == (r - 1) %% w and right == (r + 1) %% w
    # my first owned column is the left neighbour's right halo: same cell, local x differs by `owned`
    xg = x0 + 5
    assert to_local(xg, r, w, G) == 24 + 5
    assert to_local(xg, left, w, G) == 24 + 5 + seam_shift(owned, 0)
    # my last owned column is the right neighbour's left halo
    xg = x0 + owned - 3
    assert to_local(xg, right, w, G) == to_local(xg, r, w, G) + seam_shift(owned, 1)
    # the 128-byte communicator id travels from rank 0 to everyone (bench.py / scripts/slab_check.py)
    ids = [bytes(range(128)) if r == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))
    tot = aggregate(dist, [1.0 + r, 10.0 * (r + 1)], ["max", "sum"], device="cpu")
    assert tot == [float(w), 10.0 * w * (w + 1) / 2], tot
    if r == 0: print("OK", owned, lx, tot)
    dist.destroy_process_group()
""") % ROOT


def test_two_ranks(tmp_path):
    p = tmp_path / "w.py"
    p.write_text(SCRIPT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29543", str(p)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
