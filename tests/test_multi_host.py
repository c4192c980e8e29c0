"""Host-side logic of the multi-GPU path, on CPU: the slab planner (tg_tree_plan_slabs: no device needed) and the
rank rendezvous bench.py performs under torch.distributed.run (world size 2, gloo backend)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
import tangerine_b200 as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,cells", [("seaside_town", 1022), ("gear", 510), ("basic_thing", 64)])
@pytest.mark.parametrize("ranks", [1, 2, 3, 8])
def test_slab_plan_is_a_partition(name, cells, ranks):
    tree = T.Tree.load(O.model_path(name))
    lo, hi = tree.bounds()
    grid = T.export_grid(lo, hi, np.float32((hi[0] - lo[0]) / cells))
    sz = grid.shape[2]
    if sz < 8 * ranks:
        pytest.skip("grid too shallow for that many slabs")
    cuts, cost = tree.plan_slabs(grid, ranks)
    assert len(cuts) == ranks + 1 and cuts[0] == 0 and cuts[-1] == sz
    assert all(b - a >= 8 for a, b in zip(cuts, cuts[1:]))       # every slab owns at least one brick row
    assert len(cost) == sz and (cost >= 0).all() and cost.sum() > 0
    again, _ = tree.plan_slabs(grid, ranks)
    assert again == cuts                                          # deterministic: every device thread plans the same cuts
    if ranks > 1:
        # the estimate is what the cuts balance: no slab may hold more than twice its share of it
        shares = [cost[a:b].sum() / cost.sum() for a, b in zip(cuts, cuts[1:])]
        assert max(shares) <= 2.0 / ranks + 8.0 * cost.max() / cost.sum()


def test_rendezvous_world_size_2_gloo(tmp_path):
    """bench.py's join_ranks / leave_ranks with two processes on the gloo backend: both ranks pass the all-reduce, rank 1
    waits on the host until rank 0 is done."""
    script = tmp_path / "rendezvous.py"
    script.write_text(
        "import os, sys, time\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "rank, world, waiters = bench.join_ranks('gloo')\n"
        "assert world == 2\n"
        "if rank == 0:\n"
        "    time.sleep(0.5)\n"
        "    open(os.path.join(%r, 'rank0_done'), 'w').write('x')\n"
        "bench.leave_ranks(waiters)\n"
        "if rank == 1:\n"
        "    assert os.path.exists(os.path.join(%r, 'rank0_done')), 'rank 1 left before rank 0 finished'\n"
        "print('rank', rank, 'ok')\n" % (ROOT, str(tmp_path), str(tmp_path)))
    import socket
    with socket.socket() as probe:      # a free port: several test sessions may share the machine
        probe.bind(("127.0.0.1", 0))
        port = probe.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
