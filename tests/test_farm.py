"""Multi-GPU view farm: the host-side logic, exercised with world_size-2 gloo process groups on CPU."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from util import ROOT
from dvp_mvs_b200.farm import partition, run_pass


def test_partition_round_robin_covers_every_view_once():
    for n, w in ((10, 1), (10, 2), (7, 4), (3, 8), (0, 2)):
        owned = [partition(n, w, r) for r in range(w)]
        flat = sorted(v for o in owned for v in o)
        assert flat == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1


def test_partition_lpt_balances_uneven_costs_deterministically():
    costs = [10, 1, 1, 1, 1, 1, 1, 1, 1, 2]
    a = [partition(10, 2, r, costs) for r in range(2)]
    assert sorted(a[0] + a[1]) == list(range(10))
    loads = [sum(costs[v] for v in o) for o in a]
    assert abs(loads[0] - loads[1]) <= 2 and 0 in a[0]
    assert a == [partition(10, 2, r, costs) for r in range(2)]
    with pytest.raises(ValueError):
        partition(3, 2, 0, [1.0])
    with pytest.raises(ValueError):
        partition(3, 2, 5)


def test_single_process_pass():
    out = run_pass(3, lambda v: {"depth": np.full((2, 2), v, np.float32)})
    assert sorted(out) == [0, 1, 2] and (out[2]["depth"] == 2).all()


def test_two_rank_gloo_pass_gathers_every_view(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from dvp_mvs_b200.farm import run_pass, partition
        dist.init_process_group("gloo")
        rank = dist.get_rank()
        seen = []
        def process(v):
            seen.append(v)
            return {{"depth": np.full((3, 4), 10 * v + rank, np.float32), "weak": np.full((3, 4), v, np.uint8)}}
        out = run_pass(5, process)
        assert sorted(out) == [0, 1, 2, 3, 4], out.keys()
        for v in range(5):
            owner = v % 2
            assert (out[v]["depth"] == 10 * v + owner).all() and out[v]["weak"].dtype == np.uint8 and (out[v]["weak"] == v).all()
        assert seen == partition(5, 2, rank)
        open(os.path.join({str(tmp_path)!r}, f"rank{{rank}}.txt"), "w").write(f"RANK_OK {{rank}} {{seen}}")
        dist.destroy_process_group()
    """))
    import socket
    with socket.socket() as sock:   # a free rendezvous port (a fixed one can still be in TIME_WAIT from an earlier run)
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "rank0.txt").read_text() == "RANK_OK 0 [0, 2, 4]"   # one file per rank: stdout of two ranks interleaves
    assert (tmp_path / "rank1.txt").read_text() == "RANK_OK 1 [1, 3]"
