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


FAKE_SCENE = '''
import numpy as np, torch
class FakeScene:
    """Stands in for dvp_mvs_b200.Scene on the CPU: a view's new depth is a function of the pass, the seed and the depth
    maps of its sources as this rank currently holds them - enough to expose any ordering or exchange mistake."""
    def __init__(self, num_views, sizes):
        self.V, self.sizes = num_views, sizes
        self.depth = {v: torch.full((sizes[0][1], sizes[0][0]), float(v + 1)) for v in range(num_views)}
        self.log = []
    def src(self, v):
        return [(v + 1) % self.V, (v + 2) % self.V]
    def run_view(self, v, level, pass_, seed):
        w, h = self.sizes[level]
        s = sum(float(self.depth[u].double().mean()) for u in self.src(v)) if pass_ > 0 else 0.0   # pass 0: no geometric term
        self.depth[v] = torch.full((h, w), float((seed % 9973) * 1e-3 + 0.5 * s + level), dtype=torch.float32)
        self.log.append((v, level, pass_, seed))
    def depth_tensor(self, v, level, owned):
        w, h = self.sizes[level]
        if not owned and tuple(self.depth[v].shape) != (h, w):
            self.depth[v] = torch.zeros((h, w), dtype=torch.float32)
        return self.depth[v]
'''


def test_scene_schedule_single_process_order_and_seeds():
    ns = {}
    exec(FAKE_SCENE, ns)
    from dvp_mvs_b200.farm import run_scene_schedule
    sc = ns["FakeScene"](3, [(8, 6), (16, 12)])
    owner = run_scene_schedule(sc, 3, 2, seed=100)
    assert owner == {0: 0, 1: 0, 2: 0}
    # 2 levels x 4 passes x 3 views, views in index order inside a pass, seed + 1000 * pass index + view (as dvp_scene_run)
    assert sc.log == [(v, it // 4, it % 4, 100 + 1000 * it + v) for it in range(8) for v in range(3)]
    assert tuple(sc.depth[0].shape) == (12, 16)


def test_scene_schedule_two_ranks_exchange_depth_maps(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text("import os, sys, json\nsys.path.insert(0, %r)\n" % ROOT + FAKE_SCENE + textwrap.dedent(f"""
        import torch.distributed as dist
        from dvp_mvs_b200.farm import run_scene_schedule, partition
        dist.init_process_group("gloo")
        rank = dist.get_rank()
        V, sizes = 5, [(8, 6), (16, 12)]
        sc = FakeScene(V, sizes)
        owner = run_scene_schedule(sc, V, 2, seed=7)
        assert owner == {{0: 0, 1: 1, 2: 0, 3: 1, 4: 0}}, owner
        assert [e[0] for e in sc.log[:2]] == partition(V, 2, rank)[:2]
        # reference semantics: both ranks simulated in one process (own views fresh, remote views from the previous pass)
        sims = [FakeScene(V, sizes) for _ in range(2)]
        it = 0
        for level in range(2):
            for p in range(4):
                for r in range(2):
                    for v in partition(V, 2, r):
                        sims[r].run_view(v, level, p, 7 + 1000 * it + v)
                for v in range(V):
                    for r in range(2):
                        sims[r].depth[v] = sims[owner[v]].depth[v].clone()
                it += 1
        for v in range(V):
            assert torch.equal(sc.depth[v], sims[rank].depth[v]), v
        open(os.path.join({str(tmp_path)!r}, f"rank{{rank}}.json"), "w").write(json.dumps([float(sc.depth[v][0, 0]) for v in range(V)]))
        dist.destroy_process_group()
    """))
    import socket, json
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    a = json.loads((tmp_path / "rank0.json").read_text()); b = json.loads((tmp_path / "rank1.json").read_text())
    assert a == b and len(set(a)) == 5          # every rank ends with every view's depth map, and they are the same maps


def test_two_ranks_gather_their_views_for_fusion(tmp_path):
    """Row N3 after a farmed schedule: every view's plane and pixel-state maps reach the fusing rank unchanged, in view
    order, whoever owned them (stand-in scene and stand-in fusion object; gloo)."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from dvp_mvs_b200.farm import fuse_farmed_scene, partition
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        V = 5
        owner = {{v: r for r in range(world) for v in partition(V, world, r)}}
        def maps(v):
            h, w = 4 + v, 6
            planes = (np.arange(h * w * 4, dtype=np.float32).reshape(h, w, 4) + 1000 * v)
            return planes, np.full((h, w), v % 3, np.uint8), None, None
        class Scene:
            def get_view(self, v):
                assert owner[v] == rank, "a rank may only be asked for the views it owns"
                return maps(v)
        class Fusion:
            def __init__(self, views): self.views = views; self.mode = 0
            def set_mode(self, m): self.mode = m
            def run(self): return np.array([[len(self.views), self.mode, 0, 0, 0, 0]], np.float32), 0.0
        static = [dict(camera=v, image=np.zeros((4 + v, 6, 3), np.uint8), src_views=[(v + 1) % V]) for v in range(V)]
        seen = []
        def make(views):
            seen.extend(views)
            return Fusion(views)
        pts = fuse_farmed_scene(Scene(), owner, static, fuse_rank=0, mode=2, make_fusion=make)
        if rank == 0:
            assert pts.tolist() == [[5.0, 2.0, 0, 0, 0, 0]]
            for v, got in enumerate(seen):
                planes, weak, _, _ = maps(v)
                assert np.array_equal(got["planes"], planes) and np.array_equal(got["weak"], weak)
                assert got["camera"] == v and got["src_views"] == [(v + 1) % V] and got["block"] is None
        else:
            assert pts is None and not seen
        open(os.path.join({str(tmp_path)!r}, f"rank{{rank}}.txt"), "w").write("RANK_OK")
        dist.destroy_process_group()
    """))
    import socket
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "rank0.txt").read_text() == "RANK_OK" and (tmp_path / "rank1.txt").read_text() == "RANK_OK"


def _fill(obj, mv, L, iterations):
    obj.set_max_iterations(iterations)
    for v in range(len(mv.cameras)):
        obj.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        for l in range(L):
            obj.set_level(v, l, mv.levels[l][v]["image"], None, mv.levels[l][v]["label"])
            obj.compute_edges(v, l)
        obj.set_initial_planes(v, mv.planes_init[v])


@pytest.mark.gpu
def test_library_farm_on_one_gpu_equals_the_scene_driver():
    """dvp_farm_* with a single device is dvp_scene_run: same schedule, same seeds; with 0 iterations every stage is
    deterministic, so every view's maps are identical bit for bit."""
    import numpy as np
    from dvp_mvs_b200 import Farm, Scene, synth
    V, L = 4, 2
    mv = synth.make_multiview(320, 240, V, L, seed=3)
    sc = Scene(V, L); fa = Farm([0], V, L)
    _fill(sc, mv, L, 0); _fill(fa, mv, L, 0)
    sc.run(seed=11)
    wall, exch, moved = fa.run(seed=11)
    assert wall > 0 and moved == 0
    for v in range(V):
        a, b = sc.get_view(v), fa.get_view(v)
        for x, y in zip(a, b):
            assert (np.asarray(x).view(np.uint8) == np.asarray(y).view(np.uint8)).all(), v
        assert fa.owner(v) == 0
    sc.close(); fa.close()


@pytest.mark.gpu
def test_library_farm_on_two_gpus_matches_the_block_gauss_seidel_order():
    """Two GPUs: views dealt round robin, depth maps exchanged by peer copies after every pass.  With 0 iterations the result
    is deterministic and must equal what one process computes when it imposes the same visibility of depth maps (own views
    fresh, the other GPU's from the previous pass) — emulated here by two single-GPU scenes fed by hand."""
    import numpy as np
    import torch
    from dvp_mvs_b200 import Farm, Scene, synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    V, L = 4, 2
    mv = synth.make_multiview(320, 240, V, L, seed=3)
    fa = Farm([0, 1], V, L)
    _fill(fa, mv, L, 0)
    wall, exch, moved = fa.run(seed=11)
    assert moved > 0 and [fa.owner(v) for v in range(V)] == [0, 1, 0, 1]
    # emulation on one GPU: scene d plays GPU d; after every pass the fresh depth maps are handed over
    scenes = [Scene(V, L), Scene(V, L)]
    for s in scenes:
        _fill(s, mv, L, 0)
    it = 0
    for level in range(L):
        for p in range(4):
            for d in range(2):
                for v in range(d, V, 2):
                    scenes[d].run_view(v, level, p, 11 + 1000 * it + v)
            for v in range(V):
                src, dst = scenes[v % 2], scenes[1 - v % 2]
                dst.depth_tensor(v, level, False).copy_(src.depth_tensor(v, level, True))
            torch.cuda.synchronize()
            it += 1
    for v in range(V):
        a, b = scenes[v % 2].get_view(v), fa.get_view(v)
        for x, y in zip(a, b):
            assert (np.asarray(x).view(np.uint8) == np.asarray(y).view(np.uint8)).all(), v
    fa.close()
    for s in scenes:
        s.close()


_NCCL_FARM_SCRIPT = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["DVP_REPO_ROOT"])
from dvp_mvs_b200 import Scene, Farm, synth
from dvp_mvs_b200.farm import run_scene_schedule

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V, L = 4, 2
mv = synth.make_multiview(320, 240, V, L, seed=3)


def fill(o):
    o.set_max_iterations(0)                   # no propagation sweeps: every stage of a pass is deterministic
    for v in range(V):
        o.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        for l in range(L):
            o.set_level(v, l, mv.levels[l][v]["image"], None, mv.levels[l][v]["label"])
            o.compute_edges(v, l)
        o.set_initial_planes(v, mv.planes_init[v])


sc = Scene(V, L, device=local); fill(sc)
owner = run_scene_schedule(sc, V, L, seed=11)      # NCCL all-gather of the depth maps after every pass
torch.cuda.synchronize()
mine = {v: [np.ascontiguousarray(x) for x in sc.get_view(v)] for v in range(V) if owner[v] == rank}
gathered = [None, None]
dist.all_gather_object(gathered, mine)
ok = True
if rank == 0:
    views = {}
    for g in gathered:
        views.update(g)
    fa = Farm([0, 1], V, L); fill(fa)               # the library's farm: threads + peer copies, same dealing, same seeds
    fa.run(seed=11)
    for v in range(V):
        for x, y in zip(views[v], fa.get_view(v)):
            ok = ok and (np.asarray(x).view(np.uint8) == np.asarray(y).view(np.uint8)).all()
    fa.close()
open(os.path.join(os.environ["DVP_OUT_DIR"], f"rank{rank}.txt"), "w").write("RANK_OK" if ok else "MISMATCH")
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.gpu
def test_nccl_farm_on_two_gpus_equals_the_library_farm(tmp_path):
    """the NCCL exchange of farm.run_scene_schedule must be ordered against the library's own streams.
    Two ranks under torchrun run the whole schedule with 0 iterations (deterministic) and every view's maps must equal, bit for
    bit, what the in-library farm (threads + peer copies, same block Gauss-Seidel order) computes — a depth map read before its
    all-gather landed, or overwritten while NCCL was still sending it, would show up here."""
    import socket, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "nccl_farm.py"
    script.write_text(_NCCL_FARM_SCRIPT)
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", DVP_REPO_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), DVP_OUT_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert (tmp_path / "rank0.txt").read_text() == "RANK_OK" and (tmp_path / "rank1.txt").read_text() == "RANK_OK"
