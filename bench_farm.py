"""bench_farm.py — `bench.py --workload farm`, BASELINE config C4: a Tanks&Temples-shaped scene (V views of 1920x1080, two pyramid levels
480x270 and 960x540, 4 sources per view) run through the WHOLE multi-scale schedule (2 rounds x 4 passes x V views,
reference main.cpp:449-511) with the views of every pass farmed over the N ranks and the one exchange step per
pass — every view's fresh depth map, all a view needs from its sources — INSIDE the timed region.  Strong scaling: the
scene is fixed, N varies.

  ours       one resident scene per rank (dvp_scene_*), `farm.run_scene_schedule`: NCCL all-gather between device buffers
  reference  the reference's own kernels per (view, pass) with the maps chained through host arrays the way main() chains
             them through files (oracle/host_chain.py: rescales, visibility restoration on the CPU as ProcessProblem does),
             views dealt to the ranks the same way (the reference picks its GPU by argv, main.cpp:430-434), depth maps
             exchanged by broadcasts of host arrays — what N copies of the reference on a shared file system amount to.

A step = one whole schedule.  value = pixels of all (view, pass) jobs / wall time between barriers (the schedule is driven
from the host: dozens of launches and two levels of maps per job, so there is no single stream to put CUDA events on).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

V_VIEWS, FULL_W, FULL_H, LEVELS, NUM_SRC = 16, 1920, 1080, 2, 4


def _scene_cache(seed):
    return f"/tmp/dvp_farm_scene_{V_VIEWS}x{FULL_W}x{FULL_H}_L{LEVELS}_S{NUM_SRC}_seed{seed}.pkl"


def make_scene(seed=0):
    import pickle
    from dvp_mvs_b200 import synth
    path = _scene_cache(seed)
    if os.path.exists(path):
        try:
            return pickle.load(open(path, "rb"))
        except Exception:
            pass
    mv = synth.make_multiview(FULL_W, FULL_H, V_VIEWS, LEVELS, seed=seed, num_src=NUM_SRC)
    try:
        tmp = path + f".{os.getpid()}.tmp"
        pickle.dump(mv, open(tmp, "wb"), protocol=4)
        os.replace(tmp, path)
    except Exception:
        pass
    return mv


def main_library(args):
    """`--farm-driver library`: the same schedule through dvp_farm_* (dvp_mvs_b200/csrc/dvp_farm.inc) — ONE process, one host
    thread and one resident scene per GPU inside the library, depth maps exchanged by peer copies; what a C++ host would call."""
    import torch
    from dvp_mvs_b200 import Farm
    n = min(args.gpus, torch.cuda.device_count())
    if n < 1:
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    mv = make_scene(0)
    fa = Farm(list(range(n)), V_VIEWS, LEVELS)
    for v in range(V_VIEWS):
        fa.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        for l in range(LEVELS):
            fa.set_level(v, l, mv.levels[l][v]["image"], None, mv.levels[l][v]["label"])
            fa.compute_edges(v, l)
    pixels = sum(mv.levels[l][v]["w"] * mv.levels[l][v]["h"] for l in range(LEVELS) for v in range(V_VIEWS)) * 4
    walls, exchs, moved = [], [], 0
    for step in range(args.warmup + args.steps):
        for v in range(V_VIEWS):
            fa.set_initial_planes(v, mv.planes_init[v])
        for d in range(n):
            torch.cuda.synchronize(d)
        wall, exch, moved = fa.run(seed=7)
        if step >= args.warmup:
            walls.append(wall); exchs.append(exch)
    ms = float(np.mean(walls))
    print(json.dumps({"metric": "scene_schedule_mpix_per_s", "value": pixels / (ms / 1e3) / 1e6, "unit": "Mpix/s", "n_gpus": n, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "driver": "library farm (dvp_farm_run: one host thread + one scene per GPU, peer copies)",
                      "config": {"workload": f"tnt_shaped_schedule_V{V_VIEWS}_{FULL_W}x{FULL_H}_L{LEVELS}_S{NUM_SRC}_it3", "views": V_VIEWS,
                                 "full_size": [FULL_W, FULL_H], "src_views": NUM_SRC, "passes_per_view": 4 * LEVELS, "view_pass_jobs": 4 * LEVELS * V_VIEWS,
                                 "timing": "wall clock of dvp_farm_run (host-driven schedule, all GPU threads joined)"},
                      "exchange_ms_per_step": float(np.mean(exchs)), "exchange_share": float(np.mean(exchs)) / ms, "exchange_bytes_per_step": moved}))
    fa.close()


def main(args):
    if getattr(args, "farm_driver", "torchrun") == "library" and args.impl != "reference":
        return main_library(args)
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.abspath(__file__))
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1 and rank != 0:
        dist.barrier()
    mv = make_scene(0)
    if world > 1 and rank == 0:
        dist.barrier()
    pixels = sum(mv.levels[l][v]["w"] * mv.levels[l][v]["h"] for l in range(LEVELS) for v in range(V_VIEWS)) * 4
    config = {"workload": f"tnt_shaped_schedule_V{V_VIEWS}_{FULL_W}x{FULL_H}_L{LEVELS}_S{NUM_SRC}_it3", "views": V_VIEWS, "full_size": [FULL_W, FULL_H],
              "levels": [[mv.levels[l][0]["w"], mv.levels[l][0]["h"]] for l in range(LEVELS)], "src_views": NUM_SRC, "passes_per_view": 4 * LEVELS,
              "view_pass_jobs": 4 * LEVELS * V_VIEWS, "sharding": "views of a pass dealt round robin to the ranks; one depth-map exchange per pass, timed",
              "timing": "host wall clock between barriers + device synchronisation (host-driven schedule); max over ranks"}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    exch = [0.0]
    if args.impl == "reference":
        sys.path.insert(0, os.path.join(root, "oracle"))
        import ref_oracle
        import host_chain
        os.environ.setdefault("DVP_REF_K2_LIB", "libapd_ref_k2_jit.so")
        if not ref_oracle.available():
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libapd_ref.so was not built (needs /root/reference at build time)"}))
            return
        from dvp_mvs_b200.farm import partition
        owner = {v: r for r in range(world) for v in partition(V_VIEWS, world, r)}

        class Chain(host_chain.HostChain):
            def run_pass(self, level, pass_, seed):
                for v in range(self.V):
                    if owner[v] == rank:
                        self.process_problem(v, level, pass_, seed + v)
                if world > 1:
                    t0 = time.perf_counter()
                    for v in range(self.V):   # depths.dmb of every view becomes visible to every rank
                        w, h = self.mv.levels[level][v]["w"], self.mv.levels[level][v]["h"]
                        if owner[v] == rank:
                            t = torch.from_numpy(np.ascontiguousarray(self.files[v]["planes"])).cuda()
                        else:
                            t = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
                        dist.broadcast(t, src=owner[v])
                        if owner[v] != rank:
                            self.files[v] = dict(self.files[v], planes=t.cpu().numpy())
                    torch.cuda.synchronize()
                    exch[0] += time.perf_counter() - t0

        def one_schedule():
            chain = Chain(mv, lambda w, h, S, p: ref_oracle.engine(w, h, S, p, device=local))
            chain.engines = engines
            chain.run(seed=7)
        engines = {}
    else:
        from dvp_mvs_b200 import Scene
        from dvp_mvs_b200.farm import run_scene_schedule
        from dvp_mvs_b200 import farm as farm_mod
        scene = Scene(V_VIEWS, LEVELS, device=local)
        for v in range(V_VIEWS):
            scene.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
            for l in range(LEVELS):
                scene.set_level(v, l, mv.levels[l][v]["image"], None, mv.levels[l][v]["label"])
                scene.compute_edges(v, l)
        real_exchange = farm_mod.exchange_depths

        def timed_exchange(*a, **k):
            t0 = time.perf_counter()
            real_exchange(*a, **k)
            exch[0] += time.perf_counter() - t0
        farm_mod.exchange_depths = timed_exchange

        def one_schedule():
            for v in range(V_VIEWS):
                scene.set_initial_planes(v, mv.planes_init[v])
            run_scene_schedule(scene, V_VIEWS, LEVELS, seed=7)

    for _ in range(args.warmup):
        one_schedule()
    barrier()
    exch[0] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_schedule()
    barrier()
    wall = time.perf_counter() - t0
    stats = torch.tensor([wall, exch[0]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    wall, exchange = (float(x) for x in stats.tolist())
    if rank == 0:
        ms = 1e3 * wall / args.steps
        out = {"metric": "scene_schedule_mpix_per_s", "value": pixels / (wall / args.steps) / 1e6, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": config, "exchange_ms_per_step": 1e3 * exchange / args.steps, "exchange_share": exchange / wall if wall > 0 else None,
               "view_pass_jobs_per_s": config["view_pass_jobs"] / (wall / args.steps)}
        if args.impl == "reference":
            out["impl"] = "reference"
            out["notes"] = ("the reference's kernels per (view, pass), maps chained through host memory as main() chains them through files, visibility "
                            "restoration on the CPU as ProcessProblem does; views dealt to ranks like ours")
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
