#!/usr/bin/env python
"""Same schedule, one GPU, three drivers: dvp_scene_run (C loop, calling thread), farm.run_scene_schedule (Python loop over
dvp_scene_run_view), dvp_farm_run (C loop on a worker thread).  Prints wall ms per schedule for each, three repetitions."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dvp_mvs_b200 import Scene, Farm, synth
from dvp_mvs_b200.farm import run_scene_schedule

V, L = 6, 2
mv = synth.make_multiview(1920, 1080, V, L, seed=0, num_src=4)


def fill(o):
    for v in range(V):
        o.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        for l in range(L):
            o.set_level(v, l, mv.levels[l][v]["image"], None, mv.levels[l][v]["label"])
            o.compute_edges(v, l)


def planes(o):
    for v in range(V):
        o.set_initial_planes(v, mv.planes_init[v])


sc = Scene(V, L); fill(sc)
fa = Farm([0], V, L); fill(fa)
for rep in range(4):
    planes(sc); torch.cuda.synchronize(); t = time.perf_counter(); dev_ms = sc.run(seed=7); torch.cuda.synchronize(); a = 1e3 * (time.perf_counter() - t)
    planes(sc); torch.cuda.synchronize(); t = time.perf_counter(); run_scene_schedule(sc, V, L, seed=7); torch.cuda.synchronize(); b = 1e3 * (time.perf_counter() - t)
    planes(fa); torch.cuda.synchronize(); t = time.perf_counter(); wall, exch, moved = fa.run(seed=7); torch.cuda.synchronize(); c = 1e3 * (time.perf_counter() - t)
    print(f"rep {rep}: dvp_scene_run {a:.1f} ms (device {dev_ms:.1f})   python loop {b:.1f} ms   dvp_farm_run {c:.1f} ms (its own clock {wall:.1f})", flush=True)
