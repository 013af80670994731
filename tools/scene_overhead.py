#!/usr/bin/env python
"""Host overhead of the resident scene driver per pass: wall time minus device time (tools, not a test)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
from dvp_mvs_b200 import synth, Scene
mv = synth.make_multiview(3111, 2073, 5, 3, seed=0)
def fill():
    sc = Scene(5, 3)
    for v in range(5):
        sc.set_view(v, mv.cameras[v], 3111, 2073, mv.src_views[v])
        for l in range(3):
            L = mv.levels[l][v]; sc.set_level(v, l, L["image"], L["edge"], L["label"])
        sc.set_initial_planes(v, mv.planes_init[v])
    return sc
w = fill(); w.run(1); w.close()   # warm-up: every kernel launched once, local-memory reservation done
sc = fill()
it = 0; prev = 0.0
for level in range(3):
    for p in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sc.run_pass(level, p, 100 + 1000 * it)
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
        dev, n = sc.stats(); d = dev - prev; prev = dev
        print(f"level {level} pass {p}: wall {wall:8.1f} ms  device {d:8.1f} ms  overhead {wall - d:7.1f} ms ({(wall-d)/5:.1f} per view)")
        it += 1
