#!/usr/bin/env python
"""GPU diagnostic of the WEAK path: pass 1 (FIRST_INIT, all STRONG) on the reference produces the inputs of a
REFINE_INIT pass with use_APD=1; that pass is then compared stage by stage (product vs reference)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ref_oracle
from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT, REFINE_INIT, REFINE_ITER
from dvp_mvs_b200.parity import step_compare

W, H, S = (int(v) for v in (sys.argv[1:4] + [320, 240, 2][len(sys.argv) - 1:]))
geom = int(sys.argv[4]) if len(sys.argv) > 4 else 0
sc = synth.make_scene(W, H, S)
p = default_params(); p.max_iterations = 2; p.num_images = S + 1
p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
p.use_APD = 0; p.state = FIRST_INIT; p.weak_peak_radius = 6
ref = ref_oracle.engine(W, H, S, p)
ref.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
ref.run(mode=0)
planes, weak, sel, rad = ref.download()
print("pass 1 states: weak/strong/unknown", np.bincount(weak.ravel(), minlength=3))
q = default_params(); q.max_iterations = 1; q.num_images = S + 1
q.depth_min, q.depth_max = sc.depth_min, sc.depth_max
q.use_APD = 1; q.state = REFINE_ITER if geom else REFINE_INIT; q.weak_peak_radius = 6; q.use_detail = 1
q.ransac_threshold = 0.00875; q.rotate_time = 2; q.geom_consistency = geom
kw = dict(images=sc.images, depths=sc.depths if geom else None, cameras=sc.cameras, planes=planes, selected_views=sel, weak_info=weak,
          edge=sc.edge, label=sc.label, radius=rad, seed=synth.SEED_RNG + 1)
ref2 = ref_oracle.engine(W, H, S, q); prod = Engine(W, H, S, q)
ref2.upload(**kw); prod.upload(**kw)
print("weak_count", ref2.weak_count(), prod.weak_count())
res = step_compare(ref2, prod, 1, log=print)
bad = [r for r in res if r.get("error") or r["mismatched"]]
print("stages with mismatches:", sorted(set((r["stage"], r.get("buffer")) for r in bad)))

# K2 candidate offsets: the reference leaves empty sectors uninitialised (B17), so compare per (pixel, view) record
ref2.upload(**kw); prod.upload(**kw)
ref2.run_stage("K1_INIT_RANDOM_STATES"); prod.run_stage("K1_INIT_RANDOM_STATES")
ref2.run_stage("K2_GEN_EDGE_INFORM"); prod.run_stage("K2_GEN_EDGE_INFORM")
ca, cb = ref2.get("candidate"), prod.get("candidate")
rec_eq = (ca == cb).reshape(H, W, 4, -1).all(-1)[:, :, :S]
# records where the whole 11x11 window sees the view: all 12 sectors are populated, nothing is undefined
seen = np.stack([((sel >> v) & 1).astype(bool) for v in range(S)], -1)
from scipy.ndimage import minimum_filter
full = np.stack([minimum_filter(seen[..., v].astype(np.uint8), size=11, mode="constant", cval=0) > 0 for v in range(S)], -1)
print("candidate records equal: all %.4f | fully visible windows %.6f (%d records)" % (rec_eq.mean(), rec_eq[full].mean(), full.sum()))
# how much did the weak path actually do?
ref2.upload(**kw); prod.upload(**kw)
ref2.run(mode=0); prod.run()
for n in ("planes", "costs", "selected", "weak", "radius"):
    a, b = ref2.get(n), prod.get(n)
    eq = (a == b) | ((a != a) & (b != b)) if a.dtype.kind == "f" else (a == b)
    print("end-to-end", n, "differing px", int((~eq.reshape(H, W, -1).all(-1)).sum()))
print("ref ms", ref2.last_run_times()[0], [round(v, 2) for v in ref2.last_run_times()[1]])
print("prod ms", prod.last_run_times()[0], [round(v, 2) for v in prod.last_run_times()[1]])
nb = prod.get("neighbours"); print("anchors per weak px (mean)", (nb[:, 1:, 0] >= 0).sum(1).mean(), "reliable", prod.get("weak_reliable").sum())
