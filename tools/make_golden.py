#!/usr/bin/env python
"""Generates tests/golden/*.npz ON A GPU BOX from the reference's own kernels (oracle/_ref, the unmodified
APD.cu compiled for sm_100a).  The reference ships no golden vectors, so these runs of the
reference itself are what pins the CPU restatement (oracle/cpu) and, on machines without the reference
library, the product kernels.

    gpurun -- python tools/make_golden.py        # writes gpurun_out/golden/, copy into tests/golden/

Files:
  tex_probe.npz      hardware bilinear texture fetches at random coordinates (pins the filter emulation)
  c1_64x48.npz       BASELINE config C1 shrunk to 64x48: FIRST_INIT, S=2, 1 iteration, all STRONG;
                     the buffers every stage writes, in launch order (stage outputs chain into the next inputs)
  sparse_128x96.npz  K7/K8 on a sparse STRONG mask that makes the reference's racy direction-4 read
                     harmless, so the sweep is deterministic
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ref_oracle, cpu_oracle
from dvp_mvs_b200 import default_params, synth, FIRST_INIT, WEAK, STRONG
from dvp_mvs_b200.parity import sequence, STAGE_OUTPUTS, STATE_BUFS

out_dir = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out_dir, exist_ok=True)


def params_for(sc, S, iters, use_apd):
    p = default_params(); p.max_iterations = iters; p.num_images = S + 1
    p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
    p.use_APD = use_apd; p.state = FIRST_INIT; p.weak_peak_radius = 6
    return p


# ---- C1 shrunk -------------------------------------------------------------------------------------------
W, H, S = 64, 48, 2
sc = synth.make_scene(W, H, S)
p = params_for(sc, S, 1, 0)
ref = ref_oracle.engine(W, H, S, p)
kw = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
ref.upload(**kw)
gold = dict(images=sc.images, cameras=sc.cameras.view(np.uint8), planes_init=sc.planes_init, edge=sc.edge, label=sc.label,
            depth_min=np.float32(sc.depth_min), depth_max=np.float32(sc.depth_max), seed=np.uint64(synth.SEED_RNG))
for k, (stage, it) in enumerate(sequence(1)):
    ref.run_stage(stage, it)
    for n in STAGE_OUTPUTS[stage]:
        if n in ("neighbours", "label_boundary", "complex", "candidate"):
            continue
        gold[f"{k:02d}_{stage}__{n}"] = ref.get(n)
np.savez_compressed(os.path.join(out_dir, "c1_64x48.npz"), **gold)

# ---- texture probe ---------------------------------------------------------------------------------------
rng = np.random.default_rng(7)
n = 4096
xy = np.stack([rng.uniform(-3, W + 3, n), rng.uniform(-3, H + 3, n)], 1).astype(np.float32)
xy[:256] = np.round(xy[:256] * 2) / 2          # exact texel centres / edges
xy[256:512, 0] = np.floor(xy[256:512, 0]) + 0.5 + rng.integers(0, 512, 256) / 512.0   # half-steps of the 8-bit fraction
vals = cpu_oracle.tex_probe(ref, 1, xy)
np.savez_compressed(os.path.join(out_dir, "tex_probe.npz"), image=sc.images[1], xy=xy, values=vals)

# ---- sparse K7/K8 ----------------------------------------------------------------------------------------
W, H, S = 128, 96, 2
sc = synth.make_scene(W, H, S)
p = params_for(sc, S, 2, 1)
yy, xx = np.mgrid[0:H, 0:W]
weak = np.where(((xx + yy) % 128) < 2, STRONG, WEAK).astype(np.uint8)
ref = ref_oracle.engine(W, H, S, p)
ref.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, weak_info=weak, seed=synth.SEED_RNG)
for st in ("K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM", "K6_RANDOM_INITIALIZATION"):
    ref.run_stage(st)
gold = dict(images=sc.images, cameras=sc.cameras.view(np.uint8), planes_init=sc.planes_init, edge=sc.edge, label=sc.label, weak=weak,
            depth_min=np.float32(sc.depth_min), depth_max=np.float32(sc.depth_max), seed=np.uint64(synth.SEED_RNG))
for n in ("planes", "costs", "selected", "rand", "edge_neigh", "radius", "view_weight"):
    gold["pre__" + n] = ref.get(n)
k = 0
for it in range(2):
    for st in ("K7_BLACK_STRONG", "K8_RED_STRONG"):
        ref.run_stage(st, it)
        for n in STAGE_OUTPUTS[st]:
            a = ref.get(n)
            gold[f"{k:02d}_{st}_{it}__{n}"] = a[weak == STRONG]   # only STRONG pixels can change
        k += 1
np.savez_compressed(os.path.join(out_dir, "sparse_128x96.npz"), **gold)
# ---- WEAK path: pass 2 (rounds >= 1 parameters) on the output of a reference pass 1 ------------------------------
from dvp_mvs_b200 import REFINE_INIT
W, H, S = 96, 72, 2
sc = synth.make_scene(W, H, S)
p1 = params_for(sc, S, 2, 0)
ref = ref_oracle.engine(W, H, S, p1)
ref.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
ref.run(mode=0)
planes1, weak1, sel1, rad1 = ref.download()
q = params_for(sc, S, 1, 1)
q.state = REFINE_INIT; q.use_detail = 1; q.ransac_threshold = 0.00875; q.rotate_time = 2
ref2 = ref_oracle.engine(W, H, S, q)
ref2.upload(images=sc.images, cameras=sc.cameras, planes=planes1, selected_views=sel1, weak_info=weak1, edge=sc.edge, label=sc.label,
            radius=rad1, seed=synth.SEED_RNG + 1)
gold = dict(images=sc.images, cameras=sc.cameras.view(np.uint8), edge=sc.edge, label=sc.label, planes_in=planes1, weak_in=weak1,
            selected_in=sel1, radius_in=rad1, depth_min=np.float32(sc.depth_min), depth_max=np.float32(sc.depth_max),
            seed=np.uint64(synth.SEED_RNG + 1))
for k, (stage, it) in enumerate(sequence(1)):
    ref2.run_stage(stage, it)
    for n in STAGE_OUTPUTS[stage] + (("candidate",) if stage == "K2_GEN_EDGE_INFORM" else ()):
        gold[f"{k:02d}_{stage}__{n}"] = ref2.get(n)
np.savez_compressed(os.path.join(out_dir, "weak_96x72.npz"), **gold)
print("weak golden: weak pixels", int((weak1 == 0).sum()))
for f in sorted(os.listdir(out_dir)):
    print(f, os.path.getsize(os.path.join(out_dir, f)))
