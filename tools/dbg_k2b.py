import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT
W, H, S = 640, 480, 2
sc = synth.make_scene(W, H, S)
p = default_params(); p.max_iterations = 1; p.num_images = S + 1
p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
p.use_APD = 0; p.state = FIRST_INIT
kw = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
ref = Engine(W, H, S, p, lib_path=ref_oracle.REFERENCE_LIB, prefix="ref_"); prod = Engine(W, H, S, p, )
ref.upload(**kw); prod.upload(**kw)
for st in ["K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM"]:
    ref.run_stage(st); prod.run_stage(st)
a, b = ref.get("edge_neigh"), prod.get("edge_neigh")
for d in (2, 3):
    bad = (a[:, :, d] != b[:, :, d]).any(-1)
    ys, xs = np.nonzero(bad)
    dist = np.abs(b[ys, xs, d, 0] - xs)
    print("dir", d, "bad", bad.sum(), "dist to true edge: min", dist.min(), "max", dist.max(), "hist", np.bincount(np.minimum(dist, 40))[:41])
    good = ~bad & (b[:, :, d, 0] >= 0)
    ys, xs = np.nonzero(good)
    dist = np.abs(b[ys, xs, d, 0] - xs)
    print("   good-with-edge", good.sum(), "dist hist", np.bincount(np.minimum(dist, 40))[:41])
    print("   ref values in bad set unique:", np.unique(a[:, :, d][bad], axis=0)[:5])
    print("   row 63 ref x:", a[63, :40, d, 0], "prod x:", b[63, :40, d, 0])
