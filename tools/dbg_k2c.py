import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT
W, H = 640, 480
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
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
print(os.environ.get("DVP_REF_K2_LIB"), "S", S, "bad per dir", [int((a[:, :, d] != b[:, :, d]).any(-1).sum()) for d in range(8)])
