import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT
impl = sys.argv[1]
W, H, S = 320, 240, 2
sc = synth.make_scene(W, H, S)
p = default_params(); p.max_iterations = 1; p.num_images = S + 1
p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
p.use_APD = 0; p.state = FIRST_INIT; p.weak_peak_radius = 6
e = Engine(W, H, S, p, impl=impl)
e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
if impl == "reference":
    for st in ["K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM"]:
        print("stage", st, flush=True); e.run_stage(st, 0)
else:
    e.run(); print("times", e.last_run_times())
    pl = e.get("planes"); gt = sc.depths[0]
    err = np.abs(pl[..., 3] - gt) / gt
    print("depth rel err median", np.median(err), "frac<1%", (err < 0.01).mean(), "weak hist", np.bincount(e.get("weak").ravel()))
print("ok")
