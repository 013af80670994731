import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ref_oracle
from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT
from dvp_mvs_b200.parity import sequence, STATE_BUFS
W, H, S = 640, 480, 2
sc = synth.make_scene(W, H, S)
p = default_params(); p.max_iterations = 1; p.num_images = S + 1
p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
p.use_APD = 0; p.state = FIRST_INIT; p.weak_peak_radius = 6
kw = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
ref = ref_oracle.engine(W, H, S, p); prod = Engine(W, H, S, p)
ref.upload(**kw); prod.upload(**kw)
for st in sequence(1)[:-1]:
    ref.run_stage(*st)
pre = {n: ref.get(n) for n in STATE_BUFS}
ref.run_stage("K16_LOCAL_REFINE"); a = ref.get("planes")
for n, v in pre.items(): prod.set(n, v)
prod.run_stage("K16_LOCAL_REFINE"); b = prod.get("planes")
bad = np.argwhere((a != b).any(-1) & ~(np.isnan(a) & np.isnan(b)).any(-1))
print("bad", len(bad))
for y, x in bad:
    print("px", (x, y), "pre", pre["planes"][y, x], "ref", a[y, x], "prod", b[y, x], "sel", pre["selected"][y, x], "vw", pre["view_weight"][y, x, :S],
          "cost", pre["costs"][y, x], "radius", pre["radius"][y, x], "weak", pre["weak"][y, x])
