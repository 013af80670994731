#!/usr/bin/env python
"""GPU diagnostic: product vs reference oracle, stage by stage, with the reference's own run-to-run noise floor.
Writes gpurun_out/parity_report.json.  Usage: python tools/parity_report.py [W H S iters]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_oracle
from dvp_mvs_b200.parity import step_compare, compare, sequence, STATE_BUFS


def main():
    W, H, S, iters = (int(v) for v in (sys.argv[1:5] + [640, 480, 2, 1][len(sys.argv) - 1:]))
    sc = synth.make_scene(W, H, S)
    p = default_params(); p.max_iterations = iters; p.num_images = S + 1
    p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
    p.use_APD = 0; p.state = FIRST_INIT; p.geom_consistency = 0; p.weak_peak_radius = 6
    out = dict(config=dict(W=W, H=H, S=S, iters=iters))
    kw = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    ref = ref_oracle.engine(W, H, S, p); prod = Engine(W, H, S, p)
    print(ref.version(), "|", prod.version())
    ref.upload(**kw); prod.upload(**kw)
    t = time.time()
    res = step_compare(ref, prod, iters, log=print)
    out["stepwise"] = res
    print("stepwise done in %.1fs" % (time.time() - t))
    # noise floor: two reference runs from the same state, full pass
    finals = []
    for k in range(2):
        ref.upload(**kw); ref.run(mode=0)
        finals.append({n: ref.get(n) for n in ("planes", "costs", "selected", "weak")})
        print("ref run %d: total %.2f ms" % (k, ref.last_run_times()[0]))
    out["ref_times"] = ref.last_run_times()
    out["noise_floor"] = [compare(n, finals[0][n], finals[1][n]) for n in finals[0]]
    prod.upload(**kw); prod.run()
    out["prod_times"] = prod.last_run_times()
    pf = {n: prod.get(n) for n in ("planes", "costs", "selected", "weak")}
    out["end_to_end_vs_ref"] = [compare(n, finals[0][n], pf[n]) for n in pf]
    for k in ("noise_floor", "end_to_end_vs_ref"):
        for r in out[k]:
            print(k, r)
    print("ref ms", out["ref_times"][0], "prod ms", out["prod_times"][0])
    print("ref per stage", [round(v, 2) for v in out["ref_times"][1]])
    print("prod per stage", [round(v, 2) for v in out["prod_times"][1]])
    # accuracy vs ground truth, for sanity
    d_ref = finals[0]["planes"][..., 3]; d_prod = pf["planes"][..., 3]; gt = sc.depths[0]
    for nm, d in (("ref", d_ref), ("prod", d_prod)):
        err = np.abs(d - gt) / gt
        print(nm, "depth rel err median %.4f, frac<1%% %.3f" % (np.median(err), (err < 0.01).mean()))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/parity_report.json", "w"), indent=1, default=str)


if __name__ == "__main__":
    main()
