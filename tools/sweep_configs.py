#!/usr/bin/env python
"""BASELINE.json configs C3 and C5 on one B200 (device-timed, HBM-resident inputs, our engine):
  C3  ETH3D full-res 6221x4146, S=4, 3 iterations, geometric consistency + edge/label priors
  C5  3840x2160, S in {2,4,8,16} x iterations in {3,5,7}, photometric, all pixels STRONG (the reference's WEAK path
      is undefined for S > 4: candidate_cuda is dimensioned for 4 views)
Prints one JSON object per configuration.  usage: python tools/sweep_configs.py [--c3 1] [--c5 1]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def run(W, H, S, iters, state, geom, sc, steps=2, warmup=1, weak=False):
    from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT, REFINE_ITER
    p = default_params(); p.max_iterations = iters; p.num_images = S + 1
    p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
    p.geom_consistency = geom; p.weak_peak_radius = 4 if geom else 6
    kw = dict(images=sc.images[:S + 1], cameras=sc.cameras[:S + 1], edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    if state == "first_init":
        p.state = FIRST_INIT; p.use_APD = 0
        kw.update(planes=sc.planes_init)
    else:
        p.state = REFINE_ITER; p.use_APD = 1 if weak else 0
        if weak:
            p.use_detail = 1; p.rotate_time = 2; p.ransac_threshold = 0.00875
            wk = np.full((H, W), 1, np.uint8); wk[sc.plane_id == 3] = 0
            wk[:6, :] = 2; wk[-6:, :] = 2; wk[:, :6] = 2; wk[:, -6:] = 2
            kw.update(weak_info=wk)
        rng = np.random.default_rng(20250104)
        planes = sc.planes_true.copy(); planes[..., 3] *= (1.0 + rng.normal(0.0, 0.02, planes.shape[:2])).astype(np.float32)
        kw.update(planes=planes, selected_views=np.full((H, W), (1 << S) - 1, np.uint32))
        if geom:
            kw.update(depths=sc.depths[:S + 1])
    e = Engine(W, H, S, p)
    ms = []
    for i in range(warmup + steps):
        e.upload(**kw); e.run()
        total, per_stage, launches = e.last_run_times()
        if i >= warmup:
            ms.append(total)
    t = float(np.mean(ms)); N = W * H
    sweep_ms = (per_stage[6] + per_stage[7]) / (2 * iters)
    out = {"config": f"{W}x{H}_S{S}_it{iters}_{state}_geom{geom}" + ("_weak" if weak else ""), "ms_per_pass": round(t, 2), "mpix_per_s": round(N / t / 1e3, 2),
           "sweep_launch_ms": round(sweep_ms, 3), "sweep_hbm_gbs_algorithmic": round(N / 2 * (218 + 4 * S) / (sweep_ms * 1e-3) / 1e9, 1),
           "pass_hbm_gbs_algorithmic": round(N * (443 + 44 * S + 8 * geom * S + iters * (251 + 4 * S)) / (t * 1e-3) / 1e9, 1),
           "per_stage_ms": [round(v, 2) for v in per_stage]}
    e.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3", type=int, default=1); ap.add_argument("--c5", type=int, default=1)
    a = ap.parse_args()
    from dvp_mvs_b200 import synth
    if a.c3:
        t0 = time.time(); sc = synth.make_scene(6221, 4146, 4, seed=0)
        print(json.dumps({"synth_s": round(time.time() - t0, 1), "scene": "6221x4146 S=4"}), flush=True)
        print(json.dumps(run(6221, 4146, 4, 3, "refine_iter", 1, sc, weak=True)), flush=True)
        print(json.dumps(run(6221, 4146, 4, 3, "first_init", 0, sc)), flush=True)
        del sc
    if a.c5:
        t0 = time.time(); sc = synth.make_scene(3840, 2160, 16, seed=0)
        print(json.dumps({"synth_s": round(time.time() - t0, 1), "scene": "3840x2160 S=16"}), flush=True)
        for S in (2, 4, 8, 16):
            for iters in (3, 5, 7):
                print(json.dumps(run(3840, 2160, S, iters, "refine_iter", 0, sc, steps=1, warmup=1)), flush=True)


if __name__ == "__main__":
    main()
