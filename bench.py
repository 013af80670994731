#!/usr/bin/env python
"""bench.py — PatchMatch Mpix/s per view (device-timed), the metric of BASELINE.json.

One "step" = one RunPatchMatch pass (reference APD.cu:4406-4532, kernels K1..K16) over one synthetic
reference view.  Default workload (BASELINE config C2's top pyramid level): 3111x2073 (ETH3D 6221x4146 at
scale 2), 4 source views, 3 iterations, REFINE_ITER with geometric consistency, edge/label priors on.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA engine through the C ABI
  python bench.py --impl reference ...                            # the reference's own APD.cu (oracle/_ref), same GPU

value      : pixels / device time with every input already resident in HBM (state restore = device-to-device).
e2e        : same metric through the public host API: pinned host buffers -> dvp_upload -> dvp_run -> dvp_download.
roofline   : K7/K8 propagation sweep, algorithmic bytes ((218 + 4 S) B/pixel per red+black iteration)
             / measured average launch time, against the measured HBM peak.  The path is TEX/FP32-bound, not
             HBM-bound, so the fraction is small by construction; `tex` reports the binding unit.
cpu_baseline: the CPU restatement (oracle/cpu) on a bounded sample of the same workload, all host cores.
Under torchrun (N > 1) every rank runs its own view on its own GPU (NCCL-free sharding, weak scaling);
rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=3111)
    ap.add_argument("--height", type=int, default=2073)
    ap.add_argument("--src", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--state", default="refine_iter", choices=["first_init", "refine_init", "refine_iter"])
    ap.add_argument("--geom", type=int, default=1)
    ap.add_argument("--cpu-sample", default="640x480", help="WxH of the bounded CPU-baseline sample (0 = skip)")
    ap.add_argument("--next-rows", type=int, default=1, help="also time the callers' rows N3 (fusion) and N4 (edge prior) after the timed region")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_workload(args, seed):
    from dvp_mvs_b200 import synth, default_params, FIRST_INIT, REFINE_INIT, REFINE_ITER
    W, H, S = args.width, args.height, args.src
    cache = f"/tmp/dvp_bench_scene_{W}x{H}_S{S}_seed{seed}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        sc = synth.Scene(W, H, S, z["images"], z["depths"], z["cameras"].view(synth.CAMERA_DTYPE), z["planes_init"], z["planes_true"],
                         z["plane_id"], z["edge"], z["label"], float(z["depth_min"]), float(z["depth_max"]))
    else:
        sc = synth.make_scene(W, H, S, seed=seed)
        try:
            tmp = cache + f".{os.getpid()}.tmp.npz"
            np.savez(tmp, images=sc.images, depths=sc.depths, cameras=sc.cameras.view(np.uint8), planes_init=sc.planes_init,
                     planes_true=sc.planes_true, plane_id=sc.plane_id, edge=sc.edge, label=sc.label,
                     depth_min=sc.depth_min, depth_max=sc.depth_max)
            os.replace(tmp, cache)
        except Exception:
            pass
    p = default_params()
    p.max_iterations = args.iters; p.num_images = S + 1
    p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
    p.state = {"first_init": FIRST_INIT, "refine_init": REFINE_INIT, "refine_iter": REFINE_ITER}[args.state]
    p.geom_consistency = int(bool(args.geom))
    p.weak_peak_radius = 6 if not args.geom else 4
    weak = None
    if args.state == "first_init":
        p.use_APD = 0        # round 0 of the reference's schedule: every pixel STRONG (main.cpp:458-477)
        planes = sc.planes_init; selected = None
    else:
        # rounds >= 1 (main.cpp:463-477): use_APD, use_detail, rotate_time 2, ransac threshold 0.00875.  Pixel
        # states as DepthToWeak leaves them on this scene: the textureless wall WEAK, a 6 px border UNKNOWN.
        p.use_APD = 1; p.use_detail = 1; p.rotate_time = 2; p.ransac_threshold = 0.00875
        weak = np.full((H, W), 1, np.uint8)
        weak[sc.plane_id == 3] = 0
        weak[:6, :] = 2; weak[-6:, :] = 2; weak[:, :6] = 2; weak[:, -6:] = 2
        rng = np.random.default_rng(20250104 + seed)   # "previous pass" output: truth with 2 % depth noise
        planes = sc.planes_true.copy()
        planes[..., 3] *= (1.0 + rng.normal(0.0, 0.02, planes.shape[:2])).astype(np.float32)
        selected = np.full((H, W), (1 << S) - 1, np.uint32)
    inputs = dict(images=sc.images, depths=sc.depths if args.geom else None, cameras=sc.cameras, planes=np.ascontiguousarray(planes),
                  selected_views=selected, weak_info=weak, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    wfrac = 0.0 if weak is None else float((weak == 0).mean())
    name = f"eth3d_shaped_{W}x{H}_S{S}_it{args.iters}_{args.state}_geom{int(bool(args.geom))}_weak{int(round(100 * wfrac))}pct"
    return sc, p, inputs, name


def cpu_baseline(args, cores):
    import cpu_oracle
    from dvp_mvs_b200 import synth
    if args.cpu_sample in ("0", "", "none") or not cpu_oracle.available():
        return None
    w, h = (int(v) for v in args.cpu_sample.split("x"))
    a2 = argparse.Namespace(**vars(args)); a2.width, a2.height = w, h
    sc, p, inputs, _ = make_workload(a2, seed=0)
    e = cpu_oracle.engine(w, h, args.src, p)
    e.upload(**inputs)
    t0 = time.perf_counter(); e.run(); dt = time.perf_counter() - t0
    return {"value": w * h / dt / 1e6, "unit": "Mpix/s", "cores": cores, "kind": "port",
            "sample": f"one full RunPatchMatch pass of the same configuration (STRONG and WEAK paths) on a {w}x{h} view "
                      f"({dt:.1f} s, OpenMP over pixels)"}


def next_rows(local):
    """Device times of the two callers' rows that are not part of a RunPatchMatch pass, measured after the
    timed region on rank 0: N3 fusion of 5 synthetic views of 1555x1037 (4 sources each) and N4 edge prior of one
    1555x1037 level image.  Never allowed to break the bench line: any failure is reported as text."""
    try:
        import numpy as np
        from dvp_mvs_b200 import Fusion, edge_segment, synth
        mv = synth.make_multiview(3110, 2074, 5, 2, seed=2)
        views = synth.make_fusion_views(mv, 1)
        f = Fusion(views, device=local)
        f.run()
        pts, ms = f.run()
        f.close()
        img = np.clip(np.rint(mv.levels[1][0]["image"]), 0, 255).astype(np.uint8)
        edge_segment(img, device=local)
        edge, thr, edge_ms = edge_segment(img, device=local)
        npx = sum(v["depth"].size for v in views)
        return {"N3_fusion": {"views": 5, "view_size": [1555, 1037], "sources": 4, "points": int(len(pts)), "device_ms": ms,
                              "mpix_per_s": npx / 1e3 / max(ms, 1e-9)},
                "N4_edge_prior": {"size": [1555, 1037], "device_ms": edge_ms, "mpix_per_s": img.size / 1e3 / max(edge_ms, 1e-9),
                                  "edge_fraction": float((edge > 0).mean())}}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def main():
    args = parse()
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return  # the reference arm is a single-GPU baseline: rank 0 alone runs and prints it
    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    from dvp_mvs_b200 import Engine, Inputs
    from dvp_mvs_b200 import _lib
    # Every rank processes a view of the same synthetic scene (weak scaling: one view per GPU).  Rank 0 synthesises it
    # (threaded numpy, ~10 s) and leaves it in the /tmp cache; the other ranks load it after the barrier instead of
    # all ranks rendering at once on the same host cores.
    if dist is not None and rank != 0:
        dist.barrier()
    sc, p, inputs, workload = make_workload(args, seed=0)
    if dist is not None and rank == 0:
        dist.barrier()
    W, H, S = args.width, args.height, args.src
    N = W * H
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        import ref_oracle
        os.environ.setdefault("DVP_REF_K2_LIB", "libapd_ref_k2_O1.so")  # K2 at a realistic optimisation level (oracle/ref_k2_safe.cu)
        if not ref_oracle.available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libapd_ref.so was not built (needs /root/reference at build time)"}))
            return
        eng = ref_oracle.engine(W, H, S, p, device=local)
        sampler = ClockSampler(local)
        times = []
        for i in range(args.warmup + args.steps):
            eng.upload(**inputs)
            if i == args.warmup:
                torch.cuda.synchronize(); sampler.start()
            eng.run(mode=0)   # the reference's kernel sequence, cudaDeviceSynchronize after every launch as RunPatchMatch does
            total, per_stage, launches = eng.last_run_times()
            if i >= args.warmup:
                times.append(total)
        torch.cuda.synchronize(); clocks = sampler.stop()
        ms = float(np.mean(times)); val = N / ms / 1e3
        out = {"impl": "reference", "metric": "patchmatch_mpix_per_s_per_view", "value": val, "unit": "Mpix/s", "n_gpus": 1,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload, "timing": "sum of CUDA-event times of the reference's own 11+5*iters kernel launches; "
                          "K2 from the -Xptxas -O1 build (the -O3 build faults on sm_100a)", "l2": "inputs larger than L2"},
               "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": 0, "kind": "reference",
                                "sample": "the reference has no CPU path: this arm is its own CUDA kernels (APD.cu unmodified, sm_100a) on one B200"},
               "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "per_stage_ms": [round(v, 3) for v in per_stage], "gpu_launches": launches * args.steps, "clocks": clocks}
        print(json.dumps(out))
        return

    # ---------------------------------------------------------------------------------------------- ours
    eng = Engine(W, H, S, p, device=local)
    stream = torch.cuda.ExternalStream(eng._f("stream")(eng.ctx), device=torch.device("cuda", local))
    # device-resident copies of every input (what an in-memory multi-pass driver would hold)
    dev = {k: (torch.from_numpy(np.ascontiguousarray(v).view(np.uint8) if k == "cameras" else np.ascontiguousarray(v)).cuda()
               if isinstance(v, np.ndarray) else None) for k, v in inputs.items() if k != "seed"}
    def ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())
    d_in = Inputs(ptr(dev["images"]), ptr(dev.get("depths")), ptr(dev["cameras"]), ptr(dev["planes"]), ptr(dev.get("selected_views")),
                  ptr(dev.get("weak_info")), ptr(dev["edge"]), ptr(dev["label"]), None, int(inputs["seed"]))
    # pinned host copies for the end-to-end leg
    pin = {k: (torch.from_numpy(np.ascontiguousarray(v).view(np.uint8) if k == "cameras" else np.ascontiguousarray(v)).pin_memory()
               if isinstance(v, np.ndarray) else None) for k, v in inputs.items() if k != "seed"}
    def hptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())
    h_in = Inputs(hptr(pin["images"]), hptr(pin.get("depths")), hptr(pin["cameras"]), hptr(pin["planes"]), hptr(pin.get("selected_views")),
                  hptr(pin.get("weak_info")), hptr(pin["edge"]), hptr(pin["label"]), None, int(inputs["seed"]))
    h2d = sum(t.numel() * t.element_size() for t in pin.values() if t is not None) + N * 4  # + radius map built by the host side
    out_planes = torch.empty((H, W, 4), dtype=torch.float32).pin_memory(); out_weak = torch.empty((H, W), dtype=torch.uint8).pin_memory()
    out_sel = torch.empty((H, W), dtype=torch.int32).pin_memory(); out_rad = torch.empty((H, W), dtype=torch.int32).pin_memory()
    d2h = N * (16 + 1 + 4 + 4)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        eng.upload_raw(d_in, device=True)   # state restore, device-to-device
        eng.run(sync=False)

    def step_e2e():
        eng.upload_raw(h_in, device=False, overlapped=True)   # pinned buffers: the large maps stream in behind K1..K5
        eng.run(sync=False)
        eng._check(eng._f("download")(eng.ctx, hptr(out_planes), hptr(out_weak), hptr(out_sel), hptr(out_rad)), "download")

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    total_ms, per_stage, launches = eng.last_run_times()
    # ---- e2e: host buffers in, host results out
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    f0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    f1.record(stream)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(f0.elapsed_time(f1), 0.0)
    if dist is not None:
        t = torch.tensor([ms_total, e2e_ms, e2e_wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, e2e_wall_ms = (float(v) for v in t.tolist())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = world * N / ms_step / 1e3
    e2e_value = world * N / (e2e_ms / args.steps) / 1e3
    # ---- roofline of the dominant kernel (K7/K8 sweep), from the last step's per-launch CUDA events
    sweep_ms = (per_stage[6] + per_stage[7]) / (2 * args.iters)        # average duration of one sweep launch
    bytes_per_launch = N * (218 + 4 * S) / 2.0                             # half the pixels (one colour) per launch
    achieved = bytes_per_launch / (sweep_ms * 1e-3) / 1e9
    ncc_per_px_iter = 22 * S                                               # 16 candidate + 1 current + 5 refinement (hyp. 4 folded) NCCs, all views
    weak_px = 0 if inputs.get("weak_info") is None else int((inputs["weak_info"] == 0).sum())   # WEAK pixels leave the sweep at once (K10/K11 own them)
    samples_per_launch = ((N - weak_px) / 2.0) * ncc_per_px_iter * 36
    traffic = None   # measured DRAM bytes per sweep launch, from the committed ncu capture of this very workload
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(workload, {}).get("k_strong_sweep")
        if tr:
            traffic = tr["dram_read_bytes"] + tr["dram_write_bytes"]
    except Exception:
        pass
    out = {"metric": "patchmatch_mpix_per_s_per_view", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload, "views_per_gpu": 1, "l2": "inputs larger than L2 (per-view state 1.9 GB vs 126 MB L2)",
                      "timing": "CUDA events on the engine stream around K steps; per-stage = events around each launch"},
           "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall_ms / args.steps},
           "gpu_launches": int(launches * args.steps),
           "roofline": {"kernel": "k_strong_sweep (K7/K8)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes": bytes_per_launch, "peak_source": peak_src, "avg_launch_ms": sweep_ms,
                        "share_of_step": (per_stage[6] + per_stage[7]) / max(total_ms, 1e-9),
                        "note": "TEX-bound kernel: ~%d bilinear source fetches per byte of compulsory traffic; the binding roof is the texture unit "
                                "(measured 1155 Gfetch/s coherent, profiles/r01_tex_coherence_ubench.txt), not HBM" % int(samples_per_launch / bytes_per_launch),
                        "tex_gsamples_per_s": samples_per_launch / (sweep_ms * 1e-3) / 1e9, "tex_peak_gsamples_per_s": 1155.0,
                        "tex_frac": samples_per_launch / (sweep_ms * 1e-3) / 1e9 / 1155.0},
           "per_stage_ms": [round(v, 3) for v in per_stage],
           "stage_share": {k: round(v / max(sum(per_stage), 1e-9), 4) for k, v in zip(
               ("K1", "K2", "K3", "K4", "K5", "K6", "K7", "K8", "K9", "K10", "K11", "K12", "K13", "K14", "K15+K16", "K16"), per_stage) if v > 0},
           # WEAK-pixel traffic is data dependent and reported apart from the per-pixel figure:
           # neighbours 48 + label_boundary 32 + complex 4 + fit plane 16 bytes per WEAK pixel
           "weak": {"pixels": weak_px, "fraction": weak_px / N, "algorithmic_bytes_per_pass": weak_px * (48 + 32 + 4 + 16)},
           "sweep_mpix_per_s_per_iteration": N / ((per_stage[6] + per_stage[7]) / args.iters) / 1e3,   # K7 + K8 alone
           "clocks": clocks}
    cb = cpu_baseline(args, cores) if world == 1 else None   # rank 0 at N = 1 only
    if cb:
        out["cpu_baseline"] = cb
    if args.next_rows and world == 1:
        out["next_rows"] = next_rows(local)
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
