#!/usr/bin/env python
"""bench.py — PatchMatch Mpix/s per view (device-timed), the metric of BASELINE.json.

One "step" = one RunPatchMatch pass (reference APD.cu:4406-4532, kernels K1..K16) over one synthetic reference view.
Default workload = BASELINE config C3, the one the north star quotes: ETH3D full resolution 6221x4146, 4 source views,
3 iterations, REFINE_ITER with geometric consistency, depth-edge / label priors on, the textureless wall WEAK (16 % of
the pixels take the adaptive-patch-deformation path).  `--workload c2` is the top level of C2's pyramid (3111x2073), the
finest level the reference's main() ever runs; it is also measured after the timed region as `secondary`.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA engine through the C ABI
  python bench.py --impl reference ...                            # the reference's own APD.cu (oracle/_ref), same GPU(s)

value      : pixels / device time with every input already resident in HBM (state restore = device-to-device).
e2e        : same metric through the public host API: pinned host buffers -> dvp_upload -> dvp_run -> dvp_download.
roofline   : K7/K8 propagation sweep, algorithmic bytes ((218 + 4 S) B/pixel per red+black iteration)
             / measured average launch time, against the measured HBM peak.  The path is TEX-bound, not HBM-bound
             (DESIGN.md), so the fraction is small by construction; `tex` reports the binding unit with fetches COUNTED
             by the instrumented build of the same sources (libdvp_mvs_count.so) in an extra, untimed pass.
cpu_baseline: the CPU restatement (oracle/cpu) on a bounded sample of the same workload, all host cores.
derived_states: a second workload whose pixel states, planes, selected views and radii are the OUTPUT of a real
             previous pass of the same engine (DepthToWeak's labelling) instead of the hand-painted wall.
Under torchrun (N > 1) every rank runs its own view on its own GPU (NCCL-free sharding, weak scaling) in BOTH arms —
the reference picks its device by argv (main.cpp:430-434) and farms as trivially; rank 0 prints ONE JSON line.
`--workload farm` (BASELINE config C4) strong-scales a whole multi-view, multi-scale schedule with the depth-map
exchange inside the timed region (bench_farm.py, dvp_mvs_b200/farm.py).
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

WORKLOADS = {"c3": (6221, 4146), "c2": (3111, 2073)}
STAGE_NAMES = ("K1", "K2", "K3", "K4", "K5", "K6", "K7", "K8", "K9", "K10", "K11", "K12", "K13", "K14", "K15", "K16")
TEX_ROOF_GFETCH = 1155.0   # measured: coherent bilinear fp32 fetches, profiles/r01_tex_coherence_ubench.txt (4 / clk / SM)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "farm"])
    ap.add_argument("--farm-driver", default="torchrun", choices=["torchrun", "library"],
                    help="--workload farm: one process per GPU + NCCL exchange (default), or ONE process whose library farm (dvp_farm_*) runs one host thread per GPU with peer copies")
    ap.add_argument("--width", type=int, default=0, help="override the workload's width")
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--src", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--state", default="refine_iter", choices=["first_init", "refine_init", "refine_iter"])
    ap.add_argument("--geom", type=int, default=1)
    ap.add_argument("--cpu-sample", default="640x480", help="WxH of the bounded CPU-baseline sample (0 = skip)")
    ap.add_argument("--next-rows", type=int, default=1, help="also time the callers' rows N3 (fusion) and N4 (edge prior) after the timed region")
    ap.add_argument("--extras", type=int, default=1, help="secondary workload, derived-state workload and fetch count after the timed region (rank 0, N = 1)")
    a = ap.parse_args()
    if a.workload != "farm":
        w, h = WORKLOADS[a.workload]
        a.width = a.width or w; a.height = a.height or h
    return a


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
    sc = None
    if os.path.exists(cache):
        try:
            z = np.load(cache)
            sc = synth.Scene(W, H, S, z["images"], z["depths"], z["cameras"].view(synth.CAMERA_DTYPE), z["planes_init"], z["planes_true"],
                             z["plane_id"], z["edge"], z["label"], float(z["depth_min"]), float(z["depth_max"]))
        except Exception:
            sc = None
    if sc is None:
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


def workload_config(args, name, inputs):
    """The `config` object — byte-identical in both arms (the driver compares it)."""
    weak = inputs.get("weak_info")
    return {"workload": name, "size": [args.width, args.height], "src_views": args.src, "iterations": args.iters, "state": args.state,
            "geom_consistency": int(bool(args.geom)), "priors": "edge + label maps, radius map",
            "weak_fraction": 0.0 if weak is None else round(float((weak == 0).mean()), 4), "views_per_gpu": 1,
            "l2": "inputs larger than L2 (per-view state >= 1.9 GB vs 126 MB L2)",
            "timing": "CUDA events on the engine's stream; max over ranks"}


def derive_states(engine_factory, args, sc, p, inputs):
    """Inputs of a rounds >= 1 pass whose pixel states / planes / selected views / radii are what a REAL previous pass left:
    the arm's own engine runs one REFINE_INIT pass (photometric, every pixel STRONG as after round 0) from the noisy planes
    and its DepthToWeak labelling becomes the WEAK map.  Each arm derives with its own kernels (the reference arm never
    touches ours and vice versa); the two maps differ only by the sweep's race."""
    from dvp_mvs_b200 import REFINE_INIT
    q = p.copy(); q.state = REFINE_INIT; q.geom_consistency = 0; q.use_APD = 0; q.max_iterations = 1; q.weak_peak_radius = 6
    e = engine_factory(q)
    kw = dict(inputs); kw["depths"] = None; kw["weak_info"] = None
    e.upload(**kw)
    e.run(**({"mode": 0} if e.prefix == "ref_" else {}))
    planes, weak, sel, rad = e.download()
    e.close()
    out = dict(inputs)
    out.update(planes=planes, weak_info=weak, selected_views=sel, radius=rad)
    return out, float((weak == 0).mean())


def cpu_baseline(args, cores):
    import cpu_oracle
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


def ref_stage_file(name):
    return f"/tmp/dvp_bench_reference_stages_{name}.json"


def excl_k2_k3(per_stage):
    return float(sum(per_stage) - per_stage[1] - per_stage[2])


def count_fetches(W, H, S, p, inputs, local):
    """Texture fetches per stage, COUNTED by the instrumented build of the same sources in one untimed pass."""
    from dvp_mvs_b200 import Engine
    from dvp_mvs_b200.parity import sequence
    lib = os.path.join(ROOT, "dvp_mvs_b200", "libdvp_mvs_count.so")
    if not os.path.exists(lib):
        return None
    e = Engine(W, H, S, p, device=local, lib_path=lib)
    e.upload(**inputs)
    e.fetch_count()                       # arms the counter
    per = [0] * 16
    from dvp_mvs_b200._lib import STAGE
    for st, it in sequence(p.max_iterations):
        if st == "K16_LOCAL_REFINE":
            continue                      # dvp_run issues K15 and K16 as one launch (tallied under K15, like the stage times)
        e.run_stage("K15_K16_FUSED" if st == "K15_DEPTH_TO_WEAK" else st, it)
        per[STAGE[st]] += e.fetch_count(reset=True)
    e.close()
    return per


def timed_passes(eng, inputs, steps, warmup, ref_mode):
    """`steps` passes from host inputs with per-stage times (used for the untimed extras and by the reference arm)."""
    import torch
    times, stages, launches = [], None, 0
    for i in range(warmup + steps):
        eng.upload(**inputs)
        torch.cuda.synchronize()
        eng.run(**({"mode": 0} if ref_mode else {}))
        total, per_stage, launches = eng.last_run_times()
        if i >= warmup:
            times.append(total); stages = per_stage if stages is None else [a + b for a, b in zip(stages, per_stage)]
    return float(np.mean(times)), [v / max(steps, 1) for v in stages], launches


def main():
    args = parse()
    if args.workload == "farm":
        import bench_farm
        return bench_farm.main(args)
    import torch
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    from dvp_mvs_b200 import Engine, Inputs
    # Every rank processes a view of the same synthetic scene (weak scaling: one view per GPU).  Rank 0 synthesises it
    # (threaded numpy) and leaves it in the /tmp cache; the other ranks load it after the barrier instead of
    # all ranks rendering at once on the same host cores.
    if dist is not None and rank != 0:
        dist.barrier()
    sc, p, inputs, workload = make_workload(args, seed=0)
    if dist is not None and rank == 0:
        dist.barrier()
    W, H, S = args.width, args.height, args.src
    N = W * H
    config = workload_config(args, workload, inputs)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    cores = os.cpu_count() or 1

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.impl == "reference":
        import ref_oracle
        # K2 GenEdgeInform from the build whose SASS the DRIVER's JIT generates from the same PTX: correct on every ray and at
        # full speed, where ptxas 12.9's own -O3 / -O2 code faults and its -O1 code is wrong (profiles/r02_reference_k2_miscompile.md)
        os.environ.setdefault("DVP_REF_K2_LIB", "libapd_ref_k2_jit.so")
        if not ref_oracle.available():
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libapd_ref.so was not built (needs /root/reference at build time)"}))
            return
        eng = ref_oracle.engine(W, H, S, p, device=local)
        sampler = ClockSampler(local)
        times, stages, launches = [], None, 0
        for i in range(args.warmup + args.steps):
            eng.upload(**inputs)
            if i == args.warmup:
                barrier(); sampler.start()
            eng.run(mode=0)   # the reference's kernel sequence, cudaDeviceSynchronize after every launch as RunPatchMatch does
            total, per_stage, launches = eng.last_run_times()
            if i >= args.warmup:
                times.append(total); stages = per_stage if stages is None else [a + b for a, b in zip(stages, per_stage)]
        barrier(); clocks = sampler.stop()
        ms = float(np.mean(times))
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        if rank != 0:
            dist.destroy_process_group()
            return
        per_stage = [v / args.steps for v in stages]
        val = world * N / ms / 1e3
        try:
            json.dump({"workload": workload, "per_stage_ms": per_stage, "ms_per_step": ms}, open(ref_stage_file(workload), "w"))
        except Exception:
            pass
        out = {"impl": "reference", "metric": "patchmatch_mpix_per_s_per_view", "value": val, "unit": "Mpix/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
               "notes": "the reference's own APD.cu kernels (unmodified, sm_100a, its build flags) in its own launch order with a "
                        "cudaDeviceSynchronize after each, one view per GPU; time = sum of CUDA-event times of its 11 + 5*iters launches; "
                        "K2 GenEdgeInform runs from the same source and flags compiled to PTX and finished by the driver's JIT (ptxas 12.9's own "
                        "-O3 SASS of that kernel faults on sm_100a, profiles/r02_reference_k2_miscompile.md); e2e repeats value: uploads and the 25 B/pixel the real "
                        "RunPatchMatch copies back (APD.cu:4525-4530) are left OUT of this arm's time, which favours the reference",
               "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": 0, "kind": "reference",
                                "sample": "the reference has no CPU path: this arm is its own CUDA kernels (APD.cu unmodified, sm_100a) on the same B200(s)"},
               "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "per_stage_ms": [round(v, 3) for v in per_stage], "ms_excl_k2_k3": round(excl_k2_k3(per_stage), 3),
               "gpu_launches": launches * args.steps, "clocks": clocks}
        if args.extras and world == 1:
            try:   # the derived-state workload, with this arm's own kernels
                a2 = argparse.Namespace(**vars(args)); a2.width, a2.height = WORKLOADS["c2"]
                sc2, p2, in2, name2 = make_workload(a2, seed=0)
                eng.close()
                d_in, wfrac = derive_states(lambda q: ref_oracle.engine(a2.width, a2.height, S, q, device=local), a2, sc2, p2, in2)
                e2 = ref_oracle.engine(a2.width, a2.height, S, p2, device=local)
                ms2, st2, _ = timed_passes(e2, d_in, 2, 1, True)
                e2.close()
                out["derived_states"] = {"workload": name2.rsplit("_weak", 1)[0] + "_states_from_previous_pass", "weak_fraction": round(wfrac, 4),
                                         "value": a2.width * a2.height / ms2 / 1e3, "ms_per_step": ms2, "per_stage_ms": [round(v, 3) for v in st2],
                                         "ms_excl_k2_k3": round(excl_k2_k3(st2), 3)}
                json.dump(out["derived_states"], open(ref_stage_file("derived"), "w"))
            except Exception as e:  # noqa: BLE001
                out["derived_states"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        print(json.dumps(out))
        if dist is not None:
            dist.destroy_process_group()
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
    sweep_ms = (per_stage[6] + per_stage[7]) / max(2 * args.iters, 1)     # average duration of one sweep launch
    bytes_per_launch = N * (218 + 4 * S) / 2.0                             # half the pixels (one colour) per launch
    achieved = bytes_per_launch / (sweep_ms * 1e-3) / 1e9
    weak_px = 0 if inputs.get("weak_info") is None else int((inputs["weak_info"] == 0).sum())   # WEAK pixels leave the sweep at once (K10/K11 own them)
    traffic = None   # measured DRAM bytes per sweep launch, from the committed ncu capture of this very workload
    try:   # profiles/r02_traffic.json: ncu --set full of the two kernels of a sweep launch (tools/ncu_traffic.py), per workload
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(workload, {}).get("k_sweep_pair")
        if tr:
            traffic = tr["dram_read_bytes"] + tr["dram_write_bytes"]
    except Exception:
        pass
    out = {"metric": "patchmatch_mpix_per_s_per_view", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": config,
           "notes": "value: CUDA events on the engine stream around K steps, inputs resident in HBM; per-stage = events around each launch; "
                    "e2e: pinned host buffers -> dvp_upload_overlapped -> dvp_run -> dvp_download inside the timed region",
           "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": e2e_wall_ms / args.steps},
           "gpu_launches": int(launches * args.steps),
           "roofline": {"kernel": "k_sweep_score + k_sweep_update (one K7 / K8 launch = this pair)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes": bytes_per_launch, "peak_source": peak_src, "avg_launch_ms": sweep_ms,
                        "share_of_step": (per_stage[6] + per_stage[7]) / max(total_ms, 1e-9),
                        "note": "TEX-bound kernels: the binding roof is the texture unit (measured 1155 Gfetch/s for coherent bilinear fp32 fetches, "
                                "profiles/r01_tex_coherence_ubench.txt; software sampling measured slower, profiles/r02_sampling_path_ubench.txt), not HBM; "
                                "see `tex`.  traffic = measured DRAM bytes of the pair (ncu), which include the scratch area the two kernels hand over"},
           "per_stage_ms": [round(v, 3) for v in per_stage], "ms_excl_k2_k3": round(excl_k2_k3(per_stage), 3),
           "stage_share": {k: round(v / max(sum(per_stage), 1e-9), 4) for k, v in zip(
               ("K1", "K2", "K3", "K4", "K5", "K6", "K7", "K8", "K9", "K10", "K11", "K12", "K13", "K14", "K15+K16", "K16"), per_stage) if v > 0},
           # WEAK-pixel traffic is data dependent and reported apart from the per-pixel figure:
           # neighbours 48 + label_boundary 32 + complex 4 + fit plane 16 bytes per WEAK pixel
           "weak": {"pixels": weak_px, "fraction": weak_px / N, "algorithmic_bytes_per_pass": weak_px * (48 + 32 + 4 + 16)},
           "sweep_mpix_per_s_per_iteration": N / ((per_stage[6] + per_stage[7]) / max(args.iters, 1)) / 1e3,   # K7 + K8 alone
           "clocks": clocks}
    try:   # the reference arm ran first on this box (driver order): per-stage ratios, and the pass without K2 / K3
        rs = json.load(open(ref_stage_file(workload)))
        rp = rs["per_stage_ms"]
        groups = {"K1": (0,), "K2": (1,), "K3": (2,), "K4": (3,), "K6": (5,), "K7+K8": (6, 7), "K9": (8,), "K10+K11": (9, 10), "K15+K16": (14, 15)}
        out["vs_reference_stages"] = {"reference_ms_per_step": round(rs["ms_per_step"], 3),
                                      "ratio_excl_k2_k3": round(excl_k2_k3(rp) / max(excl_k2_k3(per_stage), 1e-9), 3),
                                      "per_stage_ratio": {k: round(sum(rp[i] for i in ix) / max(sum(per_stage[i] for i in ix), 1e-9), 2)
                                                          for k, ix in groups.items() if sum(per_stage[i] for i in ix) > 0},
                                      "note": "reference stage times of the --impl reference run that preceded this one on the same box; "
                                              "ratio_excl_k2_k3 leaves out the two stages whose reference time depends least on the kernels themselves: "
                                              "K2 runs from a JIT-finished build (ptxas 12.9 miscompiles it) and K3's ring search is near its worst case "
                                              "inside the solid WEAK wall of this synthetic scene"}
    except Exception:
        pass
    if args.extras and world == 1:
        del dev
        torch.cuda.empty_cache()
        try:
            per_fetch = count_fetches(W, H, S, p, inputs, local)
            if per_fetch:
                sweep_fetch = (per_fetch[6] + per_fetch[7]) / max(2 * args.iters, 1)
                tex = {"counted_by": "libdvp_mvs_count.so (same sources, every fetch site tallies itself), one untimed pass",
                       "fetches_per_pass": int(sum(per_fetch)), "fetches_per_stage": {STAGE_NAMES[i]: int(v) for i, v in enumerate(per_fetch) if v},
                       "sweep_fetches_per_launch": int(sweep_fetch), "sweep_gfetch_per_s": sweep_fetch / (sweep_ms * 1e-3) / 1e9,
                       "peak_gfetch_per_s": TEX_ROOF_GFETCH, "sweep_frac_of_tex_roof": sweep_fetch / (sweep_ms * 1e-3) / 1e9 / TEX_ROOF_GFETCH,
                       "gfetch_per_s_per_stage": {STAGE_NAMES[i]: round(per_fetch[i] / (per_stage[i] * 1e-3) / 1e9, 1)
                                                  for i in range(16) if per_fetch[i] and per_stage[i] > 0}}
                out["roofline"]["tex"] = tex
        except Exception as e:  # noqa: BLE001
            out["roofline"]["tex"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        eng.close()
        try:   # C2's top level, the finest level main() runs — and the same size with states from a real previous pass
            a2 = argparse.Namespace(**vars(args)); a2.width, a2.height = WORKLOADS["c2"]
            sc2, p2, in2, name2 = make_workload(a2, seed=0)
            e2 = Engine(a2.width, a2.height, S, p2, device=local)
            if (a2.width, a2.height) != (W, H):
                ms2, st2, _ = timed_passes(e2, in2, 3, 1, False)
                out["secondary"] = {"workload": name2, "value": a2.width * a2.height / ms2 / 1e3, "ms_per_step": ms2,
                                    "per_stage_ms": [round(v, 3) for v in st2], "ms_excl_k2_k3": round(excl_k2_k3(st2), 3)}
            d_in2, wfrac = derive_states(lambda q: Engine(a2.width, a2.height, S, q, device=local), a2, sc2, p2, in2)
            ms3, st3, _ = timed_passes(e2, d_in2, 3, 1, False)
            e2.close()
            out["derived_states"] = {"workload": name2.rsplit("_weak", 1)[0] + "_states_from_previous_pass", "weak_fraction": round(wfrac, 4),
                                     "value": a2.width * a2.height / ms3 / 1e3, "ms_per_step": ms3, "per_stage_ms": [round(v, 3) for v in st3],
                                     "ms_excl_k2_k3": round(excl_k2_k3(st3), 3)}
            try:
                rd = json.load(open(ref_stage_file("derived")))
                out["derived_states"]["vs_reference"] = {"ratio": round(rd["ms_per_step"] / ms3, 3),
                                                         "ratio_excl_k2_k3": round(rd["ms_excl_k2_k3"] / max(excl_k2_k3(st3), 1e-9), 3)}
            except Exception:
                pass
        except Exception as e:  # noqa: BLE001
            out["derived_states"] = {"error": f"{type(e).__name__}: {e}"[:300]}
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
