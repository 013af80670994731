"""Row N3 measurement: dvp_fusion_run on one B200 against the CPU restatement of RunFusion (one host thread, as the
reference's loop is sequential) on the same synthetic views.  Prints one JSON line per configuration.
  python tools/bench_fusion.py [--full-w 3110 --full-h 2074 --views 5 --reps 3] [--mixed]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))

from dvp_mvs_b200 import Fusion, synth  # noqa: E402
from fusion_oracle import FusionOracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full-w", type=int, default=3110); ap.add_argument("--full-h", type=int, default=2074)
    ap.add_argument("--views", type=int, default=5); ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--mixed", action="store_true", help="odd views one pyramid level coarser (many shared source cells)")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    mv = synth.make_multiview(a.full_w, a.full_h, a.views, 2, seed=2)
    levels = [1 - (v & 1) for v in range(a.views)] if a.mixed else 1
    views = synth.make_fusion_views(mv, levels)
    npx = sum(v["depth"].size for v in views)
    t0 = time.perf_counter()
    f = Fusion(views)
    upload_s = time.perf_counter() - t0
    f.run()                                                   # warm-up (allocations, first launches)
    dev, wall, rounds = [], [], []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        pts, ms = f.run()
        wall.append(time.perf_counter() - t0); dev.append(ms)
    f.reset()
    for v in range(a.views):
        f.run_view(v)
        rounds.append(f.last_view(v)[3])
    out = dict(row="N3 fusion", views=a.views, view_size=[int(views[0]["depth"].shape[1]), int(views[0]["depth"].shape[0])],
               mixed_sizes=bool(a.mixed), num_src=len(views[0]["src_views"]), pixels=int(npx), points=int(len(pts)),
               gpu_device_ms=float(np.median(dev)), gpu_wall_ms=1e3 * float(np.median(wall)), upload_ms=1e3 * upload_s,
               gpu_mpix_per_s=npx / 1e3 / float(np.median(dev)), reservation_rounds=rounds)
    if not a.no_cpu:
        o = FusionOracle(views)
        t0 = time.perf_counter()
        ref = o.run()
        cpu_s = time.perf_counter() - t0
        same = len(ref) == len(pts) and bool((ref[:, :3] == pts[:, :3]).all())
        out.update(cpu_ms=1e3 * cpu_s, cpu_threads=1, cpu_points=int(len(ref)), speedup_wall=cpu_s / float(np.median(wall)),
                   points_identical=same)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
