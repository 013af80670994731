"""Row N4 (edge half) measurement: dvp_edge_segment on one B200 against the CPU restatement of
EdgeSegment(scale, image, 0, true) (one host thread, as the reference runs it) and, where importable, OpenCV's own
cv2.Canny on the same image.  One JSON line per size.   python tools/bench_edges.py [--sizes 3111x2073,6221x4146]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dvp_mvs_b200 import edge_segment, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="3111x2073")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libapd_cpu.so"))
    lib.edge_cpu_segment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    for size in a.sizes.split(","):
        W, H = (int(v) for v in size.split("x"))
        sc = synth.make_scene(W, H, 1)
        img = np.clip(np.rint(sc.images[0]), 0, 255).astype(np.uint8)
        edge_segment(img)                                   # warm-up
        dev, wall = [], []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            edge, thr, ms = edge_segment(img)
            wall.append(time.perf_counter() - t0); dev.append(ms)
        want = np.empty_like(img)
        t0 = time.perf_counter()
        lib.edge_cpu_segment(img.ctypes.data, W, H, want.ctypes.data, None)
        cpu_s = time.perf_counter() - t0
        out = dict(row="N4 edge prior", size=[W, H], thresholds=list(thr), edge_fraction=float((edge > 0).mean()),
                   gpu_device_ms=float(np.median(dev)), gpu_wall_ms=1e3 * float(np.median(wall)),
                   gpu_mpix_per_s=W * H / 1e3 / float(np.median(dev)), cpu_ms=1e3 * cpu_s, cpu_threads=1,
                   identical=bool((edge == want).all()))
        try:
            import cv2
            cv2.setNumThreads(1)
            t0 = time.perf_counter()
            cv2.Canny(img, float(thr[0]), float(thr[1]), apertureSize=3, L2gradient=True)
            out["opencv_canny_ms_1_thread"] = 1e3 * (time.perf_counter() - t0)
        except ImportError:
            pass
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
