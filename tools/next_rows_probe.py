"""Small driver over the N3 (fusion) and N4 (edge prior) kernels for compute-sanitizer and ncu runs:
  compute-sanitizer --tool memcheck python tools/next_rows_probe.py --small
  ncu --set full -k regex:'k_fuse_candidates|k_edge_nms|k_edge_link' -c 3 python tools/next_rows_probe.py"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dvp_mvs_b200 import Fusion, edge_segment, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    full = (640, 480) if a.small else (3110, 2074)
    mv = synth.make_multiview(full[0], full[1], 3, 2, seed=4)
    for levels in ([1, 0, 1], 1):
        views = synth.make_fusion_views(mv, levels)
        f = Fusion(views)
        pts, ms = f.run()
        print("fusion", levels, len(pts), "points", round(ms, 3), "ms", flush=True)
        f.close()
    rng = np.random.default_rng(5)
    imgs = [np.clip(np.rint(mv.levels[1][0]["image"]), 0, 255).astype(np.uint8), rng.integers(0, 256, (67, 33)).astype(np.uint8),
            rng.integers(0, 256, (3, 3)).astype(np.uint8)]
    for img in imgs:
        edge, thr, ms = edge_segment(img)
        print("edges", img.shape, thr, int((edge > 0).sum()), round(ms, 3), "ms", flush=True)


if __name__ == "__main__":
    main()
