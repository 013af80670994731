#!/usr/bin/env python
"""What does the REFERENCE do with use_edge = 0 (the ACMH-style branch of its sweep, APD.cu:2142-2460)?
The branch gathers its candidates in `positions_tmp` but the acceptance step reads `positions[min_cost_idx]` (APD.cu:2559-2563),
an array only the use_edge branch writes: with use_edge = 0 it indexes plane_hypotheses with uninitialised stack contents.
This probe runs the reference's own K7 (oracle/_ref/libapd_ref.so) twice from the same uploaded state with use_edge = 0 and
reports (a) whether the launch survives, (b) how many pixels take a plane that is none of the planes the launch could have
produced from defined data (the pixel's own plane or any plane of the input map), (c) whether two runs agree.
  python tools/ref_use_edge_probe.py [--size 320x240]        (wrap in compute-sanitizer --tool memcheck for the read addresses)"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="320x240")
    ap.add_argument("--src", type=int, default=2)
    a = ap.parse_args()
    W, H = (int(v) for v in a.size.split("x"))
    import ref_oracle
    from dvp_mvs_b200 import default_params, synth, FIRST_INIT
    sc = synth.make_scene(W, H, a.src)
    p = default_params(); p.max_iterations = 1; p.num_images = a.src + 1
    p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
    p.use_APD = 0; p.state = FIRST_INIT; p.use_edge = 0
    kw = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    outs = []
    for run in range(2):
        e = ref_oracle.engine(W, H, a.src, p)
        e.upload(**kw)
        for st in ("K1_INIT_RANDOM_STATES", "K6_RANDOM_INITIALIZATION"):
            e.run_stage(st, 0)
        before = e.get("planes").copy()
        try:
            e.run_stage("K7_BLACK_STRONG", 0)
            after = e.get("planes").copy()
        except Exception as ex:  # noqa: BLE001
            print(f"run {run}: K7 with use_edge = 0 FAILED: {ex}", flush=True)
            return 0
        outs.append(after)
        changed = (before != after).any(-1)
        nan = ~np.isfinite(after).all(-1)
        print(f"run {run}: launch survived; {int(changed.sum())} of {W * H} pixels rewritten, {int(nan.sum())} with a non-finite plane", flush=True)
        if run == 0:
            # where do the new planes come from?  After K6 every plane of the map is a distinct random plane, so a rewritten
            # pixel whose new plane equals, bit for bit, the OLD plane of another pixel copied it from there.
            key = lambda arr: arr.view(np.uint32).astype(np.uint64) @ np.array([1, 1 << 16, 1 << 32, 1 << 48], np.uint64)  # noqa: E731
            kb = key(before.reshape(-1, 4)); ka = key(after.reshape(-1, 4))
            order = np.argsort(kb); skb = kb[order]
            idx = np.flatnonzero(changed.reshape(-1))
            pos = np.searchsorted(skb, ka[idx]); pos = np.minimum(pos, len(skb) - 1)
            hit = skb[pos] == ka[idx]
            src = order[pos]
            exact = hit & (before.reshape(-1, 4)[src] == after.reshape(-1, 4)[idx]).all(-1)
            dx = (src % W) - (idx % W); dy = (src // W) - (idx // W)
            print(f"  {int(exact.sum())} of the {len(idx)} rewritten pixels hold another pixel's previous plane; the rest hold a plane not in the input map (refinement)")
            if exact.any():
                from collections import Counter
                c = Counter(zip(dx[exact].tolist(), dy[exact].tolist()))
                print("  most frequent (dx, dy) of the copied plane:", c.most_common(24))
                print("  copies from pixel index 0:", int((src[exact] == 0).sum()), " max |dx|, |dy|:", int(np.abs(dx[exact]).max()), int(np.abs(dy[exact]).max()))
    same = np.array_equal(outs[0], outs[1], equal_nan=True)
    print("two runs from the same state agree bit for bit:", same, "" if same else f"({int((outs[0] != outs[1]).any(-1).sum())} pixels differ)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
