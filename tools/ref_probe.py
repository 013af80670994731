#!/usr/bin/env python
"""Which stage of the reference's own kernel sequence fails (or how long each takes) on a bench workload?
  python tools/ref_probe.py --workload c3 [--k2 libapd_ref_k2_O1.so]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--src", type=int, default=4)
    ap.add_argument("--iters", type=int, default=1)
    ap.add_argument("--k2", default="libapd_ref_k2_O1.so")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    a = ap.parse_args()
    os.environ["DVP_REF_K2_LIB"] = a.k2
    w, h = bench.WORKLOADS[a.workload]
    a.width = a.width or w; a.height = a.height or h
    a.state = "refine_iter"; a.geom = 1
    import ref_oracle
    from dvp_mvs_b200.parity import sequence
    sc, p, inputs, name = bench.make_workload(a, seed=0)
    e = ref_oracle.engine(a.width, a.height, a.src, p)
    e.upload(**inputs)
    print(name, "weak", e.weak_count(), flush=True)
    for st, it in sequence(a.iters):
        t0 = time.perf_counter()
        try:
            e.run_stage(st, it)
        except Exception as ex:  # noqa: BLE001
            print(f"{st}[{it}] FAILED: {ex}", flush=True)
            return 1
        print(f"{st}[{it}] ok {1e3 * (time.perf_counter() - t0):.1f} ms (wall, synchronous)", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
