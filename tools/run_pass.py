#!/usr/bin/env python
"""One or more RunPatchMatch passes of a bench workload through the C ABI — the command ncu / compute-sanitizer wrap.
  python tools/run_pass.py --workload c3 --iters 1 --passes 1 [--derived 1]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3", choices=["c3", "c2"])
    ap.add_argument("--src", type=int, default=4)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--passes", type=int, default=1)
    ap.add_argument("--state", default="refine_iter")
    ap.add_argument("--geom", type=int, default=1)
    ap.add_argument("--derived", type=int, default=0, help="pixel states from a real previous pass instead of the painted wall")
    ap.add_argument("--digest", type=int, default=1, help="print a SHA-1 over every output buffer of the last pass")
    ap.add_argument("--size", default="", help="WxH instead of the workload's size (small runs under compute-sanitizer)")
    a = ap.parse_args()
    a.width, a.height = bench.WORKLOADS[a.workload]
    if a.size:
        a.width, a.height = (int(v) for v in a.size.split("x"))
    from dvp_mvs_b200 import Engine
    sc, p, inputs, name = bench.make_workload(a, seed=0)
    if a.derived:
        inputs, wfrac = bench.derive_states(lambda q: Engine(a.width, a.height, a.src, q), a, sc, p, inputs)
        name += f"_derived_weak{int(round(100 * wfrac))}pct"
    e = Engine(a.width, a.height, a.src, p)
    for _ in range(a.passes):
        e.upload(**inputs)
        e.run()
    total, per_stage, launches = e.last_run_times()
    digest = ""
    if a.digest:   # our launches are deterministic: two builds that compute the same thing print the same digest
        import hashlib
        h = hashlib.sha1()
        for k, v in sorted(e.snapshot().items()):
            h.update(k.encode()); h.update(v.tobytes())
        digest = " sha1=" + h.hexdigest()[:16]
    print(name, f"{total:.2f} ms", [round(v, 2) for v in per_stage], launches, "launches" + digest)


if __name__ == "__main__":
    main()
