#!/usr/bin/env python
"""A/B of build variants of the library on one workload: per-stage device times of `passes` passes for each .so given.
  python tools/ab_variants.py --workload c2 [--derived 1] name=path/to/lib.so ..."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--derived", type=int, default=0)
    ap.add_argument("--passes", type=int, default=3)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--state", default="refine_iter")
    ap.add_argument("--geom", type=int, default=1)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    rows = {}
    for spec in a.libs:
        name, path = spec.split("=", 1)
        env = dict(os.environ, DVP_MVS_LIB=os.path.abspath(path))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_pass.py"), "--workload", a.workload, "--passes", str(a.passes),
                            "--iters", str(a.iters), "--derived", str(a.derived), "--state", a.state, "--geom", str(a.geom)], capture_output=True, text=True, env=env)
        line = (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1]
        rows[name] = line
        print(f"{name:28s} {line}", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
