#!/usr/bin/env python
"""The reference's K2 (GenEdgeInform) from several builds of the same unmodified source: does it run, is it right (against
the product's K2, which tests pin to the -O0 build), how long does it take?  One subprocess per build and size.
  python tools/ref_k2_matrix.py [c2 c3]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import argparse, numpy as np, bench, ref_oracle
from dvp_mvs_b200 import Engine
a = argparse.Namespace(src=4, iters=1, state="refine_iter", geom=1)
a.width, a.height = bench.WORKLOADS[%(wl)r]
sc, p, inputs, name = bench.make_workload(a, seed=0)
ref = ref_oracle.engine(a.width, a.height, 4, p); prod = Engine(a.width, a.height, 4, p)
ref.upload(**inputs); prod.upload(**inputs)
ref.run_stage("K1_INIT_RANDOM_STATES"); prod.run_stage("K1_INIT_RANDOM_STATES")
t0 = time.perf_counter(); ref.run_stage("K2_GEN_EDGE_INFORM"); dt = time.perf_counter() - t0
prod.run_stage("K2_GEN_EDGE_INFORM")
x, y = ref.get("edge_neigh"), prod.get("edge_neigh")
bad = [int((x[:, :, d] != y[:, :, d]).any(-1).sum()) for d in range(8)]
w = int((ref.get("weak") != prod.get("weak")).sum())
print("RESULT ms=%%.1f edge_neigh mismatches per direction %%s weak mismatches %%d" %% (1e3 * dt, bad, w))
'''


def main():
    sizes = sys.argv[1:] or ["c2", "c3"]
    for wl in sizes:
        for lib in ("libapd_ref_k2.so", "libapd_ref_k2_O1.so", "libapd_ref_k2_jit.so"):
            env = dict(os.environ, DVP_REF_K2_LIB=lib)
            r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT, wl=wl)], capture_output=True, text=True, env=env, timeout=900)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
            print(f"{wl} {lib:24s} {line[0] if line else 'FAILED: ' + (r.stderr.strip().splitlines() or ['?'])[-1][:200]}", flush=True)


if __name__ == "__main__":
    main()
