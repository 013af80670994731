#!/usr/bin/env python
"""Race-free K7/K8 parity: only pixels with (x+y) % 128 in {0,1} are STRONG, everything else WEAK (skipped by the
strong sweep).  Direction 4 (the reference's same-colour read) stays on the pixel's own diagonal
(x-y constant) and moves x+y by 10 + 2*k*len <= 10 + 2*21*2 < 128, so it can only ever read WEAK pixels, which no
thread writes: the sweep becomes deterministic and must match the reference bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT, REFINE_ITER, WEAK, STRONG
from dvp_mvs_b200.parity import STATE_BUFS

W, H, S = 640, 480, 2
sc = synth.make_scene(W, H, S)
p = default_params(); p.max_iterations = 1; p.num_images = S + 1
p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
p.use_APD = 1; p.state = FIRST_INIT; p.weak_peak_radius = 6
yy, xx = np.mgrid[0:H, 0:W]
weak = np.where(((xx + yy) % 128) < 2, STRONG, WEAK).astype(np.uint8)
kw = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, weak_info=weak, seed=synth.SEED_RNG)
ref = ref_oracle.engine(W, H, S, p); prod = Engine(W, H, S, p)
ref.upload(**kw); prod.upload(**kw)
ref.run_stage("K1_INIT_RANDOM_STATES"); ref.run_stage("K2_GEN_EDGE_INFORM"); ref.run_stage("K6_RANDOM_INITIALIZATION")
outs = ("planes", "costs", "selected", "view_weight", "rand")
strong = weak == STRONG
for it in range(2):
  for st in ("K7_BLACK_STRONG", "K8_RED_STRONG"):
    pre = {n: ref.get(n) for n in STATE_BUFS}
    ref.run_stage(st, it); r1 = {n: ref.get(n) for n in outs}
    for n, a in pre.items(): ref.set(n, a)
    ref.run_stage(st, it); r2 = {n: ref.get(n) for n in outs}
    for n, a in pre.items(): prod.set(n, a)
    prod.run_stage(st, it); p1 = {n: prod.get(n) for n in outs}
    def bad(u, v):
        m = np.zeros((H, W), bool)
        for n in outs:
            x, y = u[n], v[n]
            eq = (x == y) | ((x != x) & (y != y)) if x.dtype.kind == "f" else (x == y)
            m |= ~eq.reshape(H, W, -1).all(-1)
        return m
    m_rr, m_rp = bad(r1, r2), bad(r1, p1)
    changed = bad(r1, pre)
    print(st, it, "strong px", strong.sum(), "changed by ref", changed.sum(), "| ref-vs-ref", m_rr.sum(), "| ref-vs-prod", m_rp.sum(), "outside strong", (m_rp & ~strong).sum())
    for y, x in np.argwhere(m_rp)[:6]:
        print("  px", (x, y), "edge", sc.edge[y, x], "cost ref/prod", r1["costs"][y, x], p1["costs"][y, x], "pre", pre["costs"][y, x],
              "vw", r1["view_weight"][y, x, :S], p1["view_weight"][y, x, :S], "sel", r1["selected"][y, x], p1["selected"][y, x],
              "rand eq", (r1["rand"][y, x] == p1["rand"][y, x]).all())
        print("     plane ref", r1["planes"][y, x], "prod", p1["planes"][y, x], "pre", pre["planes"][y, x])
