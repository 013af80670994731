#!/usr/bin/env python
"""GPU diagnostic for individual stages: K2 edge_neigh detail, K7 race-vs-real mismatch split, K16 detail."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dvp_mvs_b200 import Engine, default_params, synth, FIRST_INIT
from dvp_mvs_b200.parity import compare, STATE_BUFS

W, H, S = 640, 480, 2
sc = synth.make_scene(W, H, S)
p = default_params(); p.max_iterations = 1; p.num_images = S + 1
p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
p.use_APD = 0; p.state = FIRST_INIT; p.weak_peak_radius = 6
kw = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
ref = Engine(W, H, S, p, lib_path=ref_oracle.REFERENCE_LIB, prefix="ref_"); prod = Engine(W, H, S, p, )
ref.upload(**kw); prod.upload(**kw)

def state(e): return {n: e.get(n) for n in STATE_BUFS}
def load(e, st):
    for n, a in st.items(): e.set(n, a)

for st in ["K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM"]:
    ref.run_stage(st); prod.run_stage(st)
a, b = ref.get("edge_neigh"), prod.get("edge_neigh")
bad = np.argwhere((a != b).any(-1))
print("K2 edge_neigh bad entries", len(bad), "per-direction counts", np.bincount(bad[:, 2], minlength=8))
for y, x, d in bad[:6]:
    print("  pixel", (x, y), "dir", d, "ref", a[y, x, d], "prod", b[y, x, d])
# brute force check of the definition on a few pixels
dirs = [(0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (1, 1), (-1, 1), (1, -1)]
def brute(x, y, d):
    dx, dy = dirs[d]; nx, ny = x + dx, y + dy
    while 0 <= nx < W and 0 <= ny < H:
        if sc.edge[ny, nx]: return (nx, ny)
        nx += dx; ny += dy
    return (-1, -1)
for y, x, d in bad[:6]:
    print("  brute", (x, y), d, brute(x, y, d))

# ---- K3..K6 on ref, then K7 analysis
for st in ["K3_FIND_NEAREST_STRONG", "K4_GEN_NEIGHBOURS", "K5_NEIGHBOUR_UPDATE", "K6_RANDOM_INITIALIZATION"]:
    ref.run_stage(st)
pre = state(ref)
outs = ("planes", "costs", "selected", "view_weight", "rand")
ref.run_stage("K7_BLACK_STRONG", 0); r1 = {n: ref.get(n) for n in outs}
load(ref, pre); ref.run_stage("K7_BLACK_STRONG", 0); r2 = {n: ref.get(n) for n in outs}
load(prod, pre); prod.run_stage("K7_BLACK_STRONG", 0); p1 = {n: prod.get(n) for n in outs}
load(prod, pre); prod.run_stage("K7_BLACK_STRONG", 0); p2 = {n: prod.get(n) for n in outs}
def badmask(u, v):
    m = np.zeros((H, W), bool)
    for n in outs:
        x, y = u[n], v[n]
        eq = (x == y) | ((x != x) & (y != y)) if x.dtype.kind == "f" else (x == y)
        m |= ~eq.reshape(H, W, -1).all(-1)
    return m
m_rr, m_pp, m_rp = badmask(r1, r2), badmask(p1, p2), badmask(r1, p1)
print("K7: ref-vs-ref differing pixels", m_rr.sum(), "| prod-vs-prod", m_pp.sum(), "| ref-vs-prod", m_rp.sum(),
      "| ref-vs-prod outside ref-noise", (m_rp & ~m_rr).sum())
# which buffer differs first for real mismatches
real = np.argwhere(m_rp & ~m_rr & ~m_pp)
print("real mismatches (stable in both):", len(real))
for y, x in real[:8]:
    print("  px", (x, y), "edge", sc.edge[y, x], "cost ref/prod", r1["costs"][y, x], p1["costs"][y, x],
          "vw ref", r1["view_weight"][y, x, :S], "prod", p1["view_weight"][y, x, :S],
          "sel", r1["selected"][y, x], p1["selected"][y, x], "rand eq", (r1["rand"][y, x] == p1["rand"][y, x]).all())
    print("     plane ref", r1["planes"][y, x], "prod", p1["planes"][y, x], "pre", pre["planes"][y, x], "precost", pre["costs"][y, x])
# border statistics of real mismatches
if len(real):
    ys, xs = real[:, 0], real[:, 1]
    print("  real mismatch x range", xs.min(), xs.max(), "y range", ys.min(), ys.max(),
          "near border(<30px):", ((xs < 30) | (ys < 30) | (xs >= W - 30) | (ys >= H - 30)).sum())

# ---- K16 detail
ref.upload(**kw); ref.run(mode=0)
