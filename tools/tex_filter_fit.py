#!/usr/bin/env python
"""Which arithmetic reproduces the B200 texture unit's bilinear fp32 filter on NON-integer texels?
Reads gpurun_out/tex_filter_probe.bin (tools/tex_filter_probe.cu) and scores candidate formulas bit for bit."""
import sys
import numpy as np

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tex_filter_probe.bin"
raw = open(path, "rb").read()
W, H, n = np.frombuffer(raw, np.int32, 3)
off = 12
img = np.frombuffer(raw, np.float32, W * H, off).reshape(H, W); off += W * H * 4
xy = np.frombuffer(raw, np.float32, 2 * n, off).reshape(n, 2); off += 8 * n
res = np.frombuffer(raw, np.float32, n, off)

x = xy[:, 0].astype(np.float64); y = xy[:, 1].astype(np.float64)
qx = np.floor((x - 0.5) * 256.0 + 0.5); qy = np.floor((y - 0.5) * 256.0 + 0.5)
ix = np.floor(qx / 256.0).astype(np.int64); iy = np.floor(qy / 256.0).astype(np.int64)
A = (qx - 256.0 * ix).astype(np.int64); B = (qy - 256.0 * iy).astype(np.int64)
cl = lambda v, hi: np.clip(v, 0, hi - 1)
t00 = img[cl(iy, H), cl(ix, W)]; t10 = img[cl(iy, H), cl(ix + 1, W)]
t01 = img[cl(iy + 1, H), cl(ix, W)]; t11 = img[cl(iy + 1, H), cl(ix + 1, W)]
w11 = (A * B + 128) >> 8; w10 = A - w11; w01 = B - w11; w00 = 256 - A - B + w11
f32 = np.float32
W00, W10, W01, W11 = (w.astype(np.float64) for w in (w00, w10, w01, w11))
T00, T10, T01, T11 = (t.astype(np.float64) for t in (t00, t10, t01, t11))


def score(name, val):
    v = np.asarray(val, np.float32)
    same = v.view(np.uint32) == res.view(np.uint32)
    region = (np.arange(n) % 3)
    clean = (np.arange(n) % 97) != 0
    out = [f"{100.0 * same[(region == r) & clean].mean():6.2f}%" for r in range(3)]
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(v.astype(np.float64) - res) / np.maximum(np.abs(res), 1e-30)
    print(f"{name:58s} frac-valued {out[0]}  wide-range {out[1]}  8-bit {out[2]}   max rel err {np.nanmax(rel[clean]):.2e}")


exact = (W00 * T00 + W10 * T10 + W01 * T01 + W11 * T11) / 256.0
score("exact sum in float64, rounded once", exact)
fa = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)   # fp32 FMA (exact product, one rounding)
z = np.zeros(n, np.float32)
w = [f32(W00), f32(W10), f32(W01), f32(W11)]; t = [t00, t10, t01, t11]
import itertools
for order in itertools.permutations(range(4)):
    acc = z
    for k in order:
        acc = fa(w[k], t[k], acc)
    score(f"fp32 FMA chain order {order} then /256", acc * f32(1 / 256))
# separable lerps with 8-bit fractions
a = f32(A.astype(np.float64) / 256); b = f32(B.astype(np.float64) / 256)
top = fa(a, (t10 - t00), t00); bot = fa(a, (t11 - t01), t01)
score("lerp x then y: t0 + a (t1 - t0), fp32", fa(b, (bot - top), top))
lft = fa(b, (t01 - t00), t00); rgt = fa(b, (t11 - t10), t10)
score("lerp y then x, fp32", fa(a, (rgt - lft), lft))
# truncation instead of rounding of the exact sum
ex32 = exact.astype(np.float32)
down = np.where(ex32.astype(np.float64) > exact, np.nextafter(ex32, f32(-np.inf)), ex32)
score("exact sum, truncated toward zero", np.where(exact >= 0, down, ex32))
