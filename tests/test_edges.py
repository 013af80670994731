"""Row N4, edge half: the depth-edge prior — EdgeSegment(scale, image, mode 0, use_canny = true), reference
APD.cpp:348-466 as called by GetProblemEdges (main.cpp:193-226); the Canny inside it is OpenCV's (not under
/root/reference).  CPU part: the restatement (oracle/cpu/edge_cpu.cpp) against golden vectors produced by the real OpenCV
(tools/make_edge_golden.py -> tests/golden/edge_canny.npz), against cv2 itself where it is importable, and against a numpy
model of the decomposition the CUDA kernels use (tile-free NMS + connected components instead of the stack flood).
GPU part: dvp_edge_segment / dvp_scene_compute_edges against the restatement, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from util import ROOT, GOLDEN
from dvp_mvs_b200 import synth

CPU_LIB = os.path.join(ROOT, "oracle", "_ref", "libapd_cpu.so")


def _lib():
    lib = C.CDLL(CPU_LIB)
    lib.edge_cpu_canny.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
    lib.edge_cpu_thresholds.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.edge_cpu_segment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def oracle_segment(img):
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    edge = np.empty_like(img); canny = np.empty_like(img)
    lib = _lib()
    assert lib.edge_cpu_segment(img.ctypes.data, W, H, edge.ctypes.data, canny.ctypes.data) == 0
    t1, t2 = C.c_int(), C.c_int()
    lib.edge_cpu_thresholds(img.ctypes.data, W, H, C.byref(t1), C.byref(t2))
    return edge, canny, (t1.value, t2.value)


def oracle_canny(img, low, high):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    _lib().edge_cpu_canny(img.ctypes.data, img.shape[1], img.shape[0], float(low), float(high), out.ctypes.data)
    return out


def test_cases(rng, count):
    out = []
    for i in range(count):
        H, W = int(rng.integers(3, 150)), int(rng.integers(3, 200))
        kind = i % 4
        if kind == 0:
            img = rng.integers(0, 256, (H, W))
        elif kind == 1:
            img = np.cumsum(rng.normal(0, 6, (H, W)), axis=1) + np.cumsum(rng.normal(0, 6, (H, W)), axis=0) + 128
        elif kind == 2:
            img = np.full((H, W), int(rng.integers(0, 256)))
            for _ in range(6):
                y0, x0 = int(rng.integers(0, H)), int(rng.integers(0, W))
                img[y0:y0 + int(rng.integers(1, H)), x0:x0 + int(rng.integers(1, W))] = int(rng.integers(0, 256))
        else:
            img = np.add.outer(np.arange(H), np.arange(W)) * rng.uniform(0.5, 4.0) % 256
        out.append(np.clip(img, 0, 255).astype(np.uint8))
    return out


test_cases.__test__ = False


# ---------------------------------------------------------------------------------------------------------- CPU
def test_restatement_matches_opencv_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "edge_canny.npz"))
    assert int(g["count"]) >= 7
    for i in range(int(g["count"])):
        edge, canny, thr = oracle_segment(g[f"image_{i}"])
        assert thr == tuple(int(v) for v in g[f"thresholds_{i}"]), i
        np.testing.assert_array_equal(canny, g[f"canny_{i}"])
        np.testing.assert_array_equal(edge, g[f"edge_{i}"])


def test_restatement_matches_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for img in test_cases(rng, 120):
        low, high = float(rng.integers(-2, 130)), float(rng.integers(-2, 255))
        np.testing.assert_array_equal(oracle_canny(img, low, high), cv2.Canny(img, low, high, apertureSize=3, L2gradient=True))


def kernels_model(img):
    """What dvp_kernels_edge.cu computes, in numpy: thresholds, Sobel with replicated borders, magnitude with a zero
    frame, suppression, hysteresis as 8-connected components holding a strong pixel, border clean-up."""
    from scipy import ndimage
    H, W = img.shape
    hist = np.minimum(np.bincount(img.ravel(), minlength=256), 1 << 24).astype(np.float32)
    half, med, tmp = H * W // 2, -1, 0
    for i in range(255):
        tmp = int(np.float32(tmp) + hist[i])
        if tmp > half:
            med = i
            break
    t1 = int(np.float32(np.float32(1) - np.float32(0.67)) * np.float32(med)); t2 = med
    low, high = (t1, t2) if t1 <= t2 else (t2, t1)
    low = low * low if low > 0 else low; high = high * high if high > 0 else high
    p = np.pad(img.astype(np.int32), 1, mode="edge")
    dx = (p[:-2, 2:] + 2 * p[1:-1, 2:] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[1:-1, :-2] + p[2:, :-2])
    dy = (p[2:, :-2] + 2 * p[2:, 1:-1] + p[2:, 2:]) - (p[:-2, :-2] + 2 * p[:-2, 1:-1] + p[:-2, 2:])
    m = np.pad(dx * dx + dy * dy, 1)                       # zero frame
    c = m[1:-1, 1:-1]
    ax, ay = np.abs(dx), np.abs(dy) << 15
    tg22 = ax * 13573; tg67 = tg22 + (ax << 16)
    horiz = (c > m[1:-1, :-2]) & (c >= m[1:-1, 2:])
    vert = (c > m[:-2, 1:-1]) & (c >= m[2:, 1:-1])
    s_neg = (dx ^ dy) < 0                                  # s = -1: compare with (y-1, x+1) and (y+1, x-1)
    diag = np.where(s_neg, (c > m[:-2, 2:]) & (c > m[2:, :-2]), (c > m[:-2, :-2]) & (c > m[2:, 2:]))
    cand = (c > low) & np.where(ay < tg22, horiz, np.where(ay > tg67, vert, diag))
    strong = cand & (c > high)
    lab, _ = ndimage.label(cand, structure=np.ones((3, 3), int))
    keep = np.zeros(lab.max() + 1, bool); keep[np.unique(lab[strong])] = True; keep[0] = False
    edge = np.where(keep[lab], 255, 0).astype(np.uint8)
    canny = edge.copy()
    edge[edge[:, 1] == 0, 0] = 0; edge[edge[:, W - 2] == 0, W - 1] = 0
    edge[0, edge[1, :] == 0] = 0; edge[H - 1, edge[H - 2, :] == 0] = 0
    return edge, canny, (t1, t2)


def test_kernel_decomposition_equals_restatement():
    rng = np.random.default_rng(12)
    g = np.load(os.path.join(GOLDEN, "edge_canny.npz"))
    imgs = test_cases(rng, 60) + [g[f"image_{i}"] for i in range(int(g["count"]))]
    for img in imgs:
        e0, c0, t0 = oracle_segment(img)
        e1, c1, t1 = kernels_model(img)
        assert t0 == t1
        np.testing.assert_array_equal(c1, c0)
        np.testing.assert_array_equal(e1, e0)


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_edge_segment_bit_exact_vs_restatement():
    from dvp_mvs_b200 import edge_segment, DvpError
    rng = np.random.default_rng(13)
    g = np.load(os.path.join(GOLDEN, "edge_canny.npz"))
    imgs = [g[f"image_{i}"] for i in range(int(g["count"]))] + test_cases(rng, 40)
    imgs += [rng.integers(0, 256, (3, 3)).astype(np.uint8), rng.integers(0, 256, (3, 64)).astype(np.uint8), rng.integers(0, 256, (70, 3)).astype(np.uint8)]
    sc = synth.make_scene(1555, 1037, 1)
    imgs.append(np.clip(np.rint(sc.images[0]), 0, 255).astype(np.uint8))      # a level-size image with real structure
    for i, img in enumerate(imgs):
        edge, thr, ms = edge_segment(img)
        want, _, thr_want = oracle_segment(img)
        assert thr == thr_want, (i, img.shape)
        np.testing.assert_array_equal(edge, want, err_msg=f"case {i} {img.shape}")
    for i in range(int(g["count"])):                                          # and OpenCV's own output directly
        np.testing.assert_array_equal(edge_segment(g[f"image_{i}"])[0], g[f"edge_{i}"])
    assert (edge > 0).mean() > 0.001 and ms > 0
    with pytest.raises(DvpError):
        edge_segment(np.zeros((2, 8), np.uint8))


@pytest.mark.gpu
def test_gpu_scene_computes_its_own_edges():
    from dvp_mvs_b200 import Scene
    mv = synth.make_multiview(320, 240, 2, 2, seed=8)
    sc = Scene(2, 2)
    for v in range(2):
        sc.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        for level in range(2):
            L = mv.levels[level][v]
            sc.set_level(v, level, L["image"] + np.float32(0.25), None, L["label"])    # non-integer grey levels: rounding matters
    for v in range(2):
        for level in range(2):
            img = mv.levels[level][v]["image"] + np.float32(0.25)
            u8 = np.clip(np.rint(img), 0, 255).astype(np.uint8)                        # convertTo(CV_8UC1): round half to even
            want, _, _ = oracle_segment(u8)
            np.testing.assert_array_equal(sc.compute_edges(v, level), want)
    sc.close()
