"""Groundwork for the label half of row N4: the CPU restatement of EdgeSegment(scale, image, mode 1)
(reference APD.cpp:348-402, 437-499; oracle/cpu/label_cpu.cpp + hough_cpu.cpp) pinned against OpenCV 4.13 — golden vectors
made by the real cv2 (tools/make_label_golden.py -> tests/golden/label_segment.npz) and, where cv2 is importable, each
restated OpenCV routine (8-bit cv::resize, cv::line, cv::HoughLinesP) on random inputs.  No device path uses it yet."""
import ctypes as C
import os

import numpy as np
import pytest

from util import ROOT, GOLDEN

LIB = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libapd_cpu.so"))
LIB.label_cpu_segment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
LIB.label_cpu_size.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
LIB.label_cpu_resize8u.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
LIB.label_cpu_line.argtypes = [C.c_void_p] + [C.c_int] * 7
LIB.label_cpu_roberts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
LIB.hough_cpu_lines_p.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]


def segment(img, scale):
    img = np.ascontiguousarray(img, np.uint8)
    H, W = img.shape
    nc, nr = C.c_int(), C.c_int()
    LIB.label_cpu_size(W, H, scale, C.byref(nc), C.byref(nr))
    labels = np.empty((nr.value, nc.value), np.int32)
    small = np.empty(((H // 2) // 2, (W // 2) // 2), np.uint8)
    assert LIB.label_cpu_segment(img.ctypes.data, W, H, scale, labels.ctypes.data, small.ctypes.data) == 0
    return labels, small


def test_label_restatement_matches_opencv_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "label_segment.npz"))
    assert int(g["count"]) >= 5
    for i in range(int(g["count"])):
        img = g[f"image_{int(g[f'image_of_{i}'])}"]
        labels, small = segment(img, int(g[f"scale_{i}"]))
        np.testing.assert_array_equal(small, g[f"edge_small_{i}"])        # after the Hough lines were drawn
        np.testing.assert_array_equal(labels, g[f"labels_{i}"])


def test_label_map_properties():
    g = np.load(os.path.join(GOLDEN, "label_segment.npz"))
    img = g["image_1"]
    labels, _ = segment(img, 1)
    H, W = img.shape
    assert labels.shape == ((H + 1) // 2, (W + 1) // 2) or labels.shape == (round(H / 2), round(W / 2))
    weak_tex_num = int(H * W / (1024 << 1 << 1))
    ids, counts = np.unique(labels[labels > 0], return_counts=True)
    assert (counts > weak_tex_num).all()                               # kept regions are larger than the limit ...
    assert ((labels == -1) | (labels == 0) | (labels > 0)).all()
    # ... and two different positive labels never touch along a row or column away from the last row / column (Label_Update)
    a, b = labels[:-1, :-1], labels[:-1, 1:]
    assert not ((a > 0) & (b > 0) & (a != b)).any()
    a, b = labels[:-1, :-1], labels[1:, :-1]
    assert not ((a > 0) & (b > 0) & (a != b)).any()
    # Roberts' frame is the constant sqrt(50^2 + 50^2) = 70
    r = np.empty_like(img)
    LIB.label_cpu_roberts(np.ascontiguousarray(img).ctypes.data, W, H, r.ctypes.data)
    assert (r[0] == 70).all() and (r[-1] == 70).all() and (r[:, 0] == 70).all() and (r[:, -1] == 70).all()


def test_restated_opencv_routines_match_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(31)
    for sw, sh in ((640, 480), (777, 518), (97, 65), (33, 31)):      # cv::resize, 8-bit, INTER_LINEAR
        src = rng.integers(0, 256, (sh, sw)).astype(np.uint8)
        for dw, dh in ((sw // 2, sh // 2), (sw // 4, sh // 4), (int(sw * 1.7), int(sh * 1.7)), (sw * 2, sh * 2), (sw * 4 + 3, sh * 4 + 1), (sw, sh)):
            out = np.empty((dh, dw), np.uint8)
            LIB.label_cpu_resize8u(src.ctypes.data, sw, sh, out.ctypes.data, dw, dh)
            np.testing.assert_array_equal(out, cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR), err_msg=f"{sw}x{sh}->{dw}x{dh}")
    for _ in range(1500):                                              # cv::line, thickness 1
        H, W = int(rng.integers(2, 60)), int(rng.integers(2, 80))
        a = np.zeros((H, W), np.uint8); b = a.copy()
        p0 = (int(rng.integers(0, W)), int(rng.integers(0, H))); p1 = (int(rng.integers(0, W)), int(rng.integers(0, H)))
        cv2.line(a, p0, p1, 255, 1)
        LIB.label_cpu_line(b.ctypes.data, W, H, p0[0], p0[1], p1[0], p1[1], 255)
        np.testing.assert_array_equal(a, b)
    for _ in range(40):                                                # cv::HoughLinesP: same lines in the same order
        H, W = int(rng.integers(40, 200)), int(rng.integers(40, 260))
        img = np.zeros((H, W), np.uint8)
        for _ in range(int(rng.integers(1, 8))):
            cv2.line(img, (int(rng.integers(0, W)), int(rng.integers(0, H))), (int(rng.integers(0, W)), int(rng.integers(0, H))), 255, 1)
        img[rng.random((H, W)) < 0.01] = 255
        thr, ml, mg = int(rng.integers(1, 30)), int(rng.integers(1, 30)), int(rng.integers(1, 10))
        ref = cv2.HoughLinesP(img, 1, np.pi / 180, thr, minLineLength=ml, maxLineGap=mg)
        ref = np.zeros((0, 4), np.int32) if ref is None else ref.reshape(-1, 4)
        out = np.zeros((8192, 4), np.int32)
        n = LIB.hough_cpu_lines_p(img.ctypes.data, W, H, 1.0, np.float32(np.pi / 180), thr, ml, mg, out.ctypes.data, 8192)
        assert n == len(ref)
        np.testing.assert_array_equal(out[:n], ref)
