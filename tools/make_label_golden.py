"""Golden vectors for the label half of row N4 (oracle groundwork; no device path yet).  Runs HERE (needs Python cv2):
a line-by-line Python transcription of EdgeSegment(scale, image, mode 1, use_canny = false) (reference APD.cpp:348-402,
437-499) in which every OpenCV call is made by the real OpenCV (cv2.resize / cv2.threshold / cv2.HoughLinesP / cv2.line);
the reference's own Connect + Label_Update (APD.cpp:138-346) come from the restatement shared with row N1
(oracle/cpu/ccl_ref.hpp, pinned by hand-computed cases in tests/test_visibility.py) -> tests/golden/label_segment.npz.
  python tools/make_label_golden.py"""
import ctypes as C
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libapd_cpu.so"))
LIB.label_cpu_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]


def connect_update(img):
    h, w = img.shape
    labels = np.empty((h, w), np.int32); counts = np.zeros(h * w + 1, np.int32)
    n = LIB.label_cpu_connect(np.ascontiguousarray(img).ctypes.data, w, h, labels.ctypes.data, counts.ctypes.data, len(counts))
    return labels, counts[:n]


def roberts(src):
    h, w = src.shape
    s = src.astype(np.int32)
    t1 = np.full((h, w), 50, np.int32); t2 = t1.copy()
    t1[1:-1, 1:-1] = s[1:-1, 1:-1] - s[2:, 2:]
    t2[1:-1, 1:-1] = s[2:, 1:-1] - s[1:-1, 2:]
    return (np.sqrt((t1 * t1 + t2 * t2).astype(np.float64)).astype(np.int32) & 255).astype(np.uint8)


def edge_segment_labels_cv2(src, scale):
    rows, cols = src.shape
    weak_tex_num = int(1.0 * rows * cols / (1024 << scale << scale))
    down = cv2.resize(src, (cols // 2, rows // 2), interpolation=cv2.INTER_LINEAR)
    down = cv2.resize(down, (down.shape[1] // 2, down.shape[0] // 2), interpolation=cv2.INTER_LINEAR)
    m = min(down.shape[1], down.shape[0])
    houthr = min_len = max_gap = int(m / 30.0)
    dst = roberts(down)
    _, dst = cv2.threshold(dst, 4, 255, cv2.THRESH_BINARY)
    lab0, cnt0 = connect_update(dst)
    for k in range(1, len(cnt0)):
        if cnt0[k] < weak_tex_num:
            continue
        inside = lab0 == k
        near = np.zeros_like(inside)
        near[:, 1:] |= inside[:, :-1]; near[:, :-1] |= inside[:, 1:]; near[1:, :] |= inside[:-1, :]; near[:-1, :] |= inside[1:, :]
        img_weak = np.where(near & ~inside, 255, 0).astype(np.uint8)
        lines = cv2.HoughLinesP(img_weak, 1, np.pi / 180, houthr, minLineLength=min_len, maxLineGap=max_gap)
        for ln in ([] if lines is None else lines.reshape(-1, 4)):
            cv2.line(dst, (int(ln[0]), int(ln[1])), (int(ln[2]), int(ln[3])), (255, 0, 0), 1)
    edge_small = dst.copy()
    factor = np.float32(1.0) / np.float32(1 << scale)
    new_cols = int(np.floor(np.float32(cols) * factor + np.float32(0.5))); new_rows = int(np.floor(np.float32(rows) * factor + np.float32(0.5)))
    up = cv2.resize(dst, (new_cols, new_rows), interpolation=cv2.INTER_LINEAR)
    _, up = cv2.threshold(up, 4, 255, cv2.THRESH_BINARY)
    d = up.reshape(-1).copy()
    for y in range(new_rows):
        if d[y * new_cols + 1] == 0:
            d[y * new_cols] = 0
        if d[y * new_cols + new_cols - 2] == 0:
            d[y * new_cols + new_cols - 1] = 0
    for x in range(new_cols):
        if d[new_cols + x] == 0:
            d[x] = 0
        if d[(new_rows - 2) * new_cols + x] == 0:
            d[(new_rows - 1) * new_cols + x] = 0
    lab, cnt = connect_update(d.reshape(new_rows, new_cols))
    small = (cnt[lab] <= weak_tex_num) & (lab != 0)
    return np.where(small, -1, lab).astype(np.int32), edge_small


def cases():
    from dvp_mvs_b200 import synth
    rng = np.random.default_rng(20250106)
    out = []
    sc = synth.make_scene(400, 300, 1)
    room = np.clip(np.rint(sc.images[0]), 0, 255).astype(np.uint8)
    out += [(room, 1), (room, 2)]                                                  # the synthetic room at two pyramid scales
    img = np.full((301, 403), 90, np.uint8)                                         # flat regions separated by lines, odd sizes
    cv2.line(img, (10, 20), (390, 250), 200, 3); cv2.line(img, (200, 0), (180, 300), 30, 5); cv2.circle(img, (300, 100), 60, 160, -1)
    img = cv2.GaussianBlur(img, (0, 0), 1.0)
    out += [(img, 0), (img, 1)]
    noise = cv2.GaussianBlur(rng.integers(0, 256, (160, 200)).astype(np.uint8), (0, 0), 5.0)
    out += [(noise, 1)]
    return out


def main():
    data = {}
    stored = []
    for i, (img, scale) in enumerate(cases()):
        labels, edge_small = edge_segment_labels_cv2(img, scale)
        which = next((k for k, s in enumerate(stored) if s is img), None)   # an image used at several scales is stored once
        if which is None:
            which = len(stored); stored.append(img); data[f"image_{which}"] = img
        data[f"image_of_{i}"] = np.array(which, np.int32); data[f"scale_{i}"] = np.array(scale, np.int32)
        data[f"labels_{i}"] = labels; data[f"edge_small_{i}"] = edge_small
        print(i, img.shape, scale, labels.shape, "regions", int(labels.max()), "small", int((labels == -1).sum()), "boundary", int((labels == 0).sum()))
    data["count"] = np.array(len(cases()), np.int32)
    data["opencv_version"] = np.array(cv2.__version__)
    path = os.path.join(ROOT, "tests", "golden", "label_segment.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
