"""Golden vectors for the depth-edge prior (row N4, edge half).  Runs HERE (needs Python cv2; the GPU box never runs it):
a line-by-line Python transcription of the reference's EdgeSegment(scale, image, mode 0, use_canny = true)
(APD.cpp:348-466) calling the real OpenCV (cv2.Canny / cv2.resize / cv2.threshold) -> tests/golden/edge_canny.npz.
  python tools/make_edge_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def edge_segment_cv2(src: np.ndarray):
    """EdgeSegment(scale, src, 0, true) with OpenCV doing what the reference asks OpenCV to do."""
    rows, cols = src.shape
    histogram = np.zeros(256, np.float32)
    for v, c in zip(*np.unique(src, return_counts=True)):
        histogram[v] = np.float32(min(int(c), 1 << 24))          # float increments stop at 2^24
    half = rows * cols // 2
    median_val, temp = -1, 0
    for i in range(255):
        temp = int(np.float32(temp) + histogram[i])
        if temp > half:
            median_val = i
            break
    sigma = np.float32(0.67)
    threshold1 = int(np.float32(np.float32(1) - sigma) * np.float32(median_val))   # C truncation toward zero
    threshold2 = median_val
    dst = cv2.Canny(src, float(threshold1), float(threshold2), apertureSize=3, L2gradient=True)
    canny = dst.copy()
    dst = cv2.resize(dst, (cols, rows), interpolation=cv2.INTER_LINEAR)
    _, dst = cv2.threshold(dst, 4, 255, cv2.THRESH_BINARY)
    d = dst.reshape(-1).copy()
    for y in range(rows):
        if d[y * cols + 1] == 0:
            d[y * cols] = 0
        if d[y * cols + cols - 2] == 0:
            d[y * cols + cols - 1] = 0
    for x in range(cols):
        if d[cols + x] == 0:
            d[x] = 0
        if d[(rows - 2) * cols + x] == 0:
            d[(rows - 1) * cols + x] = 0
    return d.reshape(rows, cols), canny, (threshold1, threshold2)


def cases():
    from dvp_mvs_b200 import synth
    rng = np.random.default_rng(20250105)
    out = []
    sc = synth.make_scene(160, 120, 1)
    out.append(np.clip(np.rint(sc.images[0]), 0, 255).astype(np.uint8))                  # the synthetic room
    out.append(rng.integers(0, 256, (61, 83)).astype(np.uint8))                           # white noise, odd size
    out.append(cv2.GaussianBlur(rng.integers(0, 256, (96, 128)).astype(np.uint8), (0, 0), 1.2))
    img = np.full((64, 96), 40, np.uint8); img[10:50, 20:70] = 200; img[0:5, :] = 120; img[:, 90:] = 250
    out.append(img)                                                                       # edges touching the border
    out.append(np.full((16, 16), 255, np.uint8))                                          # no median below 255 -> -1
    bright = np.full((32, 48), 255, np.uint8); bright[8:20, 8:30] = 7
    out.append(bright)                                                                    # median -1 with structure
    out.append((np.add.outer(np.arange(50), np.arange(70)) * 3 % 256).astype(np.uint8))   # ramps
    return out


def main():
    data = {}
    for i, img in enumerate(cases()):
        edge, canny, thr = edge_segment_cv2(img)
        data[f"image_{i}"] = img; data[f"edge_{i}"] = edge; data[f"canny_{i}"] = canny
        data[f"thresholds_{i}"] = np.array(thr, np.int32)
        print(i, img.shape, thr, int((edge > 0).sum()), int((canny > 0).sum()))
    data["count"] = np.array(len(cases()), np.int32)
    data["opencv_version"] = np.array(cv2.__version__)
    path = os.path.join(ROOT, "tests", "golden", "edge_canny.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
