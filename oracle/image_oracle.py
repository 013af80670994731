"""TEST INFRASTRUCTURE ONLY — numpy restatement of the OpenCV routines the reference's image preparation calls
(cv::resize, third-party; OpenCV >= 3.3 per the reference's README, golden vectors from 4.13.0).
resize_linear_f32: the GENERIC bilinear path for CV_32F (imgproc/src/resize.cpp, resizeGeneric_ with HResizeLinear /
VResizeLinear) as InuputInitialization / GetProblemEdges use it on the float grey image (APD.cpp:1119-1140,
main.cpp:203-209).  Pinned by tests/golden/resize_f32.npz (cv2 with its IPP back end switched off: bit-exact; an
IPP-enabled OpenCV differs by up to 0.015 grey levels, tools/make_image_golden.py prints the figure)."""
import numpy as np


def level_size(cols: int, rows: int, scale: int):
    """new_cols / new_rows of APD.cpp:1121-1123: std::round(cols * (1.0f / scale)) in float arithmetic."""
    f = np.float32(1.0) / np.float32(scale)
    return int(np.floor(np.float32(cols) * f + np.float32(0.5))), int(np.floor(np.float32(rows) * f + np.float32(0.5)))


def _taps(n_dst: int, n_src: int):
    scale = 1.0 / (np.float64(n_dst) / np.float64(n_src))
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    return s, (f - s.astype(np.float32)).astype(np.float32)


def resize_linear_f32(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.float32)
    sh, sw = src.shape
    sx, fx = _taps(dw, sw)
    lo = sx < 0; fx[lo] = 0; sx[lo] = 0                      # resize.cpp: "if (sx < 0) fx = 0, sx = 0"
    hi = sx >= sw - 1; fx[hi] = 0; sx[hi] = sw - 1           # "if (sx >= ssize.width - 1) fx = 0, sx = ssize.width - 1"
    sx1 = np.minimum(sx + 1, sw - 1)
    a0 = (np.float32(1) - fx).astype(np.float32)
    rows = (src[:, sx] * a0 + src[:, sx1] * fx).astype(np.float32)     # two products and a sum, each rounded to float
    sy, fy = _taps(dh, sh)
    y0 = np.clip(sy, 0, sh - 1); y1 = np.clip(sy + 1, 0, sh - 1)       # rows are clipped, the weights are not changed
    b0 = (np.float32(1) - fy).astype(np.float32)[:, None]
    return (rows[y0] * b0 + rows[y1] * fy[:, None]).astype(np.float32)
