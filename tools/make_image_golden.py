#!/usr/bin/env python
"""Golden vectors for the float image pyramid (tests/golden/resize_f32.npz): cv2.resize(INTER_LINEAR) of 8-bit-valued float
images at the sizes the reference's schedule produces, made by the real OpenCV with its IPP back end OFF (the generic path
the restatement and the device kernel follow).  Also prints how far an IPP-enabled build is from it."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from image_oracle import level_size, resize_linear_f32  # noqa: E402

rng = np.random.default_rng(20250201)
out = {"opencv": cv2.__version__}
cases = [(101, 77, 2), (333, 211, 8), (333, 211, 2), (160, 120, 2), (389, 259, 4), (97, 131, 4)]
worst_ipp = 0.0
for i, (w, h, scale) in enumerate(cases):
    img = rng.integers(0, 256, (h, w)).astype(np.uint8)
    dw, dh = level_size(w, h, scale)
    cv2.ipp.setUseIPP(False)
    want = cv2.resize(img.astype(np.float32), (dw, dh), interpolation=cv2.INTER_LINEAR)
    cv2.ipp.setUseIPP(True)
    ipp = cv2.resize(img.astype(np.float32), (dw, dh), interpolation=cv2.INTER_LINEAR)
    worst_ipp = max(worst_ipp, float(np.abs(ipp - want).max()))
    assert (resize_linear_f32(img.astype(np.float32), dw, dh).view(np.uint32) == want.view(np.uint32)).all(), (w, h, scale)
    out[f"image_{i}"] = img; out[f"scale_{i}"] = scale; out[f"resized_{i}"] = want
out["count"] = len(cases)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_f32.npz"), **out)
print(f"wrote {len(cases)} cases (OpenCV {cv2.__version__}); IPP-enabled cv2.resize differs from the generic path by up to {worst_ipp:.4f} grey levels")
