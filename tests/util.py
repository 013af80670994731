import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from dvp_mvs_b200 import default_params, synth, FIRST_INIT  # noqa: E402


def c1_params(depth_min, depth_max, S, iters=1, use_apd=0):
    """BASELINE config C1: photometric only, FIRST_INIT, all STRONG (reference main.cpp:458-477, round 0 pass A)."""
    p = default_params()
    p.max_iterations = iters; p.num_images = S + 1
    p.depth_min, p.depth_max = float(depth_min), float(depth_max)
    p.use_APD = use_apd; p.state = FIRST_INIT; p.geom_consistency = 0; p.weak_peak_radius = 6
    return p


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name))
    return g


def golden_inputs(g):
    return dict(images=g["images"], cameras=g["cameras"].view(synth.CAMERA_DTYPE), planes=g["planes_init"],
                edge=g["edge"], label=g["label"], seed=int(g["seed"]))


def close(a, b, rtol=1e-4, atol=0.0):
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype.kind != "f":
        return a == b
    with np.errstate(invalid="ignore"):
        return (np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol) | (np.isnan(a) & np.isnan(b)) | (a == b)


def per_pixel(ok, lead=2):
    return ok.reshape(ok.shape[:lead] + (-1,)).all(-1)
