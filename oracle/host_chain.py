"""TEST INFRASTRUCTURE ONLY — restatement of the reference's host-side pass chaining:
the schedule of main() (main.cpp:449-511), the file round trips between ProcessProblem calls
(main.cpp:297-376; InuputInitialization APD.cpp:1045-1205, 1424-1456; SupportInitialization APD.cpp:1615-1668)
with the files replaced by numpy arrays, and RescaleMatToTargetSize (APD.cpp:1773-1796).
RunPatchMatch itself is delegated to an `Engine` (product, reference or CPU restatement) given by the caller;
the visibility restoration is the CPU restatement (oracle/cpu/visibility_cpu.cpp).
Import from tests/ only."""
import ctypes as C
import os

import numpy as np

from dvp_mvs_b200._lib import Params, default_params, FIRST_INIT, REFINE_INIT, REFINE_ITER, UNKNOWN, STRONG
from dvp_mvs_b200.synth import CAMERA_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
CPU_LIB = os.path.join(HERE, "_ref", "libapd_cpu.so")


def rescale_ref(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """RescaleMatToTargetSize, APD.cpp:1773-1796: o_r = int(r / scale_x), o_c = int(c / scale_y) — the two scale
    factors are swapped; float32 arithmetic; unreachable targets are zero (uninitialised there)."""
    sh, sw = src.shape[:2]
    if sw == dw and sh == dh:
        return src.copy()
    scale_x = np.float32(dw) / np.float32(sw)
    scale_y = np.float32(dh) / np.float32(sh)
    o_r = (np.arange(dh, dtype=np.float32) / scale_x).astype(np.int64)   # truncation of non-negative values
    o_c = (np.arange(dw, dtype=np.float32) / scale_y).astype(np.int64)
    ok_r, ok_c = o_r < sh, o_c < sw
    out = np.zeros((dh, dw) + src.shape[2:], src.dtype)
    rr, cc = np.nonzero(ok_r)[0], np.nonzero(ok_c)[0]
    out[np.ix_(rr, cc)] = src[np.ix_(o_r[rr], o_c[cc])]
    return out


def level_scale(num_levels: int, level: int) -> int:
    """main.cpp:455: round i runs at scale 2^(round_num - 1 - i), round_num = num_levels + 1."""
    return 1 << (num_levels - level)


def level_size(full_w: int, full_h: int, scale: int):
    """APD.cpp:1119-1123."""
    f = np.float32(1.0) / np.float32(scale)
    rnd = lambda x: int(np.floor(x + np.float32(0.5)))
    return rnd(np.float32(full_w) * f), rnd(np.float32(full_h) * f)


def level_camera(cam_full, full_w, full_h, w, h, scale):
    """APD.cpp:1125-1140."""
    cam = np.array(cam_full, dtype=CAMERA_DTYPE, copy=True).reshape(())
    if scale != 1:
        sx = np.float32(w) / np.float32(full_w); sy = np.float32(h) / np.float32(full_h)
        K = cam["K"].copy(); K[0] *= sx; K[2] *= sx; K[4] *= sy; K[5] *= sy
        cam["K"] = K
    cam["width"] = w; cam["height"] = h
    return cam


def schedule_params(num_levels: int, level: int, pass_: int, max_iterations: int = 3) -> Params:
    """main.cpp:452-505.  pass 0 = FIRST_INIT / REFINE_INIT, passes 1..3 = REFINE_ITER (j = pass - 1)."""
    p = default_params()
    i, round_num = level, num_levels + 1
    p.max_iterations = max_iterations
    if pass_ == 0:
        p.state, p.use_APD = (FIRST_INIT, 0) if i == 0 else (REFINE_INIT, 1)
        if i > 0:
            p.ransac_threshold = 0.01 - i * 0.00125
            p.rotate_time = min(int(2 ** i), 4)
            p.use_detail = 1 if i < round_num - 1 else 0
        p.geom_consistency = 0
        p.weak_peak_radius = 6
    else:
        j = pass_ - 1
        p.state = REFINE_ITER
        p.use_APD = 0 if i == 0 else 1
        p.ransac_threshold = 0.01 - i * 0.00125
        p.rotate_time = min(int(2 ** i), 4)
        p.geom_consistency = 1
        p.weak_peak_radius = max(4 - 2 * j, 2)
        if i > 0:
            p.use_detail = 1   # problem.params persists: set by this round's REFINE_INIT pass (main.cpp:468-473)
    return p


def restore_visibility(selected, S, scale):
    lib = C.CDLL(CPU_LIB)
    fn = lib.cpu_restore_visibility
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    sel = np.ascontiguousarray(selected, np.uint32)
    out = np.zeros_like(sel)
    H, W = sel.shape
    assert fn(sel.ctypes.data, out.ctypes.data, W, H, S, scale, 0) == 0
    return out


def assemble_inputs(own: dict, sources: list, w: int, h: int, p) -> dict:
    """What InuputInitialization / SupportInitialization read back from the previous pass's files and how they bring it to
    this pass's size (APD.cpp:1147-1205, 1426-1456, 1647-1667): `own` and `sources` are the per-view "files" (planes =
    APD_normals.dmb + depths.dmb, weak = weak.bin, selected = selected_views.bin, radius = radius.bin).  None = not read.
    (The radius map's UNKNOWN -> strong_radius rule, APD.cpp:1663-1667, is applied by the engine's upload.)
    Pinned to the reference's own lines in tests/test_ref_host.py::test_input_assembly_equals_the_reference_s_own_lines."""
    depths = None
    if p.geom_consistency:
        depths = np.stack([rescale_ref(f["planes"][..., 3], w, h) for f in [own] + list(sources)])
    planes = rescale_ref(own["planes"], w, h)
    if p.state == FIRST_INIT:
        selected, weak, radius = None, None, None
    else:
        selected = rescale_ref(own["selected"], w, h)
        weak = rescale_ref(own["weak"], w, h) if p.use_APD else None
        radius = rescale_ref(own["radius"], w, h) if p.use_radius else None
    return dict(depths=depths, planes=planes, selected=selected, weak=weak, radius=radius)


class HostChain:
    """The reference's schedule over a dvp_mvs_b200.synth.MultiView, one Engine per (size, S); the per-view "files"
    (depths.dmb, APD_normals.dmb, weak.bin, selected_views.bin, radius.bin) are numpy arrays in `self.files`."""

    def __init__(self, mv, make_engine, max_iterations=3):
        self.mv, self.make_engine, self.max_iterations = mv, make_engine, max_iterations
        self.V = len(mv.cameras)
        self.files = [dict(planes=mv.planes_init[v].copy()) for v in range(self.V)]   # FIRST_INIT prior (dep/ + sfm/)
        self.engines = {}

    def _engine(self, w, h, S, p):
        key = (w, h, S)
        if key not in self.engines:
            self.engines[key] = self.make_engine(w, h, S, p)
        return self.engines[key]

    def process_problem(self, v, level, pass_, seed):
        mv = self.mv
        L = mv.levels[level][v]
        w, h = L["w"], L["h"]
        src = mv.src_views[v]
        S = len(src)
        scale = level_scale(mv.num_levels, level)
        p = schedule_params(mv.num_levels, level, pass_, self.max_iterations)
        p.num_images = S + 1
        p.depth_min = float(np.float32(mv.cameras[v]["depth_min"]) * np.float32(0.6))   # APD.cpp:1109-1110, float arithmetic
        p.depth_max = float(np.float32(mv.cameras[v]["depth_max"]) * np.float32(1.2))
        f = self.files[v]
        # InuputInitialization
        images = np.stack([L["image"]] + [mv.levels[level][s]["image"] for s in src])
        cams = np.zeros(S + 1, CAMERA_DTYPE)
        for k, vv in enumerate([v] + src):
            cams[k] = level_camera(mv.cameras[vv], mv.full_w, mv.full_h, w, h, scale)
        a = assemble_inputs(f, [self.files[vv] for vv in src], w, h, p)
        depths, planes, selected, weak, radius = a["depths"], a["planes"], a["selected"], a["weak"], a["radius"]
        e = self._engine(w, h, S, p)
        e.upload(images=images, depths=depths, cameras=cams, planes=planes, selected_views=selected, weak_info=weak,
                 edge=L["edge"], label=L["label"], radius=radius, seed=seed, params=p)
        # the reference engine replays RunPatchMatch's launch sequence with K2 from the build that runs on sm_100a (mode 0):
        # its own APD::RunPatchMatch() (mode 1) faults in the miscompiled GenEdgeInform (profiles/r02_reference_k2_miscompile.md)
        e.run(**({"mode": 0} if getattr(e, "prefix", "") == "ref_" else {}))
        planes, weak, sel, rad = e.download()
        # ProcessProblem, main.cpp:297-363
        bad = (planes[..., 3] < p.depth_min) | (planes[..., 3] > p.depth_max)
        planes[..., 3][bad] = 0.0
        weak[bad] = UNKNOWN
        sel = restore_visibility(sel, S, scale)
        self.files[v] = dict(planes=planes, weak=weak, selected=sel, radius=rad)

    def run_pass(self, level, pass_, seed):
        for v in range(self.V):
            self.process_problem(v, level, pass_, seed + v)

    def run(self, seed):
        it = 0
        for level in range(self.mv.num_levels):
            for pass_ in range(4):
                self.run_pass(level, pass_, seed + 1000 * it)
                it += 1
