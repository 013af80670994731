"""TEST INFRASTRUCTURE ONLY — opens oracle/_ref/libref_host.so: HOST functions of the reference compiled from line ranges
of its own APD.cpp (oracle/ref_host.cu, oracle/Makefile `host`): Roberts / Connect / Label_Seek / Label_Update and the
fusing loop of RunFusion.  They pin the CPU restatements of rows N1, N3 and N4.  `fast=True` opens the build with the
reference's own host flags (-O3 -ffast-math -march=native, CMakeLists.txt:31); it only runs on the machine it was built
on.  Import from tests/ only."""
import ctypes as C
import os

import numpy as np

from dvp_mvs_b200._lib import FusionView, make_fusion_view

HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path(fast: bool = False) -> str:
    return os.path.join(HERE, "_ref", "libref_host_fast.so" if fast else "libref_host.so")


def available(fast: bool = False) -> bool:
    return os.path.exists(lib_path(fast))


def _lib(fast=False):
    lib = C.CDLL(lib_path(fast))
    lib.refhost_flags.restype = C.c_char_p
    lib.refhost_restore_visibility.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.refhost_connect_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.refhost_roberts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.refhost_run_fusion.restype = C.c_longlong
    lib.refhost_run_fusion.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    if hasattr(lib, "refhost_run_fusion_tat"):
        lib.refhost_run_fusion_tat.restype = C.c_longlong
        lib.refhost_run_fusion_tat.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
    return lib


def flags(fast=False) -> str:
    return _lib(fast).refhost_flags().decode()


def restore_visibility(selected: np.ndarray, S: int, scale: int) -> np.ndarray:
    sel = np.ascontiguousarray(selected, np.uint32)
    out = np.zeros_like(sel)
    H, W = sel.shape
    assert _lib().refhost_restore_visibility(sel.ctypes.data, out.ctypes.data, W, H, S, scale) == 0
    return out


def connect_update(image: np.ndarray):
    """Connect + Label_Update of a 0 / 255 image -> (labels [H, W] int32, counts per label)."""
    img = np.ascontiguousarray(image, np.uint8)
    H, W = img.shape
    labels = np.zeros((H, W), np.int32); counts = np.zeros(H * W + 2, np.int32); n = C.c_int()
    assert _lib().refhost_connect_update(img.ctypes.data, W, H, labels.ctypes.data, counts.ctypes.data, counts.size, C.byref(n)) == 0
    return labels, counts[:n.value].copy()


def roberts(image: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(image, np.uint8)
    out = np.empty_like(img)
    assert _lib().refhost_roberts(img.ctypes.data, img.shape[1], img.shape[0], out.ctypes.data) == 0
    return out


def run_fusion(views: list, fast: bool = False):
    """The reference's own fusing loop over `views` (dicts as dvp_mvs_b200.Fusion takes them) -> (points [n, 6], masks)."""
    keep = []
    arr = (FusionView * len(views))(*[make_fusion_view(v, keep) for v in views])
    shapes = [(fv.height, fv.width) for fv in arr]
    cap = sum(h * w for h, w in shapes)
    pts = np.empty((cap, 6), np.float32)
    masks = np.zeros(cap, np.uint8)
    n = _lib(fast).refhost_run_fusion(len(views), arr, pts.ctypes.data, cap, masks.ctypes.data)
    assert n >= 0
    out_masks, off = [], 0
    for h, w in shapes:
        out_masks.append(masks[off:off + h * w].reshape(h, w).copy()); off += h * w
    return pts[:n].copy(), out_masks


def run_fusion_tat(views: list, mode: int, fast: bool = False):
    """The reference's own RunFusion_TAT_Intermediate (mode 1) / RunFusion_TAT_advanced (mode 2) loop -> (points [n, 6], masks)."""
    keep = []
    arr = (FusionView * len(views))(*[make_fusion_view(v, keep) for v in views])
    shapes = [(fv.height, fv.width) for fv in arr]
    cap = sum(h * w for h, w in shapes)
    pts = np.empty((cap, 6), np.float32)
    masks = np.zeros(cap, np.uint8)
    n = _lib(fast).refhost_run_fusion_tat(mode, len(views), arr, pts.ctypes.data, cap, masks.ctypes.data)
    assert n >= 0
    out_masks, off = [], 0
    for h, w in shapes:
        out_masks.append(masks[off:off + h * w].reshape(h, w).copy()); off += h * w
    return pts[:n].copy(), out_masks


def rescale(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """The reference's own RescaleMatToTargetSize<TYPE> (APD.cpp:1773-1796) on a uint8 / float32 / float32x3 / uint32 / int32 map."""
    a = np.ascontiguousarray(src)
    kind = {("uint8", 2): 0, ("float32", 2): 1, ("float32", 3): 2, ("uint32", 2): 3, ("int32", 2): 4}[(a.dtype.name, a.ndim)]
    assert kind != 2 or a.shape[2] == 3
    out = np.zeros((dh, dw) + a.shape[2:], a.dtype)
    lib = _lib()
    lib.refhost_rescale.restype = C.c_int
    lib.refhost_rescale.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    assert lib.refhost_rescale(kind, a.ctypes.data, a.shape[1], a.shape[0], out.ctypes.data, dw, dh) == 0
    return out


# ---- the reference's own file readers / writers (APD.cpp:548-692, main.cpp:127-170) ---------------------------------
_KINDS = {("uint8", 2): 0, ("float32", 2): 1, ("float32", 3): 2, ("uint32", 2): 3, ("int32", 2): 4}


def write_binmat(path: str, a: np.ndarray):
    a = np.ascontiguousarray(a)
    lib = _lib()
    lib.refhost_write_binmat.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    assert lib.refhost_write_binmat(os.fsencode(path), _KINDS[(a.dtype.name, a.ndim)], a.ctypes.data, a.shape[1], a.shape[0]) == 0


def read_binmat(path: str):
    """-> (rows, cols, OpenCV type code, raw bytes) as the reference's ReadBinMat sees the file; None when it refuses it."""
    lib = _lib()
    lib.refhost_read_binmat.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]
    r, c, t = C.c_int(), C.c_int(), C.c_int()
    if lib.refhost_read_binmat(os.fsencode(path), C.byref(r), C.byref(c), C.byref(t), None, 0) != 0:
        return None
    elem = {0: 1, 16: 3, 21: 12}.get(t.value, 4)
    buf = np.empty(r.value * c.value * elem, np.uint8)
    assert lib.refhost_read_binmat(os.fsencode(path), C.byref(r), C.byref(c), C.byref(t), buf.ctypes.data, buf.nbytes) == 0
    return r.value, c.value, t.value, buf


def write_dmb(path: str, a: np.ndarray):
    a = np.ascontiguousarray(a, np.float32)
    lib = _lib()
    lib.refhost_write_dmb.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    assert lib.refhost_write_dmb(os.fsencode(path), 1 if a.ndim == 2 else 3, a.ctypes.data, a.shape[1], a.shape[0]) == 0


def read_camera(path: str):
    from dvp_mvs_b200.synth import CAMERA_DTYPE
    cam = np.zeros(1, CAMERA_DTYPE)
    lib = _lib()
    lib.refhost_read_camera.argtypes = [C.c_char_p, C.c_void_p]
    assert lib.refhost_read_camera(os.fsencode(path), cam.ctypes.data) == 0
    return cam[0]


def read_pairs(dense_folder: str, cap_src: int = 64):
    lib = _lib()
    lib.refhost_read_pairs.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    n = lib.refhost_read_pairs(os.fsencode(dense_folder), None, None, None, 0, 0)
    ref = np.zeros(max(n, 1), np.int32); num = np.zeros(max(n, 1), np.int32); src = np.zeros((max(n, 1), cap_src), np.int32)
    assert lib.refhost_read_pairs(os.fsencode(dense_folder), ref.ctypes.data, num.ctypes.data, src.ctypes.data, n, cap_src) == n
    return [(int(ref[i]), [int(v) for v in src[i, :num[i]]]) for i in range(n)]


def level_camera(cam_full, full_w: int, full_h: int, scale_size: int):
    """The reference's own level-size / camera-rescale block (APD.cpp:1119-1143) -> (camera at the level, w, h)."""
    from dvp_mvs_b200.synth import CAMERA_DTYPE
    src = np.array(cam_full, dtype=CAMERA_DTYPE, copy=True).reshape(1)
    out = np.zeros(1, CAMERA_DTYPE)
    w, h = C.c_int(), C.c_int()
    lib = _lib()
    lib.refhost_level_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.refhost_level_camera(src.ctypes.data, full_w, full_h, scale_size, out.ctypes.data, C.byref(w), C.byref(h)) == 0
    return out[0], w.value, h.value


SCHEDULE_FIELDS = ("max_iterations", "num_images", "sigma_spatial", "sigma_color", "top_k", "depth_min", "depth_max", "geom_consistency",
                   "strong_radius", "strong_increment", "weak_radius", "weak_increment", "use_APD", "use_edge", "use_limit", "use_label",
                   "use_detail", "use_radius", "weak_peak_radius", "rotate_time", "ransac_threshold", "geom_factor", "state")
_FLOAT_FIELDS = {"sigma_spatial", "sigma_color", "depth_min", "depth_max", "ransac_threshold", "geom_factor"}


def schedule(round_num: int, num_problems: int):
    """main()'s own loop (main.cpp:450-512) with ProcessProblem recorded -> [dict(view, iteration, scale_size, edges, params{...})]."""
    lib = _lib()
    lib.refhost_schedule.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
    cap = (round_num - 1) * 4 * num_problems
    out = np.zeros((cap, 27), np.int32)
    n = lib.refhost_schedule(round_num, num_problems, out.ctypes.data, cap)
    assert n == cap, (n, cap)
    recs = []
    for row in out:
        prm = {}
        for k, name in enumerate(SCHEDULE_FIELDS):
            v = row[4 + k]
            prm[name] = float(np.array([v], np.int32).view(np.float32)[0]) if name in _FLOAT_FIELDS else int(v)
        recs.append(dict(view=int(row[0]), iteration=int(row[1]), scale_size=int(row[2]), edges=int(row[3]), params=prm))
    return recs


def post_pass(planes: np.ndarray, states: np.ndarray, selected: np.ndarray, num_src: int, scale_size: int, depth_min: float, depth_max: float):
    """ProcessProblem's own post-pass (main.cpp:282-363) -> (depth map, pixel states, selected views)."""
    pl = np.ascontiguousarray(planes, np.float32); st = np.ascontiguousarray(states, np.uint8); se = np.ascontiguousarray(selected, np.uint32)
    h, w = st.shape
    depth = np.empty((h, w), np.float32); st_out = np.empty((h, w), np.uint8); se_out = np.empty((h, w), np.uint32)
    lib = _lib()
    lib.refhost_post_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    assert lib.refhost_post_pass(pl.ctypes.data, st.ctypes.data, se.ctypes.data, w, h, num_src, scale_size, depth_min, depth_max,
                                 depth.ctypes.data, st_out.ctypes.data, se_out.ctypes.data) == 0
    return depth, st_out, se_out


def edge_segment(image: np.ndarray, scale: int, mode: int):
    """The reference's own EdgeSegment (APD.cpp:348-499; its OpenCV calls forwarded to the restated primitives):
    mode 0 (use_canny) -> edge map u8 [rows, cols]; mode 1 -> label map int32 at the level size."""
    img = np.ascontiguousarray(image, np.uint8)
    H, W = img.shape
    lib = _lib()
    lib.refhost_edge_segment.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    buf = np.zeros(H * W * 4 + 64, np.uint8)
    oc, orows = C.c_int(), C.c_int()
    assert lib.refhost_edge_segment(scale, img.ctypes.data, W, H, mode, 1 if mode == 0 else 0, buf.ctypes.data, C.byref(oc), C.byref(orows)) == 0
    n = oc.value * orows.value
    if mode == 0:
        return buf[:n].reshape(orows.value, oc.value).copy()
    return buf[:4 * n].view(np.int32).reshape(orows.value, oc.value).copy()


def assemble(dense_folder: str, ref_id: int, src_ids, scale_size: int, width: int, height: int, state: int, geom: int, use_apd: int,
             use_radius: int = 1, strong_radius: int = 5):
    """InuputInitialization's and SupportInitialization's own lines (APD.cpp:1147-1205, 1426-1493, 1615-1668) over the files of
    `dense_folder` -> dict(depths, weak, planes, selected, radius, edge_size, weak_count)."""
    S = len(src_ids)
    src = (C.c_int * max(S, 1))(*src_ids)
    flags = (C.c_int * 8)(state, geom, use_apd, 1, 1, 0, use_radius, strong_radius)   # use_label off: it reads another tool's MVS4/ depth maps
    depths = np.zeros((S + 1, height, width), np.float32); weak = np.zeros((height, width), np.uint8)
    planes = np.zeros((height, width, 4), np.float32); sel = np.zeros((height, width), np.uint32); rad = np.zeros((height, width), np.int32)
    edge = np.zeros((height, width), np.uint8); ewh = (C.c_int * 2)(); wc = C.c_int()
    lib = _lib()
    lib.refhost_assemble.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8
    rc = lib.refhost_assemble(os.fsencode(dense_folder), ref_id, S, src, scale_size, flags, width, height, depths.ctypes.data, weak.ctypes.data,
                              planes.ctypes.data, sel.ctypes.data, rad.ctypes.data, edge.ctypes.data, ewh, C.byref(wc))
    assert rc == 0, rc
    return dict(depths=depths, weak=weak, planes=planes, selected=sel, radius=rad, edge=edge, edge_size=(ewh[0], ewh[1]), weak_count=wc.value)


def get_problem_edges(dense_folder: str, ref_id: int, scale_size: int, image_u8: np.ndarray):
    """The reference's own GetProblemEdges (main.cpp:193-246) on `image_u8` (what cv::imread(.., IMREAD_GRAYSCALE) would return):
    writes APD/<id>/edges_<scale>.dmb and labels_<scale>.dmb under dense_folder."""
    img = np.ascontiguousarray(image_u8, np.uint8)
    lib = _lib()
    lib.refhost_get_problem_edges.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    assert lib.refhost_get_problem_edges(os.fsencode(dense_folder), ref_id, scale_size, img.ctypes.data, img.shape[1], img.shape[0]) == 0


def compute_round_num(n: int, cols: int, rows: int) -> int:
    lib = _lib()
    lib.refhost_compute_round_num.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    return int(lib.refhost_compute_round_num(b".", n, cols, rows))
