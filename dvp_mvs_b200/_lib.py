"""ctypes binding of the flat C ABI in include/dvp_mvs.h.

`Engine` drives a shared library that exports that ABI under a prefix.  By default that is the product,
dvp_mvs_b200/libdvp_mvs.so (prefix "dvp_", hand-written sm_100a CUDA).  There is no CPU fallback: if the
CUDA library is missing, constructing an Engine raises.  The test-only checkers under oracle/ export the
same ABI under other prefixes ("ref_", "cpu_") and are opened through oracle/ref_oracle.py and
oracle/cpu_oracle.py — from tests/, bench.py and smoke() only — by passing `lib_path`/`prefix` explicitly;
nothing in this package refers to them.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from .synth import CAMERA_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
# DVP_MVS_LIB selects another build of the same CUDA library (tuning variants, see csrc/Makefile); never a CPU path
PRODUCT_LIB = os.environ.get("DVP_MVS_LIB") or os.path.join(_HERE, "libdvp_mvs.so")

FIRST_INIT, REFINE_INIT, REFINE_ITER = 0, 1, 2
WEAK, STRONG, UNKNOWN = 0, 1, 2


class Params(C.Structure):
    """Mirror of dvp_params == reference PatchMatchParams (main.h:86-112)."""
    _fields_ = [
        ("max_iterations", C.c_int32), ("num_images", C.c_int32), ("sigma_spatial", C.c_float),
        ("sigma_color", C.c_float), ("top_k", C.c_int32), ("depth_min", C.c_float), ("depth_max", C.c_float),
        ("geom_consistency", C.c_int32), ("strong_radius", C.c_int32), ("strong_increment", C.c_int32),
        ("weak_radius", C.c_int32), ("weak_increment", C.c_int32), ("use_APD", C.c_int32),
        ("use_edge", C.c_int32), ("use_limit", C.c_int32), ("use_label", C.c_int32), ("use_detail", C.c_int32),
        ("use_radius", C.c_int32), ("weak_peak_radius", C.c_int32), ("rotate_time", C.c_int32),
        ("ransac_threshold", C.c_float), ("geom_factor", C.c_float), ("state", C.c_int32),
    ]

    def copy(self) -> "Params":
        p = Params()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(Params))
        return p


class Inputs(C.Structure):
    _fields_ = [
        ("images", C.c_void_p), ("depths", C.c_void_p), ("cameras", C.c_void_p), ("planes", C.c_void_p),
        ("selected_views", C.c_void_p), ("weak_info", C.c_void_p), ("edge", C.c_void_p), ("label", C.c_void_p),
        ("radius", C.c_void_p), ("seed", C.c_uint64),
    ]


BUF = dict(planes=0, costs=1, selected=2, weak=3, radius=4, view_weight=5, rand=6, fit_planes=7, edge_neigh=8,
           candidate=9, nearest_strong=10, weak_reliable=11, neighbours_map=12, neighbours=13, label_boundary=14,
           complex=15)
BUF_DTYPE = dict(planes=np.float32, costs=np.float32, selected=np.uint32, weak=np.uint8, radius=np.int32,
                 view_weight=np.uint8, rand=np.uint32, fit_planes=np.float32, edge_neigh=np.int16,
                 candidate=np.int16, nearest_strong=np.int16, weak_reliable=np.uint8, neighbours_map=np.int32,
                 neighbours=np.int16, label_boundary=np.int16, complex=np.float32)
STAGES = ["K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM", "K3_FIND_NEAREST_STRONG", "K4_GEN_NEIGHBOURS",
          "K5_NEIGHBOUR_UPDATE", "K6_RANDOM_INITIALIZATION", "K7_BLACK_STRONG", "K8_RED_STRONG",
          "K9_RANSAC_FIT_PLANE", "K10_BLACK_WEAK", "K11_RED_WEAK", "K12_DEPTH_NORMAL", "K13_BLACK_FILTER",
          "K14_RED_FILTER", "K15_DEPTH_TO_WEAK", "K16_LOCAL_REFINE"]
STAGE = {n: i for i, n in enumerate(STAGES)}
STAGE["K15_K16_FUSED"] = 16   # product only: K15 and K16 in one launch, as dvp_run issues them
STATUS = {0: "DVP_OK", -1: "DVP_ERR_ARG", -2: "DVP_ERR_CUDA", -3: "DVP_ERR_STATE", -4: "DVP_ERR_UNSUPPORTED"}

ABI_SYMBOLS = ["version", "default_params", "create", "destroy", "upload", "run", "run_stage", "download",
               "buffer_bytes", "get_buffer", "set_buffer", "last_run_times", "weak_count", "last_cuda_error", "stream"]
PRODUCT_ONLY_SYMBOLS = ["upload_device", "upload_overlapped", "restore_visibility", "rescale_map", "scene_create", "scene_destroy", "scene_level_size",
                        "scene_pass_params", "scene_set_max_iterations", "scene_set_view", "scene_set_level",
                        "scene_set_initial_planes", "scene_set_image", "scene_set_label", "scene_get_label", "scene_get_image", "scene_run_pass", "scene_run", "scene_get_view", "scene_stats",
                        "scene_run_view", "scene_depth_map", "scene_remote_depth",
                        "fusion_create", "fusion_destroy", "fusion_set_view", "fusion_set_view_planes", "scene_fuse_views", "fusion_set_mode", "fusion_reset", "fusion_run_view", "fusion_run",
                        "fusion_num_points", "fusion_get_points", "fusion_get_mask", "fusion_last_view", "fusion_last_view_index", "fusion_write_ply",
                        "edge_segment", "scene_compute_edges", "scene_get_edges",
                        "farm_create", "farm_destroy", "farm_num_devices", "farm_owner", "farm_set_max_iterations", "farm_set_view", "farm_set_level", "farm_set_image",
                        "farm_compute_edges", "farm_set_initial_planes", "farm_run", "farm_exchange_bytes", "farm_get_view",
                        "debug_race_explain", "debug_fetch_count", "resize_linear_f32", "label_size", "label_segment",
                        "io_binmat_header", "io_read_binmat", "io_write_binmat", "io_write_dmb", "io_read_camera", "io_read_pairs"]


class FusionView(C.Structure):
    """Mirror of dvp_fusion_view (include/dvp_mvs.h, row N3): one view as RunFusion holds it after its loading loop
    (reference APD.cpp:1841-1873)."""
    _fields_ = [
        ("camera", C.c_uint8 * 112), ("width", C.c_int32), ("height", C.c_int32), ("depth", C.c_void_p),
        ("normal", C.c_void_p), ("image", C.c_void_p), ("weak", C.c_void_p), ("block", C.c_void_p),
        ("num_src", C.c_int32), ("src_views", C.c_void_p),
    ]


def make_fusion_view(view: dict, keep: list) -> FusionView:
    """dict(camera, depth [h,w] f32, normal [h,w,3] f32, image [h,w,3] u8, weak [h,w] u8, src_views, block=None) ->
    FusionView; the contiguous arrays it points to are appended to `keep` (the caller keeps them alive)."""
    depth = np.ascontiguousarray(view["depth"], np.float32)
    H, W = depth.shape
    normal = _carr(view["normal"], np.float32, (H, W, 3)); image = _carr(view["image"], np.uint8, (H, W, 3))
    weak = _carr(view.get("weak"), np.uint8, (H, W)); block = _carr(view.get("block"), np.uint8, (H, W))
    src = np.ascontiguousarray(view["src_views"], np.int32)
    cam = np.ascontiguousarray(np.asarray(view["camera"], CAMERA_DTYPE).reshape(1))
    keep += [depth, normal, image, weak, block, src, cam]
    fv = FusionView()
    C.memmove(fv.camera, cam.ctypes.data, 112)
    fv.width, fv.height = W, H
    fv.depth, fv.normal, fv.image = depth.ctypes.data, normal.ctypes.data, image.ctypes.data
    fv.weak = weak.ctypes.data if weak is not None else None
    fv.block = block.ctypes.data if block is not None else None
    fv.num_src = len(src); fv.src_views = src.ctypes.data if len(src) else None
    return fv


class DvpError(RuntimeError):
    pass


def load_library(path: str, prefix: str):
    if not os.path.exists(path):
        raise DvpError(
            f"{path} not found: the CUDA library must be built first "
            f"(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(path)
    f = lambda n: getattr(lib, prefix + n)
    f("version").restype = C.c_char_p
    f("default_params").argtypes = [C.POINTER(Params)]; f("default_params").restype = None
    f("create").argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Params)]; f("create").restype = C.c_void_p
    f("destroy").argtypes = [C.c_void_p]; f("destroy").restype = None
    f("upload").argtypes = [C.c_void_p, C.POINTER(Inputs), C.POINTER(Params)]; f("upload").restype = C.c_int
    f("run").argtypes = [C.c_void_p, C.c_int]; f("run").restype = C.c_int
    f("run_stage").argtypes = [C.c_void_p, C.c_int, C.c_int]; f("run_stage").restype = C.c_int
    f("download").argtypes = [C.c_void_p] + [C.c_void_p] * 4; f("download").restype = C.c_int
    f("buffer_bytes").argtypes = [C.c_void_p, C.c_int]; f("buffer_bytes").restype = C.c_size_t
    f("get_buffer").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]; f("get_buffer").restype = C.c_int
    f("set_buffer").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]; f("set_buffer").restype = C.c_int
    f("last_run_times").argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    f("last_run_times").restype = C.c_int
    f("weak_count").argtypes = [C.c_void_p]; f("weak_count").restype = C.c_int
    f("last_cuda_error").argtypes = [C.c_void_p]; f("last_cuda_error").restype = C.c_int
    f("stream").argtypes = [C.c_void_p]; f("stream").restype = C.c_void_p
    if prefix == "dvp_":
        f("upload_device").argtypes = [C.c_void_p, C.POINTER(Inputs), C.POINTER(Params)]; f("upload_device").restype = C.c_int
        f("upload_overlapped").argtypes = [C.c_void_p, C.POINTER(Inputs), C.POINTER(Params)]; f("upload_overlapped").restype = C.c_int
        f("restore_visibility").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]; f("restore_visibility").restype = C.c_int
        f("rescale_map").argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]; f("rescale_map").restype = C.c_int
        f("scene_create").argtypes = [C.c_int, C.c_int, C.c_int]; f("scene_create").restype = C.c_void_p
        f("scene_destroy").argtypes = [C.c_void_p]; f("scene_destroy").restype = None
        f("scene_level_size").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]; f("scene_level_size").restype = C.c_int
        f("scene_pass_params").argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(Params)]; f("scene_pass_params").restype = C.c_int
        f("scene_set_max_iterations").argtypes = [C.c_void_p, C.c_int]; f("scene_set_max_iterations").restype = C.c_int
        f("scene_set_view").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]; f("scene_set_view").restype = C.c_int
        f("scene_set_level").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]; f("scene_set_level").restype = C.c_int
        f("scene_set_image").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]; f("scene_set_image").restype = C.c_int
        f("scene_set_label").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]; f("scene_set_label").restype = C.c_int
        f("scene_get_label").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]; f("scene_get_label").restype = C.c_int
        f("scene_get_image").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]; f("scene_get_image").restype = C.c_int
        f("scene_set_initial_planes").argtypes = [C.c_void_p, C.c_int, C.c_void_p]; f("scene_set_initial_planes").restype = C.c_int
        f("scene_run_pass").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64]; f("scene_run_pass").restype = C.c_int
        f("scene_run").argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_float)]; f("scene_run").restype = C.c_int
        f("scene_run_view").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64]; f("scene_run_view").restype = C.c_int
        f("scene_depth_map").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]; f("scene_depth_map").restype = C.c_int
        f("scene_remote_depth").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]; f("scene_remote_depth").restype = C.c_int
        f("scene_get_view").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)] + [C.c_void_p] * 4; f("scene_get_view").restype = C.c_int
        f("scene_stats").argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]; f("scene_stats").restype = C.c_int
        f("edge_segment").argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)]; f("edge_segment").restype = C.c_int
        f("scene_compute_edges").argtypes = [C.c_void_p, C.c_int, C.c_int]; f("scene_compute_edges").restype = C.c_int
        f("scene_get_edges").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]; f("scene_get_edges").restype = C.c_int
        f("fusion_create").argtypes = [C.c_int, C.c_int]; f("fusion_create").restype = C.c_void_p
        f("fusion_destroy").argtypes = [C.c_void_p]; f("fusion_destroy").restype = None
        f("fusion_set_view").argtypes = [C.c_void_p, C.c_int, C.POINTER(FusionView)]; f("fusion_set_view").restype = C.c_int
        f("fusion_set_view_planes").argtypes = [C.c_void_p, C.c_int, C.POINTER(FusionView), C.c_void_p]; f("fusion_set_view_planes").restype = C.c_int
        f("scene_fuse_views").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]; f("scene_fuse_views").restype = C.c_int
        f("fusion_reset").argtypes = [C.c_void_p]; f("fusion_reset").restype = C.c_int
        f("fusion_set_mode").argtypes = [C.c_void_p, C.c_int]; f("fusion_set_mode").restype = C.c_int
        f("fusion_run_view").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]; f("fusion_run_view").restype = C.c_int
        f("fusion_run").argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_float)]; f("fusion_run").restype = C.c_int
        f("fusion_num_points").argtypes = [C.c_void_p]; f("fusion_num_points").restype = C.c_longlong
        f("fusion_get_points").argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong]; f("fusion_get_points").restype = C.c_int
        f("fusion_get_mask").argtypes = [C.c_void_p, C.c_int, C.c_void_p]; f("fusion_get_mask").restype = C.c_int
        f("fusion_last_view").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]; f("fusion_last_view").restype = C.c_int
        f("fusion_write_ply").argtypes = [C.c_void_p, C.c_char_p]; f("fusion_write_ply").restype = C.c_int
        f("fusion_last_view_index").argtypes = [C.c_void_p]; f("fusion_last_view_index").restype = C.c_int
        f("label_size").argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]; f("label_size").restype = C.c_int
        f("label_segment").argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]; f("label_segment").restype = C.c_int
        f("resize_linear_f32").argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]; f("resize_linear_f32").restype = C.c_int
        f("farm_create").argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]; f("farm_create").restype = C.c_void_p
        f("farm_destroy").argtypes = [C.c_void_p]; f("farm_destroy").restype = None
        f("farm_num_devices").argtypes = [C.c_void_p]; f("farm_num_devices").restype = C.c_int
        f("farm_owner").argtypes = [C.c_void_p, C.c_int]; f("farm_owner").restype = C.c_int
        f("farm_set_max_iterations").argtypes = [C.c_void_p, C.c_int]; f("farm_set_max_iterations").restype = C.c_int
        f("farm_set_view").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]; f("farm_set_view").restype = C.c_int
        f("farm_set_level").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]; f("farm_set_level").restype = C.c_int
        f("farm_set_image").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]; f("farm_set_image").restype = C.c_int
        f("farm_compute_edges").argtypes = [C.c_void_p, C.c_int, C.c_int]; f("farm_compute_edges").restype = C.c_int
        f("farm_set_initial_planes").argtypes = [C.c_void_p, C.c_int, C.c_void_p]; f("farm_set_initial_planes").restype = C.c_int
        f("farm_run").argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_float), C.POINTER(C.c_float)]; f("farm_run").restype = C.c_int
        f("farm_exchange_bytes").argtypes = [C.c_void_p]; f("farm_exchange_bytes").restype = C.c_longlong
        f("farm_get_view").argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)] + [C.c_void_p] * 4; f("farm_get_view").restype = C.c_int
        f("debug_race_explain").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_void_p]
        f("debug_race_explain").restype = C.c_int
        f("debug_fetch_count").argtypes = [C.c_void_p, C.c_int]; f("debug_fetch_count").restype = C.c_longlong
    return lib


def default_params(lib=None, prefix="dvp_") -> Params:
    p = Params()
    if lib is not None:
        getattr(lib, prefix + "default_params")(C.byref(p))
        return p
    # reference defaults (main.h:86-112), used when no library is loaded (CPU-only host logic/tests)
    p.max_iterations, p.num_images, p.sigma_spatial, p.sigma_color, p.top_k = 3, 5, 5.0, 3.0, 4
    p.depth_min, p.depth_max, p.geom_consistency = 0.0, 1.0, 0
    p.strong_radius, p.strong_increment, p.weak_radius, p.weak_increment = 5, 2, 5, 5
    p.use_APD, p.use_edge, p.use_limit, p.use_label, p.use_detail, p.use_radius = 1, 1, 1, 1, 0, 1
    p.weak_peak_radius, p.rotate_time, p.ransac_threshold, p.geom_factor, p.state = 2, 4, 0.005, 0.2, FIRST_INIT
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _carr(a, dtype, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class Engine:
    """One PatchMatch context (one reference view on one GPU)."""

    def __init__(self, width: int, height: int, num_src: int, params: Params, device: int = 0,
                 lib_path: str | None = None, prefix: str = "dvp_"):
        self.prefix = prefix
        self.impl = {"dvp_": "product", "ref_": "reference", "cpu_": "cpu"}.get(prefix, prefix)
        self.lib = load_library(lib_path or PRODUCT_LIB, self.prefix)
        self.W, self.H, self.S, self.N = width, height, num_src, width * height
        self.params = params.copy()
        self.params.num_images = num_src + 1
        self._f = lambda n: getattr(self.lib, self.prefix + n)
        self.ctx = self._f("create")(device, width, height, num_src, C.byref(self.params))
        if not self.ctx:
            raise DvpError(f"{self.prefix}create failed (device {device}, {width}x{height}, S={num_src})")
        self._keep = None

    def version(self) -> str:
        return self._f("version")().decode()

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise DvpError(f"{self.prefix}{what} -> {STATUS.get(rc, rc)} (cudaError {self._f('last_cuda_error')(self.ctx)})")

    def close(self):
        if getattr(self, "ctx", None):
            self._f("destroy")(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, images, cameras, planes, depths=None, selected_views=None, weak_info=None, edge=None,
               label=None, radius=None, seed=0x5EED, params: Params | None = None, overlapped: bool = False):
        S, H, W = self.S, self.H, self.W
        new_params = self.params
        if params is not None:
            new_params = params.copy()
            new_params.num_images = S + 1
        keep = dict(
            images=_carr(images, np.float32, (S + 1, H, W)),
            depths=_carr(depths, np.float32, (S + 1, H, W)),
            cameras=_carr(cameras, CAMERA_DTYPE, (S + 1,)),
            planes=_carr(planes, np.float32, (H, W, 4)),
            selected_views=_carr(selected_views, np.uint32, (H, W)),
            weak_info=_carr(weak_info, np.uint8, (H, W)),
            edge=_carr(edge, np.uint8, (H, W)),
            label=_carr(label, np.int32, (H, W)),
            radius=_carr(radius, np.int32, (H, W)),
        )
        inp = Inputs(*[_ptr(keep[k]) for k in ("images", "depths", "cameras", "planes", "selected_views",
                                                  "weak_info", "edge", "label", "radius")], int(seed))
        name = "upload_overlapped" if overlapped else "upload"   # overlapped: the arrays stay referenced in self._keep
        self._check(self._f(name)(self.ctx, C.byref(inp), C.byref(new_params)), name)
        self._keep = keep
        self.params = new_params

    def upload_raw(self, inp: Inputs, device: bool = False, overlapped: bool = False):
        name = "upload_device" if device else ("upload_overlapped" if overlapped else "upload")
        self._check(self._f(name)(self.ctx, C.byref(inp), C.byref(self.params)), name)

    def run(self, sync: bool = True, mode: int | None = None):
        arg = (1 if sync else 0) if mode is None else mode
        self._check(self._f("run")(self.ctx, arg), "run")

    def run_stage(self, stage, iteration: int = 0):
        sid = STAGE[stage] if isinstance(stage, str) else int(stage)
        self._check(self._f("run_stage")(self.ctx, sid, iteration), f"run_stage({stage})")

    def last_run_times(self):
        total = C.c_float(); per = (C.c_float * 16)(); n = C.c_int()
        self._check(self._f("last_run_times")(self.ctx, C.byref(total), per, C.byref(n)), "last_run_times")
        return float(total.value), [float(x) for x in per], int(n.value)

    def restore_visibility(self, scale_size: int) -> float:
        """Post-pass of ProcessProblem on the resident maps (main.cpp:297-363); returns the device time in ms."""
        ms = C.c_float()
        self._check(self._f("restore_visibility")(self.ctx, int(scale_size), C.byref(ms)), "restore_visibility")
        return float(ms.value)

    def race_explain(self, iteration: int, red: int, offsets, planes_before, observed: dict, tear: bool = True):
        """Parity instrumentation (dvp_debug_race_explain, include/dvp_mvs.h): which pixels of an observed K7 / K8 result
        (dict of planes, costs, selected, view_weight, rand) does some outcome of the reference's direction-4 race reproduce?
        The context must hold the pre-launch state.  -> (explained [H, W] bool, (left after phase 1, left after phase 2, launches))."""
        H, W = self.H, self.W
        off = np.ascontiguousarray(offsets, np.int32)
        before = _carr(planes_before, np.float32, (H, W, 4))
        obs = dict(planes=_carr(observed["planes"], np.float32, (H, W, 4)), costs=_carr(observed["costs"], np.float32, (H, W)),
                   selected=_carr(observed["selected"], np.uint32, (H, W)), view_weight=_carr(observed["view_weight"], np.uint8, (H, W, 32)),
                   rand=_carr(observed["rand"], np.uint32, (H, W, 6)))
        explained = np.zeros((H, W), np.uint8); stats = (C.c_longlong * 3)()
        self._check(self._f("debug_race_explain")(self.ctx, iteration, red, _ptr(off), len(off), _ptr(before), _ptr(obs["planes"]), _ptr(obs["planes"]),
                                                   _ptr(obs["costs"]), _ptr(obs["selected"]), _ptr(obs["view_weight"]), _ptr(obs["rand"]),
                                                   1 if tear else 0, _ptr(explained), stats), "debug_race_explain")
        return explained.astype(bool), tuple(int(v) for v in stats)

    def fetch_count(self, reset: bool = True) -> int:
        """Texture fetches since the last reset (instrumented build only: DVP_MVS_LIB=.../libdvp_mvs_count.so)."""
        n = int(self._f("debug_fetch_count")(self.ctx, 1 if reset else 0))
        if n < 0:
            raise DvpError(f"debug_fetch_count -> {STATUS.get(n, n)} (this library is not the instrumented build)")
        return n

    def weak_count(self) -> int:
        return int(self._f("weak_count")(self.ctx))

    def get(self, name: str) -> np.ndarray:
        nbytes = int(self._f("buffer_bytes")(self.ctx, BUF[name]))
        dt = np.dtype(BUF_DTYPE[name])
        out = np.empty(nbytes // dt.itemsize, dtype=dt)
        if nbytes:
            self._check(self._f("get_buffer")(self.ctx, BUF[name], _ptr(out), nbytes), f"get_buffer({name})")
        return self._shape(name, out)

    def set(self, name: str, arr: np.ndarray):
        a = np.ascontiguousarray(arr, dtype=BUF_DTYPE[name])
        nbytes = int(self._f("buffer_bytes")(self.ctx, BUF[name]))
        if a.nbytes != nbytes:
            raise ValueError(f"{name}: expected {nbytes} bytes, got {a.nbytes}")
        if nbytes:
            self._check(self._f("set_buffer")(self.ctx, BUF[name], _ptr(a), nbytes), f"set_buffer({name})")

    def _shape(self, name, a):
        H, W = self.H, self.W
        shp = dict(planes=(H, W, 4), fit_planes=(H, W, 4), costs=(H, W), selected=(H, W), weak=(H, W), radius=(H, W),
                   view_weight=(H, W, 32), rand=(H, W, 6), edge_neigh=(H, W, 8, 2), candidate=(H, W, 4, 8, 2),
                   nearest_strong=(H, W, 2), weak_reliable=(H, W), neighbours_map=(H, W),
                   neighbours=(-1, 12, 2), label_boundary=(-1, 8, 2), complex=(-1,))[name]
        return a.reshape(shp)

    def download(self):
        H, W = self.H, self.W
        planes = np.empty((H, W, 4), np.float32); weak = np.empty((H, W), np.uint8)
        sel = np.empty((H, W), np.uint32); rad = np.empty((H, W), np.int32)
        self._check(self._f("download")(self.ctx, _ptr(planes), _ptr(weak), _ptr(sel), _ptr(rad)), "download")
        return planes, weak, sel, rad

    def snapshot(self, names=("planes", "costs", "selected", "weak", "radius", "view_weight", "rand", "fit_planes")):
        return {n: self.get(n) for n in names}

    def restore(self, snap: dict):
        for n, a in snap.items():
            self.set(n, a)


class Scene:
    """In-memory multi-scale scene driver (include/dvp_mvs.h, row N2): the schedule of the reference's main()
    (main.cpp:449-511) with every map resident on the GPU between passes."""

    def __init__(self, num_views: int, num_levels: int, device: int = 0):
        self.lib = load_library(PRODUCT_LIB, "dvp_")
        self.num_views, self.num_levels, self.device = num_views, num_levels, device
        self.h = self.lib.dvp_scene_create(device, num_views, num_levels)
        if not self.h:
            raise DvpError(f"dvp_scene_create failed (device {device}, {num_views} views, {num_levels} levels)")
        self._sizes = {}

    def _check(self, rc, what):
        if rc != 0:
            raise DvpError(f"dvp_scene_{what} -> {STATUS.get(rc, rc)}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.dvp_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def level_size(self, full_w: int, full_h: int, level: int):
        w, h = C.c_int(), C.c_int()
        self._check(self.lib.dvp_scene_level_size(self.h, full_w, full_h, level, C.byref(w), C.byref(h)), "level_size")
        return int(w.value), int(h.value)

    def pass_params(self, level: int, pass_: int) -> Params:
        p = Params()
        self._check(self.lib.dvp_scene_pass_params(self.h, level, pass_, C.byref(p)), "pass_params")
        return p

    def set_max_iterations(self, n: int):
        self._check(self.lib.dvp_scene_set_max_iterations(self.h, n), "set_max_iterations")

    def set_view(self, view: int, camera, full_w: int, full_h: int, src_views):
        self.src_views = getattr(self, "src_views", {}); self.src_views[view] = list(src_views)
        cam = np.ascontiguousarray(camera, dtype=CAMERA_DTYPE).reshape(1)
        src = (C.c_int * len(src_views))(*[int(s) for s in src_views])
        self._check(self.lib.dvp_scene_set_view(self.h, view, _ptr(cam), full_w, full_h, len(src_views), src), "set_view")
        self._sizes[view] = (full_w, full_h)

    def set_level(self, view: int, level: int, image, edge=None, label=None):
        w, h = self.level_size(*self._sizes[view], level)
        img = _carr(image, np.float32, (h, w)); e = _carr(edge, np.uint8, (h, w)); l = _carr(label, np.int32, (h, w))
        self._check(self.lib.dvp_scene_set_level(self.h, view, level, _ptr(img), _ptr(e), _ptr(l)), "set_level")

    def set_image(self, view: int, image_u8: np.ndarray, compute_edges: bool = True, compute_labels: bool = False):
        """The whole pyramid of one view from its full-resolution 8-bit grey image (dvp_scene_set_image), optionally with the
        edge and label maps GetProblemEdges derives from it."""
        W, H = self._sizes[view]
        img = _carr(image_u8, np.uint8, (H, W))
        self._check(self.lib.dvp_scene_set_image(self.h, view, _ptr(img), (1 if compute_edges else 0) | (2 if compute_labels else 0)), "set_image")

    def get_label(self, view: int, level: int) -> np.ndarray:
        w, h = self.view_level_size(view, level)
        out = np.empty((h, w), np.int32)
        self._check(self.lib.dvp_scene_get_label(self.h, view, level, _ptr(out)), "get_label")
        return out

    def set_label(self, view: int, level: int, label):
        w, h = self.view_level_size(view, level)
        lab = _carr(label, np.int32, (h, w))
        self._check(self.lib.dvp_scene_set_label(self.h, view, level, _ptr(lab)), "set_label")

    def get_image(self, view: int, level: int) -> np.ndarray:
        w, h = self.view_level_size(view, level)
        out = np.empty((h, w), np.float32)
        self._check(self.lib.dvp_scene_get_image(self.h, view, level, _ptr(out)), "get_image")
        return out

    def compute_edges(self, view: int, level: int) -> np.ndarray:
        """Row N4: the level's edge map computed on the device from the level image (it becomes the level's edge input)."""
        self._check(self.lib.dvp_scene_compute_edges(self.h, view, level), "compute_edges")
        w, h = self.level_size(*self._sizes[view], level)
        out = np.empty((h, w), np.uint8)
        self._check(self.lib.dvp_scene_get_edges(self.h, view, level, _ptr(out)), "get_edges")
        return out

    def set_initial_planes(self, view: int, planes):
        w, h = self.level_size(*self._sizes[view], 0)
        p = _carr(planes, np.float32, (h, w, 4))
        self._check(self.lib.dvp_scene_set_initial_planes(self.h, view, _ptr(p)), "set_initial_planes")

    def run_pass(self, level: int, pass_: int, seed: int):
        self._check(self.lib.dvp_scene_run_pass(self.h, level, pass_, int(seed)), "run_pass")

    def run(self, seed: int) -> float:
        ms = C.c_float()
        self._check(self.lib.dvp_scene_run(self.h, int(seed), C.byref(ms)), "run")
        return float(ms.value)

    def run_view(self, view: int, level: int, pass_: int, seed: int):
        self._check(self.lib.dvp_scene_run_view(self.h, view, level, pass_, int(seed)), "run_view")

    def view_level_size(self, view: int, level: int):
        return self.level_size(*self._sizes[view], level)

    def depth_tensor(self, view: int, level: int, owned: bool):
        """The view's depth map as a zero-copy torch CUDA tensor [h, w] f32: the buffer to broadcast from (owned) or to
        receive into (owned elsewhere; sized for `level`)."""
        import torch
        ptr, w, h = C.c_void_p(), C.c_int(), C.c_int()
        if owned:
            self._check(self.lib.dvp_scene_depth_map(self.h, view, C.byref(ptr), C.byref(w), C.byref(h)), "depth_map")
            W, H = int(w.value), int(h.value)
        else:
            W, H = self.view_level_size(view, level)
            self._check(self.lib.dvp_scene_remote_depth(self.h, view, W, H, C.byref(ptr)), "remote_depth")

        class _Buf:
            __cuda_array_interface__ = {"shape": (H, W), "typestr": "<f4", "data": (int(ptr.value), False), "version": 2}
        return torch.as_tensor(_Buf(), device=torch.device("cuda", self.device))

    def stats(self):
        ms, n = C.c_double(), C.c_longlong()
        self._check(self.lib.dvp_scene_stats(self.h, C.byref(ms), C.byref(n)), "stats")
        return float(ms.value), int(n.value)

    def get_view(self, view: int):
        w, h = C.c_int(), C.c_int()
        self._check(self.lib.dvp_scene_get_view(self.h, view, C.byref(w), C.byref(h), None, None, None, None), "get_view")
        W, H = int(w.value), int(h.value)
        planes = np.empty((H, W, 4), np.float32); weak = np.empty((H, W), np.uint8)
        sel = np.empty((H, W), np.uint32); rad = np.empty((H, W), np.int32)
        self._check(self.lib.dvp_scene_get_view(self.h, view, C.byref(w), C.byref(h), _ptr(planes), _ptr(weak), _ptr(sel), _ptr(rad)), "get_view")
        return planes, weak, sel, rad


class Farm:
    """The per-view farm inside the library (dvp_farm_*): one resident scene and one host thread per GPU, views
    dealt round robin, depth maps exchanged by peer copies.  Mirrors `Scene`; inputs are replicated on every GPU."""

    def __init__(self, devices, num_views: int, num_levels: int):
        self.lib = load_library(PRODUCT_LIB, "dvp_")
        self.devices = [int(d) for d in devices]
        self.num_views, self.num_levels = num_views, num_levels
        arr = (C.c_int * len(self.devices))(*self.devices)
        self.h = self.lib.dvp_farm_create(len(self.devices), arr, num_views, num_levels)
        if not self.h:
            raise DvpError(f"dvp_farm_create failed (devices {self.devices}, {num_views} views, {num_levels} levels)")

    def _check(self, rc, what):
        if rc != 0:
            raise DvpError(f"dvp_farm_{what} -> {STATUS.get(rc, rc)}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.dvp_farm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_max_iterations(self, n: int):
        self._check(self.lib.dvp_farm_set_max_iterations(self.h, n), "set_max_iterations")

    def set_view(self, view: int, camera, full_w: int, full_h: int, src_views):
        cam = np.ascontiguousarray(camera, dtype=CAMERA_DTYPE).reshape(1)
        src = (C.c_int * len(src_views))(*[int(s) for s in src_views])
        self._check(self.lib.dvp_farm_set_view(self.h, view, _ptr(cam), full_w, full_h, len(src_views), src), "set_view")

    def set_level(self, view: int, level: int, image, edge=None, label=None):
        img = np.ascontiguousarray(image, np.float32)
        e = None if edge is None else np.ascontiguousarray(edge, np.uint8)
        lab = None if label is None else np.ascontiguousarray(label, np.int32)
        self._check(self.lib.dvp_farm_set_level(self.h, view, level, _ptr(img), _ptr(e), _ptr(lab)), "set_level")

    def set_image(self, view: int, image_u8, compute_edges: bool = True, compute_labels: bool = False):
        img = np.ascontiguousarray(image_u8, np.uint8)
        self._check(self.lib.dvp_farm_set_image(self.h, view, _ptr(img), (1 if compute_edges else 0) | (2 if compute_labels else 0)), "set_image")

    def compute_edges(self, view: int, level: int):
        self._check(self.lib.dvp_farm_compute_edges(self.h, view, level), "compute_edges")

    def set_initial_planes(self, view: int, planes):
        pl = np.ascontiguousarray(planes, np.float32)
        self._check(self.lib.dvp_farm_set_initial_planes(self.h, view, _ptr(pl)), "set_initial_planes")

    def owner(self, view: int) -> int:
        return int(self.lib.dvp_farm_owner(self.h, view))

    def run(self, seed: int):
        """-> (wall ms, ms the slowest GPU thread spent in the exchange steps, bytes moved between GPUs)"""
        wall, exch = C.c_float(), C.c_float()
        self._check(self.lib.dvp_farm_run(self.h, int(seed), C.byref(wall), C.byref(exch)), "run")
        return float(wall.value), float(exch.value), int(self.lib.dvp_farm_exchange_bytes(self.h))

    def get_view(self, view: int):
        w, h = C.c_int(), C.c_int()
        self._check(self.lib.dvp_farm_get_view(self.h, view, C.byref(w), C.byref(h), None, None, None, None), "get_view")
        W, H = int(w.value), int(h.value)
        planes = np.empty((H, W, 4), np.float32); weak = np.empty((H, W), np.uint8); sel = np.empty((H, W), np.uint32); rad = np.empty((H, W), np.int32)
        self._check(self.lib.dvp_farm_get_view(self.h, view, C.byref(w), C.byref(h), _ptr(planes), _ptr(weak), _ptr(sel), _ptr(rad)), "get_view")
        return planes, weak, sel, rad


def resize_linear_f32(image: np.ndarray, dst_w: int, dst_h: int, device: int = 0) -> np.ndarray:
    """cv::resize(image, Size(dst_w, dst_h), INTER_LINEAR) of a float grey image on the device (dvp_resize_linear_f32): the
    pyramid level InuputInitialization / GetProblemEdges build (reference APD.cpp:1119-1140, main.cpp:203-209)."""
    lib = load_library(PRODUCT_LIB, "dvp_")
    img = np.ascontiguousarray(image, np.float32)
    H, W = img.shape
    out = np.empty((dst_h, dst_w), np.float32)
    rc = lib.dvp_resize_linear_f32(device, _ptr(img), W, H, _ptr(out), dst_w, dst_h)
    if rc != 0:
        raise DvpError(f"dvp_resize_linear_f32 -> {STATUS.get(rc, rc)}")
    return out


def label_segment(image: np.ndarray, scale: int, device: int = 0):
    """EdgeSegment(scale, image, 1) on the device (dvp_label_segment): -> (labels [new_rows, new_cols] int32, the quarter-size
    edge image after the Hough lines, device time in ms)."""
    lib = load_library(PRODUCT_LIB, "dvp_")
    img = np.ascontiguousarray(image, np.uint8)
    H, W = img.shape
    nc, nr = C.c_int(), C.c_int()
    rc = lib.dvp_label_size(W, H, int(scale), C.byref(nc), C.byref(nr))
    if rc != 0:
        raise DvpError(f"dvp_label_size -> {STATUS.get(rc, rc)}")
    labels = np.empty((nr.value, nc.value), np.int32)
    small = np.empty(((H // 2) // 2, (W // 2) // 2), np.uint8)
    ms = C.c_float()
    rc = lib.dvp_label_segment(device, _ptr(img), W, H, int(scale), _ptr(labels), _ptr(small), C.byref(ms))
    if rc != 0:
        raise DvpError(f"dvp_label_segment -> {STATUS.get(rc, rc)}")
    return labels, small, float(ms.value)


def edge_segment(image: np.ndarray, device: int = 0):
    """The depth-edge prior of one level image (include/dvp_mvs.h, row N4): EdgeSegment(scale, image, 0, true) of the
    reference (APD.cpp:348-466) on the device.  image: [h, w] uint8 -> (edge [h, w] uint8 0/255, (threshold1,
    threshold2), device ms)."""
    lib = load_library(PRODUCT_LIB, "dvp_")
    img = np.ascontiguousarray(image, np.uint8)
    if img.ndim != 2:
        raise ValueError("expected a [h, w] uint8 image")
    H, W = img.shape
    edge = np.empty((H, W), np.uint8)
    thr = (C.c_int32 * 2)()
    ms = C.c_float()
    rc = lib.dvp_edge_segment(device, _ptr(img), W, H, _ptr(edge), thr, C.byref(ms))
    if rc != 0:
        raise DvpError(f"dvp_edge_segment -> {STATUS.get(rc, rc)}")
    return edge, (int(thr[0]), int(thr[1])), float(ms.value)


class Fusion:
    """Depth-map fusion on the device (include/dvp_mvs.h, row N3): RunFusion of the reference (APD.cpp:1809-1960) with
    the reference's sequential visiting order reproduced by deterministic reservations."""

    def __init__(self, views, device: int = 0):
        """views: list of view dicts (see make_fusion_view; `planes` [h,w,4] may replace depth + normal), or the number
        of views when they are registered later (from_scene)."""
        self.lib = load_library(PRODUCT_LIB, "dvp_")
        self.device = device
        count = views if isinstance(views, int) else len(views)
        self.h = self.lib.dvp_fusion_create(device, count)
        if not self.h:
            raise DvpError(f"dvp_fusion_create failed (device {device}, {count} views)")
        self.shapes = []
        for i, v in enumerate([] if isinstance(views, int) else views):
            keep = []
            if "planes" in v:
                planes = np.ascontiguousarray(v["planes"], np.float32)
                H, W = planes.shape[:2]
                stub = dict(v, depth=np.zeros((H, W), np.float32), normal=np.zeros((H, W, 3), np.float32))
                fv = make_fusion_view(stub, keep)
                self._check(self.lib.dvp_fusion_set_view_planes(self.h, i, C.byref(fv), _ptr(planes)), "set_view_planes")
            else:
                fv = make_fusion_view(v, keep)
                self._check(self.lib.dvp_fusion_set_view(self.h, i, C.byref(fv)), "set_view")
            self.shapes.append((fv.height, fv.width, fv.num_src))

    @classmethod
    def from_scene(cls, scene: "Scene", images: list, blocks: list | None = None) -> "Fusion":
        """N2 -> N3: every view's maps go from the scene's device buffers to the fusion object (dvp_scene_fuse_views);
        images[v]: [h, w, 3] uint8 at the view's current map size."""
        V = scene.num_views
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        blks = None if blocks is None else [None if b is None else np.ascontiguousarray(b, np.uint8) for b in blocks]
        shapes = []
        for v in range(V):          # the library reads h * w * 3 bytes per image: check the sizes before handing pointers over
            w, h = C.c_int(), C.c_int()
            scene._check(scene.lib.dvp_scene_get_view(scene.h, v, C.byref(w), C.byref(h), None, None, None, None), "get_view")
            if imgs[v].shape != (h.value, w.value, 3):
                raise ValueError(f"view {v}: image {imgs[v].shape} does not match the map size {(h.value, w.value, 3)}")
            if blks is not None and blks[v] is not None and blks[v].shape != (h.value, w.value):
                raise ValueError(f"view {v}: block mask {blks[v].shape} does not match the map size {(h.value, w.value)}")
            shapes.append((int(h.value), int(w.value), len(scene.src_views[v])))
        f = cls(V, scene.device)
        img_ptrs = (C.c_void_p * V)(*[im.ctypes.data for im in imgs])
        blk_ptrs = None if blks is None else (C.c_void_p * V)(*[None if b is None else b.ctypes.data for b in blks])
        rc = f.lib.dvp_scene_fuse_views(scene.h, f.h, img_ptrs, blk_ptrs)
        if rc != 0:
            raise DvpError(f"dvp_scene_fuse_views -> {STATUS.get(rc, rc)}")
        f.shapes = shapes
        return f

    def _check(self, rc, what):
        if rc != 0:
            raise DvpError(f"dvp_fusion_{what} -> {STATUS.get(rc, rc)}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.dvp_fusion_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self._check(self.lib.dvp_fusion_reset(self.h), "reset")

    def set_mode(self, mode: int):
        """0 RunFusion (ETH, default), 1 RunFusion_TAT_Intermediate, 2 RunFusion_TAT_advanced."""
        self._check(self.lib.dvp_fusion_set_mode(self.h, mode), "set_mode")

    def run_view(self, view: int) -> float:
        ms = C.c_float()
        self._check(self.lib.dvp_fusion_run_view(self.h, view, C.byref(ms)), "run_view")
        return float(ms.value)

    def run(self):
        """-> (points [n, 6] float32 in the reference's order, device ms)."""
        n, ms = C.c_longlong(), C.c_float()
        self._check(self.lib.dvp_fusion_run(self.h, C.byref(n), C.byref(ms)), "run")
        return self.points(), float(ms.value)

    def num_points(self) -> int:
        return int(self.lib.dvp_fusion_num_points(self.h))

    def points(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.num_points() - first if count is None else count
        out = np.empty((count, 6), np.float32)
        if count:
            self._check(self.lib.dvp_fusion_get_points(self.h, _ptr(out), first, count), "get_points")
        return out

    def mask(self, view: int) -> np.ndarray:
        H, W, _ = self.shapes[view]
        out = np.empty((H, W), np.uint8)
        self._check(self.lib.dvp_fusion_get_mask(self.h, view, _ptr(out)), "get_mask")
        return out

    def _require_last(self, view: int):
        """dvp_fusion_last_view copies the buffers of the view that ran LAST, sized by that view: asking for another one
        would overrun the numpy buffers allocated for `view`."""
        last = int(self.lib.dvp_fusion_last_view_index(self.h))
        if last != view:
            raise DvpError(f"fusion stage outputs are those of view {last if last >= 0 else None} (the last one run), not of view {view}")

    def last_view(self, view: int):
        """Stage outputs of the last run_view(view): candidates (cells, terms [h*w, S]), decision (used [h*w]), rounds."""
        self._require_last(view)
        H, W, S = self.shapes[view]
        cells = np.empty((H * W, S), np.int32); terms = np.empty((H * W, S), np.float32); used = np.empty(H * W, np.uint32)
        rounds = C.c_int()
        self._check(self.lib.dvp_fusion_last_view(self.h, _ptr(cells), _ptr(terms), _ptr(used), C.byref(rounds)), "last_view")
        return cells, terms, used, int(rounds.value)

    def last_used(self, view: int) -> np.ndarray:
        """Per-pixel decision of the last run_view(view) in any mode: bit j = source j counted, 0 = no point."""
        self._require_last(view)
        H, W, _ = self.shapes[view]
        used = np.empty(H * W, np.uint32)
        self._check(self.lib.dvp_fusion_last_view(self.h, None, None, _ptr(used), None), "last_view")
        return used

    def write_ply(self, path: str):
        self._check(self.lib.dvp_fusion_write_ply(self.h, os.fsencode(path)), "write_ply")
