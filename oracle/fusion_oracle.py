"""TEST INFRASTRUCTURE ONLY — opens oracle/_ref/libapd_cpu.so for the CPU restatement of the reference's depth-map
fusion (oracle/cpu/fusion_cpu.cpp; RunFusion, APD.cpp:1809-1960).  Import from tests/ and tools/ only."""
import ctypes as C
import os

import numpy as np

from dvp_mvs_b200._lib import FusionView, make_fusion_view

HERE = os.path.dirname(os.path.abspath(__file__))
CPU_LIB = os.path.join(HERE, "_ref", "libapd_cpu.so")


class FusionOracle:
    def __init__(self, views: list):
        self.lib = C.CDLL(CPU_LIB)
        self.keep = []
        self.V = len(views)
        self.arr = (FusionView * self.V)(*[make_fusion_view(v, self.keep) for v in views])
        self.shapes = [(fv.height, fv.width, fv.num_src) for fv in self.arr]
        self.lib.fusion_cpu_run.restype = C.c_longlong
        self.lib.fusion_cpu_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]
        self.lib.fusion_cpu_candidates.restype = None
        self.lib.fusion_cpu_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        self.lib.fusion_cpu_resolve.restype = C.c_longlong
        self.lib.fusion_cpu_resolve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]
        self.lib.fusion_cpu_run_tat.restype = C.c_longlong
        self.lib.fusion_cpu_run_tat.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]
        self.reset()

    def run_tat(self, mode: int):
        """RunFusion_TAT_Intermediate (mode 1) / RunFusion_TAT_advanced (mode 2) as written -> (points, used per view)."""
        self.reset()
        cap = sum(h * w for h, w, _ in self.shapes)
        pts = np.empty((cap, 6), np.float32)
        used = [np.zeros(h * w, np.uint32) for h, w, _ in self.shapes]
        used_ptrs = (C.c_void_p * self.V)(*[u.ctypes.data for u in used])
        n = self.lib.fusion_cpu_run_tat(mode, self.V, self.arr, self.mask_ptrs, pts.ctypes.data, cap, used_ptrs)
        return pts[:n].copy(), used

    def reset(self):
        self.masks = [np.zeros((h, w), np.uint8) for h, w, _ in self.shapes]
        self.mask_ptrs = (C.c_void_p * self.V)(*[m.ctypes.data for m in self.masks])

    def run(self):
        """The reference's loop as written -> points [n, 6] float32 (masks are left in self.masks)."""
        self.reset()
        cap = sum(h * w for h, w, _ in self.shapes)
        pts = np.empty((cap, 6), np.float32)
        n = self.lib.fusion_cpu_run(self.V, self.arr, self.mask_ptrs, pts.ctypes.data, cap)
        return pts[:n].copy()

    def candidates(self, view: int):
        h, w, S = self.shapes[view]
        cells = np.empty((h * w, S), np.int32); terms = np.empty((h * w, S), np.float32)
        self.lib.fusion_cpu_candidates(self.arr, view, cells.ctypes.data, terms.ctypes.data)
        return cells, terms

    def resolve(self, view: int, cells: np.ndarray, terms: np.ndarray):
        """Sequential greedy pass of one view over given candidates against self.masks -> (used [h*w], points)."""
        h, w, S = self.shapes[view]
        cells = np.ascontiguousarray(cells, np.int32); terms = np.ascontiguousarray(terms, np.float32)
        assert cells.shape == (h * w, S) and terms.shape == (h * w, S)
        used = np.empty(h * w, np.uint32); pts = np.empty((h * w, 6), np.float32)
        n = self.lib.fusion_cpu_resolve(self.arr, view, cells.ctypes.data, terms.ctypes.data, self.mask_ptrs, used.ctypes.data, pts.ctypes.data, h * w)
        return used, pts[:n].copy()
