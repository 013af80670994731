"""TEST INFRASTRUCTURE ONLY — opens oracle/_ref/libapd_cpu.so, the CPU restatement of the reference's
PatchMatch kernels (oracle/cpu/apd_cpu.cpp).  Import from tests/, bench.py (cpu_baseline / --impl
reference fallback) and __graft_entry__.smoke() only."""
import ctypes as C
import os

import numpy as np

from dvp_mvs_b200._lib import Engine

HERE = os.path.dirname(os.path.abspath(__file__))
CPU_LIB = os.path.join(HERE, "_ref", "libapd_cpu.so")


def available() -> bool:
    return os.path.exists(CPU_LIB)


def engine(width, height, num_src, params) -> Engine:
    return Engine(width, height, num_src, params, lib_path=CPU_LIB, prefix="cpu_")


def init_random_states(width: int, height: int, seed: int) -> np.ndarray:
    """K1: cuRAND XORWOW states after curand_init(seed, y, x), [H, W, 6] uint32 {d, v0..v4}."""
    lib = C.CDLL(CPU_LIB)
    out = np.empty((height, width, 6), np.uint32)
    lib.cpu_init_random_states.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_void_p]
    rc = lib.cpu_init_random_states(width, height, seed, out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def tex_probe(eng: Engine, img: int, xy: np.ndarray) -> np.ndarray:
    xy = np.ascontiguousarray(xy, np.float32)
    out = np.empty(len(xy), np.float32)
    fn = getattr(eng.lib, eng.prefix + "tex_probe")
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    rc = fn(eng.ctx, img, xy.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), len(xy))
    assert rc == 0
    return out
