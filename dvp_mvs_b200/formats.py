"""The reference's on-disk exchange formats through the C ABI (include/dvp_mvs.h, row N4 second half; reference
ReadBinMat / WriteBinMat APD.cpp:548-648, writeDepthDmb / writeNormalDmb APD.cpp:575-628, ReadCamera APD.cpp:651-692,
GenerateSampleList main.cpp:127-170).  Host code only: no GPU is needed to call these."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import PRODUCT_LIB, STATUS, DvpError, load_library
from .synth import CAMERA_DTYPE

MAX_IMAGES = 32
# OpenCV type code = depth + 8 * (channels - 1)
_DEPTHS = {0: np.uint8, 1: np.int8, 2: np.uint16, 3: np.int16, 4: np.int32, 5: np.float32, 6: np.float64}
_CODES = {np.dtype(v): k for k, v in _DEPTHS.items()}


def _lib():
    lib = load_library(PRODUCT_LIB, "dvp_")
    lib.dvp_io_binmat_header.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.dvp_io_read_binmat.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
    lib.dvp_io_write_binmat.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.dvp_io_write_dmb.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    lib.dvp_io_read_camera.argtypes = [C.c_char_p, C.c_void_p]
    lib.dvp_io_read_pairs.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _check(rc, what):
    if rc != 0:
        raise DvpError(f"dvp_io_{what} -> {STATUS.get(rc, rc)}")


def read_binmat(path: str) -> np.ndarray:
    """A file written by the reference's WriteBinMat -> [rows, cols] or [rows, cols, channels] array."""
    lib = _lib()
    rows, cols, code = C.c_int32(), C.c_int32(), C.c_int32()
    _check(lib.dvp_io_binmat_header(os.fsencode(path), C.byref(rows), C.byref(cols), C.byref(code)), "binmat_header")
    depth, channels = code.value & 7, (code.value >> 3) + 1
    if depth not in _DEPTHS:
        raise DvpError(f"unsupported OpenCV type code {code.value}")
    out = np.empty((rows.value, cols.value, channels), _DEPTHS[depth])
    _check(lib.dvp_io_read_binmat(os.fsencode(path), out.ctypes.data, out.nbytes), "read_binmat")
    return out[..., 0] if channels == 1 else out


def write_binmat(path: str, a: np.ndarray):
    a = np.ascontiguousarray(a)
    channels = 1 if a.ndim == 2 else a.shape[2]
    code = _CODES[a.dtype] + 8 * (channels - 1)
    _check(_lib().dvp_io_write_binmat(os.fsencode(path), a.shape[0], a.shape[1], code, a.ctypes.data), "write_binmat")


def write_dmb(path: str, a: np.ndarray):
    """writeDepthDmb ([h, w]) / writeNormalDmb ([h, w, 3])."""
    a = np.ascontiguousarray(a, np.float32)
    channels = 1 if a.ndim == 2 else a.shape[2]
    _check(_lib().dvp_io_write_dmb(os.fsencode(path), a.shape[0], a.shape[1], channels, a.ctypes.data), "write_dmb")


def read_camera(path: str) -> np.ndarray:
    cam = np.zeros((), CAMERA_DTYPE)
    buf = np.zeros(1, CAMERA_DTYPE)
    _check(_lib().dvp_io_read_camera(os.fsencode(path), buf.ctypes.data), "read_camera")
    cam[...] = buf[0]
    return cam


def read_pairs(path: str):
    """pair.txt -> [(ref_id, [src ids with a positive score])]."""
    lib = _lib()
    n = C.c_int32()
    _check(lib.dvp_io_read_pairs(os.fsencode(path), 0, C.byref(n), None, None, None), "read_pairs")
    ref = np.zeros(max(n.value, 1), np.int32); num = np.zeros(max(n.value, 1), np.int32); src = np.zeros((max(n.value, 1), MAX_IMAGES), np.int32)
    _check(lib.dvp_io_read_pairs(os.fsencode(path), n.value, C.byref(n), ref.ctypes.data, num.ctypes.data, src.ctypes.data), "read_pairs")
    return [(int(ref[i]), [int(v) for v in src[i, :num[i]]]) for i in range(n.value)]
