"""Row N4, second half: the reference's on-disk exchange formats (ReadBinMat / WriteBinMat
APD.cpp:548-648, writeDepthDmb / writeNormalDmb APD.cpp:575-628, ReadCamera APD.cpp:651-692, GenerateSampleList
main.cpp:127-170) against files laid out by hand exactly as the reference writes / expects them (the same entry points are
compared with the reference's own functions, compiled from those lines, in tests/test_ref_host.py).  Host code: no GPU."""
import struct

import numpy as np
import pytest

from dvp_mvs_b200 import DvpError, formats


def test_binmat_layout_is_the_reference_s(tmp_path):
    rng = np.random.default_rng(1)
    cases = [(rng.random((5, 7)).astype(np.float32), 5), (rng.integers(0, 3, (4, 6)).astype(np.uint8), 0),
             (rng.integers(-5, 99, (3, 8)).astype(np.int32), 4), (rng.random((2, 3, 3)).astype(np.float32), 21)]
    for i, (a, code) in enumerate(cases):
        p = tmp_path / f"m{i}.bin"
        # what WriteBinMat emits: version 1, rows, cols, mat.type(), then mat.step * rows bytes
        p.write_bytes(struct.pack("<4i", 1, a.shape[0], a.shape[1], code) + a.tobytes())
        got = formats.read_binmat(str(p))
        assert got.dtype == a.dtype and got.shape == a.shape and (got == a).all()
        q = tmp_path / f"w{i}.bin"
        formats.write_binmat(str(q), a)
        assert q.read_bytes() == p.read_bytes()
    bad = tmp_path / "bad.bin"
    bad.write_bytes(struct.pack("<4i", 2, 1, 1, 0) + b"\0")               # version != 1: the reference's "Version error"
    with pytest.raises(DvpError):
        formats.read_binmat(str(bad))
    with pytest.raises(DvpError):
        formats.read_binmat(str(tmp_path / "missing.bin"))


def test_dmb_layout(tmp_path):
    depth = np.arange(12, dtype=np.float32).reshape(3, 4)
    normal = np.arange(36, dtype=np.float32).reshape(3, 4, 3)
    formats.write_dmb(str(tmp_path / "d.dmb"), depth)
    formats.write_dmb(str(tmp_path / "n.dmb"), normal)
    assert (tmp_path / "d.dmb").read_bytes() == struct.pack("<4i", 1, 3, 4, 1) + depth.tobytes()    # type, h, w, nb
    assert (tmp_path / "n.dmb").read_bytes() == struct.pack("<4i", 1, 3, 4, 3) + normal.tobytes()


def test_camera_file(tmp_path):
    p = tmp_path / "00000000_cam.txt"
    p.write_text("extrinsic\n0.5 -0.25 0.125 1.5\n0.0 1.0 0.0 -2.0\n0.25 0.0 2.0 3.0\n0.0 0.0 0.0 1.0\n\n"
                 "intrinsic\n3410.5 0.0 3110.25\n0.0 3409.5 2073.75\n0.0 0.0 1.0\n\n1.5 0.01 192 12.0\n")
    cam = formats.read_camera(str(p))
    assert cam["R"].tolist() == [0.5, -0.25, 0.125, 0.0, 1.0, 0.0, 0.25, 0.0, 2.0] and cam["t"].tolist() == [1.5, -2.0, 3.0]
    assert cam["K"].tolist() == [3410.5, 0.0, 3110.25, 0.0, 3409.5, 2073.75, 0.0, 0.0, 1.0]
    assert float(cam["depth_min"]) == 1.5 and float(cam["depth_max"]) == 12.0
    R = cam["R"].astype(np.float64).reshape(3, 3); t = cam["t"].astype(np.float64)
    np.testing.assert_array_equal(cam["c"], (-(R.T @ t)).astype(np.float32))   # -R^T t accumulated in double, stored as float
    with pytest.raises(DvpError):
        formats.read_camera(str(tmp_path / "none.txt"))


def test_pair_file(tmp_path):
    p = tmp_path / "pair.txt"
    p.write_text("3\n0\n3 1 0.75 2 0.5 7 -1.0\n4\n2 0 2.5 2 0.0\n2\n0\n")
    assert formats.read_pairs(str(p)) == [(0, [1, 2]), (4, [0]), (2, [])]   # scores <= 0 are dropped (main.cpp:163-165)
