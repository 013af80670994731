"""Row N1: visibility-map restoration after RunPatchMatch (reference main.cpp:297-363 with Connect /
Label_Seek / Label_Update, APD.cpp:138-346).  CPU part: the restatement (oracle/cpu/visibility_cpu.cpp) against
hand-computed answers, and the reference's quirky two-pass labelling against exact 4-connected components.
GPU part: dvp_restore_visibility against the restatement, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from util import ROOT

CPU_LIB = os.path.join(ROOT, "oracle", "_ref", "libapd_cpu.so")


def oracle(selected: np.ndarray, S: int, scale: int, exact_cc: int = 0) -> np.ndarray:
    lib = C.CDLL(CPU_LIB)
    fn = lib.cpu_restore_visibility
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    sel = np.ascontiguousarray(selected, np.uint32)
    out = np.zeros_like(sel)
    H, W = sel.shape
    assert fn(sel.ctypes.data, out.ctypes.data, W, H, S, scale, exact_cc) == 0
    return out


def random_masks(rng, H, W, S, p_visible):
    sel = np.zeros((H, W), np.uint32)
    for i in range(S):
        sel |= (rng.random((H, W)) < p_visible).astype(np.uint32) << i
    return sel


def blobs(rng, H, W, S, n_blobs, rmax):
    """All views visible except disc-shaped holes of assorted sizes (what real selected-view maps look like)."""
    sel = np.full((H, W), (1 << S) - 1, np.uint32)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(S):
        for _ in range(n_blobs):
            cy, cx, r = rng.integers(0, H), rng.integers(0, W), rng.integers(1, rmax)
            sel[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] &= ~np.uint32(1 << i)
    return sel


def test_known_answers():
    # scale 8 -> threshold 20 pixels: a 19-pixel hole is filled, a 20-pixel hole is kept
    sel = np.ones((12, 12), np.uint32)
    sel[1, 1:11] = 0; sel[2, 1:10] = 0            # 19 pixels, 4-connected
    sel[6, 1:11] = 0; sel[7, 1:11] = 0            # 20 pixels
    out = oracle(sel, 1, 8)
    assert out[1, 1:11].all() and out[2, 1:10].all()
    assert not out[6, 1:11].any() and not out[7, 1:11].any()
    assert (out[sel == 1] == 1).all()             # visible pixels stay visible
    # diagonal contact does not connect: two 10-pixel bars touching at a corner are two small regions
    sel = np.ones((8, 24), np.uint32)
    sel[2, 1:11] = 0; sel[3, 11:21] = 0
    assert oracle(sel, 1, 8).all()
    # ... while an edge contact makes one 20-pixel region that is kept
    sel = np.ones((8, 24), np.uint32)
    sel[2, 1:11] = 0; sel[3, 10:20] = 0
    out = oracle(sel, 1, 8)
    assert not out[2, 1:11].any() and not out[3, 10:20].any()
    # threshold scales with 20 * (8 / scale)^2 (integer division): scale 4 -> 80, scale 2 -> 320, scale 1 -> 1280
    sel = np.ones((40, 40), np.uint32)
    sel[5:14, 5:14] = 0                            # 81 pixels
    assert not oracle(sel, 1, 4)[5:14, 5:14].any()
    assert oracle(sel, 1, 2)[5:14, 5:14].all()
    # views are independent bits; bits >= S are dropped (the reference rebuilds the word from S masks)
    sel = np.full((6, 6), 0b1101, np.uint32)
    out = oracle(sel, 3, 8)
    assert (out == 0b101).all()                    # bit 1: one 36-pixel region (>= 20) stays clear; bit 3 is beyond S


def test_reference_labelling_equals_exact_components():
    """Connect overwrites union links and Label_Update skips the last row/column; on every mask tried the repaired
    partition is exactly the set of 4-connected components (which is what the CUDA path computes)."""
    rng = np.random.default_rng(7)
    n = 0
    for trial in range(4000):
        H, W = int(rng.integers(2, 24)), int(rng.integers(2, 24))
        S = int(rng.integers(1, 4))
        sel = random_masks(rng, H, W, S, rng.choice([0.1, 0.25, 0.4, 0.55, 0.7]))
        scale = int(rng.choice([8, 8, 4]))
        assert (oracle(sel, S, scale, 0) == oracle(sel, S, scale, 1)).all(), (trial, H, W)
        n += 1
    # structured: holes that reach the last row / last column, combs and arches
    for trial in range(300):
        H, W = int(rng.integers(8, 40)), int(rng.integers(8, 40))
        sel = blobs(rng, H, W, 2, int(rng.integers(1, 8)), 9)
        sel[-1, ::2] &= ~np.uint32(1)                       # comb along the last row
        sel[::3, -1] &= ~np.uint32(2)                       # and the last column
        sel[-2, :] &= ~np.uint32(rng.integers(0, 4))
        assert (oracle(sel, 2, 8, 0) == oracle(sel, 2, 8, 1)).all(), trial


def test_restoration_is_idempotent_and_monotone():
    rng = np.random.default_rng(3)
    sel = blobs(rng, 120, 160, 3, 25, 7)
    out = oracle(sel, 3, 8)
    assert ((out & sel) == sel).all()              # bits are only ever added
    assert (oracle(out, 3, 8) == out).all()        # what survives is >= threshold and survives again


# ------------------------------------------------------------------------------------------------ GPU
def _engine(W, H, S):
    from dvp_mvs_b200 import Engine, default_params, synth
    sc = synth.make_scene(W, H, S)
    p = default_params(); p.max_iterations = 1; p.num_images = S + 1
    p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
    p.use_APD = 0; p.state = 0
    e = Engine(W, H, S, p)
    e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label)
    return e, sc, p


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,S", [(64, 48, 2), (97, 61, 3), (33, 130, 1), (320, 240, 4)])
def test_gpu_restoration_matches_the_restatement(W, H, S):
    e, sc, p = _engine(W, H, S)
    rng = np.random.default_rng(W * 1000 + H)
    for scale, maker in [(8, lambda: random_masks(rng, H, W, S, 0.45)), (8, lambda: random_masks(rng, H, W, S, 0.8)),
                         (4, lambda: blobs(rng, H, W, S, 12, 9)), (8, lambda: blobs(rng, H, W, S, 30, 5)),
                         (8, lambda: np.zeros((H, W), np.uint32)), (8, lambda: np.full((H, W), (1 << S) - 1, np.uint32))]:
        sel = maker()
        planes = sc.planes_true.copy()
        planes[::7, ::5, 3] = 0.5 * p.depth_min        # out of range -> depth 0, UNKNOWN (main.cpp:300-306)
        planes[3::11, 2::9, 3] = 2.0 * p.depth_max
        e.set("selected", sel); e.set("planes", planes); e.set("weak", np.ones((H, W), np.uint8))
        ms = e.restore_visibility(scale)
        assert ms > 0
        got_planes, got_weak, got_sel, _ = e.download()
        assert (got_sel == oracle(sel, S, scale, 0)).all()
        bad = (planes[..., 3] < p.depth_min) | (planes[..., 3] > p.depth_max)
        assert (got_planes[..., 3][bad] == 0).all() and (got_planes[..., 3][~bad] == planes[..., 3][~bad]).all()
        assert (got_planes[..., :3] == planes[..., :3]).all()
        assert (got_weak[bad] == 2).all() and (got_weak[~bad] == 1).all()
        # ... and against ProcessProblem's own lines (main.cpp:282-363 compiled from the reference) where that build is present
        try:
            import ref_host
            have_ref = ref_host.available()
        except Exception:
            have_ref = False
        if have_ref and W * H <= 6000:      # Label_Seek is quadratic in the label count
            depth_ref, weak_ref, sel_ref = ref_host.post_pass(planes, np.ones((H, W), np.uint8), sel, S, scale, p.depth_min, p.depth_max)
            assert (got_sel == sel_ref).all() and (got_weak == weak_ref).all()
            assert (got_planes[..., 3].view(np.uint32) == depth_ref.view(np.uint32)).all()


@pytest.mark.gpu
def test_gpu_restoration_after_a_real_pass_and_at_full_size():
    # maps as RunPatchMatch leaves them
    e, sc, p = _engine(320, 240, 3)
    e.run()
    _, _, sel, _ = e.download()
    e.restore_visibility(8)
    _, _, got, _ = e.download()
    assert (got == oracle(sel, 3, 8, 0)).all()
    assert ((got & sel) == sel).all()
    # BASELINE top-level size (3111x2073): against exact components (the O(L^2) reference merge is too slow there)
    W, H, S = 3111, 2073, 4
    from dvp_mvs_b200 import Engine
    q = p.copy(); q.num_images = S + 1
    big = Engine(W, H, S, q)
    rng = np.random.default_rng(11)
    img = np.zeros((S + 1, H, W), np.float32)
    big.upload(images=img, cameras=np.zeros(S + 1, sc.cameras.dtype), planes=np.ones((H, W, 4), np.float32))
    sel = blobs(rng, H, W, S, 400, 60)
    sel[rng.random((H, W)) < 0.02] = 0                 # plus salt noise: many tiny regions
    big.set("selected", sel)
    ms = big.restore_visibility(2)
    _, _, got, _ = big.download()
    assert (got == oracle(sel, S, 2, 1)).all()
    assert ((got & sel) == sel).all()
    print(f"restore_visibility {W}x{H} S={S}: {ms:.2f} ms device")


@pytest.mark.gpu
def test_gpu_restoration_argument_errors():
    from dvp_mvs_b200 import Engine, default_params
    from dvp_mvs_b200._lib import DvpError
    p = default_params(); p.num_images = 3
    e = Engine(32, 32, 2, p)
    with pytest.raises(DvpError):
        e.restore_visibility(8)        # nothing uploaded yet: DVP_ERR_STATE
    e2, _, _ = _engine(32, 32, 2)
    with pytest.raises(DvpError):
        e2.restore_visibility(0)       # scale_size must be positive
