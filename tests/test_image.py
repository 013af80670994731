"""Row N2, image pyramid: cv::resize(INTER_LINEAR) of the float grey image — reference APD.cpp:1119-1140,
main.cpp:203-209 — restated (oracle/image_oracle.py), pinned against OpenCV 4.13's generic path (tests/golden/resize_f32.npz,
tools/make_image_golden.py; live against cv2 where it is importable), and run on the device (dvp_resize_linear_f32)."""
import os

import numpy as np
import pytest

from util import GOLDEN
from image_oracle import level_size, resize_linear_f32 as resize_oracle


def test_level_sizes_follow_the_reference_rounding():
    # std::round(cols * (1.0f / scale)): ETH3D 6221 x 4146 -> 778 x 518, 1555 x 1037 (1036.5 rounds up), 3111 x 2073
    assert level_size(6221, 4146, 8) == (778, 518)
    assert level_size(6221, 4146, 4) == (1555, 1037)
    assert level_size(6221, 4146, 2) == (3111, 2073)
    assert level_size(1920, 1080, 4) == (480, 270)


def test_restatement_matches_opencv_golden_vectors():
    g = np.load(os.path.join(GOLDEN, "resize_f32.npz"))
    assert int(g["count"]) >= 6
    for i in range(int(g["count"])):
        img = g[f"image_{i}"].astype(np.float32)
        want = g[f"resized_{i}"]
        got = resize_oracle(img, want.shape[1], want.shape[0])
        assert level_size(img.shape[1], img.shape[0], int(g[f"scale_{i}"])) == (want.shape[1], want.shape[0])
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), i


def test_restatement_matches_cv2_live_on_random_sizes():
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "ipp"):
        pytest.skip("this cv2 cannot switch IPP off")
    rng = np.random.default_rng(4)
    cv2.ipp.setUseIPP(False)
    try:
        for _ in range(40):
            w, h = int(rng.integers(9, 400)), int(rng.integers(9, 300))
            scale = int(rng.choice([2, 4, 8]))
            dw, dh = level_size(w, h, scale)
            if dw < 1 or dh < 1:
                continue
            img = rng.integers(0, 256, (h, w)).astype(np.float32)
            want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
            assert (resize_oracle(img, dw, dh).view(np.uint32) == want.view(np.uint32)).all(), (w, h, scale)
    finally:
        cv2.ipp.setUseIPP(True)


@pytest.mark.gpu
def test_gpu_resize_is_bit_exact_vs_restatement_and_golden():
    from dvp_mvs_b200 import resize_linear_f32
    g = np.load(os.path.join(GOLDEN, "resize_f32.npz"))
    for i in range(int(g["count"])):
        img = g[f"image_{i}"].astype(np.float32)
        want = g[f"resized_{i}"]
        got = resize_linear_f32(img, want.shape[1], want.shape[0])
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), i
    rng = np.random.default_rng(8)
    for (w, h, scale) in ((6221, 4146, 2), (6221, 4146, 8), (1920, 1080, 4), (3, 3, 2), (17, 5, 4), (640, 480, 1)):
        dw, dh = level_size(w, h, scale)
        img = rng.integers(0, 256, (h, w)).astype(np.float32)
        # fractional grey levels too: a level of the pyramid is never resized again by the reference, but the entry point takes any float image
        if w < 100:
            img += rng.random((h, w)).astype(np.float32)
        got = resize_linear_f32(img, dw, dh)
        assert (got.view(np.uint32) == resize_oracle(img, dw, dh).view(np.uint32)).all(), (w, h, scale)


@pytest.mark.gpu
def test_gpu_scene_builds_its_pyramid_from_the_full_image():
    """dvp_scene_set_image: every level image equals cv::resize of the FULL float image to the level's size (bit for bit
    with the restatement), and the edge map computed with it equals the one computed from a level image given by the caller."""
    from dvp_mvs_b200 import Scene, synth
    mv = synth.make_multiview(333, 211, 2, 2, seed=4)
    rng = np.random.default_rng(2)
    full = [rng.integers(0, 256, (mv.full_h, mv.full_w)).astype(np.uint8) for _ in range(2)]
    L = 2
    a = Scene(2, L); b = Scene(2, L)
    for v in range(2):
        a.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        b.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        a.set_image(v, full[v], compute_edges=True)
        for l in range(L):
            w, h = a.view_level_size(v, l)
            scale = 2 ** (L - l)
            assert level_size(mv.full_w, mv.full_h, scale) == (w, h)
            want = resize_oracle(full[v].astype(np.float32), w, h)
            got = a.get_image(v, l)
            assert (got.view(np.uint32) == want.view(np.uint32)).all(), (v, l)
            b.set_level(v, l, want, None, None)
            assert (b.compute_edges(v, l) == a.compute_edges(v, l)).all(), (v, l)
    a.close(); b.close()
