"""Row N3: depth-map fusion — RunFusion, the "ETH version" main() calls (reference APD.cpp:1809-1960).
CPU part: the restatement (oracle/cpu/fusion_cpu.cpp) against hand-computed answers, its split into a mask-independent
and a mask-dependent stage, and the reservation scheme of the CUDA path (emulated in numpy) against the sequential loop.
GPU part: dvp_fusion_* against the restatement, stage by stage: candidates within float tolerance (exp / acos are the
only non-IEEE operations), the order-dependent resolution bit for bit."""
import os

import numpy as np
import pytest

from util import ROOT
from dvp_mvs_b200 import synth
from fusion_oracle import FusionOracle

FREE = np.uint32(0xFFFFFFFF)


def camera(W, H, f, c=(0.0, 0.0, 0.0)):
    K = np.array([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1.0]])
    return synth.make_camera(K, np.eye(3), np.array(c, np.float64), W, H, 1.0, 10.0)


def flat_view(W, H, f, depth, src, grey=100, weak=1, c=(0.0, 0.0, 0.0)):
    """A fronto-parallel wall at z = depth seen by an axis-aligned camera."""
    n = np.zeros((H, W, 3), np.float32); n[..., 2] = -1.0
    img = np.full((H, W, 3), grey, np.uint8); img[..., 1] = grey // 2; img[..., 2] = 255 - grey
    return dict(camera=camera(W, H, f, c), depth=np.full((H, W), depth, np.float32), normal=n, image=img,
                weak=np.full((H, W), weak, np.uint8), src_views=src)


@pytest.fixture(scope="module")
def scene():
    mv = synth.make_multiview(640, 480, 4, 2, seed=1)
    return mv


# ---------------------------------------------------------------------------------------------------------- CPU
def test_known_answers_identical_cameras():
    # two identical views of one wall: every pixel of view 0 agrees with view 1 exactly (index 0 -> term 1 > 0.3), is
    # emitted with the mean colour and masks its twin, so view 1 emits nothing
    W, H = 12, 8
    views = [flat_view(W, H, 10.0, 4.0, [1], grey=100), flat_view(W, H, 10.0, 4.0, [0], grey=60)]
    o = FusionOracle(views)
    pts = o.run()
    assert len(pts) == W * H
    assert o.masks[1].all() and not o.masks[0].any()
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    np.testing.assert_array_equal(pts[:, 0].reshape(H, W), (np.float32(4.0) * (xs - np.float32(W / 2)).astype(np.float32)) / np.float32(10.0))
    np.testing.assert_array_equal(pts[:, 1].reshape(H, W), (np.float32(4.0) * (ys - np.float32(H / 2)).astype(np.float32)) / np.float32(10.0))
    assert (pts[:, 2] == 4.0).all()
    assert (pts[:, 3] == 80.0).all() and (pts[:, 4] == 40.0).all() and (pts[:, 5] == (155 + 195) / 2).all()
    # a 2 % depth disagreement fails the 1 % test in both directions: nothing is emitted for those pixels
    views[1]["depth"][2:4, :] = 4.08
    o = FusionOracle(views)
    pts = o.run()
    assert len(pts) == W * (H - 2)
    assert not o.masks[1][2:4].any() and o.masks[1][:2].all() and o.masks[1][4:].all()
    # a normal 12 degrees off fails the 10 degree test; 8 degrees passes with index 10 * 0.1396 = 1.396 -> term 0.2476,
    # which is below the 0.3 a STRONG pixel needs
    for deg, expect in ((12.0, 0), (8.0, 0), (2.0, W * H)):
        views = [flat_view(W, H, 10.0, 4.0, [1]), flat_view(W, H, 10.0, 4.0, [0])]
        a = np.deg2rad(deg)
        views[1]["normal"][...] = np.array([np.sin(a), 0.0, -np.cos(a)], np.float32)
        assert len(FusionOracle(views).run()) == expect, deg


def test_known_answers_visiting_order():
    # view 0 at twice the resolution of view 1: the four pixels of a 2x2 block project to one pixel of view 1 (or its
    # neighbour, int(x + 0.5) truncation).  The first pixel in raster order to claim a cell is emitted and masks it; the
    # later claimants see the mask, have no consistent source left and are dropped.
    W, H = 16, 12
    views = [flat_view(W, H, 20.0, 4.0, [1]), flat_view(W // 2, H // 2, 10.0, 4.0, [0])]
    o = FusionOracle(views)
    pts = o.run()
    cells, terms = o.candidates(0)
    inside = (np.arange(W * H) % W < W - 1) & (np.arange(W * H) // W < H - 1)   # the last column / row round to cell W/2, H/2
    assert ((cells[:, 0] >= 0) == inside).all()
    # expected by hand: pixel (x, y) claims cell (int(x/2 + .5), int(y/2 + .5)), which looks back at pixel (2cx, 2cy); the
    # claim's index is that distance in pixels (depths and normals agree exactly), and exp(-d) > 0.3 needs d < 1.204:
    # a diagonal neighbour (d = 1.414) is refused and leaves the cell to the next claimant
    first = np.zeros(W * H, bool)
    seen = set()
    for p in np.flatnonzero(inside):
        y, x = divmod(int(p), W)
        cx, cy = int(x / 2 + 0.5), int(y / 2 + 0.5)
        assert cells[p, 0] == cy * (W // 2) + cx
        d = np.hypot(x - 2 * cx, y - 2 * cy)
        if (cx, cy) not in seen and d < 1.2:
            seen.add((cx, cy)); first[p] = True
    n0 = int(first.sum())
    assert n0 == (W // 2) * (H // 2)                 # every cell of view 1 ends up claimed exactly once
    assert o.masks[1].sum() == n0 and len(pts) == n0  # ... so view 1 itself has nothing left to emit
    ys, xs = np.divmod(np.flatnonzero(first), W)
    np.testing.assert_array_equal(pts[:, 0], (np.float32(4.0) * (xs - np.float32(W / 2)).astype(np.float32)) / np.float32(20.0))
    np.testing.assert_array_equal(pts[:, 1], (np.float32(4.0) * (ys - np.float32(H / 2)).astype(np.float32)) / np.float32(20.0))


def test_run_equals_candidates_plus_resolve(scene):
    for levels in (1, [1, 0, 1, 0]):
        views = synth.make_fusion_views(scene, levels)
        o = FusionOracle(views)
        pts = o.run()
        masks = [m.copy() for m in o.masks]
        assert len(pts) > 1000
        o.reset()
        parts = []
        for v in range(len(views)):
            cells, terms = o.candidates(v)
            used, p = o.resolve(v, cells, terms)
            assert (used != 0).sum() == len(p)
            parts.append(p)
        np.testing.assert_array_equal(np.concatenate(parts), pts)
        for a, b in zip(masks, o.masks):
            np.testing.assert_array_equal(a, b)


def reservations_numpy(masks, ref, src_views, cells, terms, weak):
    """The CUDA path's resolve stage (dvp_kernels_fusion.cu: reserve_one / decide_one), vectorised per round."""
    N, S = cells.shape
    live = cells >= 0
    active = live.any(1) & (masks[ref].ravel() != 1)
    used = np.zeros(N, np.uint32)
    resv = {s: np.full(masks[s].size, FREE, np.uint32) for s in set(src_views)}
    rounds = 0
    while active.any():
        idx = np.flatnonzero(active)
        for j, s in enumerate(src_views):                                  # (A) drop masked cells, reserve the rest
            sel = idx[live[idx, j]]
            c = cells[sel, j]
            masked = masks[s].ravel()[c] == 1
            live[sel[masked], j] = False
            np.minimum.at(resv[s], c[~masked], sel[~masked].astype(np.uint32))
        ready = np.ones(len(idx), bool)                                    # (B) holders of all their cells decide
        for j, s in enumerate(src_views):
            lj = live[idx, j]
            ready &= ~lj | (resv[s][np.where(lj, cells[idx, j], 0)] == idx)
        r = idx[ready]
        assert len(r) > 0
        num = np.zeros(len(r), np.int32); dyn = np.zeros(len(r), np.float32); bits = np.zeros(len(r), np.uint32)
        for j in range(S):
            lj = live[r, j]
            dyn = np.where(lj, dyn + terms[r, j], dyn).astype(np.float32)
            num += lj; bits |= lj.astype(np.uint32) << np.uint32(j)
        factor = np.where(weak.ravel()[r] == 0, np.float32(0.45), np.float32(0.3)).astype(np.float32)
        acc = (num >= 1) & (dyn > factor * num.astype(np.float32))
        for j, s in enumerate(src_views):
            lj = live[r, j]
            c = cells[r[lj], j]
            resv[s][c] = FREE
            masks[s].ravel()[c[acc[lj]]] = 1
        used[r[acc]] = bits[acc]
        active[r] = False
        rounds += 1
    return used, rounds


def test_reservation_scheme_equals_sequential_loop(scene):
    # views of mixed sizes: most pixels of the fine views share their source cells with neighbours
    views = synth.make_fusion_views(scene, [1, 0, 1, 0])
    o = FusionOracle(views)
    masks = [np.zeros_like(m) for m in o.masks]
    most_rounds = 0
    for v in range(len(views)):
        cells, terms = o.candidates(v)
        used_seq, _ = o.resolve(v, cells, terms)
        used_par, rounds = reservations_numpy(masks, v, views[v]["src_views"], cells, terms, views[v]["weak"])
        np.testing.assert_array_equal(used_par, used_seq)
        for a, b in zip(masks, o.masks):
            np.testing.assert_array_equal(a, b)
        most_rounds = max(most_rounds, rounds)
    assert most_rounds > 2     # the order did matter somewhere


def test_reservation_scheme_on_adversarial_candidates():
    """Random claims with heavy collisions (many pixels per source cell, several views marking each other's cells, terms
    straddling the acceptance limit): the round-based scheme must still equal the sequential loop, whatever the chains."""
    rng = np.random.default_rng(21)
    for trial in range(30):
        V = int(rng.integers(2, 5))
        shapes = [(int(rng.integers(3, 9)), int(rng.integers(3, 9))) for _ in range(V)]
        views = []
        for v, (h, w) in enumerate(shapes):
            others = [u for u in range(V) if u != v]
            src = list(rng.permutation(others)[: int(rng.integers(1, len(others) + 1))])
            fv = flat_view(w, h, 10.0, 4.0, [int(u) for u in src])
            fv["weak"] = rng.integers(0, 3, (h, w)).astype(np.uint8)
            fv["depth"][rng.random((h, w)) < 0.1] = 0.0
            views.append(fv)
        o = FusionOracle(views)
        masks = [np.zeros_like(m) for m in o.masks]
        for v, (h, w) in enumerate(shapes):
            S = len(views[v]["src_views"])
            cells = np.full((h * w, S), -1, np.int32); terms = np.zeros((h * w, S), np.float32)
            for j, u in enumerate(views[v]["src_views"]):
                n_u = shapes[u][0] * shapes[u][1]
                claim = rng.random(h * w) < 0.8
                cells[claim, j] = rng.integers(0, max(1, n_u // 3), int(claim.sum()))     # few cells, many claimants
                terms[claim, j] = rng.choice(np.array([0.2, 0.3, 0.31, 0.45, 0.46, 0.9], np.float32), int(claim.sum()))
            cells[views[v]["depth"].ravel() <= 0] = -1        # a pixel without a depth has no candidates (stage 1 skips it)
            used_seq, _ = o.resolve(v, cells, terms)
            used_par, rounds = reservations_numpy(masks, v, views[v]["src_views"], cells, terms, views[v]["weak"])
            np.testing.assert_array_equal(used_par, used_seq)
            for a, b in zip(masks, o.masks):
                np.testing.assert_array_equal(a, b)


# ---------------------------------------------------------------------------------------------------------- GPU
def _gpu_stagewise(views):
    from dvp_mvs_b200 import Fusion
    o = FusionOracle(views)
    f = Fusion(views)
    f.reset()
    report = dict(cell_flips=0, cells=0, term_err=0.0, rounds=[])
    parts = []
    for v in range(len(views)):
        f.run_view(v)
        cells, terms, used, rounds = f.last_view(v)
        # stage 1 against the restatement: same claims, except where acos / exp land a threshold on the other side
        c_ref, t_ref = o.candidates(v)
        flips = cells != c_ref
        report["cell_flips"] += int(flips.sum()); report["cells"] += cells.size
        same = ~flips & (cells >= 0)
        if same.any():
            report["term_err"] = max(report["term_err"], float(np.max(np.abs(terms[same] - t_ref[same]) / t_ref[same])))
        # stage 2 from the device's own candidates: the sequential loop must give the same decisions bit for bit
        used_ref, p_ref = o.resolve(v, cells, terms)
        np.testing.assert_array_equal(used, used_ref)
        parts.append(p_ref)
        report["rounds"].append(rounds)
    pts = f.points()
    np.testing.assert_array_equal(pts, np.concatenate(parts))           # coordinates, colours and order
    for v in range(len(views)):
        np.testing.assert_array_equal(f.mask(v), o.masks[v])
    f.close()
    return report, pts


@pytest.mark.gpu
def test_gpu_fusion_stagewise_vs_restatement(scene):
    views = synth.make_fusion_views(scene, 1)
    report, pts = _gpu_stagewise(views)
    assert len(pts) > 1000
    assert report["cell_flips"] <= 1e-4 * report["cells"], report
    assert report["term_err"] <= 1e-6, report        # exp(-x) in double rounded to float vs libm expf: a few ulp at most


@pytest.mark.gpu
def test_gpu_fusion_mixed_sizes_many_rounds(scene):
    # fine views fused against coarse ones: shared source cells everywhere, long reservation chains, tail kernel
    views = synth.make_fusion_views(scene, [1, 0, 1, 0])
    report, pts = _gpu_stagewise(views)
    assert len(pts) > 1000
    assert max(report["rounds"]) > 4, report
    assert report["cell_flips"] <= 1e-4 * report["cells"], report


@pytest.mark.gpu
def test_gpu_fusion_whole_run_vs_restatement(scene, tmp_path):
    from dvp_mvs_b200 import Fusion
    views = synth.make_fusion_views(scene, 1, seed=3)
    views[2]["block"] = (np.arange(views[2]["depth"].size).reshape(views[2]["depth"].shape) % 7 != 0).astype(np.uint8) * 255
    ref = FusionOracle(views).run()
    f = Fusion(views)
    pts, ms = f.run()
    assert ms > 0
    # whole-run agreement: identical unless an exp / acos ulp flipped a decision (then a handful of points differ)
    a = {tuple(p) for p in np.round(pts[:, :3].astype(np.float64), 6)}
    b = {tuple(p) for p in np.round(ref[:, :3].astype(np.float64), 6)}
    assert len(a ^ b) <= max(4, 1e-4 * len(b)), (len(a), len(b), len(a ^ b))
    pts2, _ = f.run()                                                    # deterministic: a second run is identical
    np.testing.assert_array_equal(pts, pts2)
    path = str(tmp_path / "fused.ply")
    f.write_ply(path)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert b"element vertex %d\n" % len(pts) in head and b"property uchar diffuse_blue" in head
    assert len(body) == 15 * len(pts)
    rec = np.frombuffer(body, np.dtype([("xyz", "<f4", 3), ("bgr", "u1", 3)]))
    np.testing.assert_array_equal(rec["xyz"], pts[:, :3])
    np.testing.assert_array_equal(rec["bgr"], pts[:, 3:].astype(np.uint8))
    f.close()


@pytest.mark.gpu
def test_gpu_fusion_argument_errors(scene):
    from dvp_mvs_b200 import Fusion, DvpError
    views = synth.make_fusion_views(scene, 0)
    bad = [dict(v) for v in views]
    bad[1]["src_views"] = [1, 2]                      # a view cannot be its own source
    with pytest.raises(DvpError):
        Fusion(bad)
    bad = [dict(v) for v in views]
    bad[0]["src_views"] = [7]
    with pytest.raises(DvpError):
        Fusion(bad)


@pytest.mark.gpu
def test_gpu_scene_to_fusion_handoff_equals_the_host_route():
    """N2 -> N3: maps handed over inside HBM (dvp_scene_fuse_views) fuse to the same points, bit for bit, as maps that
    went through the host the way the reference passes them through depths.dmb / APD_normals.dmb / weak.bin."""
    from dvp_mvs_b200 import Fusion, Scene
    mv = synth.make_multiview(320, 240, 3, 2, seed=5)
    V = len(mv.cameras)
    sc = Scene(V, mv.num_levels)
    for v in range(V):
        sc.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        for level in range(mv.num_levels):
            L = mv.levels[level][v]
            sc.set_level(v, level, L["image"], L["edge"], L["label"])
        sc.set_initial_planes(v, mv.planes_init[v])
    sc.run(seed=7)
    fine = mv.levels[-1]
    images = [np.stack([np.clip(L["image"], 0, 255)] * 3, -1).astype(np.uint8) for L in fine]
    f = Fusion.from_scene(sc, images)
    pts, _ = f.run()
    assert len(pts) > 500
    # host route: download, split (main.cpp:300-306), camera rescaled as RescaleImageAndCamera does
    views = []
    for v in range(V):
        planes, weak, _, _ = sc.get_view(v)
        h, w = weak.shape
        cam = synth.level_camera(mv.cameras[v], mv.full_w, mv.full_h, w, h, 2)
        views.append(dict(camera=cam, depth=planes[..., 3].copy(), normal=planes[..., :3].copy(), image=images[v], weak=weak,
                          src_views=mv.src_views[v]))
    g = Fusion(views)
    pts_host, _ = g.run()
    np.testing.assert_array_equal(pts, pts_host)
    # the plane-map entry point on host memory is the same split
    k = Fusion([dict(v, planes=np.concatenate([v["normal"], v["depth"][..., None]], -1)) for v in views])
    np.testing.assert_array_equal(k.run()[0], pts)
    # and the restatement agrees on what these maps fuse to (up to exp / acos ulps)
    ref = FusionOracle(views).run()
    assert abs(len(ref) - len(pts)) <= max(4, 1e-4 * len(ref))
    # the fused cloud lies on the synthetic room: its points are the views' own back-projections
    for x in (f, g, k, sc):
        x.close()


# ------------------------------------------------------------------------------------------- T&T variants (modes 1, 2)
def _tat_views(W=12, H=8):
    return [flat_view(W, H, 10.0, 4.0, [1, 2], grey=100), flat_view(W, H, 10.0, 4.0, [2, 0], grey=60),
            flat_view(W, H, 10.0, 4.0, [0, 1], grey=20)]


def test_tat_known_answers_and_the_stale_diff_vector():
    W, H = 12, 8
    for mode in (1, 2):
        # three identical views: every pixel of view 0 has two sources in exact agreement (k = 2 needs two) and masks
        # itself; views 1 and 2 then find view 0's pixels masked, are left with one source and emit nothing
        o = FusionOracle(_tat_views())
        pts, used = o.run_tat(mode)
        assert len(pts) == W * H and (used[0] == 3).all() and not used[1].any() and not used[2].any()
        assert o.masks[0].all() and not o.masks[1].any() and not o.masks[2].any()
        want = (100 + 60 + 20) / 3 if mode == 1 else 100.0          # mode 2 keeps the reference pixel's own colour
        assert (pts[:, 3] == np.float32(want)).all()
        # holes in source 2: a pixel that cannot evaluate a source keeps the measures of the last pixel (raster order) that
        # did — so it is still emitted — unless no pixel evaluated that source before it
        views = _tat_views()
        views[2]["depth"][0, 0:3] = 0.0       # start of the image: nothing to inherit
        views[2]["depth"][3, 5:8] = 0.0       # inherits from (3, 4)
        views[2]["depth"][5, 0] = 0.0         # inherits across the row end, from (4, W - 1)
        o = FusionOracle(views)
        pts, used = o.run_tat(mode)
        u0 = used[0].reshape(H, W)
        assert not u0[0, 0:3].any() and (u0[3, 5:8] == 3).all() and u0[5, 0] == 3
        assert (u0 != 0).sum() == W * H - 3
        if mode == 1:                           # ... and mode 1 averages in the colour of the INHERITED cell: all greys equal here
            assert (pts[:, 3] == np.float32(60.0)).all()
    # a single source can never satisfy k >= 2
    v = [flat_view(W, H, 10.0, 4.0, [1]), flat_view(W, H, 10.0, 4.0, [0])]
    assert len(FusionOracle(v).run_tat(1)[0]) == 0


def _gpu_tat(views, mode):
    from dvp_mvs_b200 import Fusion
    o = FusionOracle(views)
    ref, used_ref = o.run_tat(mode)
    f = Fusion(views)
    f.set_mode(mode)
    f.reset()
    used = []
    for v in range(len(views)):
        f.run_view(v)
        used.append(f.last_used(v))
    pts = f.points()
    masks = [f.mask(v) for v in range(len(views))]
    f.close()
    return pts, used, masks, ref, used_ref, o.masks


@pytest.mark.gpu
def test_gpu_tat_known_answers():
    for mode in (1, 2):
        views = _tat_views()
        views[2]["depth"][0, 0:3] = 0.0; views[2]["depth"][3, 5:8] = 0.0; views[2]["depth"][5, 0] = 0.0
        pts, used, masks, ref, used_ref, masks_ref = _gpu_tat(views, mode)
        np.testing.assert_array_equal(pts, ref)
        for a, b in zip(used, used_ref):
            np.testing.assert_array_equal(a, b)
        for a, b in zip(masks, masks_ref):
            np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
def test_gpu_tat_variants_vs_restatement(scene):
    for levels in (1, [1, 0, 1, 0]):
        views = synth.make_fusion_views(scene, levels, seed=2, depth_noise=0.0003, normal_noise=0.01)
        for v in views:
            v.pop("weak")                                       # the T&T variants never read weak.bin
        # mode 2 tests reprojection error and depth only (IEEE operations): bit-exact, points and masks
        pts, used, masks, ref, used_ref, masks_ref = _gpu_tat(views, 2)
        assert len(ref) > 1000
        np.testing.assert_array_equal(pts, ref)
        for a, b in zip(masks, masks_ref):
            np.testing.assert_array_equal(a, b)
        # mode 1 also tests the angle (acos): a decision may flip where an angle sits within an ulp of its limit
        pts, used, masks, ref, used_ref, masks_ref = _gpu_tat(views, 1)
        assert len(ref) > 1000
        flips = sum(int((a != b).sum()) for a, b in zip(used, used_ref))
        assert flips <= max(4, 1e-4 * len(ref)), flips
        if flips == 0:
            np.testing.assert_array_equal(pts, ref)


@pytest.mark.gpu
def test_gpu_fusion_all_variants_vs_the_reference_s_own_loops(scene):
    """The device path against the reference's fusing loops THEMSELVES (APD.cpp:1875-1957, 2028-2127, 2195-2276 compiled from
    /root/reference by oracle/Makefile, tests/test_ref_host.py): same points in the same order, same masks.  The ETH loop and
    T&T mode 1 go through exp / acos, where a value within an ulp of a threshold may flip a decision (bounded at 1e-4 of the
    points); T&T mode 2 is IEEE arithmetic only and must agree bit for bit."""
    import ref_host
    from dvp_mvs_b200 import Fusion
    if not ref_host.available():
        pytest.skip("oracle/_ref/libref_host.so not built")
    views = synth.make_fusion_views(scene, 1, seed=3)
    want, want_masks = ref_host.run_fusion(views)
    f = Fusion(views)
    pts, _ = f.run()
    a = {tuple(p) for p in np.round(pts[:, :3].astype(np.float64), 6)}
    b = {tuple(p) for p in np.round(want[:, :3].astype(np.float64), 6)}
    assert len(want) > 5000 and len(a ^ b) <= max(4, 1e-4 * len(b)), (len(a), len(b), len(a ^ b))
    if len(pts) == len(want) and not (a ^ b):
        assert (pts.view(np.uint32) == want.view(np.uint32)).all()
    f.close()
    tviews = synth.make_fusion_views(scene, 1, seed=2, depth_noise=0.0003, normal_noise=0.01)
    for v in tviews:
        v.pop("weak")
    for mode in (2, 1):
        want, want_masks = ref_host.run_fusion_tat(tviews, mode)
        f = Fusion(tviews); f.set_mode(mode)
        pts, _ = f.run()
        assert len(want) > 1000
        if mode == 2:
            np.testing.assert_array_equal(pts, want)
        else:
            a = {tuple(p) for p in np.round(pts[:, :3].astype(np.float64), 6)}
            b = {tuple(p) for p in np.round(want[:, :3].astype(np.float64), 6)}
            assert len(a ^ b) <= max(4, 1e-4 * len(b)), (mode, len(a), len(b), len(a ^ b))
        f.close()
