"""The CPU restatements of the callers' rows pinned to the reference ITSELF: oracle/ref_host.cu compiles line ranges cut
out of /root/reference/APD.cpp and main.cpp at build time against the stub cv::Mat — Roberts / Connect / Label_Seek /
Label_Update (APD.cpp:120-346), the geometry helpers and the three fusing loops (APD.cpp:501-546, 1797-1806, 1875-1957,
2028-2127, 2195-2276), the file readers and writers (APD.cpp:548-692, main.cpp:127-170), the level-size block and
RescaleMatToTargetSize (APD.cpp:1119-1143, 1773-1796), ProcessProblem's post-pass (main.cpp:282-363) and main()'s schedule
loop (main.cpp:450-512).  What is compared with them here is what the GPU paths are compared with in tests/test_visibility.py,
test_fusion.py, test_labels.py, test_scene.py and test_formats.py, so device parity for rows N1-N4 rests on compiled
reference code, not on hand-computed cases."""
import ctypes as C
import os

import numpy as np
import pytest

from util import ROOT
import ref_host
from fusion_oracle import FusionOracle
from dvp_mvs_b200 import synth
from test_visibility import oracle as visibility_restatement, random_masks, blobs

pytestmark = pytest.mark.skipif(not ref_host.available(), reason="oracle/_ref/libref_host.so not built (needs /root/reference at build time)")
CPU = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libapd_cpu.so"))
CPU.label_cpu_roberts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
CPU.label_cpu_connect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]


def test_built_from_the_reference_with_plain_flags():
    assert ref_host.flags() == "-O2"


def test_visibility_restatement_equals_the_reference_functions():
    """Row N1: Connect + Label_Update + the small-region rule (main.cpp:322-363) — restatement vs the reference's own
    functions on random masks of every density and on structured ones that reach the last row / column."""
    rng = np.random.default_rng(11)
    for trial in range(1500):
        H, W = int(rng.integers(2, 28)), int(rng.integers(2, 28))
        S = int(rng.integers(1, 4))
        sel = random_masks(rng, H, W, S, rng.choice([0.05, 0.1, 0.25, 0.4, 0.55, 0.7, 0.9]))
        scale = int(rng.choice([8, 8, 4]))
        assert (visibility_restatement(sel, S, scale, 0) == ref_host.restore_visibility(sel, S, scale)).all(), (trial, H, W)
    for trial in range(150):
        H, W = int(rng.integers(8, 48)), int(rng.integers(8, 48))
        sel = blobs(rng, H, W, 2, int(rng.integers(1, 8)), 9)
        sel[-1, ::2] &= ~np.uint32(1); sel[::3, -1] &= ~np.uint32(2); sel[-2, :] &= ~np.uint32(rng.integers(0, 4))
        assert (visibility_restatement(sel, 2, 8, 0) == ref_host.restore_visibility(sel, 2, 8)).all(), trial
    # and a map of realistic size and shape: holes in otherwise visible views (Label_Seek is quadratic in the label count)
    sel = blobs(rng, 120, 160, 3, 25, 7)
    assert (visibility_restatement(sel, 3, 8, 0) == ref_host.restore_visibility(sel, 3, 8)).all()


def test_roberts_and_labelling_restatements_equal_the_reference_functions():
    """Row N4, label half: Roberts (APD.cpp:120-136) and Connect + Label_Update (APD.cpp:195-346) as EdgeSegment(mode 1)
    uses them — same label NUMBERS and counts, not merely the same partition."""
    rng = np.random.default_rng(5)
    for trial in range(200):
        H, W = int(rng.integers(3, 40)), int(rng.integers(3, 40))
        img = rng.integers(0, 256, (H, W)).astype(np.uint8)
        mine = np.empty_like(img)
        CPU.label_cpu_roberts(img.ctypes.data, W, H, mine.ctypes.data)
        assert (mine == ref_host.roberts(img)).all(), trial
        edges = np.where(rng.random((H, W)) < rng.choice([0.1, 0.3, 0.5]), 255, 0).astype(np.uint8)
        labels = np.empty((H, W), np.int32); counts = np.zeros(H * W + 2, np.int32)
        n = CPU.label_cpu_connect(edges.ctypes.data, W, H, labels.ctypes.data, counts.ctypes.data, counts.size)
        want_labels, want_counts = ref_host.connect_update(edges)
        assert n == len(want_counts) and (labels == want_labels).all() and (counts[:n] == want_counts).all(), trial


@pytest.mark.parametrize("level", [0, 1])
def test_fusion_restatement_equals_the_reference_loop(level):
    """Row N3: RunFusion's fusing loop compiled from the reference (-O2) vs oracle/cpu/fusion_cpu.cpp — the same points
    (coordinates and colours bit for bit, same order) and the same masks."""
    mv = synth.make_multiview(640, 480, 4, 2, seed=3)
    views = synth.make_fusion_views(mv, level)
    o = FusionOracle(views)
    mine = o.run()
    want, want_masks = ref_host.run_fusion(views)
    assert len(mine) == len(want) > 5000
    assert (mine.view(np.uint32) == want.view(np.uint32)).all()
    for a, b in zip(o.masks, want_masks):
        assert (a == b).all()


def test_fusion_restatement_on_views_of_mixed_sizes_and_blocks():
    """Source cells shared by several pixels (views at different resolutions) and a block mask: the order-dependent part."""
    mv = synth.make_multiview(320, 240, 3, 2, seed=9)
    v0 = synth.make_fusion_views(mv, 0); v1 = synth.make_fusion_views(mv, 1)
    views = [v1[0], v0[1], v1[2]]
    blk = np.full(views[0]["depth"].shape, 255, np.uint8); blk[:, :40] = 0
    views[0] = dict(views[0], block=blk)
    o = FusionOracle(views)
    mine = o.run()
    want, want_masks = ref_host.run_fusion(views)
    assert len(mine) == len(want) > 1000 and (mine.view(np.uint32) == want.view(np.uint32)).all()
    assert all((a == b).all() for a, b in zip(o.masks, want_masks))


@pytest.mark.skipif(not ref_host.available(fast=True), reason="fast-math build absent")
def test_reference_host_flags_do_not_change_fusion_decisions_here():
    """The reference's CMake builds APD.cpp with -O3 -ffast-math -march=native (CMakeLists.txt:31): FMA contraction and
    reassociation move coordinates by an ulp or two and could flip a decision that sits on a threshold.  Measured on this
    scene: how far apart the two builds of the SAME reference lines are — the bound on what 'parity' can mean for row N3."""
    try:
        fast_flags = ref_host.flags(fast=True)
    except OSError:
        pytest.skip("fast build not loadable on this host")
    assert "ffast-math" in fast_flags
    mv = synth.make_multiview(640, 480, 4, 2, seed=3)
    views = synth.make_fusion_views(mv, 1)
    a, ma = ref_host.run_fusion(views)
    b, mb = ref_host.run_fusion(views, fast=True)
    flipped = sum(int((x != y).sum()) for x, y in zip(ma, mb))
    assert abs(len(a) - len(b)) <= max(2, len(a) // 10000) and flipped <= max(8, len(a) // 2500), (len(a), len(b), flipped)
    if len(a) == len(b):
        assert np.abs(a[:, :3] - b[:, :3]).max() < 1e-3


@pytest.mark.parametrize("mode", [1, 2])
def test_tat_fusion_restatement_equals_the_reference_loops(mode):
    """Row N3, the two Tanks-and-Temples variants: RunFusion_TAT_Intermediate (APD.cpp:2028-2127) and RunFusion_TAT_advanced
    (APD.cpp:2195-2276) compiled from the reference vs oracle/cpu/fusion_cpu.cpp — the same points (bit for bit, same order)
    and the same masks, including the `diff` vector the reference declares once per view and therefore carries from pixel
    to pixel (a source a pixel cannot evaluate keeps the measures of the last pixel that could)."""
    for level, seed in ((1, 3), (0, 5)):
        mv = synth.make_multiview(640, 480, 4, 2, seed=seed)
        views = synth.make_fusion_views(mv, level)
        o = FusionOracle(views)
        mine, _ = o.run_tat(mode)
        want, want_masks = ref_host.run_fusion_tat(views, mode)
        assert len(mine) == len(want) > 1000, (len(mine), len(want))
        assert (mine.view(np.uint32) == want.view(np.uint32)).all()
        for a, b in zip(o.masks, want_masks):
            assert (a == b).all()


def test_tat_fusion_restatement_with_block_masks_and_mixed_sizes():
    """Views at different resolutions and a block mask, both T&T variants."""
    mv = synth.make_multiview(320, 240, 3, 2, seed=9)
    v0 = synth.make_fusion_views(mv, 0); v1 = synth.make_fusion_views(mv, 1)
    views = [v1[0], v0[1], v1[2]]
    blk = np.full(views[1]["depth"].shape, 255, np.uint8); blk[:30, :] = 0
    views[1] = dict(views[1], block=blk)
    for mode in (1, 2):
        o = FusionOracle(views)
        mine, _ = o.run_tat(mode)
        want, want_masks = ref_host.run_fusion_tat(views, mode)
        assert len(mine) == len(want) > 50 and (mine.view(np.uint32) == want.view(np.uint32)).all(), mode
        assert all((a == b).all() for a, b in zip(o.masks, want_masks)), mode


def test_rescale_restatement_equals_the_reference_function():
    """Row N2: RescaleMatToTargetSize (APD.cpp:1773-1796, scale factors swapped — SURVEY B10) compiled from the reference vs
    oracle/host_chain.rescale_ref, the restatement the scene driver's device rescale is compared with: every element type the
    reference instantiates, up- and down-scaling at the ratios of the schedule and at awkward ones."""
    import host_chain
    rng = np.random.default_rng(11)
    sizes = [((778, 518), (1555, 1037)), ((1555, 1037), (3111, 2073)), ((97, 33), (194, 67)), ((194, 67), (97, 33)), ((64, 48), (64, 48)),
             ((31, 57), (100, 41)), ((120, 90), (37, 111)), ((5, 3), (9, 7))]
    for (sw, sh), (dw, dh) in sizes:
        maps = [rng.integers(0, 255, (sh, sw)).astype(np.uint8), rng.random((sh, sw), dtype=np.float32),
                rng.random((sh, sw, 3), dtype=np.float32), rng.integers(0, 2 ** 32, (sh, sw), dtype=np.uint64).astype(np.uint32),
                rng.integers(-50, 50, (sh, sw)).astype(np.int32)]
        for m in maps:
            want = ref_host.rescale(m, dw, dh)
            mine = host_chain.rescale_ref(m, dw, dh)
            assert mine.shape == want.shape and (mine.view(np.uint8) == want.view(np.uint8)).all(), (sw, sh, dw, dh, m.dtype)


def test_on_disk_formats_against_the_reference_s_own_readers_and_writers(tmp_path):
    """Row N4, on-disk formats: dvp_io_* (csrc/dvp_io.cpp) against ReadBinMat / WriteBinMat / writeDepthDmb / writeNormalDmb /
    ReadCamera (APD.cpp:548-692) and GenerateSampleList (main.cpp:127-170) compiled from the reference: files written by one
    side are byte-identical to the other's and read back to the same values by both."""
    from dvp_mvs_b200 import formats
    rng = np.random.default_rng(4)
    mats = [rng.integers(0, 3, (37, 53)).astype(np.uint8), rng.random((19, 64), dtype=np.float32), rng.random((12, 9, 3), dtype=np.float32),
            rng.integers(0, 2 ** 31, (8, 33)).astype(np.int32)]
    for i, m in enumerate(mats):
        theirs, ours = tmp_path / f"ref{i}.bin", tmp_path / f"dvp{i}.bin"
        ref_host.write_binmat(str(theirs), m)
        formats.write_binmat(str(ours), m)
        assert theirs.read_bytes() == ours.read_bytes(), i
        got = formats.read_binmat(str(theirs))                     # their file, our reader
        assert got.dtype == m.dtype and got.shape == m.shape and (got == m).all(), i
        r, c, t, raw = ref_host.read_binmat(str(ours))             # our file, their reader
        assert (r, c) == m.shape[:2] and raw.tobytes() == m.tobytes(), i
    bad = tmp_path / "bad.bin"
    bad.write_bytes(np.array([2, 1, 1, 0], np.int32).tobytes() + b"\0")
    assert ref_host.read_binmat(str(bad)) is None                  # "Version error" on both sides
    with pytest.raises(Exception):
        formats.read_binmat(str(bad))
    depth = rng.random((23, 31), dtype=np.float32); normal = rng.random((23, 31, 3), dtype=np.float32)
    for name, a in (("d", depth), ("n", normal)):
        ref_host.write_dmb(str(tmp_path / f"ref_{name}.dmb"), a)
        formats.write_dmb(str(tmp_path / f"dvp_{name}.dmb"), a)
        assert (tmp_path / f"ref_{name}.dmb").read_bytes() == (tmp_path / f"dvp_{name}.dmb").read_bytes(), name
    # cameras: awkward decimals, exponents, negative values; the centre -R^T t is accumulated in double by both
    for k in range(20):
        R = rng.normal(size=(3, 3)); t = rng.normal(size=3) * 10; K = np.array([[3000 + rng.random(), 0, 1500.123 + k], [0, 2999.5 + rng.random(), 1000.7], [0, 0, 1]])
        txt = "extrinsic\n" + "".join(" ".join(repr(float(v)) for v in list(R[j]) + [t[j]]) + "\n" for j in range(3)) + "0.0 0.0 0.0 1.0\n\nintrinsic\n"
        txt += "".join(" ".join(f"{v:.9e}" if k % 2 else repr(float(v)) for v in K[j]) + "\n" for j in range(3)) + f"\n{0.37 + k} {0.0123} {192} {9.5 + k}\n"
        p = tmp_path / f"{k:08d}_cam.txt"
        p.write_text(txt)
        theirs, ours = ref_host.read_camera(str(p)), formats.read_camera(str(p))
        for f in ("K", "R", "t", "c", "depth_min", "depth_max"):
            assert (np.asarray(theirs[f]).view(np.uint32) == np.asarray(ours[f]).view(np.uint32)).all(), (k, f)
    (tmp_path / "pair.txt").write_text("4\n0\n3 1 0.75 2 0.5 7 -1.0\n4\n2 0 2.5 2 0.0\n2\n0\n9\n10 1 1 2 1 3 1 4 1 5 1 6 1 7 1 8 1 10 1 11 0.001\n")
    assert ref_host.read_pairs(str(tmp_path)) == formats.read_pairs(str(tmp_path / "pair.txt"))


def test_level_sizes_and_cameras_equal_the_reference_block():
    """Row N2: the level size (round(size / scale) in float) and the intrinsics scaled by the ratios of the ROUNDED sizes —
    InuputInitialization's block APD.cpp:1119-1143 compiled from the reference vs the restatement (oracle/host_chain.py) and
    vs the library's own dvp_scene_level_size, over the dataset sizes and awkward ones, scales 1 to 16."""
    import host_chain
    from dvp_mvs_b200 import Scene
    rng = np.random.default_rng(2)
    cam = np.zeros((), synth.CAMERA_DTYPE)
    sizes = [(6221, 4146), (6048, 4032), (1920, 1080), (3111, 2073), (1555, 1037), (101, 67), (33, 35), (5, 3), (4097, 2049)]
    for full_w, full_h in sizes:
        cam["K"] = np.array([3409.7 + rng.random(), 0, full_w / 2 + rng.random(), 0, 3411.3 + rng.random(), full_h / 2 + rng.random(), 0, 0, 1], np.float32)
        cam["depth_min"], cam["depth_max"] = 0.5, 9.0
        for levels in (1, 2, 3, 4):
            sc = Scene(2, levels, device=-1)            # host-only scene: schedule and size queries need no GPU
            for level in range(levels):
                scale = host_chain.level_scale(levels, level)
                want, w, h = ref_host.level_camera(cam, full_w, full_h, scale)
                assert (w, h) == host_chain.level_size(full_w, full_h, scale) == tuple(sc.level_size(full_w, full_h, level)), (full_w, full_h, scale)
                mine = host_chain.level_camera(cam, full_w, full_h, w, h, scale)
                assert (np.asarray(mine["K"]).view(np.uint32) == np.asarray(want["K"]).view(np.uint32)).all(), (full_w, full_h, scale)
                assert int(want["width"]) == w and int(want["height"]) == h
            sc.close()


@pytest.mark.parametrize("num_levels", [1, 2, 3, 4])
def test_schedule_equals_main_s_own_loop(num_levels):
    """Row N2: main()'s rounds x passes x views loop (main.cpp:450-512) compiled from the reference, with ProcessProblem and
    GetProblemEdges replaced by recorders, against the library's dvp_scene_pass_params / dvp_scene_run order and the
    restatement (oracle/host_chain.py): which view runs when, at which scale, with every PatchMatchParams field — float
    fields bit for bit (ransac_threshold is 0.01 - i * 0.00125 evaluated in double and narrowed)."""
    import host_chain
    from dvp_mvs_b200 import Scene
    V = 3
    recs = ref_host.schedule(num_levels + 1, V)
    assert len(recs) == num_levels * 4 * V
    sc = Scene(V, num_levels, device=-1)
    k = 0
    for level in range(num_levels):
        for pass_ in range(4):
            mine = sc.pass_params(level, pass_)
            restated = host_chain.schedule_params(num_levels, level, pass_)
            for v in range(V):                       # every pass visits the views in index order (Gauss-Seidel over the views)
                r = recs[k]; k += 1
                assert (r["view"], r["iteration"], r["scale_size"]) == (v, level * 4 + pass_, host_chain.level_scale(num_levels, level))
                assert r["edges"] == (1 if pass_ == 0 else 0)          # GetProblemEdges precedes the INIT pass of every round
                for name in ref_host.SCHEDULE_FIELDS:
                    if name in ("depth_min", "depth_max", "num_images"):   # set per view by InuputInitialization, not by the schedule
                        continue
                    want = r["params"][name]
                    for who, p in (("library", mine), ("restatement", restated)):
                        got = getattr(p, name)
                        if name in ("sigma_spatial", "sigma_color", "ransac_threshold", "geom_factor"):
                            assert np.float32(got).view(np.uint32) == np.float32(want).view(np.uint32), (who, level, pass_, name, got, want)
                        else:
                            assert int(got) == int(want), (who, level, pass_, name, got, want)
    sc.close()


def test_post_pass_equals_process_problem_s_own_lines():
    """Row N1 as ProcessProblem runs it: main.cpp:282-363 compiled from the reference (against an object answering the APD
    getters it calls) — out-of-range depths zeroed and marked UNKNOWN, then per source view the visibility restoration —
    vs the restatement the device path is compared with (oracle/cpu/visibility_cpu.cpp) and vs the hand-wrapped loop
    (refhost_restore_visibility) the other tests of this file use."""
    from dvp_mvs_b200 import UNKNOWN
    rng = np.random.default_rng(21)
    for trial in range(300):
        H, W = int(rng.integers(4, 40)), int(rng.integers(4, 40))
        S = int(rng.integers(1, 5))
        scale = int(rng.choice([8, 4, 2, 1]))
        sel = blobs(rng, H, W, S, int(rng.integers(1, 8)), 7) if trial % 2 else random_masks(rng, H, W, S, rng.choice([0.1, 0.4, 0.7, 0.9]))
        planes = rng.normal(size=(H, W, 4)).astype(np.float32)
        planes[..., 3] = rng.uniform(0.0, 12.0, (H, W)).astype(np.float32)
        planes[rng.random((H, W)) < 0.05, 3] = np.float32(0.0)
        states = rng.integers(0, 3, (H, W)).astype(np.uint8)
        dmin, dmax = np.float32(1.5), np.float32(9.25)
        depth, st, se = ref_host.post_pass(planes, states, sel, S, scale, dmin, dmax)
        bad = (planes[..., 3] < dmin) | (planes[..., 3] > dmax)
        want_depth = planes[..., 3].copy(); want_depth[bad] = 0
        want_states = states.copy(); want_states[bad] = UNKNOWN
        assert (depth.view(np.uint32) == want_depth.view(np.uint32)).all() and (st == want_states).all(), trial
        assert (se == visibility_restatement(sel, S, scale, 0)).all(), (trial, H, W, S, scale)
        assert (se == ref_host.restore_visibility(sel, S, scale)).all(), trial


def test_edge_segment_glue_from_the_reference_s_own_lines():
    """Row N4, both halves: EdgeSegment (APD.cpp:348-499) compiled from the reference — histogram median and thresholds,
    weak_tex_num, the region loop with its border extraction, the order the Hough segments are drawn in, the border clean-up
    and the final labelling rule are the reference's code; the five OpenCV calls it makes go to the restated primitives, each
    pinned against real OpenCV 4.13 output.  Its results equal the OpenCV-made golden vectors and the restatements the device
    paths are compared with (edge map: oracle/cpu/edge_cpu.cpp; label map: oracle/cpu/label_cpu.cpp), label numbers included."""
    golden = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(golden, "edge_canny.npz"))
    k = 0
    while f"image_{k}" in g:
        np.testing.assert_array_equal(ref_host.edge_segment(g[f"image_{k}"], 1, 0), g[f"edge_{k}"])
        k += 1
    assert k >= 5
    g = np.load(os.path.join(golden, "label_segment.npz"))
    for i in range(int(g["count"])):
        img = g[f"image_{int(g[f'image_of_{i}'])}"]
        np.testing.assert_array_equal(ref_host.edge_segment(img, int(g[f"scale_{i}"]), 1), g[f"labels_{i}"])
    # renders of the synthetic scene at other sizes and scales: against the restatements
    CPU.label_cpu_segment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    CPU.label_cpu_size.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    CPU.edge_cpu_segment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    for (W, H, scale) in ((320, 240, 1), (401, 303, 2), (640, 480, 3)):
        img = np.ascontiguousarray(np.clip(np.rint(synth.make_scene(W, H, 2, seed=W).images[0]), 0, 255).astype(np.uint8))
        edge = np.empty((H, W), np.uint8)
        assert CPU.edge_cpu_segment(img.ctypes.data, W, H, edge.ctypes.data, None) == 0
        np.testing.assert_array_equal(ref_host.edge_segment(img, scale, 0), edge)
        nc, nr = C.c_int(), C.c_int()
        CPU.label_cpu_size(W, H, scale, C.byref(nc), C.byref(nr))
        labels = np.empty((nr.value, nc.value), np.int32)
        assert CPU.label_cpu_segment(img.ctypes.data, W, H, scale, labels.ctypes.data, None) == 0
        np.testing.assert_array_equal(ref_host.edge_segment(img, scale, 1), labels)


def test_input_assembly_equals_the_reference_s_own_lines(tmp_path):
    """Row N2: how a pass's inputs are put together from the previous pass's files — InuputInitialization (APD.cpp:1147-1205,
    1426-1493) and SupportInitialization (APD.cpp:1615-1668) compiled from the reference, reading real files written by the
    reference's own WriteBinMat — vs oracle/host_chain.assemble_inputs, which the host-chain restatement (and through it the
    resident scene driver) is compared with: same size and level change (the swapped-factor rescale), REFINE_INIT and
    REFINE_ITER, with and without the adaptive-patch states."""
    import host_chain
    from dvp_mvs_b200 import default_params, FIRST_INIT, REFINE_INIT, REFINE_ITER, WEAK, UNKNOWN
    rng = np.random.default_rng(8)
    src_ids = [2, 0, 3]
    for (pw, ph), (w, h) in (((97, 61), (97, 61)), ((97, 61), (194, 122)), ((80, 60), (161, 119))):
        files = {}
        for vid in [1] + src_ids:
            planes = rng.normal(size=(ph, pw, 4)).astype(np.float32); planes[..., 3] = rng.uniform(0.5, 9.0, (ph, pw)).astype(np.float32)
            files[vid] = dict(planes=planes, weak=rng.integers(0, 3, (ph, pw)).astype(np.uint8),
                              selected=rng.integers(0, 8, (ph, pw)).astype(np.uint32), radius=rng.integers(0, 11, (ph, pw)).astype(np.int32))
            d = tmp_path / f"{pw}x{ph}_{w}" / "APD" / f"{vid:08d}"
            d.mkdir(parents=True, exist_ok=True)
            ref_host.write_binmat(str(d / "depths.dmb"), np.ascontiguousarray(planes[..., 3]))
            ref_host.write_binmat(str(d / "APD_normals.dmb"), np.ascontiguousarray(planes[..., :3]))
            ref_host.write_binmat(str(d / "weak.bin"), files[vid]["weak"])
            ref_host.write_binmat(str(d / "selected_views.bin"), files[vid]["selected"])
            ref_host.write_binmat(str(d / "radius.bin"), files[vid]["radius"])
            ref_host.write_binmat(str(d / "edges_1.dmb"), rng.integers(0, 2, (h, w)).astype(np.uint8) * 255)
        dense = str(tmp_path / f"{pw}x{ph}_{w}")
        for state, geom, use_apd in ((REFINE_INIT, 0, 1), (REFINE_ITER, 1, 1), (REFINE_ITER, 1, 0)):
            got = ref_host.assemble(dense, 1, src_ids, 2, w, h, state, geom, use_apd)
            p = default_params(); p.state = state; p.geom_consistency = geom; p.use_APD = use_apd; p.use_radius = 1
            mine = host_chain.assemble_inputs(files[1], [files[s] for s in src_ids], w, h, p)
            assert (mine["planes"].view(np.uint32) == got["planes"].view(np.uint32)).all(), (pw, w, state)
            assert (mine["selected"] == got["selected"]).all()
            if geom:
                assert (mine["depths"].view(np.uint32) == got["depths"].view(np.uint32)).all()
            weak = mine["weak"] if use_apd else np.ones((h, w), np.uint8)      # no adaptive patches: every pixel STRONG (APD.cpp:1196-1204)
            assert (weak == got["weak"]).all() and got["weak_count"] == (int((weak == WEAK).sum()) if use_apd else 0)
            want_radius = mine["radius"].copy(); want_radius[weak == UNKNOWN] = p.strong_radius     # APD.cpp:1663-1667 (the engine's upload)
            assert (want_radius == got["radius"]).all()
            assert got["edge_size"] == (w, h)                                 # edges_<scale>.dmb is taken as it is, at the pass's size


def test_get_problem_edges_from_the_reference_s_own_lines(tmp_path):
    """Rows N2 / N4 as main() prepares them: GetProblemEdges (main.cpp:193-246) compiled from the reference — the grey image
    goes to float, is resized to the level with cv::resize(INTER_LINEAR), comes back to 8 bits (saturate_cast), and THAT image
    goes through EdgeSegment(scale, ., 0, true) into edges_<scale>.dmb, while the FULL image goes through
    EdgeSegment(scale, ., 1) into labels_<scale>.dmb — vs the restatements dvp_scene_set_image is compared with
    (oracle/image_oracle.py, edge_cpu.cpp, label_cpu.cpp).  ComputeRoundNum (main.cpp:248-265) rides along."""
    import image_oracle
    from dvp_mvs_b200 import formats
    CPU.label_cpu_segment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    CPU.label_cpu_size.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    CPU.edge_cpu_segment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    for (W, H, scale_size) in ((640, 480, 2), (801, 603, 4), (333, 250, 2), (640, 480, 1)):
        img = np.ascontiguousarray(np.clip(np.rint(synth.make_scene(W, H, 2, seed=H).images[0]), 0, 255).astype(np.uint8))
        dense = tmp_path / f"{W}x{H}_{scale_size}"
        (dense / "APD" / "00000007").mkdir(parents=True)
        ref_host.get_problem_edges(str(dense), 7, scale_size, img)
        scale = {1: 0, 2: 1, 4: 2}[scale_size]
        lw, lh = image_oracle.level_size(W, H, scale_size)
        level = np.rint(image_oracle.resize_linear_f32(img.astype(np.float32), lw, lh)).clip(0, 255).astype(np.uint8) if scale_size != 1 else img
        edge = np.empty((lh, lw), np.uint8)
        assert CPU.edge_cpu_segment(np.ascontiguousarray(level).ctypes.data, lw, lh, edge.ctypes.data, None) == 0
        np.testing.assert_array_equal(formats.read_binmat(str(dense / "APD" / "00000007" / f"edges_{scale}.dmb")), edge)
        nc, nr = C.c_int(), C.c_int()
        CPU.label_cpu_size(W, H, scale, C.byref(nc), C.byref(nr))
        labels = np.empty((nr.value, nc.value), np.int32)
        assert CPU.label_cpu_segment(img.ctypes.data, W, H, scale, labels.ctypes.data, None) == 0
        np.testing.assert_array_equal(formats.read_binmat(str(dense / "APD" / "00000007" / f"labels_{scale}.dmb")), labels)
    for (cols, rows), want in (((6221, 4146), 4), ((3111, 2073), 3), ((1920, 1080), 3), ((800, 600), 1), ((801, 600), 2), ((1601, 40), 2), ((600, 3300), 4)):
        assert ref_host.compute_round_num(3, cols, rows) == want, (cols, rows)        # halve the longer side until it is <= 800
    assert ref_host.compute_round_num(0, 640, 480) == 0
