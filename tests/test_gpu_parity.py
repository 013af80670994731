"""GPU parity tests (-m gpu): the product's sm_100a kernels, called through the C ABI, against
  (1) the reference's own CUDA kernels (oracle/_ref, unmodified APD.cu) stage by stage from identical state,
  (2) golden vectors of those kernels committed under tests/golden (so the check survives without the library),
  (3) the CPU restatement where the arithmetic is exact, and
  (4) size-independent properties at BASELINE.json's full size.
Bar: bit-exact on every race-free stage; 1e-4 relative for floats is therefore met with margin.  The
reference's strong sweep (K7/K8) contains a same-colour read race: on the full image both
implementations are compared against the reference's own run-to-run noise floor, and bit-exactly on a
sparse STRONG mask that makes the race harmless."""
import numpy as np
import pytest

from util import c1_params, load_golden, golden_inputs, close, per_pixel
from dvp_mvs_b200 import Engine, synth, DvpError, STRONG, WEAK, UNKNOWN
from dvp_mvs_b200.parity import step_compare, lockstep_compare, sequence, STAGE_OUTPUTS, STATE_BUFS, compare
import ref_oracle
import cpu_oracle

pytestmark = pytest.mark.gpu

RACE_FREE = ("K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM", "K3_FIND_NEAREST_STRONG", "K5_NEIGHBOUR_UPDATE",
             "K6_RANDOM_INITIALIZATION", "K9_RANSAC_FIT_PLANE", "K12_DEPTH_NORMAL", "K13_BLACK_FILTER", "K14_RED_FILTER",
             "K15_DEPTH_TO_WEAK", "K16_LOCAL_REFINE")
needs_ref = pytest.mark.skipif(not ref_oracle.available(), reason="oracle/_ref/libapd_ref.so not built")


@pytest.fixture(scope="module")
def c1_scene():
    return synth.make_scene(640, 480, 2)   # BASELINE.json configs[0]: 640x480, 2 source views, 1 iteration


def c1_kwargs(sc):
    return dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)


@needs_ref
def test_c1_every_race_free_stage_is_bit_exact_vs_reference(c1_scene):
    sc = c1_scene
    p = c1_params(sc.depth_min, sc.depth_max, 2)
    ref = ref_oracle.engine(640, 480, 2, p); prod = Engine(640, 480, 2, p)
    ref.upload(**c1_kwargs(sc)); prod.upload(**c1_kwargs(sc))
    res = step_compare(ref, prod, 1, stages=RACE_FREE)
    assert len(res) >= 14
    bad = [r for r in res if r.get("error") or r["not_bit_exact"]]
    assert not bad, bad[:4]


def _d4_ladder_offsets(W, H):
    """Every offset m (pixels along the diagonal, after the fixed 5) at which direction 4 of the strong sweep can look:
    the edge-adaptive ladder (APD.cu:2053-2087; the distance, already divided by sqrt 2, is compared with the UNdivided
    limit max(H, W) / 30; step count 11..22, step length >= 2) and the fixed 11 x 2 px one (APD.cu:2105-2110)."""
    out = set(range(0, 22, 2))
    max_edge = max(H, W) / 30.0
    dists = list(np.arange(0.0, max_edge + 0.26, 0.125)) + [22.0, max_edge, max_edge / np.sqrt(2.0)]
    for dist in dists:
        step_num = min(max(11, int(dist / 2)), 22)
        step_len = max(int(dist / step_num), 2)
        out.update(k * step_len for k in range(step_num))
    return sorted(out)


def _race_exposed(unexplained, offsets, pre, post):
    """Of the given pixels: which have, on their direction-4 ladder, a pixel whose plane or cost the launch changed?"""
    H, W = pre["costs"].shape
    changed = (pre["planes"].view(np.uint32) != post["planes"].view(np.uint32)).any(-1) | (pre["costs"].view(np.uint32) != post["costs"].view(np.uint32))
    ys, xs = np.nonzero(unexplained)
    exposed = np.zeros(len(ys), bool)
    for m in offsets:
        qx, qy = xs - 5 - m, ys - 5 - m
        ok = (qx >= 0) & (qy >= 0)
        exposed[ok] |= changed[qy[ok], qx[ok]]
    return exposed


def _assert_race_explained(prod, pre, outs, stage, red, W, H, runs, log, slack=None):
    """Every pixel of every observed run must be reproduced by some outcome of the reference's race (dvp_debug_race_explain);
    what the enumeration does not reach (interleavings finer than the per-component / per-view reads it tries of the plane
    being replaced) may leave at most one pixel in 10 000 (one in 1 000 when only whole-plane reads are tried), and each of
    those must have a direction-4 ladder pixel that the launch rewrote — a pixel whose inputs were stable must be reproduced
    exactly.  runs: (name, observed outputs, try torn reads too)."""
    offsets = _d4_ladder_offsets(W, H)
    report = {}
    for who, observed, tear in runs:
        for n, a in pre.items():
            prod.set(n, a)                              # the context holds the pre-launch state; race_explain leaves it alone
        explained, (left1, left2, launches) = prod.race_explain(0, red, offsets, pre["planes"], observed, tear=tear)
        left = ~explained
        report[who] = dict(after_whole_plane_choices=left1, after_torn_and_per_view_reads=left2, forced_launches=launches)
        assert int(left.sum()) == left2
        assert left2 <= (slack if slack is not None else max(4, W * H // (10000 if tear else 1000))), (who, report)
        if left2:
            # Unexplained AND not even exposed to the race: pixels whose inputs were stable and which still differ from every
            # enumerated choice.  Measured at 3111x2073: 15 of 3.2 M processed pixels, all on the grazing side wall, the plane
            # offset or one candidate's cost 1 ulp apart (a contraction the reference's compiler chose differently for a term
            # that almost never reaches the last bit) and then amplified by the view sampling.  Bounded at 5 per million.
            unexposed = int((~_race_exposed(left, offsets, pre, observed)).sum())
            report[who]["unexplained_and_not_race_exposed"] = unexposed
            assert unexposed <= max(2, W * H // 200000), (who, report)
    log(f"[race] {stage} {W}x{H}: {report}")
    return report


@needs_ref
@pytest.mark.parametrize("stage,red", [("K7_BLACK_STRONG", 0), ("K8_RED_STRONG", 1)])
def test_full_image_sweep_differences_are_exactly_the_direction4_race(c1_scene, stage, red):
    """K7/K8 on the full image.  The reference disagrees with ITSELF between two runs from the same state: direction 4 of
    its ladder reads cost and plane of pixels of the colour being written (APD.cu:2039, 2071-2074).  Whatever the
    timing, that direction hands the pixel ONE candidate — the plane of a ladder pixel at some offset m, read for scoring,
    for the depth test and for the copy at acceptance, each time before or after that pixel's own update (or torn between
    the two: the reference loads planes with 32-bit loads).  dvp_debug_race_explain imposes every such choice with the
    production arithmetic; a pixel is `explained` when some choice reproduces all five of its output buffers bit for bit.
    EVERY pixel of the reference's racy run, and of ours, must be explained: all differences between the two are the
    reference's own race and nothing else.  This replaces a tolerance (round 1 accepted 20x the reference's own noise)."""
    sc = c1_scene
    W, H = 640, 480
    p = c1_params(sc.depth_min, sc.depth_max, 2)
    ref = ref_oracle.engine(W, H, 2, p); prod = Engine(W, H, 2, p)
    ref.upload(**c1_kwargs(sc)); prod.upload(**c1_kwargs(sc))
    for st in sequence(1)[:6 + red]:
        ref.run_stage(*st)
    pre = {n: ref.get(n) for n in STATE_BUFS}
    outs = STAGE_OUTPUTS[stage]

    def run(e):
        for n, a in pre.items():
            e.set(n, a)
        e.run_stage(stage, 0)
        return {n: e.get(n) for n in outs}
    r1, r2, p1 = run(ref), run(ref), run(prod)
    noise = max(compare(n, r1[n], r2[n])["frac"] for n in outs)
    ours = max(compare(n, r1[n], p1[n])["frac"] for n in outs)
    assert ours < 0.02, (ours, noise)
    assert ours > 0 or noise == 0                      # the race is real: the runs do differ somewhere
    rep = _assert_race_explained(prod, pre, outs, stage, red, W, H, (("reference", r1, True), ("reference again", r2, False), ("ours", p1, False)), print)
    print(f"[race] {stage}: reference vs itself {100 * noise:.3f} % of pixels differ, ours vs reference {100 * ours:.3f} %")
    assert rep["ours"]["after_torn_and_per_view_reads"] == 0      # our own kernel reads planes whole (LDG.128): nothing is left over


@needs_ref
def test_sparse_mask_sweep_is_bit_exact_vs_reference():
    """Race-free K7/K8 (no two processed pixels interact on a sparse STRONG mask): every output buffer identical, two iterations."""
    W, H, S = 640, 480, 2
    sc = synth.make_scene(W, H, S)
    p = c1_params(sc.depth_min, sc.depth_max, S, iters=2, use_apd=1)
    yy, xx = np.mgrid[0:H, 0:W]
    weak = np.where(((xx + yy) % 128) < 2, STRONG, WEAK).astype(np.uint8)
    kw = dict(weak_info=weak, **c1_kwargs(sc))
    ref = ref_oracle.engine(W, H, S, p); prod = Engine(W, H, S, p)
    ref.upload(**kw); prod.upload(**kw)
    for st in ("K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM", "K6_RANDOM_INITIALIZATION"):
        ref.run_stage(st)
    n_changed = 0
    for it in range(2):
        for st in ("K7_BLACK_STRONG", "K8_RED_STRONG"):
            pre = {n: ref.get(n) for n in STATE_BUFS}
            ref.run_stage(st, it)
            for n, a in pre.items():
                prod.set(n, a)
            prod.run_stage(st, it)
            for n in STAGE_OUTPUTS[st]:
                a, b = ref.get(n), prod.get(n)
                r = compare(n, a, b)
                assert r["not_bit_exact"] == 0, (st, it, r)
            n_changed += int((ref.get("costs") != pre["costs"]).sum())
    assert n_changed > 4000   # the sweep really did something


@needs_ref
@pytest.mark.parametrize("S", [8, 16])
def test_many_source_views_strong_path_vs_reference(S):
    """BASELINE config C5's view counts (8 and 16 source views; photometric, every pixel STRONG — the reference's WEAK
    path is undefined beyond 4 views): race-free stages bit-exact, then the sweep on a sparse STRONG mask."""
    W, H = 200, 152
    sc = synth.make_scene(W, H, S)
    p = c1_params(sc.depth_min, sc.depth_max, S)
    ref = ref_oracle.engine(W, H, S, p); prod = Engine(W, H, S, p)
    kw = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    ref.upload(**kw); prod.upload(**kw)
    res = step_compare(ref, prod, 1, stages=RACE_FREE)
    bad = [r for r in res if r.get("error") or r["not_bit_exact"]]
    assert len(res) >= 14 and not bad, bad[:4]
    # sparse STRONG mask: no two processed pixels interact, so the sweep is race-free and must match bit for bit
    q = c1_params(sc.depth_min, sc.depth_max, S, iters=1, use_apd=1)
    yy, xx = np.mgrid[0:H, 0:W]
    weak = np.where(((xx + yy) % 64) < 2, STRONG, WEAK).astype(np.uint8)
    ref.upload(weak_info=weak, params=q, **kw); prod.upload(weak_info=weak, params=q, **kw)
    for st in ("K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM", "K6_RANDOM_INITIALIZATION"):
        ref.run_stage(st)
    for st in ("K7_BLACK_STRONG", "K8_RED_STRONG"):
        pre = {n: ref.get(n) for n in STATE_BUFS}
        ref.run_stage(st, 0)
        for n, a in pre.items():
            prod.set(n, a)
        prod.run_stage(st, 0)
        for n in STAGE_OUTPUTS[st]:
            r = compare(n, ref.get(n), prod.get(n))
            assert r["not_bit_exact"] == 0, (S, st, r)


@needs_ref
def test_adaptive_radius_map_general_patch_path_vs_reference():
    """use_radius with a radius map that is not 5 everywhere: patches of 2..7 samples per axis take the general
    (non-hoisted) NCC path; K6, K15, K16 and the sparse-mask sweep must still match the reference bit for bit."""
    from dvp_mvs_b200 import REFINE_INIT
    W, H, S = 200, 152, 3
    sc = synth.make_scene(W, H, S)
    rng = np.random.default_rng(17)
    radius = rng.choice(np.array([1, 3, 5, 5, 6, 7, 8, 10], np.int32), size=(H, W)).astype(np.int32)
    p = c1_params(sc.depth_min, sc.depth_max, S, iters=1, use_apd=1)
    p.state = REFINE_INIT; p.use_radius = 1
    yy, xx = np.mgrid[0:H, 0:W]
    weak = np.where(((xx + yy) % 64) < 2, STRONG, WEAK).astype(np.uint8)      # sparse STRONG mask: race-free sweep
    planes = sc.planes_true.copy(); planes[..., 3] *= (1.0 + rng.normal(0, 0.02, (H, W))).astype(np.float32)
    kw = dict(images=sc.images, cameras=sc.cameras, planes=planes, selected_views=np.full((H, W), (1 << S) - 1, np.uint32),
              weak_info=weak, edge=sc.edge, label=sc.label, radius=radius, seed=synth.SEED_RNG)
    ref = ref_oracle.engine(W, H, S, p); prod = Engine(W, H, S, p)
    ref.upload(**kw); prod.upload(**kw)
    assert (prod.get("radius") == radius).all()
    res = step_compare(ref, prod, 1, stages=("K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM", "K6_RANDOM_INITIALIZATION", "K7_BLACK_STRONG",
                                             "K8_RED_STRONG", "K12_DEPTH_NORMAL", "K15_DEPTH_TO_WEAK", "K16_LOCAL_REFINE"))
    # K2 demotes WEAK pixels without anchors to UNKNOWN; those then run the strong sweep next to each other, where the
    # reference's direction-4 race applies — so the sweeps get a tolerance of a few pixels, everything else none
    racy = ("K7_BLACK_STRONG", "K8_RED_STRONG")
    bad = [r for r in res if r.get("error") or (r["not_bit_exact"] and r["stage"] not in racy)]
    assert len(res) >= 12 and not bad, bad[:4]
    assert all(r["mismatched"] <= 30 for r in res if r["stage"] in racy), [r for r in res if r["stage"] in racy]


def _second_pass_inputs(W, H, S, geom):
    """Pass 1 (FIRST_INIT, all STRONG) on the reference -> inputs of a rounds>=1 pass with WEAK pixels."""
    from dvp_mvs_b200 import REFINE_INIT, REFINE_ITER
    sc = synth.make_scene(W, H, S)
    p = c1_params(sc.depth_min, sc.depth_max, S, iters=2)
    ref = ref_oracle.engine(W, H, S, p)
    ref.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    ref.run(mode=0)
    planes, weak, sel, rad = ref.download()
    q = c1_params(sc.depth_min, sc.depth_max, S, iters=1, use_apd=1)
    q.state = REFINE_ITER if geom else REFINE_INIT
    q.use_detail = 1; q.ransac_threshold = 0.00875; q.rotate_time = 2; q.geom_consistency = geom   # main.cpp:463-503, round 1
    kw = dict(images=sc.images, depths=sc.depths if geom else None, cameras=sc.cameras, planes=planes, selected_views=sel,
              weak_info=weak, edge=sc.edge, label=sc.label, radius=rad, seed=synth.SEED_RNG + 1)
    return q, kw, weak


@needs_ref
@pytest.mark.parametrize("geom", [0, 1])
def test_weak_path_stagewise_vs_reference(geom):
    """Adaptive patch deformation: K2 (complex, label boundaries), K3, K4 GenNeighbours (anchors, RNG), K5, K9 RANSAC
    fit + radius, K10/K11 weak sweep — stepping from the reference's state.  Integer outputs and RNG states must be
    bit-exact; the weak sweep's float outputs within 1e-4 relative on all but a handful of pixels."""
    W, H, S = 320, 240, 2 + 2 * geom
    q, kw, weak = _second_pass_inputs(W, H, S, geom)
    assert (weak == WEAK).sum() > 2000
    ref = ref_oracle.engine(W, H, S, q); prod = Engine(W, H, S, q)
    ref.upload(**kw); prod.upload(**kw)
    assert ref.weak_count() == prod.weak_count() == int((weak == WEAK).sum())
    assert (ref.get("neighbours_map") == prod.get("neighbours_map")).all()
    res = step_compare(ref, prod, 1)
    by = {(r["stage"], r.get("buffer")): r for r in res}
    assert not [r for r in res if r.get("error")], [r for r in res if r.get("error")][:2]
    for key in [("K2_GEN_EDGE_INFORM", "edge_neigh"), ("K2_GEN_EDGE_INFORM", "weak"), ("K2_GEN_EDGE_INFORM", "complex"),
                ("K2_GEN_EDGE_INFORM", "label_boundary"), ("K3_FIND_NEAREST_STRONG", "nearest_strong"),
                ("K4_GEN_NEIGHBOURS", "neighbours"), ("K4_GEN_NEIGHBOURS", "weak_reliable"), ("K4_GEN_NEIGHBOURS", "rand"),
                ("K5_NEIGHBOUR_UPDATE", "weak"), ("K6_RANDOM_INITIALIZATION", "costs"), ("K6_RANDOM_INITIALIZATION", "selected"),
                ("K9_RANSAC_FIT_PLANE", "fit_planes"), ("K9_RANSAC_FIT_PLANE", "radius"), ("K9_RANSAC_FIT_PLANE", "rand"),
                ("K10_BLACK_WEAK", "view_weight"), ("K10_BLACK_WEAK", "radius"), ("K11_RED_WEAK", "view_weight"),
                ("K15_DEPTH_TO_WEAK", "weak"), ("K16_LOCAL_REFINE", "planes")]:
        assert by[key]["not_bit_exact"] == 0, (key, by[key])
    for st in ("K10_BLACK_WEAK", "K11_RED_WEAK"):
        # without the geometric term the weak sweep is bit-exact; with it a 1-ulp difference in the forward-backward
        # reprojection cost flips an accept decision on a few pixels in 100 000 (measured: <= 6 of 307 200)
        limit = 0 if geom == 0 else 8
        for n in ("planes", "costs", "selected", "rand"):
            assert by[(st, n)]["mismatched"] <= limit, (st, n, by[(st, n)])


@needs_ref
@pytest.mark.parametrize("S,rotate_time", [(8, 4), (3, 1)])
def test_weak_path_with_many_views_and_other_rotations_vs_reference(S, rotate_time):
    """The WEAK path away from the schedule's usual shape: 8 source views (the scoring kernel's 8 x S scratch rows, the
    32-entry cost columns) and rotate_time 4 / 1 (32 / 8 search directions in K4, shift ranges 1 / 8 for the exact
    multiply-based modulo), photometric: K4's anchors and RNG states, K9 and both WEAK sweeps bit for bit."""
    W, H = 256, 192
    q, kw, weak = _second_pass_inputs(W, H, S, 0)
    q.rotate_time = rotate_time
    assert (weak == WEAK).sum() > 1000
    ref = ref_oracle.engine(W, H, S, q); prod = Engine(W, H, S, q)
    ref.upload(**kw); prod.upload(**kw)
    res = step_compare(ref, prod, 1)
    by = {(r["stage"], r.get("buffer")): r for r in res}
    assert not [r for r in res if r.get("error")], [r for r in res if r.get("error")][:2]
    for key in [("K4_GEN_NEIGHBOURS", "neighbours"), ("K4_GEN_NEIGHBOURS", "weak_reliable"), ("K4_GEN_NEIGHBOURS", "rand"),
                ("K9_RANSAC_FIT_PLANE", "fit_planes"), ("K9_RANSAC_FIT_PLANE", "rand"),
                ("K10_BLACK_WEAK", "planes"), ("K10_BLACK_WEAK", "costs"), ("K10_BLACK_WEAK", "selected"), ("K10_BLACK_WEAK", "view_weight"),
                ("K10_BLACK_WEAK", "rand"), ("K11_RED_WEAK", "planes"), ("K11_RED_WEAK", "costs"), ("K11_RED_WEAK", "selected"),
                ("K11_RED_WEAK", "view_weight"), ("K11_RED_WEAK", "rand")]:
        assert by[key]["not_bit_exact"] == 0, (key, by[key])


@pytest.mark.parametrize("geom", [0, 1])
def test_fused_k15_k16_equals_the_two_kernels(geom):
    """dvp_run issues DepthToWeak and LocalRefine as one kernel that shares the NCCs of the 11 common disparity
    hypotheses; from identical state it must leave exactly what K15 followed by K16 leave."""
    from dvp_mvs_b200 import REFINE_ITER, REFINE_INIT
    W, H, S = 320, 240, 2 + 2 * geom
    sc = synth.make_scene(W, H, S)
    p = c1_params(sc.depth_min, sc.depth_max, S, iters=2)
    e = Engine(W, H, S, p)
    e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    e.run()
    planes, weak, sel, rad = e.download()
    q = c1_params(sc.depth_min, sc.depth_max, S, iters=1, use_apd=1)
    q.state = REFINE_ITER if geom else REFINE_INIT
    q.use_detail = 1; q.ransac_threshold = 0.00875; q.rotate_time = 2; q.geom_consistency = geom
    planes[40:60, 50:90, 3] = 0.0          # depth 0: both kernels leave these pixels alone
    sel[100:110, 100:140] = 0              # no selected view
    rad[::9, ::7] = 0                      # radius 0 is reset to strong_radius by DepthToWeak before LocalRefine reads it
    e.upload(images=sc.images, depths=sc.depths if geom else None, cameras=sc.cameras, planes=planes, selected_views=sel,
             weak_info=weak, edge=sc.edge, label=sc.label, radius=rad, seed=synth.SEED_RNG + 1, params=q)
    for st, it in sequence(1):
        if st in ("K15_DEPTH_TO_WEAK", "K16_LOCAL_REFINE"):
            continue
        e.run_stage(st, it)
    vw = e.get("view_weight"); vw[120:130, 30:60] = 0; e.set("view_weight", vw)   # selected views with zero weight: LocalRefine returns
    names = ("planes", "costs", "selected", "weak", "radius", "view_weight")
    s0 = e.snapshot(names)
    e.run_stage("K15_DEPTH_TO_WEAK"); e.run_stage("K16_LOCAL_REFINE")
    two = e.snapshot(names)
    e.restore(s0)
    e.run_stage("K15_K16_FUSED")
    one = e.snapshot(names)
    assert (two["planes"][..., 3] != s0["planes"][..., 3]).sum() > 100     # LocalRefine moved some depths
    assert (two["weak"] != s0["weak"]).sum() > 100                         # DepthToWeak relabelled some pixels
    for n in names:
        a, b = two[n], one[n]
        same = (a.view(np.uint32) == b.view(np.uint32)) if a.dtype == np.float32 else (a == b)
        assert same.all(), (n, int((~same).sum()))


def test_run_equals_the_stage_by_stage_sequence_with_weak_pixels():
    """dvp_run takes short cuts that the stage API does not: K2's candidate records are evaluated after K4 and only for the
    pixels some anchor list names, K15 and K16 run as one kernel.  Our own sweeps are deterministic, so a whole pass through
    dvp_run must leave every buffer bit-identical to the same stages issued one by one (each of which is what the parity
    tests compare with the reference) — on a second pass with WEAK pixels, geometric consistency and 2 iterations."""
    from dvp_mvs_b200 import REFINE_ITER
    W, H, S = 320, 240, 3
    sc = synth.make_scene(W, H, S)
    p = c1_params(sc.depth_min, sc.depth_max, S, iters=1)
    e = Engine(W, H, S, p)
    e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    e.run()
    planes, weak, sel, rad = e.download()
    assert (weak == WEAK).sum() > 2000
    q = c1_params(sc.depth_min, sc.depth_max, S, iters=2, use_apd=1)
    q.state = REFINE_ITER; q.geom_consistency = 1; q.use_detail = 1; q.ransac_threshold = 0.00875; q.rotate_time = 2
    kw = dict(images=sc.images, depths=sc.depths, cameras=sc.cameras, planes=planes, selected_views=sel, weak_info=weak, edge=sc.edge,
              label=sc.label, radius=rad, seed=synth.SEED_RNG + 5, params=q)
    names = ("planes", "costs", "selected", "weak", "radius", "rand", "view_weight", "fit_planes", "neighbours", "weak_reliable")
    e.upload(**kw); e.run()
    whole = {n: e.get(n) for n in names}
    e.upload(**kw)
    for st, it in sequence(2):
        e.run_stage(st, it)
    for n in names:
        a, b = whole[n], e.get(n)
        same = (a.view(np.uint32) == b.view(np.uint32)) if a.dtype == np.float32 else (a == b)
        assert same.all(), (n, int((~same).sum()))


@pytest.mark.parametrize("use_weak", [0, 1])
def test_overlapped_upload_gives_the_same_results(use_weak):
    """dvp_upload_overlapped streams planes, images and depth maps on a second stream behind K1..K5; with 0 iterations
    the pass is deterministic, so every output must equal the plain upload's bit for bit (pinned and pageable hosts)."""
    import torch
    from dvp_mvs_b200 import REFINE_ITER
    W, H, S = 320, 240, 3
    sc = synth.make_scene(W, H, S)
    p = c1_params(sc.depth_min, sc.depth_max, S, iters=1)
    e = Engine(W, H, S, p)
    e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label, seed=synth.SEED_RNG)
    e.run()
    planes, weak, sel, rad = e.download()
    q = c1_params(sc.depth_min, sc.depth_max, S, iters=0, use_apd=use_weak)
    q.state = REFINE_ITER; q.geom_consistency = 1; q.use_detail = 1; q.ransac_threshold = 0.00875; q.rotate_time = 2
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    for host in (lambda a: a, pin):
        kw = dict(images=host(sc.images), depths=host(sc.depths), cameras=sc.cameras, planes=host(planes), selected_views=host(sel),
                  weak_info=host(weak), edge=sc.edge, label=sc.label, radius=host(rad), seed=synth.SEED_RNG + 3, params=q)
        outs = []
        for overlapped in (False, True, True):      # twice overlapped: a re-upload must not race the previous one
            e.upload(overlapped=overlapped, **kw)
            e.run(sync=not overlapped)
            outs.append({n: e.get(n) for n in ("planes", "costs", "selected", "weak", "radius", "rand", "view_weight")})
        for n, a in outs[0].items():
            for o in outs[1:]:
                b = o[n]
                same = (a.view(np.uint32) == b.view(np.uint32)) if a.dtype == np.float32 else (a == b)
                assert same.all(), (n, int((~same).sum()))
    # an overlapped upload followed by stage stepping / buffer access also sees complete data
    e.upload(overlapped=True, **kw)
    assert (e.get("planes").view(np.uint32) == np.ascontiguousarray(planes).view(np.uint32)).all()


def test_product_matches_committed_golden_vectors():
    """Same stepping protocol against tests/golden/c1_64x48.npz (reference kernels, generated on B200)."""
    g = load_golden("c1_64x48.npz")
    p = c1_params(g["depth_min"], g["depth_max"], 2)
    e = Engine(64, 48, 2, p)
    e.upload(**golden_inputs(g))
    state = {}
    checked = 0
    for k, (stage, it) in enumerate(sequence(1)):
        for n, a in state.items():
            e.set(n, a)
        e.run_stage(stage, it)
        for n in STAGE_OUTPUTS[stage]:
            key = f"{k:02d}_{stage}__{n}"
            if key not in g.files:
                continue
            want, got = g[key], e.get(n)
            if stage in RACE_FREE:
                assert compare(n, want, got)["not_bit_exact"] == 0, (stage, n)
                checked += 1
            else:
                assert compare(n, want, got)["frac"] < 0.05, (stage, n)   # racy sweep, tiny image
            state[n] = want
    assert checked >= 14


def test_product_matches_golden_sparse_sweep():
    g = load_golden("sparse_128x96.npz")
    p = c1_params(g["depth_min"], g["depth_max"], 2, iters=2, use_apd=1)
    e = Engine(128, 96, 2, p)
    e.upload(weak_info=g["weak"], **golden_inputs(g))
    strong = g["weak"] == STRONG
    state = {n: g["pre__" + n] for n in ("planes", "costs", "selected", "rand", "edge_neigh", "radius", "view_weight")}
    k = 0
    for it in range(2):
        for st in ("K7_BLACK_STRONG", "K8_RED_STRONG"):
            for n, a in state.items():
                e.set(n, a)
            e.run_stage(st, it)
            for n in STAGE_OUTPUTS[st]:
                want, got = g[f"{k:02d}_{st}_{it}__{n}"], e.get(n)
                assert (got[~strong] == state[n][~strong]).all() or n == "costs"
                eq = (got[strong] == want) | ((got[strong] != got[strong]) & (want != want)) if want.dtype.kind == "f" else got[strong] == want
                assert eq.all(), (st, it, n, int((~eq).sum()))
                full = state[n].copy(); full[strong] = want; state[n] = full
            k += 1


def test_rng_states_match_cpu_restatement_at_full_size():
    """K1 at BASELINE full resolution (6221x4146): bit-exact vs the independent GF(2) restatement of curand_init."""
    if not cpu_oracle.available():
        pytest.skip("CPU oracle not built")
    W, H = 6221, 4146
    p = c1_params(0.9, 14.4, 1)
    e = Engine(W, H, 1, p)
    sc = synth.make_scene(64, 48, 1)
    # only K1 runs: inputs just have to be well-formed
    images = np.zeros((2, H, W), np.float32); planes = np.zeros((H, W, 4), np.float32)
    cams = np.zeros(2, synth.CAMERA_DTYPE); cams[:] = sc.cameras[0]
    e.upload(images=images, cameras=cams, planes=planes, seed=1234567)
    e.run_stage("K1_INIT_RANDOM_STATES")
    got = e.get("rand")
    want = cpu_oracle.init_random_states(W, H, 1234567)
    assert (got == want).all()


def test_full_size_properties():
    """BASELINE-sized pass (3111x2073, 4 views, geometric consistency): invariants that need no oracle."""
    W, H, S = 3111, 2073, 4
    sc = synth.make_scene(W, H, S)
    from dvp_mvs_b200 import REFINE_ITER
    p = c1_params(sc.depth_min, sc.depth_max, S, iters=2)
    p.state = REFINE_ITER; p.geom_consistency = 1; p.weak_peak_radius = 4
    planes = sc.planes_true.copy()
    planes[..., 3] *= (1 + np.random.default_rng(1).normal(0, 0.02, (H, W))).astype(np.float32)
    sel = np.full((H, W), 15, np.uint32)
    e = Engine(W, H, S, p)
    kw = dict(images=sc.images, depths=sc.depths, cameras=sc.cameras, planes=planes, selected_views=sel, edge=sc.edge, label=sc.label, seed=7)
    e.upload(**kw); e.run()
    out, weak, sel_out, rad = e.download()
    total, per_stage, launches = e.last_run_times()
    # kernels actually launched: K1, K2 (edge sweep + per-pixel pass; nothing WEAK), K3, K5, K6; K7, K8, K9 per iteration
    # (no WEAK pixel: K4, K10, K11 launch nothing); K12, K13, K14 and the fused K15+K16
    assert launches == 6 + 2 * (2 + 2 + 1) + 4 and total > 0   # each strong sweep is two kernels (scoring, update); no WEAK pixel: K10 / K11 launch nothing
    assert set(np.unique(weak)) <= {WEAK, STRONG, UNKNOWN}
    border = np.ones((H, W), bool); border[6:-6, 6:-6] = False
    assert (weak[border] == UNKNOWN).all()                      # APD.cu:3907-3910
    assert (sel_out < (1 << S)).all() and (rad == 5).all()
    n = np.linalg.norm(out[..., :3], axis=-1)
    assert np.abs(n[np.isfinite(n)] - 1).max() < 1e-3           # world normals stay unit length
    inner = ~border & np.isfinite(out[..., 3])
    err = np.abs(out[..., 3] - sc.depths[0])[inner] / sc.depths[0][inner]
    assert np.median(err) < 0.01                                # converges onto the true planes
    # determinism of everything that is race-free: K1..K6 twice from the same inputs
    snaps = []
    for _ in range(2):
        e.upload(**kw)
        for st in sequence(1)[:6]:
            e.run_stage(*st)
        snaps.append({n: e.get(n) for n in ("planes", "costs", "selected", "rand", "edge_neigh")})
    for n in snaps[0]:
        assert compare(n, snaps[0][n], snaps[1][n])["not_bit_exact"] == 0, n
    # K12 maps plane offset -> depth exactly as the definition says (idempotent geometry check on a sample)
    assert (out[..., 3][inner] > p.depth_min * 0.5).mean() > 0.99


def test_api_error_behaviour():
    p = c1_params(0.9, 14.4, 2)
    e = Engine(64, 48, 2, p)
    with pytest.raises(DvpError, match="DVP_ERR_STATE"):
        e.run()                                         # run before upload
    sc = synth.make_scene(64, 48, 2)
    q = c1_params(sc.depth_min, sc.depth_max, 2); q.geom_consistency = 1
    with pytest.raises(DvpError, match="DVP_ERR_ARG"):
        e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, params=q)   # geom without depth maps
    q = c1_params(sc.depth_min, sc.depth_max, 2); q.use_edge = 0
    with pytest.raises(DvpError, match="UNSUPPORTED"):
        e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, params=q)
    with pytest.raises(ValueError):
        e.upload(images=sc.images[:2], cameras=sc.cameras, planes=sc.planes_init)          # wrong shape
    e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init)
    with pytest.raises(ValueError):
        e.set("costs", np.zeros(5, np.float32))
    e.run()
    assert e.weak_count() == 0


def test_empty_priors_and_ragged_sizes():
    """NULL edge/label/radius/selected inputs and odd, non-multiple-of-block sizes (reference half-grid quirk)."""
    for (W, H) in ((67, 33), (33, 35), (130, 97)):
        sc = synth.make_scene(W, H, 2)
        p = c1_params(sc.depth_min, sc.depth_max, 2)
        e = Engine(W, H, 2, p)
        e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, seed=3)
        e.run()
        planes, weak, sel, rad = e.download()
        assert planes.shape == (H, W, 4) and np.isfinite(planes[..., :3]).all()
        if ref_oracle.available():
            ref = ref_oracle.engine(W, H, 2, p)
            ref.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, seed=3)
            e.upload(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, seed=3)
            res = step_compare(ref, e, 1, stages=RACE_FREE)
            bad = [r for r in res if r.get("error") or r["not_bit_exact"]]
            assert not bad, (W, H, bad[:3])


RACY = ("K7_BLACK_STRONG", "K8_RED_STRONG")


def _bench_workload(width, height, src, iters):
    """The workload bench.py times (bench.make_workload): ETH3D-shaped synthetic view, REFINE_ITER with geometric
    consistency, priors on, the textureless wall WEAK."""
    import argparse
    import bench
    ns = argparse.Namespace(width=width, height=height, src=src, iters=iters, state="refine_iter", geom=1)
    sc, p, inputs, name = bench.make_workload(ns, seed=0)
    return p, inputs, name


@needs_ref
def test_bench_workload_stage_by_stage_vs_reference():
    """The stage-wise comparison at the size and configuration bench.py reports on its C2 line (3111x2073, 4 source views,
    REFINE_ITER + geometric consistency, 16 % WEAK pixels), one iteration: large-image behaviour that 640x480 never
    exercises — K2's unbounded rays, K3's 100-ring search inside a wall, K4's search radii up to the image size and
    max(H, W)/30-step edge walks, the long adaptive ladders of the sweep.  Every race-free stage and the whole WEAK path
    bit for bit; the racy strong sweep is re-synchronised from the reference after each launch."""
    W, H, S = 3111, 2073, 4
    p, inputs, _ = _bench_workload(W, H, S, 1)
    ref = ref_oracle.engine(W, H, S, p); prod = Engine(W, H, S, p)
    ref.upload(**inputs); prod.upload(**inputs)
    assert ref.weak_count() == prod.weak_count() > 900000
    state = {}

    def before_racy(stage, it):       # keep the pre-launch state of the first racy launch for the race analysis below
        if stage == "K7_BLACK_STRONG" and it == 0:
            state.update({n: ref.get(n) for n in STAGE_OUTPUTS[stage] + ("weak", "radius")})   # everything K7 reads that later stages rewrite
    res = lockstep_compare(ref, prod, 1, racy=RACY, resync_extra={"K2_GEN_EDGE_INFORM": ("candidate",)}, before_stage=before_racy,
                           keep={("K7_BLACK_STRONG", 0): state})
    assert len(res) >= 40
    n_weak = ref.weak_count()
    for r in res:
        if r["stage"] in RACY:
            assert r["frac"] < 0.25, r                 # long ladders at this size: many pixels see a rewritten direction-4 pixel
        elif r["stage"] in ("K10_BLACK_WEAK", "K11_RED_WEAK") and r["buffer"] in ("planes", "costs", "selected", "rand"):
            # with the geometric term a 1-ulp difference in the reprojection error flips an accept decision on a few
            # WEAK pixels in 100 000 (see test_weak_path_stagewise_vs_reference)
            assert r["mismatched"] <= max(8, n_weak // 20000), r
        else:
            assert r["not_bit_exact"] == 0, r
    # the racy launch at this size: the reference's result and ours, both explained by the race model, from the kept state
    pre = {n: a for n, a in state.items() if not n.startswith("_")}
    _assert_race_explained(prod, pre, STAGE_OUTPUTS["K7_BLACK_STRONG"], "K7_BLACK_STRONG", 0, W, H,
                           (("reference", state["_ref_out"], False), ("ours", state["_prod_out"], False)), print, slack=W * H // 500)


@needs_ref
def test_full_resolution_race_free_stages_vs_reference():
    """BASELINE's full size, 6221x4146 (C3): K1, K2, K3 from the upload and K12..K16 from the reference's own state after
    one iteration, bit for bit (coordinates beyond 4096, `short2` maps, 207-step edge limits, 25.8 M RNG sequences)."""
    W, H, S = 6221, 4146, 2
    p, inputs, _ = _bench_workload(W, H, S, 1)
    ref = ref_oracle.engine(W, H, S, p); prod = Engine(W, H, S, p)
    ref.upload(**inputs); prod.upload(**inputs)
    head = ("K1_INIT_RANDOM_STATES", "K2_GEN_EDGE_INFORM", "K3_FIND_NEAREST_STRONG")
    res = lockstep_compare(ref, prod, 1, stages=head)
    for st, it in sequence(1)[3:11]:
        ref.run_stage(st, it)
    for n in ("planes", "costs", "selected", "weak", "radius", "view_weight"):
        prod.set(n, ref.get(n))
    tail = ("K12_DEPTH_NORMAL", "K13_BLACK_FILTER", "K14_RED_FILTER", "K15_DEPTH_TO_WEAK", "K16_LOCAL_REFINE")
    res += lockstep_compare(ref, prod, 1, stages=tail)
    assert len(res) >= 10
    bad = [r for r in res if r["not_bit_exact"]]
    assert not bad, bad[:4]
