"""Host-side logic that needs no GPU: synthetic scenes, stage sequence, comparison helpers."""
import numpy as np

from util import c1_params
from dvp_mvs_b200 import synth, STAGES
from dvp_mvs_b200.parity import sequence, compare, STAGE_OUTPUTS


def test_scene_is_deterministic_and_well_formed():
    a = synth.make_scene(96, 64, 2)
    b = synth.make_scene(96, 64, 2)
    assert (a.images == b.images).all() and (a.planes_init == b.planes_init).all() and (a.edge == b.edge).all()
    assert a.images.shape == (3, 64, 96) and a.images.dtype == np.float32
    assert a.images.min() >= 0 and a.images.max() <= 255 and (a.images == np.rint(a.images)).all()
    assert a.cameras.shape == (3,) and a.cameras["width"][0] == 96 and a.cameras["height"][0] == 64
    for cam in a.cameras:                      # t = -R c and R orthonormal
        R = cam["R"].reshape(3, 3).astype(np.float64)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)
        assert np.allclose(cam["t"], -R @ cam["c"], atol=1e-5)
    assert a.depths[0].min() > 1.0 and a.depths[0].max() < 9.0
    assert set(np.unique(a.edge)) <= {0, 255} and 0 < (a.edge > 0).mean() < 0.1
    assert (a.label[a.edge > 0] == 0).all() and (a.label[a.edge == 0] > 0).all()
    assert 0.05 < (a.planes_init[..., 3] == 0).mean() < 0.15   # 10 % invalid depths force random initialisation


def test_source_views_are_consistent_with_the_geometry():
    """Warp the reference image into a source view with the true depth: photo-consistent where visible."""
    sc = synth.make_scene(160, 120, 2, quantize=False)
    K = sc.cameras["K"][0].reshape(3, 3).astype(np.float64)
    R0 = sc.cameras["R"][0].reshape(3, 3).astype(np.float64); c0 = sc.cameras["c"][0].astype(np.float64)
    R1 = sc.cameras["R"][1].reshape(3, 3).astype(np.float64); t1 = sc.cameras["t"][1].astype(np.float64)
    ys, xs = np.mgrid[8:112, 8:152]
    d = sc.depths[0][ys, xs].astype(np.float64)
    Xc = np.stack([(xs - K[0, 2]) / K[0, 0] * d, (ys - K[1, 2]) / K[1, 1] * d, d], -1)
    Xw = Xc @ R0 + c0
    Xs = Xw @ R1.T + t1
    u = K[0, 0] * Xs[..., 0] / Xs[..., 2] + K[0, 2]; v = K[1, 1] * Xs[..., 1] / Xs[..., 2] + K[1, 2]
    ok = (u > 1) & (u < 158) & (v > 1) & (v < 118)
    ui = np.clip(np.rint(u).astype(int), 0, 159); vi = np.clip(np.rint(v).astype(int), 0, 119)
    visible = ok & (np.abs(sc.depths[1][vi, ui] - Xs[..., 2]) < 0.02 * Xs[..., 2])
    diff = np.abs(sc.images[1][vi, ui] - sc.images[0][ys, xs])[visible]
    assert visible.mean() > 0.4 and np.median(diff) < 12.0


def test_stage_sequence_matches_run_patch_match():
    seq = sequence(3)   # reference APD.cu:4430-4505: 6 + 5*iters + 5 launches
    assert len(seq) == 6 + 5 * 3 + 5
    assert [s for s, _ in seq[:6]] == STAGES[:6]
    assert seq[6] == ("K7_BLACK_STRONG", 0) and seq[10] == ("K11_RED_WEAK", 0) and seq[16] == ("K7_BLACK_STRONG", 2)
    assert [s for s, _ in seq[-5:]] == STAGES[11:]
    assert set(STAGE_OUTPUTS) == set(STAGES)


def test_compare_counts_pixels_not_components():
    a = np.zeros((4, 5, 3), np.float32); b = a.copy(); b[1, 2] = 1.0; b[3, 4, 0] = np.nan; a[3, 4, 0] = np.nan
    r = compare("planes", a, b)
    assert r["mismatched"] == 1 and r["pixels"] == 20 and r["first_bad"] == [1, 2]
    m = np.zeros((4, 5), bool); m[0, 0] = True
    assert compare("planes", a, b, mask=m)["mismatched"] == 0


def test_params_mirror():
    p = c1_params(0.9, 14.4, 2)
    assert p.num_images == 3 and p.state == 0 and p.use_APD == 0 and p.max_iterations == 1


def test_every_python_file_compiles():
    """bench.py, the graft entry, the tools and the oracle drivers cannot be run without a GPU here; at least they parse."""
    import glob, os, py_compile
    from util import ROOT
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for sub in ("tools", "oracle", "dvp_mvs_b200", "tests"):
        files += glob.glob(os.path.join(ROOT, sub, "*.py"))
    assert len(files) > 15
    for f in files:
        py_compile.compile(f, doraise=True)


def test_closed_form_of_the_reference_edge_walk():
    """Groundwork for K4 / K9: the positions BresenhamLine tests (reference APD.cu:281-313) have a
    closed form, so a warp can evaluate 32 steps of one walk at once.  The loop tests exactly min(max(dx, dy) + 1, max_step)
    positions — it overshoots the end point by one step — and position k (1-based) is
        x-major:  (x0 + sx k,  y0 + sy floor((k dy - e0 + dx - 1) / dx))
        y-major:  (x0 + sx min(k, floor((e0 + k dx - 1 + dy) / dy)),  y0 + sy k)        e0 = max(dx, dy) / 2
        diagonal: (x0 + sx k,  y0 + sy k);   A == B: the start point itself, once."""
    def walk_ref(x0, y0, x1, y1, max_step):
        dx, sx = abs(x1 - x0), (1 if x0 < x1 else -1)
        dy, sy = abs(y1 - y0), (1 if y0 < y1 else -1)
        erro = (dx if dx > dy else dy) // 2
        step, tagx, tagy, out = 0, True, True, []
        while tagx or tagy:
            if x0 == x1: tagx = False
            if y0 == y1: tagy = False
            e2 = erro
            if e2 > -dx: erro -= dy; x0 += sx
            if e2 < dy: erro += dx; y0 += sy
            out.append((x0, y0)); step += 1
            if step >= max_step: break
        return out

    def walk_closed(x0, y0, x1, y1, max_step):
        dx, sx = abs(x1 - x0), (1 if x0 < x1 else -1)
        dy, sy = abs(y1 - y0), (1 if y0 < y1 else -1)
        m = max(dx, dy); e0 = m // 2
        if m == 0:
            return [(x0, y0)]
        out = []
        for k in range(1, min(m + 1, max_step) + 1):
            if dx > dy:
                out.append((x0 + sx * k, y0 + sy * ((k * dy - e0 + dx - 1) // dx)))
            elif dy > dx:
                out.append((x0 + sx * min(k, (e0 + k * dx - 1 + dy) // dy), y0 + sy * k))
            else:
                out.append((x0 + sx * k, y0 + sy * k))
        return out

    for ddx in range(-36, 37):
        for ddy in range(-36, 37):
            for max_step in (1, 2, 5, 16, 200):
                assert walk_ref(3, -2, 3 + ddx, -2 + ddy, max_step) == walk_closed(3, -2, 3 + ddx, -2 + ddy, max_step), (ddx, ddy, max_step)


def test_closed_form_edge_walk_large_offsets_and_float_reciprocal_division():
    """bresenham_crosses_edge (dvp_weak.cuh) evaluates the closed form with a float-reciprocal division corrected by one
    exact integer remainder; same check as above on long walks (up to 700 pixels, step limits up to 512), with the division
    done the device's way in float32."""
    import numpy as np
    rng = np.random.default_rng(12)

    def walk_ref(x0, y0, x1, y1, max_step):
        dx, sx = abs(x1 - x0), (1 if x0 < x1 else -1)
        dy, sy = abs(y1 - y0), (1 if y0 < y1 else -1)
        erro = (dx if dx > dy else dy) // 2
        step, tagx, tagy, out = 0, True, True, []
        while tagx or tagy:
            if x0 == x1: tagx = False
            if y0 == y1: tagy = False
            e2 = erro
            if e2 > -dx: erro -= dy; x0 += sx
            if e2 < dy: erro += dx; y0 += sy
            out.append((x0, y0)); step += 1
            if step >= max_step: break
        return out

    def walk_device(x0, y0, x1, y1, max_step):
        dx, sx = abs(x1 - x0), (1 if x0 < x1 else -1)
        dy, sy = abs(y1 - y0), (1 if y0 < y1 else -1)
        M = max(dx, dy); mn = min(dx, dy); e0 = M // 2
        n_total = min(M + 1, max_step)
        xmaj, diag = dx > dy, dx == dy
        c = M - 1 - e0 if xmaj else e0 - 1 + M
        rcp = np.float32(1.0) / np.float32(M)
        out = []
        for k in range(1, n_total + 1):
            if diag:
                mi = k
            else:
                n = k * mn + c
                q = int(np.float32(n) * rcp)            # truncation, as the cast does
                r = n - q * M
                q += (1 if r >= M else 0) - (1 if r < 0 else 0)
                assert 0 <= n - q * M < M               # the corrected quotient is exact
                mi = q if xmaj else min(k, q)
            out.append((x0 + sx * mi, y0 + sy * k) if dy > dx else (x0 + sx * k, y0 + sy * mi))
        return out

    for _ in range(1500):
        ddx, ddy = int(rng.integers(-700, 701)), int(rng.integers(-700, 701))
        if ddx == 0 and ddy == 0:
            continue
        max_step = int(rng.choice([13, 21, 103, 207, 512]))
        assert walk_ref(5, 7, 5 + ddx, 7 + ddy, max_step) == walk_device(5, 7, 5 + ddx, 7 + ddy, max_step), (ddx, ddy, max_step)
    for ddx, ddy in ((600, 599), (599, 600), (600, 600), (600, 1), (1, 600), (600, 0), (0, -600), (-333, 332), (64, -63)):
        for max_step in (13, 207, 512):
            assert walk_ref(0, 0, ddx, ddy, max_step) == walk_device(0, 0, ddx, ddy, max_step), (ddx, ddy, max_step)


def test_distance_map_walk_gives_the_loop_answer():
    """bresenham_crosses_edge's skip-ahead (dvp_weak.cuh): standing on a pixel whose chessboard distance to the nearest edge
    pixel is d, the next d - 1 positions cannot be edge pixels.  Emulated here on random edge maps against the reference's
    loop (APD.cu:281-313, which skips positions outside the map): same yes / no on every segment."""
    import numpy as np
    rng = np.random.default_rng(5)

    def loop_answer(edge, ax, ay, bx, by, max_step):
        H, W = edge.shape
        x0, y0, x1, y1 = bx, by, ax, ay
        dx, sx = abs(x1 - x0), (1 if x0 < x1 else -1)
        dy, sy = abs(y1 - y0), (1 if y0 < y1 else -1)
        erro = (dx if dx > dy else dy) // 2
        step, tagx, tagy = 0, True, True
        while tagx or tagy:
            if x0 == x1: tagx = False
            if y0 == y1: tagy = False
            e2 = erro
            if e2 > -dx: erro -= dy; x0 += sx
            if e2 < dy: erro += dx; y0 += sy
            if 0 <= x0 < W and 0 <= y0 < H and edge[y0, x0]:
                return True
            step += 1
            if step >= max_step: break
        return False

    def skip_answer(edge, dist, ax, ay, bx, by, max_step):
        H, W = edge.shape
        x0, y0, x1, y1 = bx, by, ax, ay
        dx, sx = abs(x1 - x0), (1 if x0 < x1 else -1)
        dy, sy = abs(y1 - y0), (1 if y0 < y1 else -1)
        M, mn = max(dx, dy), min(dx, dy)
        e0 = M // 2
        n_total = min(M + 1, max_step)
        if M == 0:
            return False
        c = M - 1 - e0 if dx > dy else e0 - 1 + M
        k, d = 0, int(dist[y0, x0])
        while True:
            k += d
            if k > n_total:
                return False
            mi = k if dx == dy else ((k * mn + c) // M if dx > dy else min(k, (k * mn + c) // M))
            x, y = (x0 + sx * mi, y0 + sy * k) if dy > dx else (x0 + sx * k, y0 + sy * mi)
            d = 1
            if 0 <= x < W and 0 <= y < H:
                d = int(dist[y, x])
                if d == 0:
                    return True

    checked = crossing = 0
    for trial in range(60):
        H, W = int(rng.integers(20, 70)), int(rng.integers(20, 90))
        edge = (rng.random((H, W)) < rng.choice([0.002, 0.01, 0.05])).astype(np.uint8) * 255
        if trial % 5 == 0:
            edge[:, W // 2] = 255                      # a wall to cross
        ys, xs = np.nonzero(edge)
        yy, xx = np.mgrid[0:H, 0:W]
        if len(ys):
            dist = np.minimum(np.max(np.maximum(np.abs(yy[..., None] - ys), np.abs(xx[..., None] - xs)), axis=-1) * 0 +
                              np.min(np.maximum(np.abs(yy[..., None] - ys), np.abs(xx[..., None] - xs)), axis=-1), 255)
        else:
            dist = np.full((H, W), 255)
        for _ in range(120):
            ax, ay, bx, by = int(rng.integers(0, W)), int(rng.integers(0, H)), int(rng.integers(0, W)), int(rng.integers(0, H))
            if edge[ay, ax] or edge[by, bx]:
                continue                               # the function answers `false` before walking
            max_step = int(rng.choice([5, 13, 40, 200]))
            want = loop_answer(edge, ax, ay, bx, by, max_step)
            assert skip_answer(edge, dist, ax, ay, bx, by, max_step) == want, (trial, ax, ay, bx, by, max_step)
            checked += 1; crossing += int(want)
    assert checked > 3000 and crossing > 200


def test_fast_modulo_of_the_neighbour_search_is_exact():
    """dvp_common.cuh FastMod (K4's `curand() % shift_range`, APD.cu:3408-3411): q = umulhi(x, floor((2^32 - 1) / m)) plus two
    conditional subtractions equals x % m for every 32-bit x — checked on adversarial operands (multiples of m and their
    neighbours, both ends of the range) and a million random ones for every divisor the schedules can produce and beyond."""
    rng = np.random.default_rng(5)
    for m in list(range(1, 70)) + [127, 255, 256, 1000, 65535, 65536, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1]:
        M = (2 ** 32 - 1) // m
        x = np.concatenate([rng.integers(0, 2 ** 32, 1_000_000, dtype=np.uint64),
                            np.arange(0, 4 * min(m, 10 ** 5) + 4, dtype=np.uint64),
                            (2 ** 32 - 1 - np.arange(0, 4 * min(m, 10 ** 5) + 4, dtype=np.uint64)),
                            (rng.integers(0, 2 ** 32 // m + 1, 100_000).astype(np.int64) * m + rng.integers(-1, 2, 100_000)).clip(0, 2 ** 32 - 1).astype(np.uint64)])
        q = (x * np.uint64(M)) >> np.uint64(32)
        r = (x - q * np.uint64(m)) & np.uint64(0xFFFFFFFF)
        r = np.where(r >= m, r - np.uint64(m), r)
        r = np.where(r >= m, r - np.uint64(m), r)
        assert (r == x % np.uint64(m)).all(), m


def _sweep_pixel(mode, red, bx, by, tid, threads):
    """Python mirror of sweep_pixel (csrc/dvp_common.cuh): the pixel a thread of a checkerboard launch owns."""
    rows = threads // 32
    if mode == 2:
        k, r, q = bx * 8 + ((tid >> 2) & 7), by * rows + (tid >> 5), tid & 3
        return 4 * k - 2 * (r & 1) + ((q + 1) >> 1), 2 * r - 2 + red + (-1 if q == 1 else 1 if q == 2 else 0)
    if mode == 1:
        lane, w = tid & 31, tid >> 5
        x = bx * 32 + (w & 1) * 16 + 2 * (lane >> 2) + (lane & 1)
        return x, 2 * (by * rows + (w >> 1) * 2 + ((lane >> 1) & 1)) + ((x & 1) ^ red)
    x = bx * 32 + (tid & 31)
    return x, 2 * (by * rows + (tid >> 5)) + ((x & 1) ^ red)


def test_every_thread_to_pixel_map_of_the_sweep_covers_each_pixel_of_the_colour_once():
    """The quad-compact maps of the K7/K8 kernels (zigzag row, 2 x 2 in (x, y / 2), diamonds laid like bricks) with the grids
    launch_sweep_split gives them: every pixel of the colour the reference's half grid reaches exactly once, for both block sizes, odd and even sizes."""
    for W, H in ((67, 33), (64, 48), (33, 35), (130, 97), (5, 4), (1, 1)):
        yy_limit = (((H // 2) + 15) // 16) * 16
        for mode in (0, 1, 2):
            for threads in (128, 256):
                rows = threads // 32
                gx, nrows = (((W + 1) // 4 + 2 + 7) // 8, H // 2 + 2) if mode == 2 else ((W + 31) // 32, yy_limit)
                gy = (nrows + rows - 1) // rows
                for red in (0, 1):
                    seen = np.zeros((H, W), np.int32)
                    for by in range(gy):
                        for bx in range(gx):
                            for tid in range(threads):
                                x, y = _sweep_pixel(mode, red, bx, by, tid, threads)
                                if 0 <= x < W and 0 <= y < H and (y >> 1) < yy_limit:
                                    assert (x + y) & 1 == red
                                    seen[y, x] += 1
                    yy, xx = np.mgrid[0:H, 0:W]
                    # the reference's half grid stops at yy_limit row pairs (APD.cu:4421-4424): with H = 33 the last row is never swept
                    assert (seen == ((((xx + yy) & 1) == red) & ((yy >> 1) < yy_limit))).all(), (W, H, mode, threads, red)


def test_sorted_register_list_of_k4_equals_the_reference_insertion_sort():
    """k_gen_neighbours keeps the DVP_NEIGHBOUR_NUM - 1 lightest anchors in a sorted list (an anchor goes in front of the first
    kept one that is strictly heavier) instead of insertion-sorting all anchors as the reference does (sort_small_weighted,
    APD.cu:125-138) and reading the first 11.  Same anchors in the same order, ties and FLT_MAX outliers included, whenever no
    weight is NaN (the kernel takes the reference's loop otherwise)."""
    K = 11
    rng = np.random.default_rng(5)
    fmax = np.float32(np.finfo(np.float32).max)
    for trial in range(400):
        n = int(rng.integers(0, 60))
        w = rng.choice(np.array([0.0, 0.25, 0.5, 1.0, 3.0], np.float32), n) if trial % 2 else rng.random(n).astype(np.float32)
        pts = np.arange(1, n + 1)
        out = rng.random(n) < 0.3
        w = np.where(out, fmax, w).astype(np.float32)
        pts = np.where(out, -1, pts)
        # the reference: insertion sort of (weight, point) over a list pre-filled with -1, first K read back
        rw, rp = list(w), list(pts) + [-1] * K
        for i in range(1, n):
            tw, tp = rw[i], rp[i]
            j = i
            while j >= 1 and tw < rw[j - 1]:
                rw[j], rp[j] = rw[j - 1], rp[j - 1]
                j -= 1
            rw[j], rp[j] = tw, tp
        want = rp[:K]
        # the kernel: sorted list of K, +inf sentinels, branch-free insertion evaluated from the back
        tw_, tp_ = [np.float32(np.inf)] * K, [-1] * K
        for i in range(n):
            x, p = w[i], pts[i]
            if x < tw_[K - 1]:
                for j in range(K - 1, 0, -1):
                    shift, here = x < tw_[j - 1], x < tw_[j]
                    tp_[j] = tp_[j - 1] if shift else (p if here else tp_[j])
                    tw_[j] = tw_[j - 1] if shift else (x if here else tw_[j])
                if x < tw_[0]:
                    tw_[0], tp_[0] = x, p
        assert [int(v) for v in tp_] == [int(v) for v in want], (trial, n)


def test_three_pass_ransac_of_k4_equals_the_reference_single_loop():
    """k_gen_neighbours splits GenNeighbours' RANSAC loop (APD.cu:3569-3678) into pass A1 (draws, distinct, pixel inside the
    triangle), pass A2 (edge tests through a symmetric cache that is filled by whichever try asks first, with a walk that
    depends on its direction; normal and plane tests; the is_strong_plane bit) and pass B (inlier count, running best).  A model
    of both forms on random instances — asymmetric edge walks, strong / weak planes, ties in the inlier count and in the centre
    distance — must reach the same best try, the same cache and the same number of draws."""
    rng = np.random.default_rng(11)
    for trial in range(300):
        n = int(rng.integers(4, 14))
        walk = rng.random((n, n)) < 0.25                     # walk[i, j]: the walk from i to j crosses an edge (asymmetric)
        inside = rng.random((n, n, n)) < 0.3                 # pixel inside triangle (a, b, c)
        normal_ok = rng.random(n) < 0.9
        strong = rng.random((n, n, n)) < 0.5
        inliers = rng.integers(3, 10, (n, n, n))
        centre = rng.integers(0, 4, (n, n, n)).astype(np.float32)
        draws = rng.integers(0, n, (200, 3))

        def tests_without_side_effects(t):
            a, b, c = draws[t]
            return a != b and b != c and a != c and inside[a, b, c]

        def tests_with_side_effects(t, cache):
            a, b, c = draws[t]
            for i, j in ((a, b), (b, c), (c, a)):
                if (i, j) not in cache:
                    cache[(i, j)] = cache[(j, i)] = bool(walk[i, j])
            if cache[(a, b)] or cache[(b, c)] or cache[(c, a)]:
                return False
            return bool(normal_ok[a])

        def running_best(state, t):
            a, b, c = draws[t]
            is_strong = bool(strong[a, b, c])
            if state["has_strong"] and not is_strong:
                return
            cnt = int(inliers[a, b, c])
            if cnt < 6:
                return
            if cnt > state["max_count"] or (not state["has_strong"] and is_strong):
                if not state["has_strong"] and is_strong:
                    state["has_strong"] = True
                state.update(best=t, max_count=cnt, min_cost=float(centre[a, b, c]))
            elif cnt == state["max_count"] and float(centre[a, b, c]) < state["min_cost"]:
                state.update(best=t, max_count=cnt, min_cost=float(centre[a, b, c]))

        # the reference: one loop
        s1, cache1 = dict(has_strong=False, max_count=3, min_cost=np.inf, best=-1), {}
        for t in range(200):
            if tests_without_side_effects(t) and tests_with_side_effects(t, cache1):
                running_best(s1, t)
        # the kernel: three passes
        s2, cache2 = dict(has_strong=False, max_count=3, min_cost=np.inf, best=-1), {}
        survivors = [t for t in range(200) if tests_without_side_effects(t)]
        passed = [t for t in survivors if tests_with_side_effects(t, cache2)]
        for t in passed:
            running_best(s2, t)
        assert s1 == s2 and cache1 == cache2, trial
