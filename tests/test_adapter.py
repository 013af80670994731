"""include/dvp_apd_adapter.hpp must compile against the reference's own main.h with ProcessProblem's call sequence."""
import os
import shutil
import subprocess

import pytest

from util import ROOT


@pytest.mark.skipif(not os.path.exists("/root/reference/main.h") or shutil.which("nvcc") is None,
                    reason="needs the reference headers and nvcc (present in the build container only)")
def test_adapter_compiles_with_reference_headers(tmp_path):
    cmd = ["nvcc", "-std=c++14", "-w", f"-I{ROOT}/oracle/stubs", "-I/root/reference", f"-I{ROOT}/include", "-x", "cu", "-c",
           f"{ROOT}/tests/adapter/process_problem_like.cpp", "-o", str(tmp_path / "adapter.o")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]


def test_adapter_mirrors_the_reference_surface():
    src = open(os.path.join(ROOT, "include", "dvp_apd_adapter.hpp")).read()
    for name in ["InuputInitialization", "SupportInitialization", "CudaSpaceInitialization", "SetDataPassHelperInCuda", "RunPatchMatch",
                 "GetPlaneHypothesis", "GetPixelSelectedViews", "SetPixelSelectedViews", "GetEdge", "GetPixelStates", "GetSelectedViews",
                 "GetRadiusMap", "GetWidth", "GetHeight", "GetDepthMin", "GetDepthMax"]:   # reference APD.h:96-115
        assert name in src, name


def test_scene_driver_links_from_cpp_and_prints_the_reference_schedule(tmp_path):
    """main()'s loop on the dvp_scene_* entry points (tests/adapter/scene_main_like.cpp), built with g++ against the C ABI
    and run with a host-only scene: ComputeRoundNum, level sizes and the per-pass parameters of main.cpp:452-505."""
    from dvp_mvs_b200 import _lib
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = tmp_path / "scene_main_like"
    libdir = os.path.dirname(_lib.PRODUCT_LIB)
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", f"{ROOT}/tests/adapter/scene_main_like.cpp",
                        "-o", str(exe), "-L", libdir, "-l:libdvp_mvs.so", "-Wl,-rpath," + libdir], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    lines = out.stdout.strip().splitlines()
    assert lines[0] == "round_num 4 -> 3 pyramid levels"
    assert lines[1].startswith("level 0 778x518 pass 0: state 0 use_APD 0 geom 0 weak_peak_radius 6")
    assert "level 1 1555x1037 pass 0: state 1 use_APD 1 geom 0 weak_peak_radius 6 rotate_time 2 ransac 0.00875 use_detail 1" in lines
    assert "level 2 3111x2073 pass 3: state 2 use_APD 1 geom 1 weak_peak_radius 2 rotate_time 4 ransac 0.00750 use_detail 1" in lines
    assert len(lines) == 1 + 3 * 4


def _build_pipeline_exe(tmp_path):
    from dvp_mvs_b200 import _lib
    exe = tmp_path / "pipeline_main_like"
    libdir = os.path.dirname(_lib.PRODUCT_LIB)
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/include", f"{ROOT}/tests/adapter/pipeline_main_like.cpp",
                        "-o", str(exe), "-L", libdir, "-l:libdvp_mvs.so", "-Wl,-rpath," + libdir], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def test_whole_pipeline_program_links_against_the_c_abi(tmp_path):
    """main() with every device-side row (edges, schedule + visibility restoration, fusion, PLY) as a C++ program on the
    C ABI: builds with g++ alone and fails cleanly without its inputs."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = _build_pipeline_exe(tmp_path)
    out = subprocess.run([str(exe), str(tmp_path / "missing"), str(tmp_path / "x.ply")], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2      # no meta.txt


@pytest.mark.gpu
def test_whole_pipeline_program_equals_the_python_driven_run(tmp_path):
    """The C++ program and the Python binding drive the same library: with 0 PatchMatch iterations every stage is
    deterministic, so the two PLY files must be identical byte for byte."""
    import numpy as np
    from dvp_mvs_b200 import Fusion, Scene, synth
    mv = synth.make_multiview(320, 240, 3, 2, seed=6)
    V, L = 3, 2
    d = tmp_path / "scene"; d.mkdir()
    (d / "meta.txt").write_text(f"{V} {L} {mv.full_w} {mv.full_h} 0 77\n")
    fine = mv.levels[-1]
    colors = [np.stack([np.clip(x["image"], 0, 255), np.clip(x["image"] * 0.5 + 20, 0, 255), np.clip(255 - x["image"], 0, 255)], -1).astype(np.uint8) for x in fine]
    for v in range(V):
        src = np.asarray(mv.src_views[v], np.int32)
        (d / f"view{v}.cam").write_bytes(np.asarray(mv.cameras[v]).tobytes() + np.int32(len(src)).tobytes() + src.tobytes())
        for l in range(L):
            (d / f"view{v}_level{l}.image").write_bytes(np.ascontiguousarray(mv.levels[l][v]["image"], np.float32).tobytes())
            (d / f"view{v}_level{l}.label").write_bytes(np.ascontiguousarray(mv.levels[l][v]["label"], np.int32).tobytes())
        (d / f"view{v}.planes").write_bytes(np.ascontiguousarray(mv.planes_init[v], np.float32).tobytes())
        (d / f"view{v}.color").write_bytes(colors[v].tobytes())
    exe = _build_pipeline_exe(tmp_path)
    out = subprocess.run([str(exe), str(d), str(tmp_path / "cpp.ply")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    # the same through the Python binding
    sc = Scene(V, L)
    sc.set_max_iterations(0)
    for v in range(V):
        sc.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        for l in range(L):
            sc.set_level(v, l, mv.levels[l][v]["image"], None, mv.levels[l][v]["label"])
            sc.compute_edges(v, l)
        sc.set_initial_planes(v, mv.planes_init[v])
    sc.run(seed=77)
    f = Fusion.from_scene(sc, colors)
    pts, _ = f.run()
    f.write_ply(str(tmp_path / "py.ply"))
    a, b = (tmp_path / "cpp.ply").read_bytes(), (tmp_path / "py.ply").read_bytes()
    assert a == b and (f"points {len(pts)}" in out.stdout), (len(a), len(b), out.stdout)
    f.close(); sc.close()
    # and with the reference's 3 iterations per pass (not reproducible run to run: the sweep's race) the program
    # produces a real cloud
    (d / "meta.txt").write_text(f"{V} {L} {mv.full_w} {mv.full_h} 3 77\n")
    out = subprocess.run([str(exe), str(d), str(tmp_path / "cpp3.ply")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    n = int(out.stdout.split("points")[1])
    assert n > 500, out.stdout
    assert len((tmp_path / "cpp3.ply").read_bytes().split(b"end_header\n", 1)[1]) == 15 * n


ADAPTER_EXE = os.path.join(ROOT, "tests", "adapter", "_build", "process_problem_run")


def _dump_for_adapter(d, W, H, S, iters, state, geom, use_apd, seed, depth_min, depth_max, inputs, use_detail=0, rotate_time=4, ransac=0.005, wpr=6):
    import numpy as np
    has = lambda k: int(inputs.get(k) is not None)
    (d / "meta.txt").write_text(f"{W} {H} {S} {iters} {state} {geom} {use_apd} {seed} {depth_min!r} {depth_max!r} {has('radius')} {has('weak_info')} "
                                f"{has('selected_views')} {has('depths')} {use_detail} {rotate_time} {ransac!r} {wpr}\n")
    for key, name, dt in (("images", "images.f32", np.float32), ("depths", "depths.f32", np.float32), ("planes", "planes.f32", np.float32),
                          ("selected_views", "selected.u32", np.uint32), ("weak_info", "weak.u8", np.uint8), ("edge", "edge.u8", np.uint8),
                          ("label", "label.i32", np.int32), ("radius", "radius.i32", np.int32)):
        if inputs.get(key) is not None:
            (d / name).write_bytes(np.ascontiguousarray(inputs[key], dt).tobytes())
    (d / "cameras.bin").write_bytes(np.asarray(inputs["cameras"]).tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("second_pass", [0, 1])
def test_adapter_class_executes_process_problem_and_equals_the_engine(tmp_path, second_pass):
    """The `APD` adapter RUN, not just compiled: tests/adapter/process_problem_run.cpp is ProcessProblem's call sequence
    (ctor, GetDepthMin/Max, InuputInitialization, SupportInitialization, CudaSpaceInitialization, SetDataPassHelperInCuda,
    RunPatchMatch, getters) compiled against the reference's main.h with loaders that fill the reference's host members from
    raw arrays.  Its planes / pixel states / selected views / radius map must equal what the ctypes `Engine` leaves for the
    same inputs, parameters and seed: bit for bit with 0 iterations (every stage deterministic), within the sweep's race
    with 1.  second_pass: a rounds >= 1 pass (REFINE_ITER, geometric consistency, WEAK pixels, radius map)."""
    import numpy as np
    import subprocess as sp
    from dvp_mvs_b200 import Engine, synth, default_params, FIRST_INIT, REFINE_ITER
    if not os.path.exists(ADAPTER_EXE):
        pytest.skip("tests/adapter/_build/process_problem_run not built (needs the reference headers at build time)")
    W, H, S, seed = 320, 240, 3, 4242
    sc = synth.make_scene(W, H, S)
    for iters in (0, 1):
        p = default_params(); p.max_iterations = iters; p.num_images = S + 1
        p.depth_min, p.depth_max = sc.depth_min, sc.depth_max
        if not second_pass:
            p.use_APD = 1; p.state = FIRST_INIT   # the adapter passes problem.params through; with no WEAK pixel the WEAK path idles
            inputs = dict(images=sc.images, cameras=sc.cameras, planes=sc.planes_init, edge=sc.edge, label=sc.label)
        else:
            p.use_APD = 1; p.state = REFINE_ITER; p.geom_consistency = 1; p.use_detail = 1; p.rotate_time = 2; p.ransac_threshold = 0.00875; p.weak_peak_radius = 4
            weak = np.full((H, W), 1, np.uint8); weak[sc.plane_id == 3] = 0
            rng = np.random.default_rng(3)
            planes = sc.planes_true.copy(); planes[..., 3] *= (1 + rng.normal(0, 0.02, (H, W))).astype(np.float32)
            inputs = dict(images=sc.images, depths=sc.depths, cameras=sc.cameras, planes=planes, selected_views=np.full((H, W), (1 << S) - 1, np.uint32),
                          weak_info=weak, edge=sc.edge, label=sc.label, radius=np.full((H, W), 5, np.int32))
        d = tmp_path / f"case{second_pass}_{iters}"; d.mkdir()
        _dump_for_adapter(d, W, H, S, iters, p.state, p.geom_consistency, p.use_APD, seed, float(np.float32(sc.depth_min)), float(np.float32(sc.depth_max)), inputs,
                          use_detail=p.use_detail, rotate_time=p.rotate_time, ransac=float(np.float32(p.ransac_threshold)), wpr=p.weak_peak_radius)
        r = sp.run([ADAPTER_EXE, str(d)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
        got = dict(planes=np.fromfile(d / "out_planes.bin", np.float32).reshape(H, W, 4), weak=np.fromfile(d / "out_states.bin", np.uint8).reshape(H, W),
                   selected=np.fromfile(d / "out_selected.bin", np.uint32).reshape(H, W), radius=np.fromfile(d / "out_radius.bin", np.int32).reshape(H, W))
        e = Engine(W, H, S, p)
        e.upload(seed=seed, **inputs)
        e.run()
        planes, weak, sel, rad = e.download()
        want = dict(planes=planes, weak=weak, selected=sel, radius=rad)
        for n in want:
            a, b = got[n], want[n]
            same = (a.view(np.uint32) == b.view(np.uint32)).reshape(H, W, -1).all(-1) if a.dtype == np.float32 else (a == b)
            if iters == 0:
                assert same.all(), (second_pass, n, int((~same).sum()))
            else:
                assert same.mean() > 0.9, (second_pass, n, float(same.mean()))     # the strong sweep's race (two runs of either differ as much)
        e.close()


@pytest.mark.gpu
def test_whole_pipeline_program_from_grey_images_only(tmp_path):
    """The C++ program given nothing per view but the full-resolution grey image, the camera, the source list and the FIRST_INIT
    prior: pyramid (row N2), edge maps and label maps (both halves of row N4) are built on the device by dvp_scene_set_image.
    Byte-identical PLY with the Python-driven run of the same scene (0 iterations: every stage deterministic), and the
    scene's label maps are dvp_label_segment of the full image at each level's scale."""
    import numpy as np
    from dvp_mvs_b200 import Fusion, Scene, synth, label_segment
    mv = synth.make_multiview(320, 240, 3, 2, seed=6)
    V, L = 3, 2
    d = tmp_path / "scene"; d.mkdir()
    (d / "meta.txt").write_text(f"{V} {L} {mv.full_w} {mv.full_h} 0 77\n")
    sc0 = synth.make_scene(mv.full_w, mv.full_h, V - 1, seed=6)     # full-resolution grey images with the room's textureless wall
    grey = [np.clip(np.rint(sc0.images[v]), 0, 255).astype(np.uint8) for v in range(V)]
    sc = Scene(V, L)
    sc.set_max_iterations(0)
    for v in range(V):
        src = np.asarray(mv.src_views[v], np.int32)
        (d / f"view{v}.cam").write_bytes(np.asarray(mv.cameras[v]).tobytes() + np.int32(len(src)).tobytes() + src.tobytes())
        (d / f"view{v}.gray").write_bytes(grey[v].tobytes())
        (d / f"view{v}.planes").write_bytes(np.ascontiguousarray(mv.planes_init[v], np.float32).tobytes())
        sc.set_view(v, mv.cameras[v], mv.full_w, mv.full_h, mv.src_views[v])
        sc.set_image(v, grey[v], compute_edges=True, compute_labels=True)
        sc.set_initial_planes(v, mv.planes_init[v])
        for l in range(L):
            want, _, _ = label_segment(grey[v], L - l)
            assert (sc.get_label(v, l) == want).all(), (v, l)
    fw, fh = sc.view_level_size(0, L - 1)
    colors = [np.stack([np.full((fh, fw), 40 + 50 * v, np.uint8)] * 3, -1) for v in range(V)]
    for v in range(V):
        (d / f"view{v}.color").write_bytes(colors[v].tobytes())
    exe = _build_pipeline_exe(tmp_path)
    out = subprocess.run([str(exe), str(d), str(tmp_path / "cpp.ply")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    sc.run(seed=77)
    f = Fusion.from_scene(sc, colors)
    pts, _ = f.run()
    f.write_ply(str(tmp_path / "py.ply"))
    a, b = (tmp_path / "cpp.ply").read_bytes(), (tmp_path / "py.ply").read_bytes()
    assert a == b and (f"points {len(pts)}" in out.stdout), (len(a), len(b), out.stdout)
    f.close(); sc.close()
