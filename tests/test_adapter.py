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
