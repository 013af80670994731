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
