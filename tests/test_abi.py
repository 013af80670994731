"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/dvp_mvs.h declares; argument validation that needs no GPU behaves as documented."""
import ctypes as C
import os
import re

import pytest

from util import ROOT
from dvp_mvs_b200 import _lib, default_params, Params


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dvp_mvs.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dvp_[a-z_]+)\s*\(", hdr)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for name in ["dvp_create", "dvp_destroy", "dvp_upload", "dvp_upload_device", "dvp_run", "dvp_run_stage",
                 "dvp_download", "dvp_get_buffer", "dvp_set_buffer", "dvp_buffer_bytes", "dvp_last_run_times",
                 "dvp_default_params", "dvp_version", "dvp_weak_count", "dvp_last_cuda_error", "dvp_stream",
                 # rows N1 / N2 and the overlapped upload
                 "dvp_restore_visibility", "dvp_upload_overlapped", "dvp_rescale_map", "dvp_scene_create", "dvp_scene_destroy",
                 "dvp_scene_level_size", "dvp_scene_pass_params", "dvp_scene_set_max_iterations", "dvp_scene_set_view",
                 "dvp_scene_set_level", "dvp_scene_set_initial_planes", "dvp_scene_run_pass", "dvp_scene_run_view", "dvp_scene_run",
                 "dvp_scene_get_view", "dvp_scene_stats", "dvp_scene_depth_map", "dvp_scene_remote_depth"]:
        assert name in syms


def test_product_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.PRODUCT_LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(_lib.PRODUCT_LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_defaults_match_reference_main_h():
    lib = _lib.load_library(_lib.PRODUCT_LIB, "dvp_")
    assert b"sm_100a" in lib.dvp_version()
    p = default_params(lib)
    q = default_params()  # python-side copy of main.h:86-112
    for f, _ in Params._fields_:
        assert getattr(p, f) == pytest.approx(getattr(q, f)), f
    assert (p.max_iterations, p.top_k, p.strong_radius, p.weak_peak_radius, p.rotate_time) == (3, 4, 5, 2, 4)
    assert p.ransac_threshold == pytest.approx(0.005) and p.geom_factor == pytest.approx(0.2)


def test_struct_layouts():
    from dvp_mvs_b200.synth import CAMERA_DTYPE
    assert CAMERA_DTYPE.itemsize == 112          # reference `struct Camera`, main.h:58-67
    assert C.sizeof(Params) == 23 * 4
    assert C.sizeof(_lib.Inputs) == 9 * 8 + 8


def test_null_arguments_are_rejected_without_a_gpu():
    lib = _lib.load_library(_lib.PRODUCT_LIB, "dvp_")
    assert lib.dvp_run(None, 1) == -1                 # DVP_ERR_ARG
    assert lib.dvp_run_stage(None, 0, 0) == -1
    assert lib.dvp_buffer_bytes(None, 0) == 0
    assert lib.dvp_weak_count(None) == -1
    p = default_params()
    assert not lib.dvp_create(0, 0, 10, 2, C.byref(p))       # bad size
    assert not lib.dvp_create(0, 10, 10, 40, C.byref(p))     # more than MAX_IMAGES views
    assert not lib.dvp_create(0, 10, 10, 2, None)


def test_no_cpu_fallback_when_library_is_missing(tmp_path):
    with pytest.raises(_lib.DvpError):
        _lib.load_library(str(tmp_path / "libdvp_mvs.so"), "dvp_")


def test_product_package_never_references_the_oracle():
    pkg = os.path.join(ROOT, "dvp_mvs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libapd_" not in src, f                                   # never names an oracle library
                assert not re.search(r"^\s*(import|from)\s+(cpu_oracle|ref_oracle)", src, flags=re.M), f
                assert "dlopen" not in src and "/oracle/_ref" not in src, f


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """The boundary is a C ABI: include/dvp_mvs.h must compile as C99 and a C program must link against
    libdvp_mvs.so and reach the entry points that need no GPU."""
    import shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "dvp_mvs.h"\n'
                   'int main(void) { dvp_params p; dvp_default_params(&p);\n'
                   '  if (dvp_create(0, 0, 0, 0, &p) != NULL) return 2;          /* invalid size: rejected before any CUDA call */\n'
                   '  if (dvp_restore_visibility(NULL, 8, NULL) != DVP_ERR_ARG) return 3;\n'
                   '  if (dvp_scene_create(0, 1, 1) != NULL) return 4;           /* a scene needs at least two views */\n'
                   '  if (dvp_fusion_create(0, 0) != NULL) return 5;             /* rows N3 / N4: argument checks need no GPU */\n'
                   '  if (dvp_fusion_run(NULL, NULL, NULL) != DVP_ERR_ARG || dvp_fusion_set_mode(NULL, 1) != DVP_ERR_ARG) return 6;\n'
                   '  { dvp_fusion_view v; unsigned char px[9] = {0}; v.width = 3; v.height = 3; (void)v;\n'
                   '    if (dvp_edge_segment(0, px, 2, 3, px, NULL, NULL) != DVP_ERR_ARG) return 7; }  /* images below 3 x 3 are refused */\n'
                   '  printf("%s %d %d %.3f\\n", dvp_version(), p.max_iterations, p.strong_radius, p.ransac_threshold); return 0; }\n')
    exe = tmp_path / "t"
    libdir = os.path.dirname(_lib.PRODUCT_LIB)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-l:libdvp_mvs.so", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.split()[-3:] == ["3", "5", "0.005"] and "sm_100a" in out.stdout


def test_documents_only_name_entry_points_that_exist():
    """Every dvp_* function INTEGRATION.md, DESIGN.md or README.md mentions is declared in include/dvp_mvs.h (or is a
    type / file / wildcard of the same family) — the documents are part of the boundary."""
    import re
    header = open(os.path.join(ROOT, "include", "dvp_mvs.h")).read()
    declared = set(re.findall(r"\b(dvp_[a-z0-9_]+)\s*\(", header)) | set(re.findall(r"\b(dvp_[a-z0-9_]+)\b(?=;|\s*\{|\s+[a-z_*]+[;,)])", header))
    known_other = {"dvp_mvs", "dvp_mvs_b200", "dvp_ctx", "dvp_scene", "dvp_fusion", "dvp_params", "dvp_inputs", "dvp_camera", "dvp_status",
                   "dvp_fusion_view", "dvp_apd_adapter", "dvp_stage", "dvp_buffer", "dvp_api", "dvp_ncc", "dvp_strong", "dvp_weak", "dvp_common",
                   "dvp_launch", "dvp_unionfind", "dvp_io", "dvp_kernels_post", "dvp_kernels_edge", "dvp_kernels_fusion", "dvp_kernels_prep",
                   "dvp_kernels_strong", "dvp_kernels_weak", "dvp_kernels_image", "dvp_kernels_", "dvp_b200", "dvp_fuse", "dvp_farm"}
    missing = {}
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        text = open(os.path.join(ROOT, doc)).read()
        for name in set(re.findall(r"\b(dvp_[a-z0-9_]+)\b", text)):
            if name in declared or name in known_other or name.endswith("_"):
                continue
            if any(d.startswith(name) for d in declared):      # a family prefix such as dvp_scene_* / dvp_io_*
                continue
            missing.setdefault(doc, []).append(name)
    assert not missing, missing
