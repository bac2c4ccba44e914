"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/flame_b200.h declares, mirrors the reference's default parameters, and fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "flame_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree(capi):
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(capi):
    lib = C.CDLL(capi.lib_path())
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_default_params_match_reference_yaml(capi):
    """cfg/flame_nodelet.yaml:69-75,86-89,31-46 of the reference."""
    r = capi.default_nltgv2_params()
    assert (round(r.data_factor, 6), round(r.step_x, 6), r.step_q, r.theta) == (0.15, 0.001, 125.0, 0.25)
    e = capi.default_epi_params()
    assert (e.win_size, e.min_grad_mag, e.epipolar_line_var, e.max_dropouts) == (5, 5.0, 4.0, 5)
    t = capi.default_tri_filter_params()
    assert (t.do_oblique, t.do_edge_length, t.do_idepth) == (1, 1, 1)
    assert abs(t.oblique_normal_thresh - 1.57) < 1e-6 and abs(t.edge_length_thresh - 0.333) < 1e-6
    assert abs(t.oblique_idepth_diff_factor - 0.35) < 1e-6 and abs(t.min_triangle_idepth - 0.01) < 1e-6


def test_oracle_and_product_param_structs_have_same_layout(capi, oracle):
    for a, b in ((capi.EpiParams, oracle.EpiParams), (capi.NLTGV2Params, oracle.NLTGV2Params),
                 (capi.TriFilterParams, oracle.TriFilterParams)):
        assert [(n, t) for n, t in a._fields_] == [(n, t) for n, t in b._fields_]
        assert C.sizeof(a) == C.sizeof(b)


def test_no_gpu_means_loud_failure_not_fallback(capi):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    with pytest.raises(capi.FlameError) as ei:
        capi.Context(1, 64, 48, 2, 16, 16, 16)
    assert "no CUDA device" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "flame_ros_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "flame_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "flame_oracle" not in open(p).read()
