"""CPU: the C-ABI library loads, exports every symbol include/spx.h declares, and refuses to compute without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

from sp_slam_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "spx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spx_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(api.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(api.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (spx_[a-z_0-9]+)", out))
    assert exported == set(declared_symbols())


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header(tmp_path):
    # the header must compile as plain C, and the ctypes / numpy mirrors must have the compiler's sizes
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "spx.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(spx_config),sizeof(spx_plane),sizeof(spx_frame_header),sizeof(spx_point),sizeof(spx_model_info),'
                   'sizeof(spx_line_info),sizeof(spx_batch_result),sizeof(spx_device_result),'
                   '(size_t)SPX_MAX_MODELS);return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes == [ctypes.sizeof(api.SpxConfig), api.PLANE_DTYPE.itemsize, api.HEADER_DTYPE.itemsize,
                     api.POINT_DTYPE.itemsize, api.MODEL_DTYPE.itemsize, api.LINE_DTYPE.itemsize,
                     ctypes.sizeof(api.SpxBatchResult), ctypes.sizeof(api.SpxDeviceResult), api.SPX_MAX_MODELS]
    cfg = api.default_config()
    assert (cfg.cloud_dis, cfg.min_size, cfg.ransac_max_iter) == (3, 500, 1000)       # TUM1.yaml:73-74, Frame.cc:945
    assert abs(cfg.angle_thr_deg - 3.0) < 1e-7 and abs(cfg.dist_thr - 0.05) < 1e-7   # TUM1.yaml:75-76
    assert abs(cfg.line_ratio - 0.2) < 1e-12 and abs(cfg.line_dist_thr - 0.01) < 1e-7  # TUM1.yaml:99-100


def _has_gpu():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_has_gpu(), reason="this checks the loud failure on a box WITHOUT a GPU")
def test_no_cpu_fallback():
    with pytest.raises(api.SpxError) as e:
        api.PlaneExtractor()
    assert e.value.code == api.SPX_ERR_CUDA and "no CPU path" in str(e.value)


def test_bad_arguments_are_rejected_before_touching_cuda():
    lib = api.lib()
    h = ctypes.c_void_p()
    assert lib.spx_create(None, ctypes.byref(h)) == api.SPX_ERR_ARG
    cfg = api.default_config(max_cols=4000)          # organized cloud wider than the row buffers
    assert lib.spx_create(ctypes.byref(cfg), ctypes.byref(h)) == api.SPX_ERR_ARG
    cfg = api.default_config(normal_smoothing_size=7.0)
    assert lib.spx_create(ctypes.byref(cfg), ctypes.byref(h)) == api.SPX_ERR_ARG
    assert b"smoothing" in lib.spx_last_error(None)
    assert lib.spx_extract(None, None, 1, 1, 4, None) == api.SPX_ERR_ARG


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "sp_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f) + " mentions the oracle"
