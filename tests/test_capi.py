"""CPU: the C-ABI library loads without a GPU, exports every symbol include/lgx.h declares, and the product
path fails loudly (never falls back to a CPU implementation) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "lgx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lgx_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(lgx):
    lib = lgx._lib.load()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lgx.h but not exported by liblgx.so"
    # and the ctypes binding covers them all
    assert set(names) <= set(lgx._lib.PROTOTYPES), set(names) - set(lgx._lib.PROTOTYPES)


def test_pure_host_entry_points(lgx):
    lib = lgx._lib.load()
    assert lib.lgx_version() == 100
    assert lib.lgx_plane_pitch(2448) == 2448 and lib.lgx_plane_pitch(1279) == 1280
    assert lib.lgx_bits_pitch(2448) == 77 and lib.lgx_bits_pitch(32) == 1 and lib.lgx_bits_pitch(33) == 2
    assert lib.lgx_strerror(0) == b"ok" and b"no CPU path" in lib.lgx_strerror(-3)
    assert lib.lgx_workspace_bytes(2448, 2048, 1, 0) > 3 * 8 * 2448 * 2048
    assert lib.lgx_workspace_bytes(1, 1, 1, 0) == 0


def test_no_device_is_an_error_not_a_fallback(lgx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    lib = lgx._lib.load()
    h = C.c_void_p()
    assert lib.lgx_create(0, 640, 480, 1, 0, C.byref(h)) == -3 and not h.value
    with pytest.raises(lgx._lib.LgxError):
        lgx.Frontend(640, 480)
    img = np.zeros((64, 64), np.uint8)
    with pytest.raises(lgx._lib.LgxError):
        lgx.load_and_preprocess_image(img)
    with pytest.raises(lgx._lib.LgxError):
        lgx.extract_joints(img)
    with pytest.raises(lgx._lib.LgxError):
        lgx.detect_points_batch(img[None])


def test_argument_validation_happens_before_any_device_work(lgx):
    with pytest.raises(ValueError):
        lgx.load_and_preprocess_image(np.zeros((4, 4, 3, 1), np.uint8))       # reference: ValueError on ndim
    with pytest.raises(TypeError):
        lgx.load_and_preprocess_image(np.zeros((8, 8), np.float32))
    lib = lgx._lib.load()
    assert lib.lgx_set_option(None, 1, 1) == -1
    assert lib.lgx_bgr2gray(None, 8, 1, 4, 4, None, None) == -1
    assert lib.lgx_render_noisy(None, 1, 1, 4, 4, C.c_float(1.0), C.c_uint64(0), 8, None, None) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cylinder-pose-estimation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
