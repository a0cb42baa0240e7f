"""CPU tests of the drop-in boundary: libmmidx.so loads, exports every symbol include/mmidx.h declares, and
fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import mmidx_b200 as M
from multimedia_indexing_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mmidx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmidx_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 25
    lib = C.CDLL(_capi.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mmidx.h but not exported by libmmidx.so"
    assert sorted(_capi.EXPORTS) == syms, "ctypes binding and header disagree"


def test_no_torch_or_oracle_in_the_library():
    out = os.popen(f"ldd {_capi.LIB_PATH}").read()
    assert "torch" not in out and "oracle" not in out


def test_version_and_error_string():
    assert "sm_100a" in M.pkg.version()


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour WITHOUT a GPU")
def test_fails_loudly_without_gpu():
    with pytest.raises(M.MmidxError) as e:
        M.IVFPQ(16, 100, 4, 16, M.TransformationType.None_, 4)
    assert e.value.code == _capi.ERR_CUDA and "no CPU path" in str(e.value)
    agg = M.VladAggregator(np.zeros((4, 8)))
    with pytest.raises(M.MmidxError) as e:
        agg.aggregate(np.zeros((3, 8)))
    assert e.value.code == _capi.ERR_CUDA


def test_argument_validation_before_any_device_work():
    h = C.c_void_p()
    p = _capi.Params(_capi.MMIDX_PQ, 10, 100, 3, 16, 0, 0, -1, 0, 0)  # 10 % 3 != 0 (PQ.java:148-150)
    assert _capi.lib.mmidx_create(C.byref(p), C.byref(h)) == _capi.ERR_DIM
    assert b"subvectors" in _capi.lib.mmidx_last_error()
    p = _capi.Params(7, 10, 100, 2, 16, 0, 0, -1, 0, 0)
    assert _capi.lib.mmidx_create(C.byref(p), C.byref(h)) == _capi.ERR_INVALID
    assert _capi.lib.mmidx_create(None, C.byref(h)) == _capi.ERR_INVALID
    assert _capi.lib.mmidx_destroy(None) == _capi.OK


def test_random_permutation_host_logic():
    g = np.load(os.path.join(ROOT, "tests", "golden", "perm.npz"))
    assert (M.random_permutation(1, 10) == g["p10"]).all()
    assert (M.random_permutation(1, 128) == g["p128"]).all()
    assert (M.random_permutation(7, 1024) == g["p1024_seed7"]).all()
