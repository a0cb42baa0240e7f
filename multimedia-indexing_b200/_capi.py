"""ctypes binding of libmmidx.so (include/mmidx.h).  No fallback: importing this module fails loudly when
the CUDA library has not been built -- there is no CPU implementation of the product path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MMIDX_LIB_PATH: development override used by the kernel-variant sweeps under profiles/ (still a CUDA build of csrc/)
LIB_PATH = os.environ.get("MMIDX_LIB_PATH") or os.path.join(_HERE, "libmmidx.so")

MMIDX_LINEAR, MMIDX_PQ, MMIDX_IVFPQ = 0, 1, 2
OK, ERR_INVALID, ERR_DIM, ERR_FULL, ERR_STATE, ERR_CUDA, ERR_UNSUPPORTED, ERR_W = range(8)
MAX_K = 1024
COMM_HANDLE_BYTES = 128


class Params(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("d", C.c_int32),
        ("max_n", C.c_int64),
        ("m", C.c_int32),
        ("ks", C.c_int32),
        ("nlist", C.c_int32),
        ("w", C.c_int32),
        ("device", C.c_int32),
        ("shard_rank", C.c_int32),
        ("shard_count", C.c_int32),
    ]


class MmidxError(Exception):
    """Mirror of the reference's checked `Exception(message)`; `.code` is the C status."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). libmmidx has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
_SIGS = {
    "mmidx_create": [C.POINTER(Params), C.POINTER(_vp)],
    "mmidx_destroy": [_vp],
    "mmidx_set_product_quantizer": [_vp, _vp],
    "mmidx_set_coarse_quantizer": [_vp, _vp],
    "mmidx_set_permutation": [_vp, _vp],
    "mmidx_set_transform": [_vp, _i32, _vp, _vp],
    "mmidx_set_w": [_vp, _i32],
    "mmidx_set_shard_map": [_vp, _vp],
    "mmidx_add": [_vp, _i64, _vp, _vp, _vp],
    "mmidx_add_codes": [_vp, _i64, _vp, _vp],
    "mmidx_encode": [_vp, _i64, _vp, _vp, _vp],
    "mmidx_add_dev": [_vp, _i64, _vp, _vp, _vp],
    "mmidx_comm_create": [_vp, _i32, _i32, _i32, _i64, _i32, _vp],
    "mmidx_comm_attach": [_vp, _vp],
    "mmidx_comm_destroy": [_vp],
    "mmidx_search_multi_dev": [_vp, _i64, _vp, _i32, _i32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                               C.POINTER(_i64), C.POINTER(_i64), _vp],
    "mmidx_search_multi": [_vp, _i64, _vp, _i32, _vp, _vp, _vp, C.POINTER(_i64), C.POINTER(_i64)],
    "mmidx_search": [_vp, _i64, _vp, _i32, _vp, _vp, _vp],
    "mmidx_search_dev": [_vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp],
    "mmidx_search_shard_dev": [_vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mmidx_coarse_probe_dev": [_vp, _i64, _vp, _i32, _vp, _vp],
    "mmidx_merge_topk_dev": [_i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mmidx_tie_collect_shard_dev": [_vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mmidx_tie_finish_dev": [_i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mmidx_coarse_probe": [_vp, _i64, _vp, _i32, _vp],
    "mmidx_pq_lut": [_vp, _i64, _vp, _vp],
    "mmidx_size": [_vp, C.POINTER(_i64)],
    "mmidx_list_sizes": [_vp, _vp],
    "mmidx_get_vector": [_vp, _i64, _vp],
    "mmidx_scan_bytes": [_vp, _i64, _vp, C.POINTER(_i64)],
    "mmidx_last_timings": [_vp, _vp],
    "mmidx_last_timings_multi": [_vp, _vp],
    "mmidx_enable_timings": [_vp, _i32],
    "mmidx_last_launches": [_vp, C.POINTER(_i32)],
    "mmidx_debug_stats": [_vp, _vp],
    "mmidx_vlad": [_vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _i32],
    "mmidx_vlad_dev": [_vp, _i32, _i32, _i64, _vp, _i64, _vp, _vp, _vp, _vp],
    "mmidx_normalize_rows": [_vp, _i64, _i64, _i32, C.c_double, _i32, _i32],
    "mmidx_normalize_rows_dev": [_vp, _i64, _i64, _i32, _i32, C.c_double, _i32, _vp],
    "mmidx_vlad_multi": [_vp, _i32, _vp, _i32, _i64, _vp, _vp, _i32, _vp, _i32],
    "mmidx_vlad_multi_dev": [_vp, _i32, _vp, _i32, _i64, _vp, _i64, _vp, _i32, _vp, _vp],
    "mmidx_pca_project": [_vp, _vp, _i32, _i32, _i64, _vp, _i32, _vp, _i32],
    "mmidx_pca_project_dev": [_vp, _vp, _i32, _i32, _i64, _vp, _i32, _vp, _vp],
}
for _name, _args in _SIGS.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = C.c_int
lib.mmidx_last_error.restype = C.c_char_p
lib.mmidx_last_error.argtypes = []
lib.mmidx_version.restype = C.c_char_p
lib.mmidx_version.argtypes = []

EXPORTS = sorted(list(_SIGS) + ["mmidx_last_error", "mmidx_version"])


def check(rc):
    if rc != OK:
        raise MmidxError(rc, lib.mmidx_last_error().decode())
    return rc
