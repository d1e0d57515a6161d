"""ctypes binding of libformoniq_b200.so (the C ABI of include/formoniq_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device
is usable, the calls raise."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FQ_LIB_PATH") or os.path.join(HERE, "lib", "libformoniq_b200.so")  # FQ_LIB_PATH: tuning builds

FQ_MASS, FQ_DIF_TRIAL, FQ_DIF_TEST, FQ_DIF_BOTH, FQ_LUMPED = 0, 1, 2, 3, 4


class FormoniqError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"formoniq_b200 error {code}: {msg}")
        self.code = code


_lib = None

# (name, restype, argtypes): every symbol include/formoniq_b200.h declares
_vp, _sz, _i, _d, _i64 = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int64
_P = C.POINTER
SIGNATURES = [
    ("fq_last_error", C.c_char_p, []),
    ("fq_device_count", _i, []),
    ("fq_ctx_create", _i, [_i, _P(_vp)]),
    ("fq_ctx_destroy", _i, [_vp]),
    ("fq_ctx_set_stream", _i, [_vp, _vp]),
    ("fq_ctx_synchronize", _i, [_vp]),
    ("fq_ctx_launch_count", _i64, [_vp]),
    ("fq_device_cache_trim", _i, []),
    ("fq_ctx_set_timing", _i, [_vp, _i]),
    ("fq_ctx_timing_report", _i, [_vp, _vp, _sz]),
    ("fq_mesh_create", _i, [_vp, _i, _sz, _vp, _vp, _vp, _P(_vp)]),
    ("fq_mesh_create_part", _i, [_vp, _i, _sz, _vp, _vp, _vp, _vp, _vp, _P(_vp)]),
    ("fq_mesh_create_kuhn", _i, [_vp, _i, _vp, _vp, _vp, _vp, _d, _sz, _sz, _P(_vp)]),
    ("fq_mesh_destroy", _i, [_vp]),
    ("fq_mesh_dim", _i, [_vp]),
    ("fq_mesh_ncells", _sz, [_vp]),
    ("fq_mesh_nsimplices", _sz, [_vp, _i]),
    ("fq_mesh_nowned_cells", _sz, [_vp]),
    ("fq_mesh_owned_range", _i, [_vp, _i, _P(_sz), _P(_sz)]),
    ("fq_mesh_held_range", _i, [_vp, _i, _P(_sz), _P(_sz)]),
    ("fq_mesh_set_lengths", _i, [_vp, _vp, _vp]),
    ("fq_mesh_download_cell_faces", _i, [_vp, _vp, _i, _vp]),
    ("fq_mesh_download_lengths", _i, [_vp, _vp, _vp]),
    ("fq_kuhn_cell_faces_host", _i, [_i, _vp, _i, _vp]),
    ("fq_kuhn_counts", _i, [_i, _vp, _vp]),
    ("fq_kuhn_slab_ranges", _i, [_i, _vp, _sz, _sz, _i, _vp]),
    ("fq_elmat_shape", _i, [_i, _i, _i, _P(_i), _P(_i)]),
    ("fq_elmat_batch", _i, [_vp, _vp, _i, _i, _sz, _sz, _i, _vp]),
    ("fq_assemble_symbolic", _i, [_vp, _vp, _i, _i, _sz, _sz, _P(_vp)]),
    ("fq_assemble_numeric", _i, [_vp, _vp, _vp, _i]),
    ("fq_assemble", _i, [_vp, _vp, _i, _i, _i, _P(_vp)]),
    ("fq_hodge_symbolic", _i, [_vp, _vp, _i, _sz, _sz, _sz, _sz, _P(_vp)]),
    ("fq_hodge_numeric", _i, [_vp, _vp, _vp, _i]),
    ("fq_hodge_block", _vp, [_vp, _i]),
    ("fq_hodge_destroy", _i, [_vp]),
    ("fq_hodge_mixed_laplacian", _i, [_vp, _vp, _P(_vp)]),
    ("fq_hodge_mixed_kkt_symmetric", _i, [_vp, _vp, _P(_vp)]),
    ("fq_csr_row_abs_sums", _i, [_vp, _vp, _vp]),
    ("fq_csr_inv_diagonal", _i, [_vp, _vp, _vp]),
    ("fq_vec_mul", _i, [_vp, _vp, _vp, _vp]),
    ("fq_csr_add", _i, [_vp, _vp, _vp, _P(_vp)]),
    ("fq_csr_transpose", _i, [_vp, _vp, _P(_vp)]),
    ("fq_csr_restrict", _i, [_vp, _vp, _vp, _sz, _vp, _sz, _P(_vp)]),
    ("fq_csr_shape", _i, [_vp, _P(_sz), _P(_sz), _P(_sz)]),
    ("fq_csr_row_range", _i, [_vp, _P(_sz), _P(_sz)]),
    ("fq_csr_download", _i, [_vp, _vp, _vp, _vp, _vp]),
    ("fq_csr_download_async", _i, [_vp, _vp, _vp, _vp, _vp]),
    ("fq_ctx_wait_downloads", _i, [_vp]),
    ("fq_csr_upload", _i, [_vp, _sz, _sz, _vp, _vp, _vp, _P(_vp)]),
    ("fq_csr_destroy", _i, [_vp]),
    ("fq_csr_assembly_bytes", _i64, [_vp]),
    ("fq_csr_assembly_shared_bytes", _i64, [_vp]),
    ("fq_csr_plan_build_ms", C.c_double, [_vp]),
    ("fq_csr_plan_cell_visits", _sz, [_vp]),
    ("fq_csr_spmv_bytes", _i64, [_vp]),
    ("fq_vec_create", _i, [_vp, _sz, _P(_vp)]),
    ("fq_vec_wrap", _i, [_vp, _vp, _sz, _P(_vp)]),
    ("fq_vec_destroy", _i, [_vp]),
    ("fq_vec_len", _sz, [_vp]),
    ("fq_vec_upload", _i, [_vp, _vp, _vp]),
    ("fq_vec_download", _i, [_vp, _vp, _vp]),
    ("fq_vec_copy", _i, [_vp, _vp, _vp]),
    ("fq_vec_dot", _i, [_vp, _vp, _vp, _P(_d)]),
    ("fq_vec_scale", _i, [_vp, _vp, _d]),
    ("fq_vec_axpy", _i, [_vp, _vp, _d, _vp]),
    ("fq_vec_device_ptr", _vp, [_vp]),
    ("fq_spmv", _i, [_vp, _vp, _vp, _vp]),
    ("fq_spmv_window", _i, [_vp, _vp, _vp, _sz, _vp]),
    ("fq_matfree_create", _i, [_vp, _vp, _i, _i, _P(_vp)]),
    ("fq_matfree_refresh", _i, [_vp, _vp]),
    ("fq_matfree_destroy", _i, [_vp]),
    ("fq_matfree_shape", _i, [_vp, _P(_sz), _P(_sz)]),
    ("fq_matfree_apply", _i, [_vp, _vp, _vp, _vp]),
    ("fq_matfree_diagonal", _i, [_vp, _vp, _vp]),
    ("fq_linear_form_create", _i, [_vp, _vp, _i, _P(_vp)]),
    ("fq_linear_form_assemble", _i, [_vp, _vp, _vp, _vp]),
    ("fq_linear_form_destroy", _i, [_vp]),
    ("fq_source_form_assemble", _i, [_vp, _vp, _i, _vp, _vp, _vp, _vp]),
    ("fq_weighted_mass_numeric", _i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i]),
    ("fq_vec_ipc_export", _i, [_vp, _vp, _vp]),
    ("fq_vec_ipc_import", _i, [_vp, _vp, _sz, _P(_vp)]),
    ("fq_spmv_peer", _i, [_vp, _vp, _vp, _sz, _sz, _sz, _vp, _sz, _vp, _sz, _vp]),
    ("fq_spmv_peer_epoch", _i, [_vp, _vp, _vp, _sz, _sz, _sz, _vp, _sz, _vp, _vp, _sz, _vp, C.c_double, _vp, _vp]),
    ("fq_flag_signal", _i, [_vp, _vp, _d]),
    ("fq_flag_wait", _i, [_vp, _vp, _d]),
    ("fq_flag_check", _i, [_vp]),
    ("fq_cg_op", _i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _d, _sz, _vp, _P(_sz), _P(_d), _P(_i)]),
    ("fq_minres_op", _i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _d, _sz, _vp, _P(_sz), _P(_d), _P(_i)]),
    ("fq_minres_blockdiag", _i, [_vp, _vp, _i, _P(_vp), _P(_sz), _d, _sz, _vp, _d, _sz, _vp, _P(_sz), _P(_d), _P(_i), _P(_sz)]),
    ("fq_vec_view", _i, [_vp, _vp, _sz, _sz, _P(_vp)]),
    ("fq_cg", _i, [_vp, _vp, _i, _vp, _d, _sz, _vp, _P(_sz), _P(_d), _P(_i)]),
    ("fq_minres", _i, [_vp, _vp, _i, _vp, _d, _sz, _vp, _P(_sz), _P(_d), _P(_i)]),
]


def lib():
    """Load the CUDA library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FormoniqError(-2, f"{LIB_PATH} not found: run `python -m formoniq_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SIGNATURES:
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise FormoniqError(rc, lib().fq_last_error().decode(errors="replace"))
