"""ctypes binding of libautofunc_b200.so -- one prototype per symbol in include/autofunc_b200.h.

The library is the product; there is no CPU fallback.  Importing this module only loads the
shared object (which works on a host without a GPU); creating a context needs a B200.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AUTOFUNC_B200_LIB") or os.path.join(_HERE, "libautofunc_b200.so")  # override: A/B builds only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "autofunc_b200.h")

AF_OK, AF_ESHAPE, AF_ECUDA, AF_ENOMEM, AF_EINVAL, AF_ENOTSUP, AF_ENCCL = 0, -1, -2, -3, -4, -5, -6
AF_DP_OFF, AF_DP_BLOCKING, AF_DP_OVERLAP, AF_DP_DEFERRED = 0, 1, 2, 3
AF_DP_ID_BYTES = 128

P = C.c_void_p          # opaque handles and device pointers
D = C.c_void_p          # device double*
H = C.c_void_p          # host double*
I64 = C.c_int64
INT = C.c_int

PROTOTYPES = {
    "af_ctx_create": (INT, [INT, C.POINTER(P)]),
    "af_ctx_create_on_stream": (INT, [INT, P, C.POINTER(P)]),
    "af_ctx_destroy": (INT, [P]),
    "af_sync": (INT, [P]),
    "af_last_error": (C.c_char_p, []),
    "af_version": (INT, []),
    "af_launch_count": (I64, [P]),
    "af_set_gemm_path": (INT, [P, INT]),
    "af_set_option": (INT, [P, C.c_char_p, INT]),
    "af_trace_begin": (INT, [P, INT]),
    "af_trace_end": (INT, [P, INT, C.POINTER(INT), C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(INT)]),
    "af_graph_begin": (INT, [P]),
    "af_graph_end": (INT, [P, C.POINTER(P)]),
    "af_graph_launch": (INT, [P]),
    "af_graph_launches": (I64, [P]),
    "af_graph_destroy": (INT, [P]),
    "af_malloc": (INT, [P, I64, C.POINTER(D)]),
    "af_free": (INT, [P, D]),
    "af_upload": (INT, [P, D, H, I64]),
    "af_download": (INT, [P, H, D, I64]),
    "af_zero": (INT, [P, D, I64]),
    "af_copy": (INT, [P, D, D, I64]),
    "af_lintran_fwd": (INT, [P, D, D, D, I64, I64, I64]),
    "af_lintran_fwd_r": (INT, [P, D, D, D, D, D, D, I64, I64, I64]),
    "af_lintran_wgrad": (INT, [P, D, D, D, I64, I64, I64]),
    "af_lintran_wgrad_r": (INT, [P, D, D, D, D, D, D, I64, I64, I64]),
    "af_lintran_igrad": (INT, [P, D, D, D, I64, I64, I64]),
    "af_lintran_igrad_r": (INT, [P, D, D, D, D, D, D, I64, I64, I64]),
    "af_lintran_backward": (INT, [P, D, D, D, D, D, I64, I64, I64, INT]),
    "af_lintran_backward_r": (INT, [P, D, D, D, D, D, D, D, D, D, D, I64, I64, I64, INT]),
    "af_bias_fwd": (INT, [P, D, D, D, I64, I64]),
    "af_bias_fwd_r": (INT, [P, D, D, D, D, D, D, I64, I64]),
    "af_bias_grad": (INT, [P, D, D, I64, I64]),
    "af_bias_grad_r": (INT, [P, D, D, D, D, I64, I64]),
    "af_sigmoid_fwd": (INT, [P, D, D, I64]),
    "af_sigmoid_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_sigmoid_bwd": (INT, [P, D, D, I64]),
    "af_sigmoid_bwd_r": (INT, [P, D, D, D, D, I64]),
    "af_exp_fwd": (INT, [P, D, D, I64]),
    "af_exp_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_exp_bwd": (INT, [P, D, D, I64]),
    "af_exp_bwd_r": (INT, [P, D, D, D, D, I64]),
    "af_add": (INT, [P, D, D, D, I64]),
    "af_sub": (INT, [P, D, D, D, I64]),
    "af_mul": (INT, [P, D, D, D, I64]),
    "af_mul_r": (INT, [P, D, D, D, D, D, I64]),
    "af_mul_bwd": (INT, [P, D, D, D, D, D, D, I64]),
    "af_scale": (INT, [P, D, C.c_double, I64]),
    "af_axpy": (INT, [P, C.c_double, D, D, I64]),
    "af_square_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_square_bwd_r": (INT, [P, D, D, D, D, I64]),
    "af_sum_all": (INT, [P, D, D, D, D, I64]),
    "af_fill": (INT, [P, D, C.c_double, I64]),
    "af_log_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_log_bwd_r": (INT, [P, D, D, D, D, I64]),
    "af_logsigmoid_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_logsigmoid_bwd_r": (INT, [P, D, D, D, D, I64]),
    "af_sin_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_sin_bwd_r": (INT, [P, D, D, D, D, I64]),
    "af_tanh_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_tanh_bwd_r": (INT, [P, D, D, D, D, I64]),
    "af_inverse_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_inverse_bwd_r": (INT, [P, D, D, D, D, I64]),
    "af_pow_fwd_r": (INT, [P, D, D, C.c_double, D, D, I64]),
    "af_pow_bwd_r": (INT, [P, D, D, C.c_double, D, D, I64]),
    "af_div": (INT, [P, D, D, D, I64]),
    "af_div_bwd": (INT, [P, D, D, D, D, D, I64]),
    "af_add_scaler": (INT, [P, D, C.c_double, D, I64]),
    "af_add_dev": (INT, [P, D, D, C.c_double, D, I64]),
    "af_scale_dev": (INT, [P, D, D, D, D, D, D, I64]),
    "af_fill_dev": (INT, [P, D, D, I64]),
    "af_dot": (INT, [P, D, D, D, D, D, I64]),
    "af_max_abs": (INT, [P, D, D, INT, I64]),
    "af_norm_fwd_r": (INT, [P, D, D, D, D, I64]),
    "af_norm_bwd_r": (INT, [P, D, D, D, D, D, D, D, D, I64]),
    "af_softmax_fwd_r": (INT, [P, D, D, C.c_double, D, D, I64, I64]),
    "af_softmax_bwd_r": (INT, [P, D, D, C.c_double, D, D, I64, I64]),
    "af_sqerr_fwd_r": (INT, [P, D, D, D, D, D, D, I64]),
    "af_sqerr_bwd_r": (INT, [P, D, D, D, D, D, D, C.c_double, D, D, I64]),
    "af_dense_fwd": (INT, [P, D, D, D, D, D, D, D, D, I64, I64, I64, INT]),
    "af_dense_igrad": (INT, [P, D, D, D, D, D, D, D, D, I64, I64, I64]),
    "af_mlp_create": (INT, [P, C.POINTER(I64), INT, I64, C.POINTER(P)]),
    "af_mlp_destroy": (INT, [P]),
    "af_mlp_set_param": (INT, [P, INT, H, I64]),
    "af_mlp_set_rvector": (INT, [P, INT, H, I64]),
    "af_mlp_param_ptr": (D, [P, INT]),
    "af_mlp_rvector_ptr": (D, [P, INT]),
    "af_mlp_grad_ptr": (D, [P, INT]),
    "af_mlp_rgrad_ptr": (D, [P, INT]),
    "af_mlp_input_ptr": (D, [P]),
    "af_mlp_output_ptr": (D, [P]),
    "af_mlp_routput_ptr": (D, [P]),
    "af_mlp_upstream_ptr": (D, [P]),
    "af_mlp_upstream_r_ptr": (D, [P]),
    "af_mlp_zero_grads": (INT, [P]),
    "af_mlp_add_to_vars": (INT, [P, C.c_double]),
    "af_mlp_get_param": (INT, [P, INT, H, I64]),
    "af_mlp_step": (INT, [P, INT, INT]),
    "af_mlp_step_fresh": (INT, [P, INT]),
    "af_mlp_set_graph": (INT, [P, INT]),
    "af_mlp_step_host": (INT, [P, INT, H, H, H, H, H, C.POINTER(H), C.POINTER(H)]),
    "af_mlp_submit_host": (INT, [P, INT, H, H, H, H, H, C.POINTER(H), C.POINTER(H)]),
    "af_mlp_wait": (INT, [P]),
    "af_mlp_submit_compute_host": (INT, [P, INT, H, H, H]),
    "af_mlp_submit_download_host": (INT, [P, INT, H, H, C.POINTER(H), C.POINTER(H)]),
    "af_mlp_bind_slot": (INT, [P, INT]),
    "af_mlp_set_grad_ready": (INT, [P, C.c_void_p, C.c_void_p, INT]),
    "af_dp_unique_id": (INT, [C.c_void_p]),
    "af_dp_init": (INT, [P, INT, INT, C.c_void_p]),
    "af_dp_init_local": (INT, [C.POINTER(P), INT]),
    "af_dp_destroy": (INT, [P]),
    "af_dp_info": (INT, [P, C.POINTER(INT), C.POINTER(INT), C.POINTER(INT)]),
    "af_dp_group_begin": (INT, []),
    "af_dp_group_end": (INT, []),
    "af_dp_calls": (I64, [P]),
    "af_allreduce_sum": (INT, [P, D, I64]),
    "af_mlp_set_dp": (INT, [P, INT]),
    "af_mlp_step_dp": (INT, [P, INT]),
    "af_mlp_step_fresh_dp": (INT, [P, INT]),
    "af_mlp_dp_join": (INT, [P]),
    "af_mlp_grad_slab": (I64, [P, C.POINTER(D)]),
    "af_mlp_set_host_shard": (INT, [P, INT]),
    "af_lstm_step_dp": (INT, [P, INT]),
    "af_copy_2d": (INT, [P, D, I64, D, I64, I64, I64]),
    "af_lstm_create": (INT, [P, I64, I64, I64, I64, I64, C.POINTER(P)]),
    "af_lstm_destroy": (INT, [P]),
    "af_lstm_set_param": (INT, [P, INT, H, I64]),
    "af_lstm_set_rvector": (INT, [P, INT, H, I64]),
    "af_lstm_param_ptr": (D, [P, INT]),
    "af_lstm_grad_ptr": (D, [P, INT]),
    "af_lstm_rgrad_ptr": (D, [P, INT]),
    "af_lstm_output_ptr": (D, [P]),
    "af_lstm_routput_ptr": (D, [P]),
    "af_lstm_upstream_ptr": (D, [P]),
    "af_lstm_upstream_r_ptr": (D, [P]),
    "af_lstm_grad_slab": (I64, [P, C.POINTER(D)]),
    "af_lstm_zero_grads": (INT, [P]),
    "af_lstm_add_to_vars": (INT, [P, C.c_double]),
    "af_lstm_set_inputs": (INT, [P, D]),
    "af_lstm_forward": (INT, [P, INT]),
    "af_lstm_backward": (INT, [P, INT]),
    "af_lstm_run_host": (INT, [P, INT, H, H, H, H, H, C.POINTER(H), C.POINTER(H)]),
}

GRAD_READY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int64)

_lib = None


class AutofuncError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"autofunc_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class ShapeError(AutofuncError, ValueError):
    """AF_ESHAPE -- the reference panics with the same message (lin_tran.go:26-27 etc.)."""


def load() -> C.CDLL:
    """Load the shared library (building nothing).  Fails loudly when it is missing: there is
    no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(autofunc_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != AF_OK:
        msg = (load().af_last_error() or b"").decode("utf-8", "replace")
        raise (ShapeError if rc == AF_ESHAPE else AutofuncError)(rc, msg)
