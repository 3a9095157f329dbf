"""ctypes binding of ``libbayeslm_b200.so`` (the C ABI declared in ``include/bayeslm_b200.h``).

The library is the only compute path of this package: there is no PyTorch / CPU
fallback.  Importing this module never touches the GPU; :func:`lib` loads the shared
object (raising :class:`BlmError` with build instructions if it is missing) and
:func:`init` checks that the device really is an sm_100 part.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbayeslm_b200.so")

BLM_OK = 0
ERR_NAMES = {-1: "BLM_ERR_SHAPE", -2: "BLM_ERR_ALIGN", -3: "BLM_ERR_ARCH", -4: "BLM_ERR_CUDA", -5: "BLM_ERR_ARG"}

ACT_NONE, ACT_GELU, ACT_GPMIX, ACT_SOFTMAX_GRAD, ACT_GELU_GRAD, ACT_GPMIX_GRAD, ACT_GELU_FAST, ACT_GPMIX_FAST = 0, 1, 2, 3, 4, 5, 6, 7
EPS_NONE, EPS_PTR, EPS_PHILOX = 0, 1, 2
MAX_SEG = 6


class BlmError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("N", C.c_int64), ("nseg", C.c_int32), ("act", C.c_int32),
        ("A", C.c_void_p * MAX_SEG), ("B", C.c_void_p * MAX_SEG),
        ("K", C.c_int64 * MAX_SEG), ("lda", C.c_int64 * MAX_SEG), ("ldb", C.c_int64 * MAX_SEG),
        ("bias", C.c_void_p), ("coef", C.c_void_p),
        ("col_scale", C.c_float), ("col_scale_cols", C.c_int32),
        ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("ldc", C.c_int64),
        ("lse", C.c_void_p), ("targets", C.c_void_p), ("grad_scale", C.c_float), ("k_chunk", C.c_int32),
        ("out_pre", C.c_void_p), ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("a_f16", C.c_int32), ("a_mn", C.c_int32), ("b_mn", C.c_int32), ("fast_act", C.c_int32), ("f32_rows32", C.c_int32), ("drop", C.c_void_p),
    ]


class GemmLnDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("A", C.c_void_p), ("lda", C.c_int64), ("B", C.c_void_p), ("ldb", C.c_int64),
        ("bias", C.c_void_p), ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("reserved", C.c_int32),
        ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("ldc", C.c_int64),
    ]


class GemmSampledDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("A", C.c_void_p), ("lda", C.c_int64), ("mu", C.c_void_p), ("ldmu", C.c_int64),
        ("sigma", C.c_void_p), ("eps", C.c_void_p), ("eps_mode", C.c_int32), ("act", C.c_int32),
        ("seed", C.c_uint64), ("stream_id", C.c_uint64),
        ("bias", C.c_void_p), ("coef", C.c_void_p), ("resid", C.c_void_p), ("ldr", C.c_int64),
        ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("ldc", C.c_int64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("mu_f32", C.c_void_p), ("ldmu_f32", C.c_int64), ("lgstd_f32", C.c_void_p),
    ]


class VocabNllDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("V", C.c_int64), ("nseg", C.c_int32), ("reserved", C.c_int32),
        ("H", C.c_void_p * MAX_SEG), ("E", C.c_void_p * MAX_SEG),
        ("K", C.c_int64 * MAX_SEG), ("ldh", C.c_int64 * MAX_SEG), ("lde", C.c_int64 * MAX_SEG),
        ("bias", C.c_void_p), ("targets", C.c_void_p), ("nll", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("lse", C.c_void_p),
    ]


class DropoutDesc(C.Structure):
    _fields_ = [("mask", C.c_void_p), ("p", C.c_float), ("reserved", C.c_int32), ("seed", C.c_uint64),
                ("seed_dev", C.c_void_p), ("stream_id", C.c_uint64)]


_p, _i32, _i64, _u64, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float

# name -> (restype, argtypes); must list every symbol include/bayeslm_b200.h declares
SIGNATURES = {
    "blm_version": (C.c_int, []),
    "blm_last_error": (C.c_char_p, []),
    "blm_init": (C.c_int, [C.c_int]),
    "blm_num_sms": (C.c_int, []),
    "blm_gemm": (C.c_int, [C.POINTER(GemmDesc), _p]),
    "blm_gemm_ln": (C.c_int, [C.POINTER(GemmLnDesc), _p]),
    "blm_gemm_sampled": (C.c_int, [C.POINTER(GemmSampledDesc), _p]),
    "blm_gemm_sampled_workspace_bytes": (_i64, [_i64, _i64]),
    "blm_sigma_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "blm_vocab_nll_workspace_bytes": (_i64, [_i64, _i64]),
    "blm_vocab_nll": (C.c_int, [C.POINTER(VocabNllDesc), _p]),
    "blm_segment_sum": (C.c_int, [_p, _p, _i64, _p, _p]),
    "blm_mc_combine": (C.c_int, [_p, _i64, _i64, _p, _p]),
    "blm_split_bf16": (C.c_int, [_p, _p, _p, _i64, _p]),
    "blm_embed": (C.c_int, [_p, _p, _p, _p, _f, _i64, _i32, _p, _p, _p, _p]),
    "blm_layernorm": (C.c_int, [_p, _p, _p, _f, _i64, _i32, _p, _p, _p, _p]),
    "blm_reparam": (C.c_int, [_p, _i64, _p, _p, _i32, _u64, _u64, _i64, _i64, _p, _p, _p, _p]),
    "blm_philox_normal": (C.c_int, [_u64, _u64, _i64, _p, _p]),
    "blm_philox_normal_scaled": (C.c_int, [_u64, _u64, _i64, _f, _p, _p]),
    "blm_mha_causal": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p]),
    "blm_mha_causal_bf16": (C.c_int, [_p, _p, _i64, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _i64, _p]),
    "blm_kl_workspace_bytes": (_i64, []),
    "blm_kl_gauss": (C.c_int, [_p, _i64, _p, _i64, _i64, _i32, _f, _i32, _p, _p, _p]),
    "blm_gp_lstm_cell": (C.c_int, [_p, _i64, _p, _i32, _i32, _p, _i32, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "blm_gp3_bwd": (C.c_int, [_p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _p]),
    "blm_lstm_cell_step": (C.c_int, [_p, _i64, _p, _p, _i32, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "blm_transpose_split": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _p]),
    "blm_split_transpose": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _i64, _p, _p, _i64, _p]),
    "blm_transpose_bf16": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _i64, _p]),
    "blm_colsum_workspace_bytes": (_i64, [_i64, _i64]),
    "blm_colsum": (C.c_int, [_p, _i64, _i64, _i64, _f, _i32, _p, _p, _p]),
    "blm_colsum_bf16": (C.c_int, [_p, _p, _i64, _i64, _i64, _f, _i32, _p, _p, _p]),
    "blm_layernorm_bwd_workspace_bytes": (_i64, [_i64, _i32]),
    "blm_layernorm_bwd": (C.c_int, [_p, _p, _p, _f, _i64, _i32, _p, _p, _p, _i32, _p, _p]),
    "blm_layernorm_bwd_ex": (C.c_int, [_p, _p, _p, _f, _i64, _i32, _p, _p, _p, _p, _p, _i32, _p, _i32, _p, _p]),
    "blm_mha_causal_bwd": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i32, _i32, _i32, _f, _p, _i64, _p]),
    "blm_mha_causal_bwd_tc": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i32, _i32, _i32, _f, _i32, _p, _i64, _p]),
    "blm_dropout": (C.c_int, [_p, _i64, C.POINTER(DropoutDesc), _p, _p, _p, _p, _p]),
    "blm_mha_causal_bf16_dropout": (C.c_int, [_p, _p, _i64, _p, _i64, _i32, _i32, _i32, C.POINTER(DropoutDesc), _p, _p, _p,
                                              _i64, _p]),
    "blm_mha_causal_bwd_tc_dropout": (C.c_int, [_p, _i64, _p, _i64, _p, _i64, _i32, _i32, _i32, _f, _i32,
                                                C.POINTER(DropoutDesc), _p, _i64, _p]),
    "blm_gpmix_dcoef": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _p, _p]),
    "blm_vnoise_fwd": (C.c_int, [_p, _p, _p, _i32, _u64, _u64, _f, _p, _i64, _i32, _i32, _p, _p]),
    "blm_vnoise_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _u64, _u64, _f, _i64, _i32, _i32, _f, _p, _p, _p, _p, _p]),
    "blm_embed_bwd": (C.c_int, [_p, _p, _f, _i64, _i32, _p, _p]),
    "blm_kl_gauss_bwd": (C.c_int, [_p, _i64, _p, _i64, _i64, _f, _p, _i64, _p, _p]),
    "blm_reparam_bwd": (C.c_int, [_p, _i64, _p, _p, _i32, _u64, _u64, _i64, _i64, _i32, _p, _i64, _p, _p]),
    "blm_reduce_workspace_bytes": (_i64, []),
    "blm_reduce": (C.c_int, [_p, _i64, _i32, _f, _i32, _p, _p, _p]),
    "blm_sgd_momentum": (C.c_int, [_p, _p, _p, _i64, _f, _f, _p, _f, _f, _p]),
    "blm_sgd_momentum_split": (C.c_int, [_p, _p, _p, _i64, _f, _f, _p, _f, _f, _p, _p, _p]),
    "blm_lstm_gates_act": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p]),
    "blm_lstm_bwd_step": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i64, _i64, _p, _p, _p, _p]),
    "blm_vnn_noise": (C.c_int, [_p, _p, _i64, _i32, _p, _p]),
    "blm_rowgroup_add": (C.c_int, [_p, _p, _i64, _i64, _i32, _p, _p, _p, _p]),
    "blm_rowgroup_sum": (C.c_int, [_p, _i64, _i64, _i32, _p, _p]),
    "blm_vnn_kl": (C.c_int, [_p, _p, _i64, _i32, _f, _p, _p, _p, _p]),
    "blm_vnn_drho": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _p]),
    "blm_gp_lstm_bwd_step": (C.c_int, [_p, _i64, _p, _i32, _i32, _p, _p, _p, _p, _p, _i32, _i64, _i32, _p, _p, _p, _i64,
                                       _p, _p]),
    "blm_vocab_from_text": (_p, [C.c_char_p, _i64]),
    "blm_vocab_size": (_i64, [_p]),
    "blm_vocab_id": (_i32, [_p, C.c_char_p, _i64]),
    "blm_vocab_free": (None, [_p]),
    "blm_nbest_scan": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _p, _p, _i32]),
    "blm_nbest_tokenize": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _p, _p, _p, _i32]),
    "blm_nbest_group": (C.c_int, [_p, _p, _i64, _p, _p, _p, _p, _p]),
    "blm_scores_format": (_i64, [_p, _p, _p, _p, _p, _i64, _p, _p, _i64]),
    "blm_lstm_workspace_bytes": (_i64, [_i64, _i64]),
    "blm_lstm_layer": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "blm_lstm_layer_seq": (C.c_int, [_p, C.c_int32, _p, _p, _p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
}

_lib = None
_lock = threading.Lock()
_inited_device = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise BlmError(
                    f"{LIB_PATH} is missing: build it with `tools/build_lib.sh` (or "
                    "`python -c 'import __graft_entry__ as g; g.build()'`). bayeslms_b200 has no "
                    "CPU or PyTorch fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != BLM_OK:
        msg = lib().blm_last_error().decode("utf-8", "replace")
        raise BlmError(f"{what or 'bayeslm_b200'} failed with {ERR_NAMES.get(rc, rc)}: {msg}")


def init(device: int = 0) -> None:
    """Check the device is sm_100 and set kernel attributes.  Idempotent per device."""
    global _inited_device
    if _inited_device == device:
        return
    check(lib().blm_init(int(device)), "blm_init")
    _inited_device = device
