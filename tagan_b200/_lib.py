"""ctypes binding of libtagan_b200.so (the C ABI declared in include/tagan_b200.h).

No torch types cross this boundary: callers pass raw device pointers (``tensor.data_ptr()``),
sizes and the raw ``cudaStream_t`` of torch's current stream.  There is deliberately no CPU
fallback: if the library is missing or a tensor is not on a CUDA device the call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtagan_b200.so")

_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f32 = C.c_float
_sz = C.c_size_t


class Epilogue(C.Structure):
    """struct tagan_epilogue (fused GEMM epilogues, include/tagan_b200.h)"""
    _fields_ = [("mode", _i32), ("split", _i32), ("in0", _p), ("ld_in0", _i64), ("in1", _p), ("ld_in1", _i64),
                ("out0", _p), ("ld_out0", _i64), ("out1", _p), ("ld_out1", _i64), ("out2", _p), ("ld_out2", _i64),
                ("gamma", _p), ("beta", _p), ("mean", _p), ("rstd", _p)]


EPI_STORE, EPI_RES_LN, EPI_GATES, EPI_BLEND, EPI_GATES_BWD, EPI_STORE_BF16 = range(6)


class HeadWeights(C.Structure):
    """struct tagan_head_weights (classification head parameters or their gradient buffers)"""
    _fields_ = [(n, _p) for n in ("attn0_weight", "attn0_bias", "attn2_weight", "fc0_weight", "fc0_bias", "ln_weight", "ln_bias",
                                  "fc1_weight", "fc1_bias")]


class TimeParams(C.Structure):
    """struct tagan_time_params"""
    _fields_ = [("range", _p), ("mu", _p), ("inv2sig2", _p), ("wc", _p), ("bc", _p), ("nb", _i32)]


# name -> (restype, argtypes); must list every symbol include/tagan_b200.h declares.
SIGNATURES = {
    "tagan_abi_version": (_i32, []),
    "tagan_csr_workspace_bytes": (_sz, [_i64, _i32]),
    "tagan_csr_build": (_i32, [_p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "tagan_csr_build_batched": (_i32, [_p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "tagan_csr_build_part": (_i32, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "tagan_geo_attn_fwd_part": (_i32, [_p, _i64, _p, _p, _i64, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "tagan_geo_attn_bwd_part": (_i32, [_p, _i64, _p, _p, _i64, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p,
                                       _p, _p, _i64, _p, _p, _i64, _p, _p, _p, _p]),
    "tagan_geo_attn_fwd": (_i32, [_p, _p, _p, _i64, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "tagan_geo_attn_bwd": (_i32, [_p, _p, _p, _i64, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p,
                                  _p, _p, _p, _i64, _p, _p, _p, _p]),
    "tagan_geo_attn_fwd_bf16": (_i32, [_p, _p, _p, _i64, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "tagan_geo_attn_bwd_bf16": (_i32, [_p, _p, _p, _i64, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p,
                                       _p, _p, _p, _i64, _p, _p, _p, _p]),
    "tagan_geo_attn_fwd_part_bf16": (_i32, [_p, _i64, _p, _p, _i64, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "tagan_geo_attn_bwd_part_bf16": (_i32, [_p, _i64, _p, _p, _i64, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p,
                                            _p, _p, _i64, _p, _p, _i64, _p, _p, _p, _p]),
    "tagan_layernorm_fwd": (_i32, [_p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _p, _p, _i64, _i32, _p]),
    "tagan_layernorm_bwd_workspace_bytes": (_sz, [_i64, _i32]),
    "tagan_layernorm_bwd": (_i32, [_p, _i64, _p, _i64, _p, _p, _p, _p, _p, _i64, _i32, _p, _p, _p, _sz,
                                   _i64, _i32, _p]),
    "tagan_colsum_workspace_bytes": (_sz, [_i64, _i32]),
    "tagan_colsum": (_i32, [_p, _i64, _p, _p, _sz, _i64, _i32, _p]),
    "tagan_gemm_workspace_bytes": (_sz, [_i32, _i64, _i64, _i64]),
    "tagan_gemm": (_i32, [_i32, _i64, _i64, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _i32, _i32, _p, _sz, _p]),
    "tagan_gemm_tn_colsum_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "tagan_gemm_tn_colsum": (_i32, [_i64, _i64, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _i32, _p, _sz, _p]),
    "tagan_gemm_fused_workspace_bytes": (_sz, [_i32, _i64, _i64, _i64]),
    "tagan_gemm_set_tuning": (None, [_i32, _i32]),
    "tagan_geo_attn_set_l2_policy": (_i32, [_i32]),
    "tagan_geo_attn_set_l2_policy_bf16": (_i32, [_i32]),
    "tagan_gemm_set_trace": (None, [_p]),
    "tagan_gemm_fused": (_i32, [_i32, _i64, _i64, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _p, C.POINTER(Epilogue), _i32,
                                _p, _sz, _p]),
    "tagan_ln_pair_fwd": (_i32, [_p, _i64, _p, _p, _p, _p, _p, _i64, _i32, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _i64,
                                 _i32, _p]),
    "tagan_ln_pair_bwd_workspace_bytes": (_sz, [_i64, _i32]),
    "tagan_ln_pair_bwd": (_i32, [_p, _i64, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _i32, _p, _sz,
                                 _i64, _i32, _p]),
    "tagan_gelu_ln_fwd": (_i32, [_p, _i64, _p, _p, _p, _i64, _p, _p, _i64, _i32, _p]),
    "tagan_gelu_ln_bwd_workspace_bytes": (_sz, [_i64, _i32]),
    "tagan_gelu_ln_bwd": (_i32, [_p, _i64, _p, _i64, _p, _p, _p, _p, _i64, _p, _p, _sz, _i64, _i32, _p]),
    "tagan_window_gelu_fwd": (_i32, [_p, _p, _i32, _i64, _i32, _i32, _p]),
    "tagan_window_gelu_bwd": (_i32, [_p, _p, _p, _i32, _i64, _i32, _i32, _p]),
    "tagan_gru_blend_ln_fwd": (_i32, [_p, _i64, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _i64, _i32, _p, _i64, _p, _i64, _p, _p,
                                      _p, _p, _p, _i64, _i32, _p]),
    "tagan_gru_blend_ln_bwd": (_i32, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _p, _p, _p, _p, _p, _p, _p,
                                      _p, _p, _i32, _p, _sz, _i64, _i32, _p]),
    "tagan_gru_reset_bwd": (_i32, [_p, _p, _p, _i64, _p, _i64, _p, _i64, _i64, _i32, _p]),
    "tagan_mse_workspace_bytes": (_sz, []),
    "tagan_mse_fwd": (_i32, [_p, _i64, _p, _p, _sz, _p]),
    "tagan_mse_bwd": (_i32, [_p, _i64, _p, _p, _p]),
    "tagan_ts_range": (_i32, [_p, _i64, _i32, _p, _p]),
    "tagan_time_bias_fwd": (_i32, [_p, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "tagan_time_bias_bwd_workspace_bytes": (_sz, [_i32, _i32]),
    "tagan_time_bias_bwd": (_i32, [_p, _i64, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _sz, _p]),
    "tagan_tattn_fwd_strided": (_i32, [_p, _p, _p, _i64, _i64, _i32, _i32, _i32, _i64, _i64, _p, _p, _i64, _p, _i32, _f32, _p,
                                       _p, _i32, _i32, _p, _p, _p, _p]),
    "tagan_tattn_bwd_strided": (_i32, [_p, _p, _p, _i64, _i64, _i32, _i32, _i32, _i64, _i64, _p, _p, _i64, _p, _i32, _f32, _p,
                                       _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _sz, _p]),
    "tagan_pack_padded_fwd": (_i32, [_p, _p, _p, _i32, _i32, _i32, _p]),
    "tagan_pack_padded_bwd": (_i32, [_p, _p, _p, _i32, _i32, _i32, _p]),
    "tagan_pool_blocks_workspace_bytes": (_sz, [_i32, _i32]),
    "tagan_pool_blocks_fwd": (_i32, [_p, _i64, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "tagan_pool_blocks_bwd": (_i32, [_p, _i64, _i32, _i32, _i32, _p, _p]),
    "tagan_head_fwd": (_i32, [C.POINTER(HeadWeights), _p, _i32, _i32, _i32, _i32, _i32, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p,
                              _p, _p]),
    "tagan_head_bwd_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "tagan_head_bwd": (_i32, [C.POINTER(HeadWeights), _p, _i32, _i32, _i32, _i32, _i32, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p,
                              _p, _p, _p, C.POINTER(HeadWeights), _p, _sz, _p]),
    "tagan_adam_clip_step": (_i32, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _p, _p, _p]),
    "tagan_tattn_fwd": (_i32, [_p, _p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p, _p, _i64, _p, _i32, _f32, _p,
                               _p, _i32, _i32, _p, _p, _p, _p]),
    "tagan_tattn_bwd_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "tagan_tattn_bwd": (_i32, [_p, _p, _p, _i64, _i64, _i32, _i32, _i32, _i32, _p, _p, _i64, _p, _i32, _f32, _p,
                               _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _sz, _p]),
    "tagan_tattn_mask_allones": (_i32, [_p, _i64, _i32, _f32, _p, _i64, _p, _p]),
    "tagan_gates_fwd": (_i32, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _i64, _i32, _p]),
    "tagan_gates_bwd": (_i32, [_p, _i64, _p, _p, _p, _p, _i64, _p, _i64, _p, _i64, _i32, _i64, _i32, _p]),
    "tagan_blend_fwd": (_i32, [_p, _i64, _p, _p, _i64, _p, _p, _i32, _i64, _i32, _p]),
    "tagan_blend_bwd": (_i32, [_p, _p, _p, _p, _i64, _p, _i64, _p, _p, _i64, _i32, _i32, _i64, _i32, _p]),
    "tagan_skip_window_fwd": (_i32, [_p, _p, _i32, _i64, _i32, _i32, _p]),
    "tagan_skip_window_bwd": (_i32, [_p, _p, _p, _p, _i32, _i64, _i32, _i32, _p]),
    "tagan_bank_gather": (_i32, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _p, _p]),
    "tagan_bank_update": (_i32, [_p, _p, _p, _p, _p, _p, _p, _i64, _p, _i64, _i64, _i32, _i32, _i32, _p, _p, _i32,
                                 C.c_double, _i32, _p, _p, _p, _p]),
    "tagan_bank_decay_all": (_i32, [_p, _p, _f32, _i32, _i32, _p]),
    "tagan_axpby": (_i32, [_p, _f32, _p, _f32, _p, _i64, _p]),
    "tagan_gelu_fwd": (_i32, [_p, _p, _i64, _p]),
    "tagan_gelu_bwd": (_i32, [_p, _p, _p, _i64, _p]),
    "tagan_scale_rows": (_i32, [_p, _i64, _p, _p, _i64, _i64, _i32, _i32, _p]),
    "tagan_decay_scale": (_i32, [_p, _i64, _i32, _p, _i64, _p]),
}

_lib = None


class TaganLibraryError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TaganLibraryError(
            f"{LIB_PATH} not found: build it with `python -m tagan_b200.build` "
            "(there is no CPU / PyTorch fallback for the TAGAN hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise TaganLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_ERRORS = {-1: "invalid argument", -2: "unsupported shape", -3: "workspace too small"}


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"{what}: {_ERRORS.get(rc, rc)} (libtagan_b200 code {rc})")
    raise RuntimeError(f"{what}: CUDA error {rc}")
