"""ctypes binding of libhealswin_b200.so (include/healswin_b200.h).

The library is required: importing this module raises if it cannot be loaded.  It is built
in-tree by ``python -m heal_swin_b200.build`` (``__graft_entry__.build()`` does that); there is
no fallback implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (HEALSWIN_B200_LIB: another build of the same library, for same-box A/B runs of build variants)
LIB_PATH = os.environ.get("HEALSWIN_B200_LIB") or os.path.join(_HERE, "libhealswin_b200.so")

HS_OK, HS_ERR_ARG, HS_ERR_CUDA, HS_ERR_UNSUPPORTED = 0, 1, 2, 3
SHIFT_NONE, SHIFT_NEST_ROLL, SHIFT_NEST_GRID, SHIFT_RING = 0, 1, 2, 3
ATTN_COS = 1
ATTN_NO_TC = 2
ATTN_NO_TRUNC_COMP = 4
MLP_GRAD16 = 256
GEMM_PLAIN, GEMM_ADD, GEMM_GELU, GEMM_GELU_GRAD = 0, 1, 2, 3
GEMM_GELU_C, GEMM_GELU_GRAD_C = 6, 7  # compact forms: the activation's derivative as FP16 instead of z
PREC_BF16X3, PREC_TF32, PREC_BF16 = 0, 1, 2

STRATEGY_CODES = {"nest_roll": SHIFT_NEST_ROLL, "nest_grid_shift": SHIFT_NEST_GRID, "ring_shift": SHIFT_RING}

_p = C.c_void_p
_i, _i64, _u32, _f, _u64 = C.c_int, C.c_int64, C.c_uint32, C.c_float, C.c_uint64

# name -> argtypes; restype is int (status) unless listed in _RESTYPES.  tests/test_abi.py checks
# that every function declared in include/healswin_b200.h appears here and is exported.
SIGNATURES = {
    "hs_last_error": [],
    "hs_version": [],
    "hs_device_info": [_p, _p, _p, _p, _i],
    "hs_nest_win_idcs": [_i, _p],
    "hs_rel_pos_index": [_i, _p],
    "hs_nest2ring": [_i64, _p, _p, _i64],
    "hs_ring2nest": [_i64, _p, _p, _i64],
    "hs_shift_tables": [_i, _i64, _i, _i, _i, _p, _p, _p],
    "hs_attn_mask_from_groups": [_p, _i64, _i, _p],
    "hs_gather_rows": [_p, _p, _p, _i, _i64, _i, _p],
    "hs_rel_bias_expand": [_p, _p, _p, _i, _i, _i, _p],
    "hs_rel_bias_reduce": [_p, _p, _p, _i, _i, _i, _p],
    "hs_layernorm_fwd": [_p, _p, _p, _p, _p, _p, _i64, _f, _u64, _p, _p, _p, _i64, _i, _f, _p],
    "hs_layernorm_bwd": [_p, _p, _p, _p, _p, _p, _p, _i64, _f, _u64, _p, _p, _p, _p, _i64, _i, _p],
    "hs_linear_wgrad_supported": [_i64, _i, _i],
    "hs_linear_wgrad": [_p, _p, _p, _p, _i64, _i, _i, _u32, _p],
    "hs_mlp_dgrad_gelu_supported": [_i64, _i, _i],
    "hs_mlp_dgrad_gelu": [_p, _p, _p, _p, _f, _u64, _p, _i64, _i, _i, _u32, _p],
    "hs_weight_split": [_p, _i, _i, _i, _i, _i, _p, _p],
    "hs_weight_split_batch": [_p, _i, _i, _p],
    "hs_gemm3_supported": [_i64, _i, _i],
    "hs_gemm3": [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _f, _u64, _p],
    "hs_gemm3_ln_supported": [_i64, _i, _i, _i],
    "hs_gemm3_ln": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _f, _i, _p],
    "hs_gemm3_lnin_supported": [_i64, _i, _i],
    "hs_gemm3_lnin": [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _f, _i, _p],
    "hs_ln_head_supported": [_i64, _i, _i],
    "hs_ln_head_fwd": [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _i, _f, _p],
    "hs_ln_head_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i, _i, _p],
    "hs_cross_entropy_supported": [_i],
    "hs_cross_entropy": [_p, _p, _i, _p, _p, _i, _i, _i64, _i64, _p],
    "hs_bias_gelu_supported": [_i64, _i],
    "hs_bias_gelu_fwd": [_p, _p, _f, _u64, _p, _i64, _i, _p],
    "hs_bias_gelu_bwd": [_p, _p, _p, _f, _u64, _p, _p, _i64, _i, _p],
    "hs_window_attn_fwd": [_p, _p, _p, _p, _p, _p, _f, _f, _u64, _p, _p, _i, _i64, _i, _i, _i, _u32, _p],
    "hs_window_attn_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _f, _f, _u64, _p, _p, _p, _i, _i64, _i, _i, _i, _u32, _p],
}
_RESTYPES = {"hs_last_error": C.c_char_p}


def _load():
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m heal_swin_b200.build` "
            "(needs nvcc; there is no CPU fallback for the HEAL-SWIN hot path)."
        )
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift, fail loudly
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name, C.c_int)
    return lib


lib = _load()


def last_error() -> str:
    return lib.hs_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Status -> exception, mirroring the reference's error convention: bad
    shapes/arguments are AssertionError like the reference's asserts, the rest RuntimeError."""
    if rc == HS_OK:
        return
    msg = last_error()
    if rc == HS_ERR_ARG:
        raise AssertionError(msg)
    raise RuntimeError(msg)


def ptr(t):
    """Device/host pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "healswin_b200 kernels need contiguous tensors"
    return C.c_void_p(t.data_ptr())


def current_stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    import torch

    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "heal_swin_b200 runs on CUDA tensors only (B200 / sm_100a); there is no CPU fallback. "
                "Move the module and its inputs to a CUDA device."
            )
        if t.device.index != torch.cuda.current_device():
            # kernels are launched on the current device's stream: a tensor of another device would be a silent
            # wrong-device access (one process per GPU is the supported layout, heal_swin/train.py:187)
            raise RuntimeError(
                f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: call "
                "torch.cuda.set_device(...) (or wrap the call in torch.cuda.device(...)) first"
            )
