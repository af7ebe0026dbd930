"""ctypes binding of liblitparrot_b200.so (the C ABI declared in include/lp_abi.h).

This file is the complete "reference-side binding": the reference is Python, so what a maintainer adds to
call the library is exactly these ctypes prototypes.  There is NO fallback: if the
shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LP_LIB_PATH") or os.path.join(_HERE, "liblitparrot_b200.so")  # LP_LIB_PATH: A/B builds

# enums of lp_abi.h
LP_F32, LP_BF16 = 0, 1
LP_W_F32, LP_W_BF16, LP_W_INT4, LP_W_NF4, LP_W_INT8 = 0, 1, 2, 3, 4
LP_EPI_NONE, LP_EPI_GELU, LP_EPI_SWIGLU, LP_EPI_RESIDUAL = 0, 1, 2, 3
LP_NORM_LAYERNORM, LP_NORM_RMS = 0, 1
LP_ABI_VERSION = 6
LP_WF_AUX_PACKED = 1
LP_WF_AUX_TILED = 2
LP_STEP_LINEAR, LP_STEP_ATTENTION, LP_STEP_EXCHANGE, LP_STEP_SLAB = 0, 1, 2, 3

c_void_p, c_int, c_float, c_size_t, c_u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_uint64


class LpWeight(ctypes.Structure):
    """struct lp_weight"""

    _fields_ = [("w", c_void_p), ("aux0", c_void_p), ("aux1", c_void_p), ("aux2", c_void_p), ("bias", c_void_p),
                ("fmt", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("group", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("reserved", ctypes.c_int32), ("out_bias", c_void_p), ("out_scale", c_void_p)]


class LpStepOp(ctypes.Structure):
    """struct lp_step_op: one entry of the decode-step op table"""

    _fields_ = [("kind", ctypes.c_int32), ("dep", ctypes.c_int32), ("W", ctypes.POINTER(LpWeight)), ("x", c_void_p),
                ("x_is_attention", ctypes.c_int32), ("norm_kind", ctypes.c_int32), ("norm_w", c_void_p), ("norm_b", c_void_p),
                ("eps", c_float), ("epilogue", ctypes.c_int32), ("residual", c_void_p), ("out", c_void_p),
                ("qkv", c_void_p), ("k_cache", c_void_p), ("v_cache", c_void_p),
                ("tp_buf_ptrs", c_void_p), ("tp_pad_ptrs", c_void_p), ("tp_state", c_void_p), ("tp_buf_offset", ctypes.c_uint64),
                ("tp_pad_base", ctypes.c_int32), ("tp_rank", ctypes.c_int32), ("tp_size", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("slab_image", c_void_p), ("slab_meta", c_void_p), ("slab_src", ctypes.c_int32), ("keep_local", ctypes.c_int32)]


class LpSlabMeta(ctypes.Structure):
    """struct lp_slab_meta"""

    _fields_ = [("off", ctypes.c_int64)] + [(n, ctypes.c_int32) for n in
                                            ("nunits", "units_a", "nseg", "row0", "nrb", "unit0", "group_a", "stage_bytes")]


class LpStepGeom(ctypes.Structure):
    """struct lp_step_geom"""

    _fields_ = [("pos", c_void_p), ("idx", c_void_p), ("idx_offset", c_void_p), ("wte", c_void_p), ("x0", c_void_p),
                ("cos", c_void_p), ("sin", c_void_p), ("workspace", c_void_p), ("workspace_bytes", c_size_t),
                ("idx_is_int64", ctypes.c_int32), ("wte_dtype", ctypes.c_int32), ("E", ctypes.c_int32), ("H", ctypes.c_int32),
                ("G", ctypes.c_int32), ("hs", ctypes.c_int32), ("n_elem", ctypes.c_int32), ("max_seq", ctypes.c_int32),
                ("kv_dtype", ctypes.c_int32), ("scale", c_float), ("zero_ptr", c_void_p), ("zero_bytes", c_size_t)]


class LpStepHandle(ctypes.Structure):
    """struct lp_step_handle (opaque)"""

    _fields_ = [("opaque", ctypes.c_uint64 * 32)]


# name -> (restype, argtypes); every symbol declared in include/lp_abi.h
PROTOTYPES = {
    "lp_abi_version": (c_int, []),
    "lp_status_str": (ctypes.c_char_p, [c_int]),
    "lp_last_cuda_error": (ctypes.c_char_p, []),
    "lp_launch_count": (ctypes.c_ulonglong, []),
    "lp_set_pdl": (c_int, [c_int]),
    "lp_set_linear_path": (c_int, [c_int]),
    "lp_debug_stream_trace": (c_int, [c_void_p]),
    "lp_init": (c_int, [c_int]),
    "lp_validate_inputs": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "lp_stage_inputs": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "lp_embed": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "lp_norm": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "lp_linear": (c_int, [c_void_p, c_int, ctypes.POINTER(LpWeight), c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "lp_norm_linear": (c_int, [c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, ctypes.POINTER(LpWeight), c_int, c_void_p,
                               c_void_p, c_int, c_void_p]),
    "lp_set_gemm_pair": (c_int, [c_int]),
    "lp_split_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_int, c_void_p]),
    "lp_gemm_bf16_tc": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                c_int, c_void_p]),
    "lp_gemm_bf16_tc_affine": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                       c_void_p, c_int, c_int, c_void_p]),
    "lp_adapter_attn": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "lp_lora_merge": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_float, c_void_p]),
    "lp_gptq_hessian_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "lp_gptq_hessian_update": (c_int, [c_void_p, c_int, c_void_p, c_int, c_float, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "lp_gptq_find_params": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "lp_gptq_block_sweep": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "lp_gptq_trailing_update": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "lp_swiglu": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "lp_dequant_bf16": (c_int, [ctypes.POINTER(LpWeight), c_void_p, c_void_p]),
    "lp_rope_kv_append": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lp_attn_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "lp_attn_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int,
                               c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "lp_attn_fused_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "lp_attn_decode_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                     c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]),
    "lp_set_attn_prefill_path": (c_int, [c_int]),
    "lp_attn_prefill": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_float, c_int, c_void_p]),
    "lp_decode_step_plan_bytes": (c_size_t, [c_int]),
    "lp_decode_step_workspace_bytes": (c_size_t, [c_int, c_int]),
    "lp_decode_step_plan": (c_int, [ctypes.POINTER(LpStepOp), c_int, ctypes.POINTER(LpStepGeom), c_void_p, c_size_t,
                                    ctypes.POINTER(LpStepHandle)]),
    "lp_decode_step": (c_int, [ctypes.POINTER(LpStepHandle), c_void_p]),
    "lp_decode_step_slab_layout": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(LpSlabMeta), c_int,
                                           ctypes.POINTER(c_int), ctypes.POINTER(c_size_t)]),
    "lp_decode_step_slab_build": (c_int, [ctypes.POINTER(LpWeight), c_void_p, c_int, c_void_p, c_void_p]),
    "lp_decode_step_status": (c_int, [ctypes.POINTER(LpStepHandle), ctypes.POINTER(ctypes.c_int32)]),
    "lp_decode_step_cooperative": (c_int, [ctypes.POINTER(LpStepHandle)]),
    "lp_debug_step_trace": (c_int, [c_void_p]),
    "lp_debug_gemm_stats": (c_int, [c_void_p]),
    "lp_tp_allreduce_residual": (c_int, [c_void_p, c_void_p, c_int, c_int, c_size_t, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                         c_void_p]),
    "lp_sample": (c_int, [c_void_p, c_int, c_int, c_float, c_int, c_u64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lp_sample_bf16": (c_int, [c_void_p, c_int, c_int, c_float, c_int, c_u64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lp_int4_row_bytes": (c_size_t, [c_int]),
    "lp_repack_gptq_int4": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
}

_lock = threading.Lock()
_lib: Optional[ctypes.CDLL] = None
_inited_devices = set()


def load() -> ctypes.CDLL:
    """dlopen the library and bind every prototype.  Does not touch the GPU."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m lit_parrot_b200.build` (needs nvcc, sm_100a). "
                    "lit_parrot_b200 has no CPU or PyTorch fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)  # AttributeError if the symbol is not exported
                fn.restype, fn.argtypes = res, args
            if lib.lp_abi_version() != LP_ABI_VERSION:
                raise RuntimeError(f"ABI mismatch: library {lib.lp_abi_version()} vs binding {LP_ABI_VERSION}")
            _lib = lib
    return _lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        lib = load()
        msg = lib.lp_status_str(status).decode()
        detail = lib.lp_last_cuda_error().decode() if status == -3 else ""
        raise RuntimeError(f"liblitparrot_b200: {what} failed with {msg} {detail}".rstrip())


def init(device_index: int) -> ctypes.CDLL:
    lib = load()
    if device_index not in _inited_devices:
        check(lib.lp_init(device_index), "lp_init")
        _inited_devices.add(device_index)
    return lib
