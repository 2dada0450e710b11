"""ctypes binding of ``libvqa_sm100.so`` (declared in ``include/vqa_b200.h``).

The library is the product: if it is missing or cannot be loaded the import of the compute path
fails loudly -- there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvqa_sm100.so")

ABI_VERSION = 11
PREC_TF32X3 = 0
PREC_TF32 = 1
PREC_TF32X3_HP = 2
GEMM_RELU = 1
GEMM_ACCUMULATE = 8
GEMM_NO_CLUSTER = 16
GC_RELU = 1

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_f = C.c_float
_u64 = C.c_ulonglong

# name -> argument types (all return int unless noted); mirrors include/vqa_b200.h one-to-one
SIGNATURES = {
    "vqa_gemm_f32": [_p, _ll, _i, _p, _ll, _i, _p, _ll, _i, _i, _i, _p, _p, _ll, _i, _p, _ll, _f, _i, _i, _i, _i, _p],
    "vqa_split_bf16_f32": [_p, _ll, _p, _p, _ll, _ll, _i, _p],
    "vqa_dropout_split_f32": [_p, _ll, _p, _p, _ll, _ll, _i, _f, _u64, _u64, _p, _p],
    "vqa_gemm_bf16s": [_p, _p, _ll, _i, _p, _p, _ll, _i, _p, _ll, _p, _p, _ll, _i, _i, _i, _p, _p, _ll, _i, _p, _ll, _p, _ll,
                       _f, _i, _i, _i, _i, _p, _i, _p],
    "vqa_dropout_f32": [_p, _p, _ll, _f, _u64, _u64, _p, _p],
    "vqa_weight_norm_fwd_f32": [_p, _p, _p, _i, _i, _p],
    "vqa_weight_norm_split_f32": [_p, _p, _i, _i, _i, _i, _p, _p, _ll, _p],
    "vqa_weight_norm_bwd_f32": [_p, _p, _p, _p, _p, _i, _i, _p],
    "vqa_colsum_f32": [_p, _ll, _p, _p, _ll, _i, _p, _p],
    "vqa_segment_sum_f32": [_p, _p, _i, _i, _i, _p],
    "vqa_adjacency_topk_fwd_f32": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "vqa_topk_softmax_f32": [_p, _p, _p, _i, _i, _i, _p],
    "vqa_adjacency_topk_bwd_f32": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "vqa_graphconv_fwd_f32": [_p, _ll, _p, _p, _p, _ll, _p, _p, _ll, _i, _i, _i, _i, _i, _i, _f, _u64, _u64, _p, _p],
    "vqa_graphconv_pool_fwd_f32": [_p, _ll, _p, _p, _ll, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "vqa_graphconv_bwd_f32": [_p, _ll, _p, _p, _p, _ll, _p, _p, _p, _ll, _p, _p, _ll, _p, _i, _i, _i, _i, _i, _p],
    "vqa_graphconv_edge_coef": [_p, _p, _p, _ll, _p, _p, _p, _i, _i, _i, _i, _p],
    "vqa_graphconv_mma_fwd": [_p, _p, _ll, _p, _p, _p, _ll, _p, _p, _p, _ll, _i, _i, _i, _i, _i, _i, _f, _u64, _u64, _p, _p, _p, _p],
    "vqa_graphconv_mma_pool_fwd": [_p, _p, _ll, _p, _p, _ll, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p],
    "vqa_graphconv_mma_bwd_data": [_p, _p, _ll, _p, _p, _p, _ll, _p, _p, _p, _ll, _i, _i, _i, _i, _i, _p, _p, _p],
    "vqa_graphconv_pool_bwd_data": [_p, _p, _p, _p, _p, _p, _ll, _i, _i, _i, _i, _i, _p],
    "vqa_graphconv_mma_bwd_edges": [_p, _p, _ll, _p, _p, _p, _p, _ll, _p, _p, _p, _ll, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "vqa_graphconv_edge_blocks": [_i, _i, _i],
    "vqa_graphconv_edge_bwd_f32": [_p, _p, _p, _p, _ll, _p, _p, _p, _i, _i, _i, _i, _p],
    "vqa_gaussian_weights_f32": [_p, _p, _p, _ll, _i, _p],
    "vqa_embed_gather_split": [_p, _ll, _p, _ll, _i, _p, _p, _ll, _i, _i, _p, _p],
    "vqa_embed_scatter_add_f32": [_p, _ll, _p, _ll, _p, _p, _ll, _i, _i, _i, _p],
    "vqa_gru_cell_fwd_f32": [_p, _ll, _p, _p, _p, _p, _i, _p, _p, _p, _ll, _p, _i, _i, _p],
    "vqa_gru_step_fused": [_p, _p, _ll, _p, _p, _ll, _p, _ll, _p, _p, _p, _i, _p, _p, _p, _ll, _p, _p, _i, _i, _p],
    "vqa_gru_seq_fused": [_p, _p, _ll, _p, _p, _ll, _p, _ll, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "vqa_gru_cell_bwd_f32": [_p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _ll, _p, _i, _i, _p],
    "vqa_gate_bwd_f32": [_p, _p, _p, _p, _p, _ll, _p],
    "vqa_mlsm_loss_blocks": [_ll],
    "vqa_set_sm_budget": [_i],
    "vqa_patch_operator_fwd_f32": [_p, _p, _p, _ll, _i, _i, _i, _p],
    "vqa_patch_operator_bwd_f32": [_p, _p, _p, _p, _p, _ll, _i, _i, _i, _p],
    "vqa_adam_flat_p2p": [_p, _p, _ll, _i, _p, _p, _p, _p, _i, _i, _p, _f, _f, _f, _f, _f, _p, _p],
    "vqa_memcpy2d_async": [_p, _ll, _p, _ll, _ll, _ll, _p],
    "vqa_adam_flat_mc": [_p, _p, _p, _p, _p, _ll, _ll, _p, _f, _f, _f, _f, _f, _p, _p],
    "vqa_p2p_barrier": [_p, _i, _i, _p, _p],
    "vqa_mlsm_loss_fwd_f32": [_p, _p, _ll, _f, _p, _p, _p, _p],
    "vqa_mlsm_loss_bwd_f32": [_p, _p, _p, _p, _ll, _f, _p],
    "vqa_adam_flat_f32": [_p, _i, _p, _p, _p, _p, _f, _f, _f, _f, _f, _p, _p],
    "vqa_gather_image_f32": [_p, _i, _p, _p, _ll, _p, _i, _i, _i, _p, _p],
    "vqa_scatter_targets_f32": [_p, _p, _p, _p, _i, _i, _p, _p],
}
EXPORTS = ["vqa_last_error", "vqa_abi_version"] + list(SIGNATURES)

_lib = None


class VqaKernelError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library once; raise if it is absent or its ABI does not match."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C vqa-project_b200/csrc). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    lib.vqa_last_error.restype = C.c_char_p
    lib.vqa_last_error.argtypes = []
    lib.vqa_abi_version.restype = _i
    lib.vqa_abi_version.argtypes = []
    if lib.vqa_abi_version() != ABI_VERSION:
        raise ImportError(f"libvqa_sm100.so ABI {lib.vqa_abi_version()} != expected {ABI_VERSION}; rebuild")
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = _i
        fn.argtypes = args
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke an entry point; non-zero status -> VqaKernelError with the library's message."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise VqaKernelError(f"{name} failed ({rc}): {lib.vqa_last_error().decode(errors='replace')}")
