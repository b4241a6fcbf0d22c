"""ctypes binding of libchaorec_b200.so (the C ABI declared in include/chaorec_b200.h).

There is no fallback: if the library is missing the import of any compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
# (CHAOREC_LIB_PATH: diagnostic builds of the same ABI, e.g. the pipeline-floor variants of scripts/gpu_eval_floor.sh)
import os as _os
LIB_PATH = Path(_os.environ["CHAOREC_LIB_PATH"]) if _os.environ.get("CHAOREC_LIB_PATH") else _PKG / "libchaorec_b200.so"

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f32 = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/chaorec_b200.h one to one
SIGNATURES = {
    "chaorec_abi_version": (C.c_int, []),
    "chaorec_error_string": (C.c_char_p, [C.c_int]),
    "chaorec_coo_to_csr_workspace_bytes": (_sz, [_i64, _i64]),
    "chaorec_coo_to_csr": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _p, _p, _p, _p, _sz, _p]),
    "chaorec_degree_inv_sqrt": (C.c_int, [_p, _i64, _f32, _p, _p]),
    "chaorec_spmm_hub_threshold": (_i32, []),
    "chaorec_spmm_hub_chunk": (_i32, []),
    "chaorec_spmm_hub_plan": (C.c_int, [_i64, _p, _p]),
    "chaorec_spmm_hub_row_chunks": (_i64, [_i64, _i32]),
    "chaorec_spmm_csr_f32": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _f32, _i32, _p, _p, _i32, _i64, _i32, _i32, _p,
                                       _p, _p, _i64, _i64, _i32, _p]),
    "chaorec_spmm_csr_push_f32": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i32, _i64, _i32, _i32, _p, _p, _p, _i64, _i64, _i32,
                                            _p, _i32, _i64, _i32, _p]),
    "chaorec_reduce_slots_broadcast_f32": (C.c_int, [_p, _i32, _i64, _i64, _i32, _i64, _p, _p, _f32, _i32, _p, _i64, _i32, _p]),
    "chaorec_sddmm_csr_f32": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, _p]),
    "chaorec_sample_negatives": (C.c_int, [_p, _i64, _p, _p, _i64, _i64, C.c_uint64, C.c_uint64, _p, _p]),
    "chaorec_bpr_loss_grad_f32": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i32, _i64, _f32, _p, _p, _p]),
    "chaorec_score_mask_topk_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "chaorec_uses_tensor_engine": (C.c_int, [_i64, _i64, _i32, _i32]),
    "chaorec_score_mask_topk_fallback_rows": (_p, [_p, _i64, _i64, _i32, _i32]),
    "chaorec_score_mask_topk_f32": (C.c_int, [_p, _p, _p, _p, _f32, _i64, _i64, _i32, _i32, _i64, _p, _p, _p, _sz,
                                              _p]),
    "chaorec_topk_hits": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _p, _p]),
    "chaorec_metrics_from_hits": (C.c_int, [_p, _p, _i64, _p, _i32, _p, _p]),
    "chaorec_l2_normalize_f32": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "chaorec_knn_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "chaorec_knn_cosine_topk": (C.c_int, [_p, _i64, _i32, _i32, C.c_int, _i64, _i64, _p, _p, _p, _sz, _p]),
    "chaorec_cooccurrence_max_users": (_i64, []),
    "chaorec_cooccurrence_topk": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i32, _i64, _i64, _p, _p, _p, _p]),
    "chaorec_probe_fp32_fma": (C.c_int, [_i64, _p, _p, _p]),
    "chaorec_probe_tmem_read": (C.c_int, [_i64, _i32, _p, _p, _p]),
}

_lib = None


ABI_VERSION = 3      # CHAOREC_ABI_VERSION of include/chaorec_b200.h this binding was written against


def load() -> C.CDLL:
    """Loads the CUDA library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m chaorec_b200.build` "
                "(chaorec_b200 has no CPU or PyTorch fallback path)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.chaorec_abi_version() != ABI_VERSION:
            raise RuntimeError("libchaorec_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().chaorec_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def ptr(t: torch.Tensor | None):
    """Device (or host) pointer of a contiguous tensor, None -> NULL."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor passed to the C ABI must be contiguous")
    return C.c_void_p(t.data_ptr())


def stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: chaorec_b200 runs on the GPU only (no CPU fallback)")


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """256-byte aligned scratch (torch's caching allocator aligns to 512 bytes)."""
    t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    assert t.data_ptr() % 256 == 0
    return t
