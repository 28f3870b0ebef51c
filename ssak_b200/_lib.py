"""ctypes binding of libssak_b200.so (the C ABI declared in include/ssak_b200.h).

There is NO fallback: if the CUDA library is missing the import of any op fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSAK_B200_LIB") or os.path.join(_HERE, "libssak_b200.so")   # (override: A/B runs of two builds)

_p, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t

# name -> (restype, argtypes); kept in one table so tests can check it against the header.
SIGNATURES = {
    "ssak_b200_version": (C.c_int, []),
    "ssak_b200_strerror": (C.c_char_p, [C.c_int]),
    "ssak_b200_last_cuda_error": (C.c_int, []),
    "ssak_ctc_loss_workspace_bytes": (_sz, [_i64, _i64, _i64, C.c_int]),
    "ssak_ctc_loss_workspace_bytes_v": (_sz, [_i64, _i64, _i64, _i64, C.c_int]),
    "ssak_ctc_loss_launches": (C.c_int, [_i64, _i64, _i64, _i32]),
    "ssak_ctc_loss_path_flags": (C.c_int, [_p, _i64, _i64, _i64, _i64, _i32, _p, _p]),
    "ssak_ctc_loss_supported": (C.c_int, [_i64, _i64, _i64, _i64]),
    "ssak_ctc_loss_forward": (C.c_int, [_p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i32, _i32,
                                        _p, _p, _sz, _p]),
    "ssak_ctc_loss_backward": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i32,
                                         _i32, _p, _p, _i64, _i64, _p, _sz, _p]),
    "ssak_ctc_logits_forward": (C.c_int, [_p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i32, _i32,
                                          _p, _p, _sz, _p]),
    "ssak_ctc_logits_backward": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _i64, _i32,
                                           _i32, _p, _p, _i64, _i64, _p, _sz, _p]),
    "ssak_ctc_loss_reduce": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p, _p]),
    "ssak_ctc_shard_pack": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p, _p]),
    "ssak_ctc_shard_finish": (C.c_int, [_p, _i32, _i64, _p, _p, _p]),
    "ssak_ctc_shard_grad_scale": (C.c_int, [_p, _p, _p, _i64, _p, _p]),
    "ssak_ctc_loss_nll_is_provisional": (C.c_int, [_i64, _i64, _i64]),
    "ssak_ctc_grad_scale": (C.c_int, [_p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "ssak_align_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "ssak_forced_align": (C.c_int, [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _i64, _p, _p, _i32, _i32, _p,
                                    _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ssak_ctc_greedy": (C.c_int, [_p, _i64, _i64, _i64, _i64, _i64, _p, _i32, _p, _p, _p, _p]),
    "ssak_context_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "ssak_context_destroy": (None, [_p]),
    "ssak_ctc_loss_host": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _i64, _p, _p, _i32, _i32, _p, _p, _p]),
    "ssak_forced_align_host": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _i64, _p, _p, _i32, _i32, _p, _p, _p,
                                         _p, _p, _p]),
    "ssak_ctc_greedy_host": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _i32, _p, _p, _p]),
}

_lib = None


class SsakB200Error(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises ImportError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA library first (python ssak_b200/build.py, needs "
                "nvcc with sm_100a support).  ssak_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        L = lib()
        msg = L.ssak_b200_strerror(status).decode()
        extra = ""
        if status == -4:
            extra = f" (cudaError {L.ssak_b200_last_cuda_error()})"
        raise SsakB200Error(f"{what}: {msg}{extra}")


def require_cuda(t, name: str):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"ssak_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    return t


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
