"""ctypes binding of libvq_b200.so (the C ABI in include/vq_b200.h).

There is no CPU path and no PyTorch fallback: if the library is missing the import of any
compute entry point raises, loudly, with the command that builds it.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("VQ_B200_LIB") or os.path.join(_PKG_DIR, "lib", "libvq_b200.so")   # env: diagnostic builds
BUILD_SCRIPT = os.path.join(_PKG_DIR, "csrc", "build.py")

# constants mirrored from include/vq_b200.h
ABI_VERSION = 6
FORM_VIT, FORM_VQGAN, FORM_VQGAN_L2 = 0, 1, 2
LAYOUT_TOKEN_MAJOR, LAYOUT_NCHW = 0, 1
FLAG_INDICES_ONLY, FLAG_EXACT_SCAN, FLAG_KEEP_STATS, FLAG_IDX32, FLAG_IDX16 = 1, 2, 4, 8, 16
STAT_NEAR_TIE_ROWS, STAT_AMBIGUOUS_ROWS, STAT_FALLBACK_ROWS, STAT_LOSS_FIXED, STAT_BAD_INDEX, STAT_NONFINITE = range(6)
STAT_PEER_TIMEOUT = 6
PEER_MAX_RANKS, IPC_HANDLE_BYTES = 16, 64
(PROFILE_PREP_CODEBOOK, PROFILE_PREP_TOKENS, PROFILE_SEARCH, PROFILE_EXACT_FINISH, PROFILE_TAIL, PROFILE_BACKWARD_TOKENS,
 PROFILE_CODEBOOK_GRAD) = range(7)
STATS_LEN = 8
SEG_SHIFT = 30

# every symbol include/vq_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "vq_abi_version": (c_int, []),
    "vq_last_error": (c_char_p, []),
    "vq_uses_tensor_cores": (c_int, [c_int64, c_int, c_int]),
    "vq_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "vq_codebook_bytes": (c_int, [c_int, c_int, POINTER(c_size_t)]),
    "vq_codebook_prepare": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vq_codebook_prepare_raw": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "vq_workspace_bytes": (c_int, [c_int64, c_int, c_int, c_int, POINTER(c_size_t)]),
    "vq_forward": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_int64,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_size_t, c_void_p]),
    "vq_prequant_supported": (c_int, [c_int, c_int]),
    "vq_forward_projected": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                     c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vq_project_codebook": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "vq_gather_projected": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p]),
    "vq_loss_finalize": (c_int, [c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p]),
    "vq_backward_workspace_bytes": (c_int, [c_int64, c_int, c_int, POINTER(c_size_t)]),
    "vq_backward_tokens": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "vq_backward_codebook": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int64, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "vq_backward": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                            c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vq_exchange_bytes": (c_int, [c_int, c_int, POINTER(c_size_t)]),
    "vq_exchange_slot": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p)]),
    "vq_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p), c_void_p]),
    "vq_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "vq_peer_close": (c_int, [c_void_p]),
    "vq_peer_free": (c_int, [c_void_p]),
    "vq_peer_configure": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "vq_peer_resync": (c_int, [c_void_p, c_void_p]),
    "vq_backward_codebook_sharded": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, ctypes.c_uint32, c_void_p, c_int, c_int,
                                             c_int, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p]),
    "vq_backward_sharded": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, ctypes.c_uint32, c_void_p, c_int64, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_int64, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "vq_gather": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                          c_void_p, c_void_p]),
    "vq_gather_tokens": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p]),
    "vq_tokens_convert": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p]),
    "vq_token_embed_tokens": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vq_token_embed_causal_tokens": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p]),
    "vq_token_embed": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vq_token_embed_causal": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "vq_embedding_backward_bytes": (c_int, [c_int64, c_int, POINTER(c_size_t)]),
    "vq_embedding_backward": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "vq_kernel_launches": (c_int64, []),
    "vq_profile_begin": (c_int, [c_int, ctypes.c_uint32]),
    "vq_profile_end": (c_int, [POINTER(ctypes.c_double), POINTER(c_int64), POINTER(c_int64)]),
    "vq_profile_slot": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(c_int64)]),
    "vq_host_step_arena_bytes": (c_int, [c_int64, c_int, c_int, POINTER(c_size_t)]),
    "vq_host_step": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}


class VQLibraryError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """dlopen libvq_b200.so and type every entry point.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VQLibraryError(
            f"libvq_b200.so not found at {LIB_PATH}. vq_b200 has no CPU or PyTorch fallback; build the CUDA "
            f"library first:  python {BUILD_SCRIPT}   (or __graft_entry__.build())")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    if lib.vq_abi_version() != ABI_VERSION:
        raise VQLibraryError(f"libvq_b200.so ABI {lib.vq_abi_version()} != expected {ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise VQLibraryError(f"libvq_b200 error {status}: {load().vq_last_error().decode(errors='replace')}")


_size_cache = {}


def size_query(fn_name: str, *args) -> int:
    """Byte-size queries of the ABI (pure host arithmetic), memoised: they sit on the per-step path of the modules."""
    key = (fn_name, args)
    hit = _size_cache.get(key)
    if hit is not None:
        return hit
    out = c_size_t(0)
    check(getattr(load(), fn_name)(*args, ctypes.byref(out)))
    _size_cache[key] = int(out.value)
    return _size_cache[key]
