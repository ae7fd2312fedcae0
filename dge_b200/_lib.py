"""ctypes binding of libdge_b200.so (include/dge_b200.h).

The product path has NO fallback: if the CUDA library is missing this module
raises, and nothing under dge_b200/ imports oracle/.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DGE_B200_LIB: an alternative build of the SAME library (A/B measurements of kernel variants)
LIB_PATH = os.environ.get("DGE_B200_LIB") or os.path.join(_HERE, "_build", "libdge_b200.so")
ABI_VERSION = 14

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

_f = C.c_float
_i = C.c_int
_p = C.c_void_p

_SIGNATURES = {
    "dge_last_error": (C.c_char_p, []),
    "dge_abi_version": (_i, []),
    "dge_rasterize_forward": (_i, [ALLOC_FN, ALLOC_FN, ALLOC_FN, _p, _i, _i, _i, _p, _i, _i, _p, _p, _p,
                                   _p, _p, _f, _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _i, _p]),
    "dge_rasterize_backward": (_i, [ALLOC_FN, _p, _i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _f, _p, _p,
                                    _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                    _p, _p, _i, _i, _p]),
    "dge_apply_weights": (_i, [ALLOC_FN, ALLOC_FN, ALLOC_FN, _p, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p,
                               _p, _f, _p, _p, _p, _p, _p, _f, _f, _i, _p, _p, _p, _i, _i, _p]),
    "dge_mark_visible": (_i, [_i, _p, _p, _p, _p, _p]),
    "dge_geom_bytes": (C.c_size_t, [_i]),
    "dge_binning_bytes": (C.c_size_t, [_i, _i, _i]),
    "dge_image_bytes": (C.c_size_t, [_i, _i]),
    "dge_backward_scratch_bytes": (C.c_size_t, [_i]),
    "dge_geom_pointers": (None, [_p, _i, C.POINTER(_p)]),
    "dge_binning_pointers": (None, [_p, _i, _i, _i, C.POINTER(_p)]),
    "dge_image_pointers": (None, [_p, _i, _i, C.POINTER(_p)]),
    "dge_debug_sorted_keys": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "dge_launch_count": (C.c_ulonglong, []),
    "dge_clock_probe": (_i, [_p, _p]),
    "dge_profile_enable": (None, [C.c_uint]),
    "dge_profile_read": (_i, [C.POINTER(C.c_float), C.POINTER(_i)]),
    "dge_fit_forward": (_i, [ALLOC_FN, ALLOC_FN, ALLOC_FN, _p, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _f, _p, _p,
                             _f, _f, _p, _p, _p, _p, _p, _p]),
    "dge_fit_backward_blend": (_i, [_i, _i, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "dge_fit_views_forward": (_i, [ALLOC_FN, ALLOC_FN, ALLOC_FN, _p, _i, _i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _f, _p,
                                   _p, _p, _p, _p, _p, C.c_size_t, _p, C.c_size_t, _p, _p, _p, _i, _p]),
    "dge_fit_views_front": (_i, [ALLOC_FN, ALLOC_FN, ALLOC_FN, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _f, _p, _p, _p,
                                 _p, C.c_size_t, _p, _i, _p]),
    "dge_fit_views_colour": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "dge_fit_views_blend": (_i, [_i, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, C.c_size_t, _p, _p, _p]),
    "dge_fit_views_backward_blend": (_i, [_i, _i, _i, _p, _i, _i, _i, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "dge_fit_binning_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "dge_fit_views_apply_weights": (_i, [ALLOC_FN, ALLOC_FN, ALLOC_FN, _p, _i, _i, _i, _i, _p, _p, _p, _f, _p, _p, _p, _i,
                                         _p, _p, _p, _i, _p]),
    "dge_fit_backward_geom": (_i, [_i, _i, _i, _i, _p, _i, _i, _f, _p, C.c_size_t, _p, C.c_size_t, _p, _p, _p, _p, _p, _p,
                                   _p, _p, _p, _p, _i, _p]),
    "dge_fit_activate": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "dge_fit_backward_geom_raw": (_i, [_i, _i, _i, _p, _i, _i, _f, _p, C.c_size_t, _p, C.c_size_t, _p, _p, _p, _p, _p,
                                       _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "dge_densify_select": (_i, [_i, _p, _p, _p, _p, _f, _f, _f, _f, _i, _p, _p, _p]),
    "dge_densify_gather": (_i, [_i, _i, _p, _p, _i, _i, _i, _i, C.POINTER(_p), C.POINTER(_p), C.POINTER(_i), _i, _i, _i,
                                _p, _p, _p, _p]),
    "dge_l1_loss_grad": (_i, [_p, _p, C.c_size_t, _f, _p, _p, _p]),
    "dge_fit_update_stats": (_i, [_i, _p, _p, _p, _p, _p, _p]),
    "dge_fused_adam": (_i, [_p, _p, _p, _p, C.c_size_t, _f, _f, _f, _f, _i, _p, _i, _p]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    """Load libdge_b200.so; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"dge_b200: CUDA library not built ({LIB_PATH} missing). Run `python -m dge_b200.build` "
            "or __graft_entry__.build(); there is no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dge_abi_version() != ABI_VERSION:
        raise RuntimeError("dge_b200: libdge_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def last_error():
    return load().dge_last_error().decode()


def check(rc, what):
    if rc < 0:
        raise RuntimeError(f"{what} failed: {last_error()}")
    return rc


def ptr(t):
    """Device pointer of a tensor; None (NULL) for None / empty tensors, as the reference's
    kernels treat empty tensors (DGR/cuda_rasterizer/forward.cu:205,241)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


class Arena:
    """Allocator callbacks backed by torch's caching allocator (the reference's
    resizeFunctional, DGR/rasterize_points.cu:27-33). Each callback remembers the uint8 tensor
    it handed out so the caller can keep it alive / save it for backward."""

    def __init__(self, device, n=3):
        self.device = device
        self.bufs = [None] * n
        self.cbs = [ALLOC_FN(self._make(i)) for i in range(n)]

    def _make(self, i):
        def alloc(_ctx, nbytes):
            t = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self.bufs[i] = t
            return t.data_ptr()
        return alloc


_ARENAS = {}


def arena(device, n=3):
    """A per-device Arena whose ctypes trampolines are created ONCE (building a CFUNCTYPE costs tens of
    microseconds, and the per-view API used to build three per call). The caller takes the tensors of this call
    with take(): the arena then forgets them, so their lifetime is the caller's (ctx.save_for_backward)."""
    key = (str(device), n, threading.get_ident())  # per thread: autograd runs backwards on its own
    a = _ARENAS.get(key)
    if a is None:
        a = _ARENAS[key] = Arena(device, n)
    a.bufs = [None] * n
    return a


def take(a):
    bufs, a.bufs = a.bufs, [None] * len(a.bufs)
    return bufs


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
