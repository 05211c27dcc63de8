"""ctypes binding of libb200corr.so (the C ABI declared in include/b200corr.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200CORR_LIB") or os.path.join(_HERE, "libb200corr.so")   # env override: tuning variants

_lib = None

c_int, c_void_p, c_size_t, c_float = ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float

# name -> (restype, argtypes); every symbol include/b200corr.h declares
PROTOTYPES = {
    "b200corr_version": (c_int, []),
    "b200corr_last_error": (ctypes.c_char_p, []),
    "b200corr_sampler_out_size": (c_int, [c_int] * 5),
    "b200corr_sampler_forward_workspace_bytes": (c_size_t, [c_int] * 17),
    "b200corr_sampler_backward_workspace_bytes": (c_size_t, [c_int] * 17),
    "b200corr_sampler_forward": (c_int, [c_void_p] * 4 + [c_size_t] + [c_int] * 17 + [c_void_p]),
    "b200corr_sampler_backward": (c_int, [c_void_p] * 6 + [c_size_t] + [c_int] * 17 + [c_void_p]),
    "b200corr_sampler_uses_fast_path": (c_int, [c_int] * 18),
    "b200corr_sampler_backward_plan": (c_int, [c_int] * 17 + [c_void_p, c_size_t]),
    "b200corr_merge_supported": (c_int, [c_int] * 6),
    "b200corr_merge_forward": (c_int, [c_void_p] * 3 + [c_int] * 8 + [c_float, c_void_p]),
    "b200corr_merge_backward_scratch_bytes": (c_size_t, [c_int] * 4),
    "b200corr_merge_backward": (c_int, [c_void_p] * 8 + [c_size_t] + [c_int] * 8 + [c_float, c_void_p]),
    "b200corr_channelnorm_forward": (c_int, [c_void_p] * 2 + [c_int] * 5 + [c_void_p]),
    "b200corr_channelnorm_backward": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "b200corr_resample2d_forward": (c_int, [c_void_p] * 3 + [c_int] * 6 + [c_void_p]),
    "b200corr_resample2d_backward": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "b200corr_warp_forward": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
    "b200corr_warp_backward": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "b200corr_allpairs_workspace_bytes": (c_size_t, [c_int] * 5),
    "b200corr_allpairs_pyramid": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_int, c_int, c_int,
                                          c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "b200corr_allpairs_rect_workspace_bytes": (c_size_t, [c_int] * 7),
    "b200corr_allpairs_pyramid_rect": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int,
                                               c_int, c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "b200corr_allpairs_blocked_levels": (c_int, [c_int] * 4),
    "b200corr_blocked_level_dims": (None, [c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "b200corr_allpairs_pyramid_layout": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int,
                                                 c_int, c_int, c_int, c_float, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "b200corr_lookup_forward_layout": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                               c_int, c_int, c_int, c_int, c_void_p]),
    "b200corr_allpairs_pyramid_storage": (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int,
                                                  c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_size_t,
                                                  c_void_p]),
    "b200corr_lookup_forward_storage": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                                c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200corr_lookup_forward": (c_int, [ctypes.POINTER(c_void_p), c_int, c_void_p, c_void_p, c_int, c_int,
                                        c_int, c_int, c_int, c_void_p]),
    "b200corr_lookup_forward_from": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                             c_int, c_int, c_int, c_void_p]),
    "b200corr_lookup_convc1_weight_bytes": (c_size_t, [c_int] * 3),
    "b200corr_lookup_convc1_prepare": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200corr_lookup_convc1_forward": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p, c_void_p,
                                               c_void_p] + [c_int] * 7 + [c_void_p]),
    "b200corr_conv1x1_forward": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "b200corr_lookup_backward": (c_int, [ctypes.POINTER(c_void_p), c_int, c_void_p, c_void_p, c_int, c_int,
                                         c_int, c_int, c_int, c_void_p]),
    "b200corr_pyramid_backward": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_void_p]),
    "b200corr_volume_backward_workspace_bytes": (c_size_t, [c_int] * 5),
    "b200corr_volume_backward": (c_int, [ctypes.POINTER(c_void_p), c_int] + [c_void_p] * 4 + [c_int] * 4
                                 + [c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "b200corr_altcorr_forward": (c_int, [c_void_p] * 4 + [c_int] * 8 + [c_void_p]),
    "b200corr_altcorr_backward": (c_int, [c_void_p] * 7 + [c_int] * 8 + [c_void_p]),
    "b200corr_patch_compose_forward": (c_int, [c_void_p] * 7 + [c_int] * 4 + [ctypes.c_longlong] * 4 + [c_void_p]),
    "b200corr_patch_compose_backward_scratch_bytes": (c_size_t, [c_int] * 2),
    "b200corr_patch_compose_backward": (c_int, [c_void_p] * 9 + [c_size_t] + [c_int] * 4 + [ctypes.c_longlong] * 4
                                        + [c_void_p]),
    "b200corr_measure_fp32_peak": (c_int, [c_int, ctypes.POINTER(c_float), c_void_p]),
    "b200corr_launch_count": (ctypes.c_uint64, []),
}


def lib():
    """Load the shared library once.  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m understanding_flow_robustness_b200.build` "
                "(there is no fallback implementation)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().b200corr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({code}): {msg}")


def current_stream(device):
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def ptr_array(tensors):
    """Host array of device pointers (the `h_levels` arguments of the C ABI)."""
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
