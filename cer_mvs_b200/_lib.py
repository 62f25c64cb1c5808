"""ctypes binding of libcer_mvs_b200.so (C ABI: include/cer_mvs_b200.h).  Fails loudly when the
library is missing -- there is deliberately no fallback path."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcer_mvs_b200.so")

c_void_p, c_int, c_float, c_double, c_size_t, c_ll = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_longlong


class PlanConfig(C.Structure):
    _fields_ = [("h", c_int), ("w", c_int), ("max_views", c_int), ("n_stages", c_int), ("D", c_int * 4),
                ("incre", c_double * 4), ("iters", c_int * 4), ("feats_f16", c_int), ("use_graph", c_int)]


# name -> (restype, argtypes); every symbol include/cer_mvs_b200.h declares
SIGNATURES = {
    "cer_abi_version": (c_int, []),
    "cer_last_error": (C.c_char_p, []),
    "cer_device_check": (c_int, []),
    "cer_corr_forward_f32": (c_int, [c_void_p] * 4 + [c_int] * 8 + [c_void_p]),
    "cer_nchw_to_nhwc": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "cer_nchw_to_nhwc_pad": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                     c_void_p]),
    "cer_nhwc_to_nchw": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cer_projection_matrices": (c_int, [c_void_p] * 4 + [c_int, c_void_p, c_void_p]),
    "cer_build_volume": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                 c_float, c_float, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_void_p]),
    "cer_build_volume_part": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                      c_float, c_float, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "cer_set_build_variant": (c_int, [c_int]),
    "cer_debug_set_build_profile": (c_int, [c_void_p]),
    "cer_pool_pairs": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "cer_lookup": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_void_p, c_int,
                           c_int, c_void_p]),
    "cer_lookup_strided": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_ll, c_int, c_float, c_int, c_int, c_void_p,
                                   c_int, c_int, c_void_p]),
    "cer_normalize_images": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "cer_resize_bilinear_ac": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cer_disp_to_depth": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cer_multires_merge": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]),
    "cer_geo_mats_bytes": (c_size_t, [c_int]),
    "cer_geometric_filter": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_double, c_double] + [c_void_p] * 10),
    "cer_encoder_blob_bytes": (c_size_t, [c_int]),
    "cer_encoder_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cer_pack_encoder_weights": (c_int, [c_void_p, c_int, c_void_p]),
    "cer_encoder_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_float, c_void_p]),
    "cer_set_conv_variant": (c_int, [c_int]),
    "cer_lookup_encode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_void_p, c_void_p]),
    "cer_set_lookup_variant": (c_int, [c_int]),
    "cer_debug_set_conv_profile": (c_int, [c_void_p]),
    "cer_gru_step": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p]),
    "cer_update_blob_bytes": (c_size_t, []),
    "cer_pack_update_weights": (c_int, [C.POINTER(c_void_p), c_void_p]),
    "cer_update_workspace_bytes": (c_size_t, [c_int, c_int]),
    "cer_update_step": (c_int, [c_void_p] * 6 + [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "cer_plan_create": (c_int, [C.POINTER(PlanConfig), C.POINTER(c_void_p)]),
    "cer_plan_destroy": (None, [c_void_p]),
    "cer_plan_workspace_bytes": (c_size_t, [c_void_p]),
    "cer_plan_set_weights": (c_int, [c_void_p, c_void_p]),
    "cer_plan_run_device": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                    c_float, c_void_p, c_void_p]),
    "cer_plan_run_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                  c_float, c_void_p, c_void_p]),
    "cer_plan_submit_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                     c_float, c_void_p, c_void_p]),
    "cer_plan_wait_host": (c_int, [c_void_p]),
    "cer_plan_prepare": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                 c_int, c_int, c_void_p]),
    "cer_plan_build_stage": (c_int, [c_void_p, c_int, c_void_p]),
    "cer_plan_build_stage_units": (c_int, [c_void_p, c_int, c_ll, c_ll, c_void_p]),
    "cer_plan_feature_buffer": (c_void_p, [c_void_p, c_int]),
    "cer_plan_net_buffer": (c_void_p, [c_void_p]),
    "cer_plan_inp_buffer": (c_void_p, [c_void_p]),
    "cer_plan_prepare_inplace": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "cer_plan_partial_volume": (c_void_p, [c_void_p, c_int, C.POINTER(c_size_t)]),
    "cer_plan_iterate_stage": (c_int, [c_void_p, c_int, c_void_p]),
    "cer_plan_finish": (c_int, [c_void_p, c_float, c_void_p, c_void_p]),
    "cer_plan_set_kernel_timing": (c_int, [c_void_p, c_int]),
    "cer_plan_kernel_times": (c_int, [c_void_p, C.POINTER(c_double), C.POINTER(c_ll), c_int, c_void_p]),
    "cer_plan_last_launch_count": (c_ll, [c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises RuntimeError (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cer_mvs_b200.build` "
                "(cer_mvs_b200 has no CPU / PyTorch fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().cer_last_error().decode(errors="replace")
        raise RuntimeError(f"cer_mvs_b200 {what} failed (code {rc}): {msg}")


def stream_ptr():
    """The current torch CUDA stream as a cudaStream_t value."""
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    """Reference error convention (alt_cuda_corr/correlation.cpp:19-21)."""
    for i, t in enumerate(tensors):
        if not t.is_cuda:
            raise RuntimeError(f"argument {i} must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"argument {i} must be contiguous")
