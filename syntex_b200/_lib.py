"""ctypes binding of include/syntex_b200.h. There is no fallback: if the library is missing, importing a
compute entry point raises, and every non-zero return code becomes a RuntimeError carrying stx_last_error()."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsyntex_b200.so")

_lib = None


class SyntexLibraryError(RuntimeError):
    pass


def _declare(lib):
    vp, fp, ip = c_void_p, c_void_p, c_void_p   # device pointers travel as integers
    sig = {
        "stx_last_error": (c_char_p, []),
        "stx_abi_version": (c_int, []),
        "stx_launch_count": (ctypes.c_longlong, []),
        "stx_crc32c": (ctypes.c_uint, [c_void_p, c_size_t, ctypes.c_uint]),
        "stx_debug_set": (c_int, [c_int, c_int]),
        "stx_debug_pair_probe": (c_int, [vp, vp, fp, vp]),
        "stx_debug_patch_probe": (c_int, [vp, c_int, vp, c_int, c_int, c_int, fp, vp]),
        "stx_vgg_create": (c_int, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_int,
                                   POINTER(c_float), vp]),
        "stx_vgg_destroy": (None, [vp]),
        "stx_vgg_weights": (c_int, [vp, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p)]),
        "stx_plan_create": (c_int, [POINTER(c_void_p), vp, c_int, c_int, c_int, c_int, c_int, POINTER(c_int),
                                    POINTER(c_float), c_int]),
        "stx_plan_destroy": (None, [vp]),
        "stx_plan_device_bytes": (c_size_t, [vp]),
        "stx_plan_forward": (c_int, [vp, fp, c_int, c_int, vp]),
        "stx_plan_layer_shape": (c_int, [vp, c_int, POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
        "stx_plan_copy_feature_f32": (c_int, [vp, c_int, fp, vp]),
        "stx_plan_copy_gram": (c_int, [vp, c_int, fp, vp]),
        "stx_plan_update_target_from_forward": (c_int, [vp, vp]),
        "stx_plan_set_target_gram": (c_int, [vp, c_int, fp, c_int, vp]),
        "stx_plan_copy_target_gram": (c_int, [vp, c_int, fp, vp]),
        "stx_plan_clear_target": (c_int, [vp]),
        "stx_conv1x1_per_image_bf16": (c_int, [vp, c_int, c_int, c_int, c_int, vp, c_int, vp, vp]),
        "stx_lbfgs_phase_cycles": (c_int, [vp, POINTER(ctypes.c_longlong)]),
        "stx_adam_tf": (c_int, [fp, fp, fp, fp, c_size_t, fp, fp, c_float, c_float, c_float, vp]),
        "stx_plan_generation": (ctypes.c_longlong, [vp]),
        "stx_plan_step": (c_int, [vp, fp, c_int, fp, fp, fp, vp]),
        "stx_plan_step_timed": (c_int, [vp, fp, c_int, fp, fp, POINTER(c_float), POINTER(c_int), vp]),
        "stx_plan_step_host": (c_int, [vp, fp, c_int, fp, fp, vp]),
        "stx_plan_layer_grad": (c_int, [vp, c_int, fp, vp]),
        "stx_plan_layer_grad_bf16": (c_int, [vp, c_int, vp, vp]),
        "stx_plan_copy_layer_total_grad": (c_int, [vp, c_int, fp, vp]),
        "stx_plan_feature_bf16": (c_int, [vp, c_int, POINTER(c_void_p), POINTER(c_void_p)]),
        "stx_plan_backward_custom": (c_int, [vp, fp, c_int, POINTER(c_void_p), POINTER(c_void_p), fp, vp]),
        "stx_lbfgs_create": (c_int, [POINTER(c_void_p), vp, c_int, c_int, c_float, c_float]),
        "stx_lbfgs_destroy": (None, [vp]),
        "stx_lbfgs_reset": (c_int, [vp, fp, c_int, vp]),
        "stx_lbfgs_run": (c_int, [vp, c_int, c_int, vp]),
        "stx_lbfgs_result": (c_int, [vp, fp, fp, ip, ip, vp]),
        "stx_lbfgs_minimize_host": (c_int, [vp, fp, c_int, c_int, fp, fp, ip, ip, vp]),
        "stx_lbfgs_minimize_host_f64": (c_int, [vp, fp, c_int, c_int, vp, fp, ip, ip, vp]),
        "stx_pack_conv_weights": (c_int, [fp, c_int, c_int, vp, vp, vp]),
        "stx_pack_conv_weights_split": (c_int, [fp, c_int, c_int, vp, vp, vp, vp]),
        "stx_conv_bf16x3": (c_int, [vp, vp, c_int, c_int, c_int, c_int, vp, vp, c_int, fp, c_int, vp, vp, c_int, vp]),
        "stx_conv_bf16": (c_int, [vp, c_int, c_int, c_int, c_int, vp, c_int, c_int, fp, vp, vp, c_int, vp, c_int,
                                  vp]),
        "stx_conv1_1_fwd": (c_int, [fp, c_int, c_int, c_int, c_int, POINTER(c_float), fp, fp, vp, vp, vp]),
        "stx_conv1_1_bwd": (c_int, [vp, fp, c_int, c_int, c_int, c_int, fp, fp, vp]),
        "stx_avgpool2_fwd": (c_int, [vp, vp, c_int, c_int, c_int, c_int, vp, vp, vp]),
        "stx_avgpool2_bwd": (c_int, [vp, vp, c_int, c_int, c_int, c_int, vp, vp]),
        "stx_gram_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
        "stx_gram_bf16": (c_int, [vp, c_int, c_int, c_int, c_float, fp, vp, vp]),
        "stx_gram_cross_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
        "stx_gram_cross_bf16": (c_int, [vp, vp, c_int, c_int, c_int, c_float, fp, vp, vp]),
        "stx_conv_bf16_pad": (c_int, [vp, c_int, c_int, c_int, c_int, vp, c_int, c_int, fp, vp, c_int, vp, c_int, vp]),
        "stx_conv_wgrad_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
        "stx_conv_wgrad_bf16": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, fp, vp, vp]),
        "stx_conv_wgrad_oihw_bf16": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, fp, vp, vp]),
        "stx_pack_conv_weights_oihw": (c_int, [fp, c_int, c_int, c_int, vp, vp, vp]),
        "stx_pad_replicate_bf16": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, vp, vp]),
        "stx_pad_replicate_bwd_bf16": (c_int, [vp, vp, c_int, c_int, c_int, c_int, vp, vp]),
        "stx_inorm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
        "stx_inorm_stats_bf16": (c_int, [vp, c_int, c_int, c_int, c_float, fp, fp, vp, vp]),
        "stx_pad2d_bf16": (c_int, [vp, c_int, c_int, c_int, c_int, c_int, c_int, c_float, vp, vp]),
        "stx_pad2d_bwd_bf16": (c_int, [vp, vp, c_int, c_int, c_int, c_int, c_int, c_float, vp, vp]),
        "stx_inorm_apply_bf16": (c_int, [vp, c_int, c_int, c_int, fp, fp, fp, fp, c_int, c_float, vp, vp, vp]),
        "stx_inorm_bwd_bf16": (c_int, [vp, vp, c_int, c_int, c_int, fp, fp, fp, fp, c_int, c_float, fp, fp, vp, vp,
                                       vp]),
        "stx_f32_to_bf16": (c_int, [fp, c_size_t, vp, vp]),
        "stx_bf16_to_f32": (c_int, [vp, c_size_t, fp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED = None


def lib():
    """Load (once) and return the ctypes library. Raises if it has not been built."""
    global _lib, EXPORTED
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SyntexLibraryError(
                "syntex_b200 CUDA library not built: %s is missing (run `python -m syntex_b200.build`); "
                "there is no CPU fallback" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        EXPORTED = _declare(l)
        _lib = l
        # bring-up switches: STX_DEBUG="key=value,key=value" (see stx_debug_set)
        for kv in filter(None, os.environ.get("STX_DEBUG", "").split(",")):
            k, v = kv.split("=")
            l.stx_debug_set(int(k), int(v))
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().stx_last_error()
        raise SyntexLibraryError("syntex_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def stream_ptr():
    """Raw cudaStream_t of torch's current stream (torch is used for device memory and streams only)."""
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SyntexLibraryError("syntex_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
