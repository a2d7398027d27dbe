"""ctypes binding of ``libdiffsplit_b200.so`` (the C ABI declared in ``include/diffsplit_b200.h``).

There is no CPU fallback: if the shared library cannot be loaded (or built), importing any compute entry
point raises; if no CUDA device is present, calls fail with the library's own error text.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libdiffsplit_b200.so")

DS_OK = 0
UNET_SR3, UNET_DDPM = 0, 1
PREC_FP32, PREC_BF16, PREC_TF32 = 0, 1, 2
TILE_TRIM, TILE_PAD, TILE_SHIFT = 0, 1, 2
MAX_LEVELS = 8


class UNetDesc(C.Structure):
    _fields_ = [("variant", C.c_int32), ("in_channel", C.c_int32), ("out_channel", C.c_int32),
                ("inner_channel", C.c_int32), ("norm_groups", C.c_int32),
                ("n_mults", C.c_int32), ("channel_mults", C.c_int32 * MAX_LEVELS),
                ("n_attn_res", C.c_int32), ("attn_res", C.c_int32 * MAX_LEVELS),
                ("res_blocks", C.c_int32), ("image_size", C.c_int32), ("with_time_emb", C.c_int32),
                ("tf32_weights", C.c_int32)]


class TensorView(C.Structure):
    _fields_ = [("name", C.c_char_p), ("d_data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class OpProfile(C.Structure):
    _fields_ = [("kind", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("ksize", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("launches", C.c_int32), ("ms", C.c_float), ("flops", C.c_double), ("bytes", C.c_double)]


class SamplerState(C.Structure):      # device-resident; mirrored on the host only to build the init bytes
    _fields_ = [("seed", C.c_uint64), ("offset", C.c_uint64), ("step", C.c_int32), ("done", C.c_uint32)]


class StepArgs(C.Structure):
    _fields_ = [("d_x", C.c_void_p), ("d_net", C.c_void_p), ("d_out", C.c_void_p), ("numel", C.c_int64),
                ("mode", C.c_int32), ("clip", C.c_int32), ("d_coef", C.c_void_p), ("n_steps", C.c_int32),
                ("step", C.c_int32), ("d_state", C.c_void_p), ("d_noise", C.c_void_p), ("seed", C.c_uint64),
                ("offset", C.c_uint64), ("offset_inc", C.c_uint64), ("rng_threads", C.c_int32),
                ("skip_rng_if_zero", C.c_int32), ("d_time_table", C.c_void_p), ("d_time_out", C.c_void_p),
                ("time_len", C.c_int32), ("per_sample_numel", C.c_int64)]


class TileNorm(C.Structure):
    _fields_ = [("mean_target", C.c_double * 2), ("std_target", C.c_double * 2), ("mean_input", C.c_double),
                ("std_input", C.c_double), ("w0", C.c_float), ("w1", C.c_float),
                ("input_from_normalized_target", C.c_int32)]


class PsnrArgs(C.Structure):
    _fields_ = [("d_gt", C.c_void_p), ("d_pred", C.c_void_p),
                ("gt_frame_stride", C.c_int64), ("gt_channel_stride", C.c_int64), ("gt_pixel_stride", C.c_int64),
                ("pred_frame_stride", C.c_int64), ("pred_channel_stride", C.c_int64), ("pred_pixel_stride", C.c_int64),
                ("n_frames", C.c_int32), ("C", C.c_int32), ("npix", C.c_int64),
                ("unnormalize", C.c_int32), ("quantize_u16", C.c_int32),
                ("scale", C.POINTER(C.c_double)), ("offset", C.POINTER(C.c_double)),
                ("d_out", C.c_void_p), ("d_workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


_I3 = C.c_int32 * 3
_SIGS = {
    "ds_last_error": (C.c_char_p, []),
    "ds_version": (C.c_int, []),
    "ds_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "ds_unet_create": (C.c_int, [C.POINTER(UNetDesc), C.POINTER(C.c_void_p)]),
    "ds_unet_destroy": (None, [C.c_void_p]),
    "ds_unet_num_weights": (C.c_int, [C.c_void_p]),
    "ds_unet_weight_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "ds_unet_weight_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "ds_unet_load_weights": (C.c_int, [C.c_void_p, C.POINTER(TensorView), C.c_int, C.c_void_p]),
    "ds_unet_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ds_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ds_unet_forward_profiled": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.POINTER(OpProfile), C.c_int, C.POINTER(C.c_int)]),
    "ds_unet_flops": (C.c_double, [C.c_void_p, C.c_int, C.c_int]),
    "ds_unet_launches": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ds_unet_read_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ds_groupnorm_swish_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ds_groupnorm_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "ds_conv2d_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ds_conv2d_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ds_conv2d_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "ds_conv2d_bf16_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ds_gnconv_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "ds_gnconv_bf16_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ds_conv2d_tf32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                 C.c_void_p]),
    "ds_conv2d_tf32_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ds_gnconv_tf32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "ds_gnconv_tf32_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ds_debug_chain_phases": (C.c_int, [C.POINTER(C.c_longlong), C.c_int, C.POINTER(C.c_int)]),
    "ds_debug_halo_phases": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "ds_debug_stream_phases": (C.c_int, [C.POINTER(C.c_longlong)]),
    "ds_debug_trace_reset": (C.c_int, [C.c_int]),
    "ds_debug_trace_read": (C.c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.c_int]),
    "ds_attention_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ds_chain2_bf16_scratch_bytes": (C.c_size_t, [C.c_int] * 8),
    "ds_chain2_bf16": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_void_p, C.c_void_p]
                       + [C.c_int] * 6 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "ds_attention_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ds_sampler_step": (C.c_int, [C.POINTER(StepArgs), C.c_void_p]),
    "ds_randn_axpy": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p,
                                C.c_uint64, C.c_int, C.c_void_p]),
    "ds_tile_counts": (C.c_int, [_I3, _I3, _I3, C.c_int, _I3, C.POINTER(C.c_int64)]),
    "ds_tile_patch_locations": (C.c_int, [_I3, _I3, _I3, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int32)]),
    "ds_crop_tiles": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _I3, _I3, _I3, C.c_int, C.c_int64, C.c_int64,
                                C.c_void_p, C.c_void_p]),
    "ds_stitch_tiles": (C.c_int, [C.c_void_p, C.c_int, _I3, _I3, _I3, C.c_int, C.c_void_p, C.c_void_p]),
    "ds_tile_regions": (C.c_int, [_I3, _I3, _I3, C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int32)]),
    "ds_pack_tile_regions": (C.c_int, [C.c_void_p, C.c_int, _I3, _I3, _I3, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_void_p]),
    "ds_stitch_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, _I3, _I3, _I3, C.c_int, C.c_void_p, C.c_void_p]),
    "ds_tile_batch": (C.c_int, [C.c_void_p, C.c_int, _I3, _I3, _I3, C.c_int, C.c_int64, C.c_int64,
                                C.POINTER(TileNorm), C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_psnr_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int64]),
    "ds_psnr": (C.c_int, [C.POINTER(PsnrArgs), C.c_void_p]),
    "ds_time_head_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ds_time_head_f32": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib() -> C.CDLL:
    """Load (once) the native library; build it with nvcc if the .so is absent. Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from .build import build
        build()
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(handle, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int):
    if rc != DS_OK:
        msg = lib().ds_last_error()
        raise RuntimeError(f"diffsplit_b200 error {rc}: {msg.decode() if msg else '?'}")


def i3(v):
    return _I3(*[int(x) for x in v])


_dev_info = {}


def device_info(device_index: int):
    """(sm_count, max_threads_per_sm, cc_major, cc_minor) of a CUDA device."""
    import torch
    if device_index not in _dev_info:
        with torch.cuda.device(device_index):
            vals = [C.c_int() for _ in range(4)]
            check(lib().ds_device_info(*[C.byref(v) for v in vals]))
        _dev_info[device_index] = tuple(v.value for v in vals)
    return _dev_info[device_index]


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"diffsplit_b200: {what} is on {t.device}; this path has no CPU fallback - "
                           f"move the model and its inputs to a CUDA device (sm_100a)")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
