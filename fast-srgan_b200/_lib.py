"""ctypes binding of libfsr_b200.so (include/fsr_b200.h).  No CPU fallback: a missing library or a
non-zero return code raises - the product path must fail loudly (never route through oracle/)."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfsr_b200.so")
FSR_MAX_LAYERS = 32
FSR_F16, FSR_BF16 = 0, 1
EPI_RAW_STATS, EPI_BIAS_ACT, EPI_PS_PRELU, EPI_HEAD_TANH, EPI_F32 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_PRELU = 0, 1, 2, 3
K_NONE, K_NECK, K_CONV_RES, K_IN_APPLY, K_CONV_UP, K_CONV_HEAD, K_CONV_BIAS_ACT, K_CONV_GEN, K_CONV_WGRAD = -1, 0, 1, 2, 3, 4, 5, 6, 7

(OPT_HALO1, OPT_WS, OPT_FUSE_IN, OPT_FUSE_RES, OPT_UP_2CTA, OPT_GEN_WS, OPT_GEN_2CTA, OPT_SMALL_MMA, OPT_IN_BWD_FUSED,
 OPT_OVERLAP_STREAMS) = range(10)

_vp, _fp, _i, _f, _sz = C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_size_t


class FsrGeneratorParams(C.Structure):
    _fields_ = [
        ("n_filters", _i), ("n_layers", _i), ("dtype", _i), ("reserved", _i),
        ("neck_w", _vp), ("neck_b", _vp), ("neck_alpha", _vp),
        ("stem_w1", _vp * FSR_MAX_LAYERS), ("stem_alpha", _vp * FSR_MAX_LAYERS), ("stem_w2", _vp * FSR_MAX_LAYERS),
        ("bott_w", _vp),
        ("up_w", _vp * 2), ("up_b", _vp * 2), ("up_alpha", _vp * 2),
        ("head_w", _vp), ("head_b", _vp),
    ]


class FsrPackTask(C.Structure):
    _fields_ = [("w", _vp), ("out", _vp), ("bias", _vp), ("bias_out", _vp), ("row_scale", _vp),
                ("cout", _i), ("cin", _i), ("pad", _i), ("flags", _i)]


PACK_T, PACK_PS, PACK_FLIP = 1, 2, 4

_SIGS = {
    "fsr_abi_version": (_i, []),
    "fsr_error_string": (C.c_char_p, [_i]),
    "fsr_pack_conv3x3_weight": (_i, [_fp, _fp, _vp, _fp, _i, _i, _i, _i, _i, _vp]),
    "fsr_pack_multi": (_i, [C.POINTER(FsrPackTask), _i, _i, _vp]),
    "fsr_conv3x3_c64": (_i, [_vp, _vp, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "fsr_conv3x3_gen": (_i, [_vp, _vp, _vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "fsr_pack_conv3x3_weight_t": (_i, [_fp, _vp, _i, _i, _i, _i, _i, _fp, _i, _vp]),
    "fsr_wgrad_workspace_bytes": (_sz, []),
    "fsr_conv3x3_wgrad": (_i, [_vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _i, _vp]),
    "fsr_conv3x3_wgrad_grouped": (_i, [_vp, _vp, _vp, _i, C.c_longlong, C.c_longlong, _i, _i, _i, _i, _i, _vp, _sz, _i, _vp]),
    "fsr_parity_layout": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "fsr_maxpool2": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fsr_maxpool2_relu_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fsr_relu_bwd": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "fsr_add": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "fsr_conv1x1_to1_fwd": (_i, [_vp, _fp, _fp, _fp, _i, _i, _i, _vp]),
    "fsr_conv1x1_to1_bwd": (_i, [_vp, _fp, _fp, _vp, _fp, _fp, _i, _i, _i, _vp]),
    "fsr_bce_logits": (_i, [_fp, _fp, _f, _f, _i, _fp, _fp, _f, _vp]),
    "fsr_smooth_l1": (_i, [_vp, _vp, _sz, _fp, _vp, _f, _i, _vp]),
    "fsr_instnorm_bwd": (_i, [_vp, _fp, _vp, _fp, _vp, _fp, _fp, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "fsr_instnorm_bwd_parity": (_i, [_vp, _fp, _vp, _vp, _fp, _fp, _i, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "fsr_act_bwd": (_i, [_vp, _vp, _vp, _sz, _fp, _f, _i, _fp, _i, _vp]),
    "fsr_ps_prelu_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _fp, _fp, _i, _vp]),
    "fsr_tanh_bwd": (_i, [_fp, _fp, _fp, _sz, _vp]),
    "fsr_wgrad_c3": (_i, [_fp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fsr_bias_grad": (_i, [_vp, _fp, _sz, _i, _i, _i, _vp]),
    "fsr_bias_grad_nchw": (_i, [_fp, _fp, _i, _i, _sz, _vp]),
    "fsr_adamw": (_i, [_fp, _fp, _fp, _fp, _sz, _f, _f, _f, _f, _f, _i, _f, _vp]),
    "fsr_adamw_dev": (_i, [_fp, _fp, _fp, _fp, _sz, _f, _f, _f, _f, _f, _vp, _f, _vp]),
    "fsr_conv3x3_head": (_i, [_vp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _vp]),
    "fsr_neck_conv3x3": (_i, [_vp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _i, _vp]),
    "fsr_instnorm_apply": (_i, [_vp, _fp, _vp, _vp, _fp, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "fsr_instnorm_apply_parity": (_i, [_vp, _fp, _vp, _fp, _i, _i, _i, _i, _i, _f, _f, _i, _vp]),
    "fsr_pixel_shuffle2": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "fsr_nchw_f32_to_nhwc": (_i, [_fp, _vp, _i, _i, _i, _i, _vp]),
    "fsr_nhwc_to_nchw_f32": (_i, [_vp, _fp, _i, _i, _i, _i, _vp]),
    "fsr_profile_enable": (_i, [_i]),
    "fsr_profile_read": (_i, [_vp, _i]),
    "fsr_profile_enable_mask": (_i, [C.c_uint]),
    "fsr_profile_read_ids": (_i, [_vp, _vp, _i]),
    "fsr_profile_read_ex": (_i, [_vp, _vp, _vp, _i]),
    "fsr_launch_count": (C.c_ulonglong, []),
    "fsr_set_halo_mode": (_i, [_i]),
    "fsr_set_ws_mode": (_i, [_i]),
    "fsr_set_small_mma": (_i, [_i]),
    "fsr_set_gen_ws": (_i, [_i]),
    "fsr_set_gen_2cta": (_i, [_i]),
    "fsr_conv3x3_gen_flat": (_i, [_vp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "fsr_maxpool2_padded": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fsr_maxpool2_relu_bwd_padded": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fsr_set_fuse_in": (_i, [_i]),
    "fsr_conv3x3_c64_in": (_i, [_vp, _vp, _fp, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fsr_conv3x3_c64_res_in": (_i, [_vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fsr_set_fuse_res": (_i, [_i]),
    "fsr_set_up_2cta": (_i, [_i]),
    "fsr_psnr_ssim": (_i, [_fp, _fp, _i, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "fsr_crop_resize_aa": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _fp, _fp, _vp]),
    "fsr_set_overlap_streams": (_i, [_i]),
    "fsr_ctx_create": (_i, [C.POINTER(_vp)]),
    "fsr_ctx_destroy": (_i, [_vp]),
    "fsr_ctx_set": (_i, [_vp, _i, _i]),
    "fsr_ctx_bind": (_i, [_vp]),
    "fsr_split_f32": (_i, [_fp, _vp, _vp, _sz, _vp]),
    "fsr_neck_conv3x3_f32": (_i, [_fp, _fp, _fp, _fp, _fp, _i, _i, _i, _vp]),
    "fsr_in_stats_f32": (_i, [_fp, _vp, _i, _i, _vp]),
    "fsr_in_stats_fold_pair": (_i, [_vp, _i, _vp]),
    "fsr_set_pair_rows": (_i, [_i]),
    "fsr_set_pdl": (_i, [_i]),
    "fsr_conv3x3_c64_head_pair": (_i, [_vp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _vp]),
    "fsr_neck_conv3x3_c32": (_i, [_vp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    "fsr_in_apply_f32": (_i, [_fp, _vp, _fp, _fp, _vp, _vp, _fp, _i, _i, _i, _f, _vp]),
    "fsr_ps_prelu_f32": (_i, [_fp, _fp, _fp, _fp, _vp, _vp, _i, _i, _i, _vp]),
    "fsr_tanh_f32": (_i, [_fp, _vp, _i, _i, _vp]),
    "fsr_nccl_available": (_i, []),
    "fsr_nccl_version": (_i, []),
    "fsr_nccl_unique_id": (_i, [_vp]),
    "fsr_nccl_init": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "fsr_nccl_allreduce": (_i, [_vp, _fp, _sz, _vp]),
    "fsr_nccl_broadcast": (_i, [_vp, _fp, _sz, _i, _vp]),
    "fsr_nccl_destroy": (_i, [_vp]),
    "fsr_generator_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "fsr_generator_forward": (_i, [C.POINTER(FsrGeneratorParams), _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and set prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing - build it with `python fast-srgan_b200/build.py` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = "fsr call"):
    if rc != 0:
        msg = load().fsr_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float16:
        return FSR_F16
    if dt == torch.bfloat16:
        return FSR_BF16
    raise ValueError(f"compute dtype must be float16 or bfloat16, got {dt}")
