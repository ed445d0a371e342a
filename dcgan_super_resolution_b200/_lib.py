"""ctypes binding of libdcgansr.so (include/dcgansr.h).  No CPU fallback: if the CUDA library is
missing or no B200 is visible, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdcgansr.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "dcgansr.h")

OK = 0
STRICT_FP32, FAST_TF32 = 0, 1
CONV, FULLCONV, BN, RELU, LRELU, TANH, SIGMOID, UPNEAREST, VIEW = 1, 2, 3, 4, 5, 6, 7, 8, 9
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
LOSS_BCE, LOSS_MSE = 0, 1


class Cfg(C.Structure):
    _fields_ = [("device", C.c_int), ("precision", C.c_int), ("world_size", C.c_int), ("rank", C.c_int),
                ("sync_bn", C.c_int), ("use_graph", C.c_int)]


class Layer(C.Structure):
    _fields_ = [("kind", C.c_int), ("cin", C.c_int), ("cout", C.c_int),
                ("kh", C.c_int), ("kw", C.c_int), ("sh", C.c_int), ("sw", C.c_int), ("ph", C.c_int), ("pw", C.c_int),
                ("adjh", C.c_int), ("adjw", C.c_int),
                ("negval", C.c_float), ("eps", C.c_float), ("momentum", C.c_float), ("scale", C.c_int)]


class StepCfg(C.Structure):
    _fields_ = [("loss", C.c_int), ("real_label", C.c_float), ("fake_label", C.c_float), ("gen_label", C.c_float),
                ("pixel_label", C.c_int), ("pixel_div", C.c_float),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double)]


class DcgansrError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_F = C.POINTER(C.c_float)
_I64 = C.c_int64
_SIGS = {
    "dcgansr_version": (C.c_int, []),
    "dcgansr_ctx_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(_P)]),
    "dcgansr_ctx_destroy": (None, [_P]),
    "dcgansr_last_error": (C.c_char_p, [_P]),
    "dcgansr_synchronize": (C.c_int, [_P]),
    "dcgansr_timer_begin": (C.c_int, [_P]),
    "dcgansr_timer_end": (C.c_int, [_P, _F]),
    "dcgansr_launch_count": (C.c_int, [_P, C.POINTER(_I64)]),
    "dcgansr_flush_l2": (C.c_int, [_P]),
    "dcgansr_profile_begin": (C.c_int, [_P]),
    "dcgansr_profile_end": (C.c_int, [_P, C.c_char_p, _I64]),
    "dcgansr_comm_get_unique_id": (C.c_int, [_P, _P]),
    "dcgansr_comm_init": (C.c_int, [_P, _P]),
    "dcgansr_comm_peer_enabled": (C.c_int, [_P]),
    "dcgansr_net_create": (C.c_int, [_P, C.POINTER(Layer), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "dcgansr_net_destroy": (None, [_P]),
    "dcgansr_net_out_shape": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dcgansr_net_num_params": (C.c_int, [_P, C.POINTER(_I64)]),
    "dcgansr_net_set_params": (C.c_int, [_P, _P]),
    "dcgansr_net_get_params": (C.c_int, [_P, _P]),
    "dcgansr_net_get_grads": (C.c_int, [_P, _P]),
    "dcgansr_net_num_bn_channels": (C.c_int, [_P, C.POINTER(_I64)]),
    "dcgansr_net_get_bn_running": (C.c_int, [_P, _P, _P]),
    "dcgansr_net_set_bn_running": (C.c_int, [_P, _P, _P]),
    "dcgansr_net_get_adam_state": (C.c_int, [_P, _P, _P, C.POINTER(_I64)]),
    "dcgansr_net_set_adam_state": (C.c_int, [_P, _P, _P, _I64]),
    "dcgansr_net_forward": (C.c_int, [_P, _P, C.c_int, _P]),
    "dcgansr_net_backward": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "dcgansr_net_update_grad_input": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "dcgansr_net_zero_grads": (C.c_int, [_P]),
    "dcgansr_net_adam": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_double]),
    "dcgansr_train_step": (C.c_int, [_P, _P, _P, C.POINTER(StepCfg), _P, C.c_int, _P]),
    "dcgansr_stage_batch": (C.c_int, [_P, _P, _P, C.c_int, C.c_int]),
    "dcgansr_train_step_staged": (C.c_int, [_P, _P, _P, C.POINTER(StepCfg), C.c_int, C.c_int, _P]),
    "dcgansr_generate": (C.c_int, [_P, _P, _P, C.c_int, _P]),
}
_CONV_SIG = (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 8)
for _n in ("conv2d_fwd", "conv2d_dgrad", "conv2d_wgrad", "fullconv2d_fwd", "fullconv2d_dgrad", "fullconv2d_wgrad"):
    _SIGS["dcgansr_" + _n] = _CONV_SIG
_SIGS["dcgansr_extract_patches"] = (C.c_int, [_P, _P, _P] + [C.c_int] * 7)
_SIGS["dcgansr_assemble_patches"] = (C.c_int, [_P, _P, _P] + [C.c_int] * 7)
_SIGS["dcgansr_stitch_overlap"] = (C.c_int, [_P, _P, _P] + [C.c_int] * 6)
_SIGS["dcgansr_scale_bilinear"] = (C.c_int, [_P, _P, _P] + [C.c_int] * 5)
_SIGS["dcgansr_stage_patches"] = (C.c_int, [_P, _P, _P] + [C.c_int] * 8)
_SIGS["dcgansr_psnr"] = (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 3)
_SIGS["dcgansr_ssim"] = (C.c_int, [_P, _P, _P, _P] + [C.c_int] * 3)
_SIGS["dcgansr_bench_conv"] = (C.c_int, [_P] + [C.c_int] * 11 + [_F])
_SIGS.update({
    "dcgansr_bn_fwd_train": (C.c_int, [_P] * 9 + [C.c_int] * 4 + [C.c_float, C.c_float]),
    "dcgansr_bn_bwd": (C.c_int, [_P] * 9 + [C.c_int] * 4),
    "dcgansr_act_fwd": (C.c_int, [_P, _P, _P, _I64, C.c_int, C.c_float]),
    "dcgansr_act_bwd": (C.c_int, [_P, _P, _P, _P, _I64, C.c_int, C.c_float]),
    "dcgansr_upnearest2_fwd": (C.c_int, [_P, _P, _P] + [C.c_int] * 4),
    "dcgansr_upnearest2_bwd": (C.c_int, [_P, _P, _P] + [C.c_int] * 4),
    "dcgansr_avgpool2_fwd": (C.c_int, [_P, _P, _P] + [C.c_int] * 4),
    "dcgansr_bce": (C.c_int, [_P, _P, _P, _I64, _P, _P]),
    "dcgansr_mse": (C.c_int, [_P, _P, _P, _I64, _P, _P]),
    "dcgansr_pixel_mse_per_sample": (C.c_int, [_P, _P, _P, _P, C.c_int, _I64, C.c_float]),
    "dcgansr_adam_step": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I64, C.c_double, C.c_double, C.c_double, C.c_double]),
})


def symbols():
    """Names of every entry point this binding declares (== what include/dcgansr.h declares)."""
    return sorted(_SIGS)


def load():
    """dlopen libdcgansr.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DcgansrError(f"{LIB_PATH} is missing: run `python -m dcgan_super_resolution_b200.build` "
                           "(libdcgansr.so is the only compute path; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, ctx=None):
    if rc != 0:
        msg = load().dcgansr_last_error(ctx)
        raise DcgansrError(f"libdcgansr error {rc}: {msg.decode() if msg else '?'}")
