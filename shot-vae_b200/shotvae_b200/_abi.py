"""ctypes binding of libshotvae.so (C ABI in include/shotvae.h).

There is deliberately no fallback: if the CUDA library is missing or its ABI does not match, import
fails loudly.  Tensors cross the boundary as raw device pointers; torch only owns the memory and
the stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libshotvae.so")
MAX_TAPS = 16

if not os.path.exists(LIB_PATH):
    raise ImportError("libshotvae.so not found at %s -- run `python __graft_entry__.py` (nvcc, sm_100a) first; "
                      "there is no CPU fallback" % LIB_PATH)
lib = C.CDLL(LIB_PATH)

vp, f32p, i64p, i8p = C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int8)
i32, i64, f32 = C.c_int32, C.c_int64, C.c_float


class IgemmArgs(C.Structure):
    _fields_ = [("A", vp), ("Wt", vp), ("out_bf16", vp), ("out_f32", vp), ("residual", vp), ("bias", vp), ("stats", vp),
                ("NB", i32), ("H", i32), ("W", i32), ("C", i32), ("OH", i32), ("OW", i32), ("N", i32), ("T", i32),
                ("in_stride", i32), ("out_stride", i32), ("out_off_y", i32), ("out_off_x", i32),
                ("OHf", i32), ("OWf", i32), ("n_valid", i32), ("group_images", i32),
                ("dy", C.c_int8 * MAX_TAPS), ("dx", C.c_int8 * MAX_TAPS), ("impl", i32), ("w_layout", i32),
                ("bn_y", vp), ("bn_scale", vp), ("bn_shift", vp), ("bn_mean", vp), ("bn_var", vp), ("bn_slope", f32), ("bn_eps", f32)]


class WgradArgs(C.Structure):
    _fields_ = [("A", vp), ("Gr", vp), ("partial", vp),
                ("NB", i32), ("H", i32), ("W", i32), ("C", i32), ("OH", i32), ("OW", i32), ("N", i32), ("T", i32),
                ("in_stride", i32), ("splits", i32), ("dy", C.c_int8 * MAX_TAPS), ("dx", C.c_int8 * MAX_TAPS), ("impl", i32)]


class ReduceDesc(C.Structure):
    _fields_ = [("partial", vp), ("grad", vp), ("sn", i64), ("sc", i64), ("st", i64),
                ("splits", i32), ("N", i32), ("C", i32), ("T", i32), ("n_real", i32), ("c_real", i32),
                ("tap_index", C.c_int8 * MAX_TAPS)]


class Heads(C.Structure):
    _fields_ = [("W", vp * 4), ("bias", vp * 4), ("out", vp * 4), ("g", vp * 4), ("dW", vp * 4), ("dbias", vp * 4),
                ("N", i32 * 4), ("n", i32)]


class PackDesc(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("sn", i64), ("sc", i64), ("st", i64),
                ("N", i32), ("C", i32), ("T", i32), ("n_real", i32), ("c_real", i32), ("layout", i32),
                ("tap", C.c_int8 * MAX_TAPS)]


class RunDesc(C.Structure):
    _fields_ = [("mean", vp * 4), ("var", vp * 4), ("running_mean", vp), ("running_var", vp), ("nbt", vp),
                ("count", f32), ("npass", i32), ("C", i32)]


class BnBwdTerm(C.Structure):
    _fields_ = [("g_a", vp), ("g_feat", vp), ("scale", vp), ("shift", vp), ("mean", vp), ("var", vp),
                ("dgamma", vp), ("dbeta", vp), ("grad_gamma", vp), ("grad_beta", vp), ("slope", f32), ("c_real", i32)]


_PROTOS = {
    "sv_abi_version": (C.c_int, []),
    "sv_last_error": (C.c_char_p, []),
    "sv_has_tcgen05": (C.c_int, []),
    "sv_launch_count": (C.c_longlong, []),
    "sv_set_cta_limit": (C.c_int, [i32]),
    "sv_sizeof_igemm_args": (C.c_int, []),
    "sv_sizeof_wgrad_args": (C.c_int, []),
    "sv_sizeof_bn_bwd_term": (C.c_int, []),
    "sv_igemm_fprop": (C.c_int, [C.POINTER(IgemmArgs), vp]),
    "sv_igemm_fprop_supports": (C.c_int, [C.POINTER(IgemmArgs), i32]),
    "sv_igemm_fprop_batch": (C.c_int, [C.POINTER(IgemmArgs), i32, vp]),
    "sv_igemm_wgrad": (C.c_int, [C.POINTER(WgradArgs), vp]),
    "sv_igemm_wgrad_splits": (C.c_int, [C.POINTER(WgradArgs)]),
    "sv_wgrad_reduce": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, i64, i64, i64, i8p, vp]),
    "sv_wgrad_reduce_batched": (C.c_int, [C.POINTER(ReduceDesc), i32, vp]),
    "sv_sizeof_wgrad_reduce_desc": (C.c_int, []),
    "sv_pack_weight": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i64, i64, i64, i8p, i32, vp]),
    "sv_pack_weights_batched": (C.c_int, [vp, i32, i32, vp]),
    "sv_sizeof_pack_desc": (C.c_int, []),
    "sv_pack_image": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
    "sv_nhwc_to_nchw_f32": (C.c_int, [vp, vp, i32, i32, i32, vp]),
    "sv_bn_finalize": (C.c_int, [vp, vp, vp, f32, f32, i32, i32, i32, vp, vp, vp, vp, vp]),
    "sv_bn_act_fwd": (C.c_int, [vp, vp, vp, vp, f32, i64, i32, i32, vp]),
    "sv_bn_finalize_act_fwd": (C.c_int, [vp, vp, vp, vp, vp, f32, f32, f32, i64, i32, i32, vp, vp, vp, vp, vp]),
    "sv_bn_running_update_batched": (C.c_int, [vp, i32, i32, f32, vp]),
    "sv_sizeof_run_desc": (C.c_int, []),
    "sv_bn_act_gap_fwd": (C.c_int, [vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]),
    "sv_bn_bwd_reduce": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, f32, f32, i64, i32, i32, i32, vp, vp, vp]),
    "sv_bn_bwd_apply": (C.c_int, [C.POINTER(BnBwdTerm), i32, vp, vp, vp, f32, i64, i32, i32, i32, vp]),
    "sv_bn_running_update": (C.c_int, [C.POINTER(vp), C.POINTER(vp), i32, f32, f32, i32, vp, vp, vp, vp]),
    "sv_colsum_bf16": (C.c_int, [vp, vp, i64, i32, i32, vp]),
    "sv_pad_channels_f32": (C.c_int, [vp, vp, i64, i32, i32, vp]),
    "sv_colsum_f32": (C.c_int, [vp, vp, i64, i32, i32, vp]),
    "sv_pack_image_f32": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
    "sv_bn_act_fwd_f32": (C.c_int, [vp, vp, vp, vp, f32, i64, i32, i32, vp]),
    "sv_bn_finalize_act_fwd_f32": (C.c_int, [vp, vp, vp, vp, vp, f32, f32, f32, i64, i32, i32, vp, vp, vp, vp, vp]),
    "sv_bn_act_gap_fwd_f32": (C.c_int, [vp, vp, vp, vp, f32, i32, i32, i32, i32, vp]),
    "sv_bn_bwd_reduce_f32": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, f32, f32, i64, i32, i32, i32, vp, vp, vp]),
    "sv_bn_bwd_apply_f32": (C.c_int, [C.POINTER(BnBwdTerm), i32, vp, vp, vp, f32, i64, i32, i32, i32, vp]),
    "sv_linear_fwd": (C.c_int, [vp, i32, vp, i32, i32, vp, vp, vp, i32, vp, i32, i32, i32, i32, vp]),
    "sv_linear_bwd_input": (C.c_int, [vp, vp, i32, vp, i32, i32, vp, i32, i32, i32, i32, i32, vp]),
    "sv_linear_bwd_weight": (C.c_int, [vp, vp, i32, vp, i32, vp, i32, i32, vp, i32, i32, i32, vp]),
    "sv_sizeof_heads": (C.c_int, []),
    "sv_heads_fwd": (C.c_int, [vp, i32, C.POINTER(Heads), i32, i32, vp]),
    "sv_heads_bwd_input": (C.c_int, [C.POINTER(Heads), vp, i32, i32, i32, vp]),
    "sv_heads_bwd_weight": (C.c_int, [C.POINTER(Heads), vp, i32, i32, i32, vp]),
    "sv_log_softmax_fwd": (C.c_int, [vp, vp, i32, i32, vp]),
    "sv_log_softmax_bwd": (C.c_int, [vp, vp, vp, i32, i32, vp]),
    "sv_sample_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i32, f32, i32, i32, i32, vp, i32, vp]),
    "sv_sample_bwd": (C.c_int, [vp, i32, vp, vp, vp, i32, f32, i32, i32, i32, vp, vp, vp, i32, vp]),
    "sv_elbo_rec_fwd_bwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, i32, vp, vp]),
    "sv_elbo_kl_fwd": (C.c_int, [vp, vp, vp, i32, i32, i32, vp, vp]),
    "sv_elbo_kl_bwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, vp]),
    "sv_posterior_fwd_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp]),
    "sv_inference_kl": (C.c_int, [vp, vp, i32, i32, vp, vp]),
    "sv_mixup_lerp": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp]),
    "sv_noise_fill": (C.c_int, [vp, i64, vp, i64, vp, vp]),
    "sv_sizeof_noise_state": (C.c_int, []),
    "sv_fill_zero": (C.c_int, [vp, i64, vp]),
    "sv_pairwise_kl_second_nearest": (C.c_int, [vp, vp, i32, i32, vp, vp, vp]),
    "sv_sgd_step": (C.c_int, [vp, vp, vp, vp, i64, vp]),
    "sv_pairwise_dist": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    "sv_kl_pair_fwd_bwd": (C.c_int, [i32, vp, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp]),
    "sv_augment_batch": (C.c_int, [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
    "sv_debug_halo_trace": (C.c_int, [vp]),
}
EXPORTS = sorted(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)          # AttributeError here = symbol missing from the library
    _fn.restype, _fn.argtypes = _res, _args

if lib.sv_abi_version() != 1:
    raise ImportError("libshotvae ABI version %d, binding expects 1" % lib.sv_abi_version())
for _fn, _st in (("sv_sizeof_heads", Heads), ("sv_sizeof_wgrad_reduce_desc", ReduceDesc), ("sv_sizeof_run_desc", RunDesc), ("sv_sizeof_pack_desc", PackDesc), ("sv_sizeof_igemm_args", IgemmArgs), ("sv_sizeof_wgrad_args", WgradArgs), ("sv_sizeof_bn_bwd_term", BnBwdTerm)):
    if getattr(lib, _fn)() != C.sizeof(_st):
        raise ImportError("ctypes mirror of %s is %d bytes, library says %d" % (_st.__name__, C.sizeof(_st), getattr(lib, _fn)()))


class ShotVaeError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise ShotVaeError("libshotvae error %d: %s" % (rc, lib.sv_last_error().decode()))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL).  Refuses CPU tensors: no CPU path exists."""
    if t is None:
        return None
    if not t.is_cuda:
        raise ShotVaeError("libshotvae needs CUDA tensors (got a %s tensor); there is no CPU fallback" % t.device)
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def taps_array(vals):
    a = (C.c_int8 * MAX_TAPS)()
    for i, v in enumerate(vals):
        a[i] = v
    return a


def launch_count():
    return int(lib.sv_launch_count())
