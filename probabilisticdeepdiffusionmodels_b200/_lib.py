"""ctypes binding of libpddm_b200.so (include/pddm.h).  No torch types cross the ABI: tensors are passed as
raw device pointers + sizes and work is enqueued on ``torch.cuda.current_stream()``.

There is NO fallback: if the library is missing or the device is not sm_100 every call raises.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libpddm_b200.so")

F32, BF16 = 0, 1
MAX_TAPS = 9

c_i32, c_i64, c_f32, c_vp, c_sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class QSampleParams(C.Structure):
    _fields_ = [("x0", c_vp), ("noise", c_vp), ("x_t", c_vp), ("t", c_vp), ("t_const", c_i32),
                ("alphas_hat_sqrt", c_vp), ("one_min_alphas_hat_sqrt", c_vp), ("B", c_i32), ("chw", c_i32)]


class SqErrParams(C.Structure):
    _fields_ = [("pred", c_vp), ("noise", c_vp), ("per_sample", c_vp), ("grad_pred", c_vp), ("gscale", c_vp),
                ("grad_v_unit", c_vp), ("v_scale", c_f32),
                ("B", c_i32), ("C", c_i32), ("c_total", c_i32), ("hw", c_i32)]


class Tables(C.Structure):
    _fields_ = [(n, c_vp) for n in (
        "betas", "alphas_sqrt", "posterior_variance", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
        "posterior_mean_coef1", "posterior_mean_coef2", "denoising_coef", "alphas_hat_sqrt",
        "one_min_alphas_hat_sqrt", "posterior_log_variance_clipped", "log_betas")] + [("T", c_i32)]


class PSampleParams(C.Structure):
    _fields_ = [("x_t", c_vp), ("model_out", c_vp), ("z", c_vp), ("x_prev", c_vp), ("tab", Tables),
                ("t_step_dev", c_vp), ("t_step", c_i32), ("B", c_i32), ("C", c_i32), ("c_out", c_i32), ("hw", c_i32),
                ("clip", c_i32), ("sigma_mode", c_i32)]


class VlbParams(C.Structure):
    _fields_ = [("x0", c_vp), ("x_t", c_vp), ("model_out", c_vp), ("t", c_vp), ("out", c_vp), ("grad_v", c_vp),
                ("tab", Tables), ("B", c_i32), ("C", c_i32), ("c_out", c_i32), ("hw", c_i32), ("mode", c_i32),
                ("sigma_mode", c_i32)]


class GnFwdParams(C.Structure):
    _fields_ = [("x", c_vp), ("x_dtype", c_i32), ("gamma", c_vp), ("beta", c_vp), ("scale", c_vp), ("shift", c_vp),
                ("ld_ss", c_i32), ("y", c_vp), ("mean", c_vp), ("rstd", c_vp), ("B", c_i32), ("HW", c_i32),
                ("C", c_i32), ("G", c_i32), ("eps", c_f32), ("silu", c_i32),
                ("x2", c_vp), ("C_a", c_i32), ("ldx", c_i32), ("ldx2", c_i32), ("ldy", c_i32)]


class GnBwdParams(C.Structure):
    _fields_ = [("x", c_vp), ("x_dtype", c_i32), ("dy", c_vp), ("gamma", c_vp), ("beta", c_vp), ("scale", c_vp),
                ("shift", c_vp), ("ld_ss", c_i32), ("mean", c_vp), ("rstd", c_vp), ("dx", c_vp), ("dx_dtype", c_i32),
                ("dgamma", c_vp), ("dbeta", c_vp), ("dx_colsum", c_vp), ("dscale", c_vp), ("dshift", c_vp),
                ("B", c_i32), ("HW", c_i32), ("C", c_i32), ("G", c_i32), ("silu", c_i32),
                ("x2", c_vp), ("C_a", c_i32), ("ldx", c_i32), ("ldx2", c_i32), ("lddy", c_i32),
                ("gres", c_vp), ("ld_gres", c_i32), ("dx2", c_vp), ("ld_dx", c_i32), ("ld_dx2", c_i32),
                ("dx_accumulate", c_i32), ("dx2_accumulate", c_i32), ("part_dgamma", c_vp), ("part_dbeta", c_vp),
                ("ld_part", c_i32), ("dx_colsum2", c_vp), ("ld_colsum", c_i32), ("ld_colsum2", c_i32),
                ("colsum_accumulate", c_i32), ("colsum2_accumulate", c_i32)]


class ConvParams(C.Structure):
    _fields_ = [("x", c_vp), ("w", c_vp), ("bias", c_vp), ("bcast", c_vp), ("residual", c_vp), ("y", c_vp),
                ("ld_bcast", c_i32), ("res_dtype", c_i32), ("y_dtype", c_i32),
                ("x_NB", c_i32), ("B", c_i32), ("H", c_i32), ("W", c_i32), ("Cin", c_i32), ("ldx", c_i32),
                ("Cout", c_i32), ("ntaps", c_i32),
                ("tap_db", c_i32 * MAX_TAPS), ("tap_dh", c_i32 * MAX_TAPS), ("tap_dw", c_i32 * MAX_TAPS),
                ("tap_w", c_i32 * MAX_TAPS), ("w_ntaps", c_i32),
                ("out_H", c_i32), ("out_W", c_i32), ("out_sh", c_i32), ("out_sw", c_i32), ("out_oh", c_i32),
                ("out_ow", c_i32), ("x2", c_vp), ("Cin_a", c_i32), ("ldx2", c_i32)]


class WgradParams(C.Structure):
    _fields_ = [("x", c_vp), ("dy", c_vp), ("dw", c_vp),
                ("x_NB", c_i32), ("B", c_i32), ("H", c_i32), ("W", c_i32), ("Cin", c_i32), ("ldx", c_i32),
                ("Cout", c_i32), ("lddy", c_i32), ("ntaps", c_i32),
                ("tap_db", c_i32 * MAX_TAPS), ("tap_dh", c_i32 * MAX_TAPS), ("tap_dw", c_i32 * MAX_TAPS),
                ("dw_layout", c_i32), ("accumulate", c_i32), ("dw_ldc", c_i32), ("dw_c0", c_i32)]


class PackDesc(C.Structure):
    _fields_ = [("src", c_vp), ("dst", c_vp), ("Cout", c_i32), ("Cin", c_i32), ("ntaps", c_i32), ("mode", c_i32),
                ("Cout_pad", c_i32), ("Cin_pad", c_i32), ("ld_dst", c_i32), ("reserved", c_i32), ("dst2", c_vp)]


PACK_TILE = 32  # PDDM_PACK_TILE


class AttnFwdParams(C.Structure):
    _fields_ = [("qkv", c_vp), ("out", c_vp), ("lse", c_vp), ("B", c_i32), ("T", c_i32), ("heads", c_i32), ("d", c_i32)]


class AttnBwdParams(C.Structure):
    _fields_ = [("qkv", c_vp), ("out", c_vp), ("dout", c_vp), ("lse", c_vp), ("dqkv", c_vp), ("B", c_i32),
                ("T", c_i32), ("heads", c_i32), ("d", c_i32), ("ws", c_vp), ("ws_bytes", c_i64)]


class AdamParams(C.Structure):
    _fields_ = [("param", c_vp), ("grad", c_vp), ("exp_avg", c_vp), ("exp_avg_sq", c_vp), ("ema", c_vp), ("n", c_i64),
                ("lr", c_f32), ("beta1", c_f32), ("beta2", c_f32), ("eps", c_f32), ("weight_decay", c_f32),
                ("ema_decay", c_f32), ("grad_scale", c_f32), ("step", c_i32), ("step_dev", c_vp), ("lr_dev", c_vp)]


# symbol -> (restype, argtypes); this table is also what tests/test_abi.py checks against include/pddm.h
P = C.POINTER
SIGNATURES = {
    "pddm_version": (c_i32, []),
    "pddm_strerror": (C.c_char_p, [c_i32]),
    "pddm_check_device": (c_i32, []),
    "pddm_sm_count": (c_i32, []),
    "pddm_set_sm_reserve": (c_i32, [c_i32]),
    "pddm_get_sm_reserve": (c_i32, []),
    "pddm_q_sample": (c_i32, [P(QSampleParams), c_vp]),
    "pddm_sq_err": (c_i32, [P(SqErrParams), c_vp]),
    "pddm_p_sample_step": (c_i32, [P(PSampleParams), c_vp]),
    "pddm_step_advance": (c_i32, [c_vp, c_vp, c_i32, c_vp]),
    "pddm_vlb_terms": (c_i32, [P(VlbParams), c_vp]),
    "pddm_timestep_embedding": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_f32, c_vp]),
    "pddm_nchw_to_nhwc": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_nhwc_to_nchw": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "pddm_copy_channels": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_i32, c_i32, c_i64, c_i32, c_vp]),
    "pddm_upsample2x": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_upsample2x_bwd": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_phase_split": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_phase_merge": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_add_bf16": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pddm_convert": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i64, c_vp]),
    "pddm_silu": (c_i32, [c_vp, c_vp, c_i32, c_i64, c_vp]),
    "pddm_silu_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pddm_images_to_uint8": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "pddm_silu_map": (c_i32, [c_vp, c_vp, c_i64, c_vp]),
    "pddm_silu_map_bwd": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pddm_colsum_workspace": (C.c_size_t, [c_i64, c_i32, c_i32]),
    "pddm_colsum": (c_i32, [c_vp, c_i32, c_i64, c_i32, c_vp, c_i32, c_vp, C.c_size_t, c_vp]),
    "pddm_colsum_per_sample": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, C.c_size_t, c_vp]),
    "pddm_gn_silu_fwd_workspace": (c_sz, [c_i32, c_i32]),
    "pddm_gn_silu_fwd": (c_i32, [P(GnFwdParams), c_vp, c_sz, c_vp]),
    "pddm_gn_silu_bwd_workspace": (c_sz, [c_i32, c_i32]),
    "pddm_gn_silu_bwd": (c_i32, [P(GnBwdParams), c_vp, c_sz, c_vp]),
    "pddm_gn_pipe_slots": (c_i32, [c_i32, c_i32, c_i32, c_i32, c_i32, c_i32]),
    "pddm_conv2d_fwd": (c_i32, [P(ConvParams), c_vp]),
    "pddm_conv2d_wgrad_workspace": (c_sz, [P(WgradParams)]),
    "pddm_conv2d_wgrad": (c_i32, [P(WgradParams), c_vp, c_sz, c_vp]),
    "pddm_pack_conv_weight": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_pack_weights_multi": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_vp]),
    "pddm_colsum_f32": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_vp]),
    "pddm_colsum_rows": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_i32, c_vp, c_i32, c_i32, c_vp]),
    "pddm_batch_fold": (c_i32, [c_vp, c_i32, c_i32, c_vp, c_i32, c_vp, c_vp]),
    "pddm_convert_rows": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_i32, c_vp]),
    "pddm_im2col3x3": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_nchw_to_nhwc_padded": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_nhwc_slice_to_nchw": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_attn_fwd": (c_i32, [P(AttnFwdParams), c_vp]),
    "pddm_attn_bwd": (c_i32, [P(AttnBwdParams), c_vp]),
    "pddm_attn_bwd_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32, c_i32]),
    "pddm_split_bf16": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_vp]),
    "pddm_gn_split_f32": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_f32, c_i32,
                                  c_vp]),
    "pddm_attn_fwd_f32": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_vp]),
    "pddm_adam_ema_step": (c_i32, [P(AdamParams), c_vp]),
    "pddm_adam_ema_multi": (c_i32, [c_vp, c_vp, c_i32, P(AdamParams), c_vp]),
    "pddm_counter_add": (c_i32, [c_vp, c_i32, c_vp]),
}

_lib = None


def load():
    """Load the shared library (no CUDA call is made). Raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m probabilisticdeepdiffusionmodels_b200.build` "
                "(there is no CPU / eager fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


_device_ok = {}


def require_device(t=None):
    """Fail loudly unless we are on an sm_100 CUDA device with the library loaded."""
    lib = load()
    if not torch.cuda.is_available():
        raise RuntimeError("pddm_b200 requires a CUDA (sm_100a / B200) device; there is no CPU fallback")
    dev = t.device.index if (t is not None and t.is_cuda) else torch.cuda.current_device()
    if dev not in _device_ok:
        with torch.cuda.device(dev):
            rc = lib.pddm_check_device()
        _device_ok[dev] = rc
    if _device_ok[dev] != 0:
        raise RuntimeError("pddm_b200: " + lib.pddm_strerror(_device_ok[dev]).decode())
    return lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"pddm {what} failed: {load().pddm_strerror(rc).decode()} ({rc})")


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def dt(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"unsupported dtype {t.dtype}")


LAUNCHES = [0]  # number of C-ABI compute calls issued from Python
KERNELS = [0]   # number of kernels those calls launched (bench.py's gpu_launches claim is derived from it)
# entry points that launch more than one kernel (memsets are not counted)
_KERNELS_PER_CALL = {"pddm_gn_silu_fwd": 1, "pddm_gn_silu_bwd": 2, "pddm_conv2d_wgrad": 2, "pddm_colsum": 2,
                     "pddm_colsum_per_sample": 2}


def call(name, *args):
    lib = load()
    LAUNCHES[0] += 1
    KERNELS[0] += _KERNELS_PER_CALL.get(name, 1)
    check(getattr(lib, name)(*args), name)
