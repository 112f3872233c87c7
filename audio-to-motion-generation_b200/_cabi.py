"""ctypes binding of liba2m_b200.so (include/a2m_b200.h).  There is no fallback: if the library is
missing or a call fails, the caller gets an exception."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liba2m_b200.so")

c_i64, c_int, c_float, c_double, c_void_p = (ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double,
                                             ctypes.c_void_p)


class A2MError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("liba2m_b200 error %d: %s" % (code, message))
        self.code = code


class Metrics(ctypes.Structure):
    """a2m_metrics (64 bytes): 5 x int64 counters, 2 x fp64 sums, 1 reserved."""
    _fields_ = [("pck_hits", c_i64), ("n_keypoints", c_i64), ("n_frames", c_i64), ("n_pose", c_i64),
                ("n_motion", c_i64), ("abs_pose", c_double), ("abs_motion", c_double), ("reserved", c_i64)]


class TensorDesc(ctypes.Structure):
    """a2m_tensor_desc"""
    _fields_ = [("name", ctypes.c_char_p), ("data", c_void_p), ("dtype", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("shape", c_i64 * 4)]


class GemmDesc(ctypes.Structure):
    """a2m_gemm_desc"""
    _fields_ = [("n_src", ctypes.c_int32), ("a_rank", ctypes.c_int32 * 2), ("a_ptr", c_void_p * 2),
                ("a_dims", (c_i64 * 5) * 2), ("a_strides", (c_i64 * 5) * 2), ("box", ctypes.c_int32 * 4),
                ("m_extent", ctypes.c_int32 * 4), ("n_taps", ctypes.c_int32), ("tap_src", ctypes.c_int32 * 24),
                ("tap_off", (ctypes.c_int32 * 4) * 24), ("tap_channels", ctypes.c_int32 * 24),
                ("tap_w_off", c_i64 * 24), ("N", ctypes.c_int32), ("act", ctypes.c_int32),
                ("out_type", ctypes.c_int32), ("tile_hint", ctypes.c_int32), ("out_stride", c_i64 * 4),
                ("out_base", c_i64)]


# name -> (restype, argtypes); every symbol declared in include/a2m_b200.h
SIGNATURES = {
    "a2m_version": (c_int, []),
    "a2m_last_error": (ctypes.c_char_p, []),
    "a2m_launch_count": (c_i64, []),
    "a2m_launch_count_reset": (None, []),
    "a2m_mel_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_double, c_int,
                                    ctypes.POINTER(c_void_p)]),
    "a2m_mel_plan_create_ex": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_double, c_int, c_int,
                                       ctypes.POINTER(c_void_p)]),
    "a2m_mel_plan_destroy": (None, [c_void_p]),
    "a2m_mel_num_frames": (c_i64, [c_void_p, c_i64]),
    "a2m_logmel_f32": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "a2m_logmel_i16": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "a2m_mel_schedule_host": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "a2m_stft_magnitude_f32": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "a2m_melspec_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_double, c_int, c_int,
                                        ctypes.POINTER(c_void_p)]),
    "a2m_melspec_plan_destroy": (None, [c_void_p]),
    "a2m_melspec_num_frames": (c_i64, [c_void_p, c_i64]),
    "a2m_melspec_f32": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "a2m_resample_plan_create": (c_int, [c_double, c_double, c_int, ctypes.POINTER(c_void_p)]),
    "a2m_resample_plan_destroy": (None, [c_void_p]),
    "a2m_resample_out_length": (c_i64, [c_void_p, c_i64]),
    "a2m_resample_f32": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_void_p, c_void_p]),
    "a2m_eval_l1_pck_f32": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "a2m_eval_l1_pck_f64": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "a2m_motion_smoothness_f32": (c_int, [c_void_p, c_i64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "a2m_pose_normalize_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p]),
    "a2m_pose_denormalize_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_void_p]),
    "a2m_pose_stats_f64": (c_int, [c_void_p, c_i64, c_void_p, c_void_p]),
    "a2m_pose_stats_ex_f64": (c_int, [c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "a2m_comm_unique_id": (c_int, [c_void_p]),
    "a2m_comm_init": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "a2m_allreduce_metrics": (c_int, [c_void_p, c_void_p, c_void_p]),
    "a2m_allreduce_smoothness": (c_int, [c_void_p, c_void_p, c_void_p]),
    "a2m_comm_destroy": (None, [c_void_p]),
    "a2m_model_create": (c_int, [ctypes.POINTER(TensorDesc), c_int, c_int, ctypes.POINTER(c_void_p)]),
    "a2m_model_destroy": (None, [c_void_p]),
    "a2m_model_forward": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "a2m_model_forward_windows": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p]),
    "a2m_model_set_output_denorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "a2m_model_timeline_begin": (c_int, [c_void_p, c_i64, c_int, c_int, c_int]),
    "a2m_model_timeline_read": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int]),
    "a2m_model_encoder_forward": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p, c_void_p]),
    "a2m_model_encoder_forward_ex": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "a2m_model_plan_generation": (c_i64, [c_void_p]),
    "a2m_block_create": (c_int, [c_int, ctypes.POINTER(TensorDesc), c_int, c_int, c_int, c_int, c_int,
                                 ctypes.POINTER(c_void_p)]),
    "a2m_block_forward": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "a2m_disc_create": (c_int, [ctypes.POINTER(TensorDesc), c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "a2m_disc_out_length": (c_int, [c_int, c_int]),
    "a2m_disc_forward": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "a2m_model_unet_forward": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "a2m_model_gnn_forward": (c_int, [c_void_p, c_int, c_void_p, c_i64, c_void_p, c_void_p]),
    "a2m_model_status": (c_int, [c_void_p]),
    "a2m_model_gemm_flops": (c_i64, [c_void_p, c_i64, c_int, c_int]),
    "a2m_model_profile": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p]),
    "a2m_model_profile_ops": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_void_p, c_int,
                                      c_void_p, c_void_p]),
    "a2m_model_op_name": (ctypes.c_char_p, [c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "a2m_gemm_taps": (c_int, [ctypes.POINTER(GemmDesc), c_void_p, c_i64, c_i64, c_void_p, c_void_p, c_void_p,
                              c_void_p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python audio-to-motion-generation_b200/build.py` "
                "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)            # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code):
    if code != 0:
        raise A2MError(code, lib().a2m_last_error().decode("utf-8", "replace"))


def require_cuda(what):
    if not torch.cuda.is_available():
        raise RuntimeError("%s needs a CUDA device (B200, sm_100a); there is no CPU fallback" % what)


def stream_ptr(device=None):
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
