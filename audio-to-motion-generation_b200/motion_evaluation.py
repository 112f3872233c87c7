"""B200 drop-in for ``motion_evaluation.py`` of the reference: PCK and its radius, plus the fused
L1 / PCK partial-sum evaluation the multi-GPU driver all-reduces.

``compute_pck`` / ``compute_pck_radius`` keep the reference's signatures (motion_evaluation.py:4,17)
and its numpy-in / numpy-out behaviour; arithmetic is numpy's fp32 operation order, reproduced on the
GPU bit for bit by csrc/eval.cu.  The reference computes in the dtype of its inputs: float32 arrays
in fp32 (the hot path), float64 arrays in fp64 -- both widths have a kernel instantiation, so hit counts
equal the reference's for either; other dtypes (ints, halves) are converted to fp32 first.  torch tensors
are accepted and yield torch tensors on their own device.
"""
import ctypes

import numpy as np
import torch

from . import _cabi

K_JOINTS = 52            # the radius is tiled to 52 keypoints, motion_evaluation.py:22
METRIC_FIELDS = ("pck_hits", "n_keypoints", "n_frames", "n_pose", "n_motion")


def _cuda_device_of(*tensors):
    """The CUDA device the work runs on: the inputs' own device if any of them lives on a GPU (never silently the
    *current* device), else the current device."""
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    return torch.device("cuda", torch.cuda.current_device())


def _same_device(device, **tensors):
    for name, t in tensors.items():
        if t is not None and t.device != device:
            raise ValueError("%s is on %s but the poses are on %s" % (name, t.device, device))


def _frames_on_device(x, name, device=None, dtype=None):
    """[N, 2, 52] (numpy / torch, any device) -> contiguous CUDA tensor, fp64 if the input is fp64 else fp32."""
    was_numpy = not isinstance(x, torch.Tensor)
    t = torch.as_tensor(np.asarray(x)) if was_numpy else x
    if t.dim() != 3 or t.shape[1] != 2 or t.shape[2] != K_JOINTS:
        raise ValueError("%s must have shape [N, 2, %d], got %s" % (name, K_JOINTS, tuple(t.shape)))
    if dtype is None:
        dtype = torch.float64 if t.dtype == torch.float64 else torch.float32
    t = t.to(device=device or _cuda_device_of(t), dtype=dtype).contiguous()
    return t, was_numpy


def new_metrics(device="cuda"):
    """Zeroed 64-byte accumulator (a2m_metrics) as an int64[8] CUDA tensor."""
    return torch.zeros(8, dtype=torch.int64, device=device)


def read_metrics(buf):
    """Device accumulator -> dict (one 64-byte D2H copy)."""
    host = buf.cpu()
    out = {k: int(host[i]) for i, k in enumerate(METRIC_FIELDS)}
    sums = host[5:7].view(torch.float64)
    out["abs_pose"], out["abs_motion"] = float(sums[0]), float(sums[1])
    return out


def finalize_metrics(m):
    """hits / keypoints and sum / n -- identical on every rank after the all-reduce."""
    return {"pck": m["pck_hits"] / max(m["n_keypoints"], 1),
            "l1_pose": m["abs_pose"] / max(m["n_pose"], 1),
            "l1_motion": m["abs_motion"] / max(m["n_motion"], 1)}


def evaluate_poses(pred, gt, alpha=0.2, accum=None, pck_per_frame=None, radius_per_frame=None):
    """Fused evaluation of a batch of pose sequences [B, T, 104] (CUDA tensors, fp32 -- or fp64, computed in fp64 like
    the reference does for float64 arrays): accumulates PCK hits, |pred-gt| and |motion(pred)-motion(gt)| sums into
    ``accum`` (device, 64 bytes)."""
    _cabi.require_cuda("evaluate_poses")
    if pred.shape != gt.shape or pred.dim() != 3 or pred.shape[-1] != 2 * K_JOINTS:
        raise ValueError("pred and gt must both be [B, T, %d]" % (2 * K_JOINTS))
    wide = pred.dtype == torch.float64 or gt.dtype == torch.float64
    dtype = torch.float64 if wide else torch.float32
    dev = _cuda_device_of(pred, gt)
    pred = pred.to(device=dev, dtype=dtype).contiguous()
    gt = gt.to(device=dev, dtype=dtype).contiguous()
    if accum is None:
        accum = new_metrics(dev)
    _same_device(dev, accum=accum, pck_per_frame=pck_per_frame, radius_per_frame=radius_per_frame)
    if radius_per_frame is not None and radius_per_frame.dtype != dtype:
        raise ValueError("radius_per_frame must be %s like the poses" % dtype)
    entry = _cabi.lib().a2m_eval_l1_pck_f64 if wide else _cabi.lib().a2m_eval_l1_pck_f32
    with torch.cuda.device(dev):
        _cabi.check(entry(_cabi.ptr(pred), _cabi.ptr(gt), pred.shape[0], pred.shape[1], float(alpha),
                          _cabi.ptr(pck_per_frame), _cabi.ptr(radius_per_frame), _cabi.ptr(accum),
                          _cabi.stream_ptr(dev)))
    return accum


SMOOTH_FIELDS = ("sum_accel_norm", "sum_jerk_norm", "n_accel", "n_jerk")


def new_smoothness(device="cuda"):
    """Zeroed 32-byte accumulator (a2m_smooth_metrics) as an int64[4] CUDA tensor (two fp64 sums, two counts)."""
    return torch.zeros(4, dtype=torch.int64, device=device)


def evaluate_smoothness(seq, accum=None, from_pose=False):
    """Accumulate the smoothness / jerk partial sums of a batch [B, L, F] (fp32 CUDA): ``seq`` is the motion
    (velocities) the reference functions take, or poses when ``from_pose`` (their first difference is taken on
    the fly, pos_to_motion)."""
    _cabi.require_cuda("evaluate_smoothness")
    if seq.dim() != 3:
        raise ValueError("motion sequence must be [B, L, F], got %s" % (tuple(seq.shape),))
    seq = seq.to(device=_cuda_device_of(seq), dtype=torch.float32).contiguous()
    if accum is None:
        accum = new_smoothness(seq.device)
    _same_device(seq.device, accum=accum)
    with torch.cuda.device(seq.device):
        _cabi.check(_cabi.lib().a2m_motion_smoothness_f32(
            _cabi.ptr(seq), seq.shape[0], seq.shape[1], seq.shape[2], int(bool(from_pose)), _cabi.ptr(accum),
            _cabi.stream_ptr(seq.device)))
    return accum


def read_smoothness(buf):
    """Device accumulator -> {'smoothness', 'jerk', sums, counts} (one 32-byte D2H copy); an empty mean is NaN as
    in the reference (torch.mean of an empty tensor)."""
    host = buf.cpu()
    sums = host[0:2].view(torch.float64)
    out = {"sum_accel_norm": float(sums[0]), "sum_jerk_norm": float(sums[1]), "n_accel": int(host[2]), "n_jerk": int(host[3])}
    out["smoothness"] = out["sum_accel_norm"] / out["n_accel"] if out["n_accel"] else float("nan")
    out["jerk"] = out["sum_jerk_norm"] / out["n_jerk"] if out["n_jerk"] else float("nan")
    return out


def pos_to_motion(pose_batch):
    """First difference along time (version5_model_train.py:208-213); a torch op on the caller's device -- the fused
    kernels take poses directly (``evaluate_poses``, ``evaluate_smoothness(from_pose=True)``)."""
    return torch.diff(pose_batch, n=1, dim=1)


def _smooth_scalar(motion_seq, key):
    was_numpy = not isinstance(motion_seq, torch.Tensor)
    t = torch.as_tensor(np.asarray(motion_seq)) if was_numpy else motion_seq
    val = read_smoothness(evaluate_smoothness(t))[key]
    return np.float32(val) if was_numpy else torch.tensor(val, dtype=torch.float32, device=t.device)


def compute_temporal_smoothness_loss(motion_seq):
    """mean_{b,t} ||motion[b,t+1] - motion[b,t]||_2 (version5_model_train.py:216-230), a 0-dim fp32 tensor."""
    _cabi.require_cuda("compute_temporal_smoothness_loss")
    return _smooth_scalar(motion_seq, "smoothness")


def compute_jerk_loss(motion_seq):
    """mean_{b,t} ||accel[b,t+1] - accel[b,t]||_2 (version5_model_train.py:233-248), a 0-dim fp32 tensor."""
    _cabi.require_cuda("compute_jerk_loss")
    return _smooth_scalar(motion_seq, "jerk")


def _per_frame(pred, gt, alpha, want):
    g, was_numpy = _frames_on_device(gt, "gt")
    wide = g.dtype == torch.float64 or (isinstance(pred, torch.Tensor) and pred.dtype == torch.float64) or \
        (isinstance(pred, np.ndarray) and pred.dtype == np.float64)
    dtype = torch.float64 if wide else torch.float32
    g = g.to(dtype)
    p, _ = _frames_on_device(pred, "pred", g.device, dtype)
    if p.shape != g.shape:
        raise ValueError("pred and gt differ in shape: %s vs %s" % (tuple(p.shape), tuple(g.shape)))
    n = g.shape[0]
    pck = torch.empty(n, dtype=torch.float64, device=g.device) if want == "pck" else None
    rad = torch.empty(n, dtype=dtype, device=g.device) if want == "radius" else None
    # every frame is its own "clip" of length 1: no motion term, per-frame outputs only
    evaluate_poses(p.view(n, 1, 2 * K_JOINTS), g.view(n, 1, 2 * K_JOINTS), alpha, None, pck, rad)
    return (pck if want == "pck" else rad), was_numpy


def compute_pck(pred, gt, alpha=0.2):
    """Per-sample fraction of the 52 keypoints with ||gt - pred||_2 <= alpha * max(bbox w, bbox h)
    -> float64 [N] (motion_evaluation.py:4-14)."""
    _cabi.require_cuda("compute_pck")
    out, was_numpy = _per_frame(pred, gt, alpha, "pck")
    return out.cpu().numpy() if was_numpy else out


def compute_pck_radius(gt, alpha):
    """alpha * max(|max x - min x|, |max y - min y|) per sample, tiled to [N, 52] (:17-23)."""
    _cabi.require_cuda("compute_pck_radius")
    out, was_numpy = _per_frame(gt, gt, alpha, "radius")
    out = out[:, None].expand(-1, K_JOINTS)
    return out.cpu().numpy().copy() if was_numpy else out.contiguous()
