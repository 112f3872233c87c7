"""B200 drop-in for the pose-normalisation steps around the generator (reference ``normalization_tools.py`` and the
inline code of ``version5_model_train.py:296-307`` / ``generate_motion_video.py:247-260``).

The reference computes dataset statistics by iterating its HDF5 data loader (out of scope here) and then applies,
per batch: view ``[B, T, 2, 52]``, subtract the neck (joint 0) from every joint, ``(x - mean) / std``; after the
generator: ``x * std + mean``.  These element-wise steps -- and the accumulation behind ``get_mean_std_necksub``
for callers that stream batches themselves -- run as CUDA kernels of liba2m_b200 (csrc/pose_norm.cu) with the
reference's operation order, so results equal torch's CPU results bit for bit.  No CPU fallback.
"""
import torch

from . import _cabi

FEATS = 104


def _prep(pose, name):
    _cabi.require_cuda(name)
    t = torch.as_tensor(pose)
    if t.shape[-1] != FEATS:
        raise ValueError("%s: last dimension must be %d (52 x then 52 y), got %s" % (name, FEATS, tuple(t.shape)))
    return t.to(device="cuda", dtype=torch.float32).contiguous()


def _vec(v, device, name):
    v = torch.as_tensor(v).to(device=device, dtype=torch.float32).contiguous()
    if v.numel() != FEATS:
        raise ValueError("%s must have %d elements" % (name, FEATS))
    return v


def normalize_pose_necksub(pose, pose_mean, pose_std):
    """[..., 104] -> neck-subtracted, standardised poses (version5_model_train.py:300-307)."""
    x = _prep(pose, "normalize_pose_necksub")
    mean, std = _vec(pose_mean, x.device, "pose_mean"), _vec(pose_std, x.device, "pose_std")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().a2m_pose_normalize_f32(_cabi.ptr(x), _cabi.ptr(mean), _cabi.ptr(std), x.numel() // FEATS,
                                                       _cabi.ptr(out), _cabi.stream_ptr(x.device)))
    return out


def denormalize_pose(pose_norm, pose_mean, pose_std):
    """[..., 104] -> x * std + mean (generate_motion_video.py:259-260)."""
    x = _prep(pose_norm, "denormalize_pose")
    mean, std = _vec(pose_mean, x.device, "pose_mean"), _vec(pose_std, x.device, "pose_std")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().a2m_pose_denormalize_f32(_cabi.ptr(x), _cabi.ptr(mean), _cabi.ptr(std), x.numel() // FEATS,
                                                         _cabi.ptr(out), _cabi.stream_ptr(x.device)))
    return out


class PoseStats:
    """Streaming version of get_mean_std_necksub (normalization_tools.py:24-45) for equal-sized batches:
    ``update(pose_batch)`` per batch, then ``finalize() -> (pose_mean, pose_std)`` with std[0] = std[52] = 1."""

    def __init__(self, device="cuda"):
        _cabi.require_cuda("PoseStats")
        self.accum = torch.zeros(2 * FEATS + 1, dtype=torch.float64, device=device)

    def update(self, pose):
        x = _prep(pose, "PoseStats.update")
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().a2m_pose_stats_f64(_cabi.ptr(x), x.numel() // FEATS, _cabi.ptr(self.accum),
                                                       _cabi.stream_ptr(x.device)))
        return self

    def finalize(self):
        a = self.accum.cpu()
        n = max(float(a[2 * FEATS]), 1.0)
        mean = a[:FEATS] / n
        std = (a[FEATS:2 * FEATS] / n - mean ** 2).clamp_min(0.0) ** 0.5
        std[0] = 1.0
        std[52] = 1.0
        return mean.float(), std.float()
