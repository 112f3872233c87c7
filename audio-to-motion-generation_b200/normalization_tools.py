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
    dev = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())    # a GPU input stays on its own GPU
    return t.to(device=dev, dtype=torch.float32).contiguous()


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
    ``update(pose_batch)`` per batch, then ``finalize() -> (pose_mean, pose_std)`` with std[0] = std[52] = 1.
    ``neck_sub=False`` gives get_mean_std (:5-20): no neck subtraction and no std override."""

    def __init__(self, device=None, neck_sub=True):
        _cabi.require_cuda("PoseStats")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.neck_sub = bool(neck_sub)
        self.accum = torch.zeros(2 * FEATS + 1, dtype=torch.float64, device=self.device)

    def update(self, pose):
        t = torch.as_tensor(pose)
        x = _prep(t if t.is_cuda else t.to(self.device), "PoseStats.update")
        if x.device != self.accum.device:
            raise ValueError("PoseStats lives on %s, the batch on %s" % (self.accum.device, x.device))
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.lib().a2m_pose_stats_ex_f64(_cabi.ptr(x), x.numel() // FEATS, int(self.neck_sub),
                                                          _cabi.ptr(self.accum), _cabi.stream_ptr(x.device)))
        return self

    def finalize(self):
        a = self.accum.cpu()
        n = max(float(a[2 * FEATS]), 1.0)
        mean = a[:FEATS] / n
        std = (a[FEATS:2 * FEATS] / n - mean ** 2).clamp_min(0.0) ** 0.5
        if self.neck_sub:
            std[0] = 1.0
            std[52] = 1.0
        return mean.float(), std.float()


def _dataset_stats(dataloader, neck_sub):
    """The reference's loop: the MEAN over batches of each batch's own mean and mean square (so a ragged last batch
    weighs as much as a full one, exactly as in normalization_tools.py:8-17,28-41), finished in fp32 like the
    reference.  Each batch's sums come from the CUDA statistics kernel (fp64 accumulation of fp32 values)."""
    _cabi.require_cuda("get_mean_std")
    mean_sum = sq_sum = None
    n_batches = 0
    for n_batches, batch in enumerate(dataloader.train, 1):
        stats = PoseStats(neck_sub=neck_sub).update(batch["pose/data"])
        a = stats.accum
        n = a[2 * FEATS].clamp_min(1.0)
        m, q = (a[:FEATS] / n).float(), (a[FEATS:2 * FEATS] / n).float()
        mean_sum = m if mean_sum is None else mean_sum + m
        sq_sum = q if sq_sum is None else sq_sum + q
    if n_batches == 0:
        raise ValueError("dataloader.train yielded no batches")
    pose_mean = (mean_sum / n_batches).cpu()
    pose_std = ((sq_sum / n_batches).cpu() - pose_mean ** 2) ** 0.5
    return pose_mean, pose_std


def get_mean_std(dataloader):
    """normalization_tools.py:5-20: (pose_mean [104], pose_std [104]) over ``dataloader.train`` batches
    (dicts carrying ``'pose/data'`` [B, T, 104])."""
    return _dataset_stats(dataloader, neck_sub=False)


def get_mean_std_necksub(dataloader):
    """normalization_tools.py:24-45: the same on neck-subtracted poses, std of the two neck features set to 1."""
    pose_mean, pose_std = _dataset_stats(dataloader, neck_sub=True)
    pose_std[0] = 1.
    pose_std[52] = 1.
    return pose_mean, pose_std
