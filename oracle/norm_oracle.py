"""Oracle (test infrastructure): the reference's pose (de)normalisation, restated with torch CPU ops in the same
order as version5_model_train.py:296-307, generate_motion_video.py:247-260 and normalization_tools.py:24-45."""
import torch


def normalize_necksub(pose, pose_mean, pose_std):
    pose = torch.as_tensor(pose, dtype=torch.float32)
    p = pose.reshape(-1, 2, 52)                                    # x block, y block (version5_model_train.py:301)
    neck = p[:, :, 0].reshape(-1, 2, 1)                            # joint 0 of each block (:302)
    p = torch.sub(p, neck).reshape(pose.shape)                     # :303-304
    return torch.div(torch.sub(p, pose_mean), pose_std)            # :305-306


def denormalize(pose_norm, pose_mean, pose_std):
    return torch.add(torch.mul(torch.as_tensor(pose_norm, dtype=torch.float32), pose_std), pose_mean)


def mean_std_necksub(batches):
    """get_mean_std_necksub over an iterable of [B, T, 104] batches (normalization_tools.py:24-45)."""
    s = torch.zeros(104)
    q = torch.zeros(104)
    n = 0
    for n, pose in enumerate(batches, 1):
        pose = torch.as_tensor(pose, dtype=torch.float32)
        pose = pose.reshape(pose.shape[0], pose.shape[1], 2, -1)
        neck = pose[:, :, :, 0].reshape(pose.shape[0], pose.shape[1], 2, 1)
        pose = torch.sub(pose, neck).reshape(pose.shape[0], pose.shape[1], -1)
        s += torch.mean(pose, dim=[0, 1])
        q += torch.mean(pose ** 2, dim=[0, 1])
    mean = s / n
    std = (q / n - mean ** 2) ** 0.5
    std[0] = 1.
    std[52] = 1.
    return mean, std


def mean_std(batches):
    """get_mean_std over an iterable of [B, T, 104] batches (normalization_tools.py:5-20): no neck subtraction."""
    s = torch.zeros(104)
    q = torch.zeros(104)
    n = 0
    for n, pose in enumerate(batches, 1):
        pose = torch.as_tensor(pose, dtype=torch.float32)
        s += torch.mean(pose, dim=[0, 1])
        q += torch.mean(pose ** 2, dim=[0, 1])
    mean = s / n
    std = (q / n - mean ** 2) ** 0.5
    return mean, std
