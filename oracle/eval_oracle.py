"""Oracle (test infrastructure): numpy restatement of the reference evaluation metrics.

Follows ``/root/reference/motion_evaluation.py``:
  pck_radius <- compute_pck_radius() :17-23
  pck        <- compute_pck()        :4-14
and the L1 metric of ``version5_model_train.py`` (``nn.L1Loss()`` :264, applied to
``pos_to_motion`` = first difference along time :208-213, used at :367,467).

dtype behaviour is the reference's: fp32 in -> fp32 distances / radius / compare
(numpy keeps float32 through ``max``, ``abs``, ``* alpha`` [python float is a weak scalar]
and ``linalg.norm``), float64 per-frame mean out.
"""
import numpy as np

K_JOINTS = 52     # hard-coded tile width, motion_evaluation.py:22


def pck_radius(gt, alpha):
    """alpha * max(bbox width, bbox height) per frame, broadcast to the 52 keypoints."""
    xs, ys = gt[:, 0, :], gt[:, 1, :]
    width = np.abs(xs.max(axis=1) - xs.min(axis=1))
    height = np.abs(ys.max(axis=1) - ys.min(axis=1))
    side = np.maximum(width, height)
    return np.repeat(side[:, None], K_JOINTS, axis=1) * alpha


def pck_hits(pred, gt, alpha=0.2):
    """Boolean [N, K] hit mask: ||gt - pred||_2 <= radius (motion_evaluation.py:12)."""
    d = gt - pred
    # np.linalg.norm(axis=2) on [N,K,2] == sqrt(dx*dx + dy*dy) evaluated in the input dtype
    dist = np.sqrt(d[:, 0, :] * d[:, 0, :] + d[:, 1, :] * d[:, 1, :])
    return dist <= pck_radius(gt, alpha)


def pck(pred, gt, alpha=0.2):
    """Per-frame mean hit rate, float64 [N] (motion_evaluation.py:13)."""
    return np.mean(pck_hits(pred, gt, alpha), axis=1)


def poses_as_frames(pose):
    """[B,T,104] -> [B*T,2,52]: x block then y block (version5_model_train.py:301; D7)."""
    pose = np.asarray(pose)
    return pose.reshape(-1, 2, K_JOINTS)


def l1(a, b):
    """nn.L1Loss(): mean |a-b| over all elements; fp64 accumulation here."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.mean(np.abs(a.astype(np.float64) - b.astype(np.float64))))


def motion(pose):
    """pos_to_motion: first difference along time (dim 1), in the input dtype."""
    return np.diff(np.asarray(pose), n=1, axis=1)


def metric_partials(pred, gt, alpha=0.2):
    """The reducible partial sums the GPU path all-reduces:
    integer hit / keypoint / frame counts, fp64 abs-sums for pose- and motion-L1."""
    pred = np.asarray(pred, dtype=np.float32)
    gt = np.asarray(gt, dtype=np.float32)
    hits = pck_hits(poses_as_frames(pred), poses_as_frames(gt), alpha)
    mp, mg = motion(pred), motion(gt)
    return dict(
        pck_hits=int(hits.sum()),
        n_keypoints=int(hits.size),
        n_frames=int(hits.shape[0]),
        abs_pose=float(np.abs(pred - gt).astype(np.float64).sum()),      # fp32 |a-b|, fp64 sum
        n_pose=int(pred.size),
        abs_motion=float(np.abs(mp - mg).astype(np.float64).sum()),
        n_motion=int(mp.size),
    )


def finalize(partials):
    """hits/keypoints, sum/n -- what every rank computes after the all-reduce."""
    return dict(
        pck=partials["pck_hits"] / max(partials["n_keypoints"], 1),
        l1_pose=partials["abs_pose"] / max(partials["n_pose"], 1),
        l1_motion=partials["abs_motion"] / max(partials["n_motion"], 1),
    )


def smoothness(motion_seq):
    """compute_temporal_smoothness_loss (version5_model_train.py:216-230): mean over (b, t) of the L2 norm over the
    features of the second difference; fp32 differences, fp64 norm accumulation here."""
    m = np.asarray(motion_seq, dtype=np.float32)
    acc = m[:, 1:] - m[:, :-1]
    return float(np.mean(np.sqrt(np.sum(acc.astype(np.float64) ** 2, axis=-1))))


def jerk(motion_seq):
    """compute_jerk_loss (version5_model_train.py:233-248): the same on the third difference."""
    m = np.asarray(motion_seq, dtype=np.float32)
    acc = m[:, 1:] - m[:, :-1]
    j = acc[:, 1:] - acc[:, :-1]
    return float(np.mean(np.sqrt(np.sum(j.astype(np.float64) ** 2, axis=-1))))
