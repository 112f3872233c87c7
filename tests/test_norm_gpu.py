"""GPU parity of the pose (de)normalisation kernels (csrc/pose_norm.cu) against the oracle restatement of
version5_model_train.py:296-307 / generate_motion_video.py:247-260 / normalization_tools.py:24-45: bit-exact for
the element-wise transforms, fp64-accumulated statistics within 1e-6 of the reference's fp32 batch means."""
import numpy as np
import pytest
import torch

from oracle import norm_oracle, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nt(pkg):
    return pkg.install_dropin()["normalization_tools"]


def test_normalize_and_denormalize_bit_exact(nt):
    pose = torch.from_numpy(synth.gt_pose_batch(0, 5))                   # [5, 64, 104]
    g = torch.Generator().manual_seed(0)
    mean = torch.randn(104, generator=g)
    std = 0.5 + torch.rand(104, generator=g)
    got = nt.normalize_pose_necksub(pose.cuda(), mean, std).cpu()
    ref = norm_oracle.normalize_necksub(pose, mean, std)
    assert got.shape == ref.shape and torch.equal(got, ref)
    assert torch.all(got.reshape(5, 64, 2, 52)[..., 0] == ((0 - mean.reshape(2, 52)[:, 0]) / std.reshape(2, 52)[:, 0]))
    back = nt.denormalize_pose(got.cuda(), mean, std).cpu()
    assert torch.equal(back, norm_oracle.denormalize(ref, mean, std))
    assert nt.normalize_pose_necksub(torch.zeros(0, 64, 104), mean, std).shape == (0, 64, 104)
    with pytest.raises(ValueError):
        nt.normalize_pose_necksub(torch.zeros(3, 64, 100), mean, std)


def test_streaming_statistics_match_reference(nt):
    batches = [torch.from_numpy(synth.gt_pose_batch(8 * i, 8)) for i in range(4)]
    stats = nt.PoseStats()
    for b in batches:
        stats.update(b.cuda())
    mean, std = stats.finalize()
    ref_mean, ref_std = norm_oracle.mean_std_necksub(batches)
    np.testing.assert_allclose(mean.numpy(), ref_mean.numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(std.numpy(), ref_std.numpy(), rtol=1e-5, atol=1e-4)
    assert std[0] == 1.0 and std[52] == 1.0


def test_get_mean_std_dropins_follow_the_reference_loop(nt):
    """normalization_tools.get_mean_std / get_mean_std_necksub(dataloader): the reference's names and loop (mean over
    batches of per-batch means, so the ragged last batch counts as much as a full one); fp64-accumulated sums vs the
    reference's fp32 torch.mean: rtol 1e-5 / atol 1e-4 as above."""
    class Loader:                                                        # what the reference iterates: dataloader.train
        def __init__(self, batches):
            self.train = [{"pose/data": b} for b in batches]
    batches = [torch.from_numpy(synth.gt_pose_batch(8 * i, 8)) for i in range(3)] + \
              [torch.from_numpy(synth.gt_pose_batch(100, 3))]            # ragged tail
    mean, std = nt.get_mean_std_necksub(Loader(batches))
    ref_mean, ref_std = norm_oracle.mean_std_necksub(batches)
    assert mean.shape == (104,) and mean.dtype == torch.float32 and not mean.is_cuda
    np.testing.assert_allclose(mean.numpy(), ref_mean.numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(std.numpy(), ref_std.numpy(), rtol=1e-5, atol=1e-4)
    assert std[0] == 1.0 and std[52] == 1.0
    mean, std = nt.get_mean_std(Loader([b.cuda() for b in batches]))     # device-resident batches work too
    ref_mean, ref_std = norm_oracle.mean_std(batches)
    np.testing.assert_allclose(mean.numpy(), ref_mean.numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(std.numpy(), ref_std.numpy(), rtol=1e-5, atol=1e-4)
    assert std[0] != 1.0
    with pytest.raises(ValueError):
        nt.get_mean_std(Loader([]))


def test_fused_output_denorm_equals_separate_pass(pkg, nt):
    """SelfAttention_G.set_output_denorm: the forward's output pass applies pose * std + mean; bit-equal to the
    separate kernel (and to the oracle's torch restatement) on the same normalised poses; losses unchanged;
    survives a repack; off again with set_output_denorm()."""
    from oracle import weights
    mods = pkg.install_dropin()
    model = mods["real_motion_model"].SelfAttention_G().cuda().eval()
    model.load_state_dict(weights.make_state_dict(3, "stress"))
    g = torch.Generator().manual_seed(11)
    x = (-1.5 + 1.5 * torch.randn(3, 64, 64, generator=g)).cuda()
    mean, std = torch.randn(104, generator=g) * 40 + 300, 5 + 20 * torch.rand(104, generator=g)
    plain, losses0 = model(x)
    model.set_output_denorm(mean, std)
    fused, losses1 = model(x)
    assert torch.equal(fused, nt.denormalize_pose(plain, mean, std))
    assert torch.equal(fused.cpu(), norm_oracle.denormalize(plain.cpu(), mean, std))
    assert torch.equal(losses0[0], losses1[0])
    model.repack()                                      # a new native handle must receive the setting again
    assert torch.equal(model(x)[0], fused)
    model.set_output_denorm()
    assert torch.equal(model(x)[0], plain)
    with pytest.raises(ValueError):
        model.set_output_denorm(mean, None)
    with pytest.raises(ValueError):
        model.set_output_denorm(mean[:50], std[:50])
