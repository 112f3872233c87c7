"""GPU parity of the discriminator drop-in (SelfAttention_D through a2m_disc_forward) against the unmodified reference
class (tests/golden/disc_reference.npz) and the oracle restatement.

Tolerance: bf16 operands, fp32 accumulation.  The score is a 4096 x 3-term dot product whose terms largely cancel, so
the bar is on the absolute error relative to the scale of the scores: max|a - b| <= 2e-2 * max|b| + 2e-3."""
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle, weights
from oracle.make_golden import DISC_CASES, real_pose_input

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(got, ref):
    got, ref = got.detach().cpu().double().numpy(), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-2 * np.abs(ref).max() + 2e-3, (np.abs(got - ref).max(), np.abs(ref).max())


@pytest.fixture(scope="module")
def disc_cls(pkg):
    return pkg.install_dropin()["real_motion_model"].SelfAttention_D


def test_discriminator_matches_reference_golden(disc_cls):
    g = np.load(os.path.join(ROOT, "tests", "golden", "disc_reference.npz"))
    d = disc_cls().cuda().eval()
    for name, seed, B, T in DISC_CASES:
        d.load_state_dict(weights.make_state_dict(seed, "stress", discriminator=True))
        y, losses = d(real_pose_input(seed, B, T).cuda())
        assert losses == [] and y.is_cuda and y.dtype == torch.float32
        close(y, g[name])


def test_discriminator_matches_oracle_other_shapes(disc_cls):
    sd = weights.make_state_dict(21, "stress", discriminator=True)
    d = disc_cls().cuda().eval()
    d.load_state_dict(sd)
    for B, T in ((5, 64), (256, 64), (2, 100), (3, 38)):           # 38: odd lengths at every stride-2 stage
        pose = real_pose_input(30 + T, B, T)
        y, _ = d(pose.cuda())
        ref = model_oracle.discriminator_forward(sd, pose)
        assert y.shape == ref.shape
        close(y, ref.numpy())
    big, _ = d(real_pose_input(94, 256, 64).cuda())                # a clip's score does not depend on its batch
    one, _ = d(real_pose_input(94, 256, 64)[7:9].cuda())
    assert torch.equal(one, big[7:9])


def test_discriminator_contracts(disc_cls):
    d = disc_cls().cuda()
    with pytest.raises(RuntimeError, match="eval"):
        d(torch.zeros(1, 64, 104, device="cuda"))
    d.eval()
    with pytest.raises(NotImplementedError):
        d(torch.zeros(1, 64, 104, device="cuda"), audio=torch.zeros(1, 256, 64, device="cuda"))
    with pytest.raises(ValueError):
        d(torch.zeros(1, 64, 100, device="cuda"))
    with pytest.raises(ValueError):
        d(torch.zeros(1, 20, 104, device="cuda"))                  # 20 -> 10 -> 9 -> 4 -> 3 -> 1 -> 0: too short
    with pytest.raises(NotImplementedError):
        disc_cls(groups=2)
    assert d(torch.zeros(0, 64, 104, device="cuda"))[0].shape == (0, 4)
