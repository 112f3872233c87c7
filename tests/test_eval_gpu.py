"""GPU parity: fused L1/PCK kernel through the drop-in -> C ABI against the oracle and the reference
goldens.  PCK hit counts bit-exact; L1 rtol 1e-5 (D8)."""
import numpy as np
import pytest
import torch

from oracle import eval_oracle, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ev(pkg):
    return pkg.install_dropin()["motion_evaluation"]


def test_compute_pck_matches_reference_golden(ev, golden):
    g = golden["eval"]
    gt, pred = synth.gt_pose_batch(0, 4), synth.noisy_pred_batch(0, 4)
    gtf, prf = gt.reshape(-1, 2, 52), pred.reshape(-1, 2, 52)
    got = ev.compute_pck(prf, gtf, 0.2)
    assert isinstance(got, np.ndarray) and got.dtype == np.float64
    np.testing.assert_array_equal(got, g["pck_alpha02"])
    np.testing.assert_array_equal(ev.compute_pck(prf, gtf, 0.1), g["pck_alpha01"])
    np.testing.assert_array_equal(ev.compute_pck(gtf, gtf), g["pck_identity"])
    rad = ev.compute_pck_radius(gtf, 0.2)
    assert rad.shape == (256, 52) and rad.dtype == np.float32
    np.testing.assert_array_equal(rad[:, 0], g["radius_alpha02"])
    np.testing.assert_array_equal(rad, eval_oracle.pck_radius(gtf, 0.2))


def test_fused_metrics_match_oracle(ev, golden):
    for n_clips, T in ((4, 64), (37, 64), (3, 7), (1, 1)):
        gt, pred = synth.gt_pose_batch(50, n_clips, T), synth.noisy_pred_batch(50, n_clips, T)
        ref = eval_oracle.metric_partials(pred, gt, 0.2)
        acc = ev.evaluate_poses(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), 0.2)
        got = ev.read_metrics(acc)
        for k in ("pck_hits", "n_keypoints", "n_frames", "n_pose", "n_motion"):
            assert got[k] == ref[k], k
        np.testing.assert_allclose(got["abs_pose"], ref["abs_pose"], rtol=1e-12)
        np.testing.assert_allclose(got["abs_motion"], ref["abs_motion"], rtol=1e-12)
    gt, pred = synth.gt_pose_batch(0, 4), synth.noisy_pred_batch(0, 4)
    fin = ev.finalize_metrics(ev.read_metrics(ev.evaluate_poses(torch.from_numpy(pred).cuda(),
                                                                torch.from_numpy(gt).cuda())))
    np.testing.assert_allclose(fin["l1_pose"], golden["eval"]["l1_pose"], rtol=1e-5)
    np.testing.assert_allclose(fin["l1_motion"], golden["eval"]["l1_motion"], rtol=1e-5)


def test_accumulation_and_edge_cases(ev):
    gt, pred = synth.gt_pose_batch(0, 6), synth.noisy_pred_batch(0, 6)
    tg, tp = torch.from_numpy(gt).cuda(), torch.from_numpy(pred).cuda()
    whole = ev.read_metrics(ev.evaluate_poses(tp, tg))
    acc = ev.new_metrics()
    ev.evaluate_poses(tp[:2], tg[:2], accum=acc)
    ev.evaluate_poses(tp[2:], tg[2:], accum=acc)
    parts = ev.read_metrics(acc)
    assert parts["pck_hits"] == whole["pck_hits"] and parts["n_motion"] == whole["n_motion"]
    np.testing.assert_allclose(parts["abs_pose"], whole["abs_pose"], rtol=1e-13)
    empty = ev.read_metrics(ev.evaluate_poses(tp[:0], tg[:0]))
    assert empty["n_frames"] == 0 and empty["pck_hits"] == 0
    assert ev.compute_pck(np.zeros((0, 2, 52), np.float32), np.zeros((0, 2, 52), np.float32)).shape == (0,)
    # degenerate bbox: all keypoints identical -> radius 0 -> hit only where pred == gt exactly
    g0 = np.full((2, 2, 52), 3.0, np.float32)
    p0 = g0.copy(); p0[1, 0, 5] += 1e-3
    np.testing.assert_array_equal(ev.compute_pck(p0, g0), eval_oracle.pck(p0, g0))
    with pytest.raises(ValueError):
        ev.compute_pck(np.zeros((4, 2, 48), np.float32), np.zeros((4, 2, 48), np.float32))


def test_full_size_properties(ev):
    """Config-3 scale slice (16 384 clips = 1 M frames): hits are integers, shard sums equal the whole,
    pred == gt gives PCK 1 and L1 0."""
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    gt = 50 * torch.randn(16384, 64, 104, device="cuda", generator=g)
    pred = gt + 12 * torch.randn(16384, 64, 104, device="cuda", generator=g)
    whole = ev.read_metrics(ev.evaluate_poses(pred, gt))
    acc = ev.new_metrics()
    for lo in range(0, 16384, 4096):
        ev.evaluate_poses(pred[lo:lo + 4096], gt[lo:lo + 4096], accum=acc)
    parts = ev.read_metrics(acc)
    assert parts["pck_hits"] == whole["pck_hits"] and whole["n_keypoints"] == 16384 * 64 * 52
    np.testing.assert_allclose(parts["abs_pose"], whole["abs_pose"], rtol=1e-12)
    ident = ev.finalize_metrics(ev.read_metrics(ev.evaluate_poses(gt, gt)))
    assert ident["pck"] == 1.0 and ident["l1_pose"] == 0.0 and ident["l1_motion"] == 0.0
    ref = eval_oracle.metric_partials(pred[:64].cpu().numpy(), gt[:64].cpu().numpy())
    assert ev.read_metrics(ev.evaluate_poses(pred[:64], gt[:64]))["pck_hits"] == ref["pck_hits"]
