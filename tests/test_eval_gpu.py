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


def test_compute_pck_float64_inputs_match_reference_golden(ev, golden):
    """The reference computes in the dtype of its inputs (motion_evaluation.py:11-23); golden `pck_alpha02_f64` is
    compute_pck on the float64 casts of the same poses.  fp64 kernel instantiation: bit-exact per-frame rates."""
    g = golden["eval"]
    gt, pred = synth.gt_pose_batch(0, 4), synth.noisy_pred_batch(0, 4)
    gtf, prf = gt.reshape(-1, 2, 52).astype(np.float64), pred.reshape(-1, 2, 52).astype(np.float64)
    got = ev.compute_pck(prf, gtf, 0.2)
    assert got.dtype == np.float64
    np.testing.assert_array_equal(got, g["pck_alpha02_f64"])
    np.testing.assert_array_equal(got, eval_oracle.pck(prf, gtf, 0.2))
    # inputs that are exactly representable in fp32 but sit on the radius in fp64 vs fp32 arithmetic: the two widths
    # are separate code paths, each equal to numpy in its own width
    rng = np.random.default_rng(3)
    g64 = rng.normal(0, 50, (512, 2, 52))
    p64 = g64 + rng.normal(0, 12, g64.shape)
    np.testing.assert_array_equal(ev.compute_pck(p64, g64, 0.1), eval_oracle.pck(p64, g64, 0.1))
    np.testing.assert_array_equal(ev.compute_pck(p64.astype(np.float32), g64.astype(np.float32), 0.1),
                                  eval_oracle.pck(p64.astype(np.float32), g64.astype(np.float32), 0.1))
    rad = ev.compute_pck_radius(g64, 0.2)
    assert rad.dtype == np.float64 and rad.shape == (512, 52)
    np.testing.assert_array_equal(rad, eval_oracle.pck_radius(g64, 0.2))
    # torch float64 on the device -> torch float64 out, fused partial sums in fp64
    t = ev.compute_pck(torch.from_numpy(p64).cuda(), torch.from_numpy(g64).cuda(), 0.2)
    assert t.is_cuda and t.dtype == torch.float64
    np.testing.assert_array_equal(t.cpu().numpy(), eval_oracle.pck(p64, g64, 0.2))


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


def test_smoothness_and_jerk_match_reference_golden(ev, golden):
    """version5_model_train.py:216-248 through the drop-in names; fp32 norms, fp64 sums: rtol 1e-6 (written here)."""
    from oracle.make_golden import SMOOTH_CASES
    g = golden["smooth"]
    for name, first, n in SMOOTH_CASES:
        pose = torch.from_numpy(synth.noisy_pred_batch(first, n)).cuda()
        motion = ev.pos_to_motion(pose)
        s, j = ev.compute_temporal_smoothness_loss(motion), ev.compute_jerk_loss(motion)
        assert isinstance(s, torch.Tensor) and s.dim() == 0 and s.dtype == torch.float32 and s.is_cuda
        np.testing.assert_allclose(s.item(), g[name + "_smoothness"], rtol=1e-6)
        np.testing.assert_allclose(j.item(), g[name + "_jerk"], rtol=1e-6)
        fused = ev.read_smoothness(ev.evaluate_smoothness(pose, from_pose=True))     # poses in, motion on the fly
        np.testing.assert_allclose(fused["smoothness"], g[name + "_smoothness"], rtol=1e-6)
        np.testing.assert_allclose(fused["jerk"], g[name + "_jerk"], rtol=1e-6)
        assert fused["n_accel"] == n * 62 and fused["n_jerk"] == n * 61
    short = synth.noisy_pred_batch(9, 2)[:, :4]
    got = ev.compute_jerk_loss(np.diff(short, axis=1))                                # numpy in -> numpy scalar out
    assert isinstance(got, np.floating)
    np.testing.assert_allclose(got, g["short_jerk"], rtol=1e-6)


def test_smoothness_shapes_and_accumulation(ev):
    """Partial sums add up over shards; odd lengths / feature counts; too-short sequences give the reference's NaN."""
    pose = synth.noisy_pred_batch(3, 37, 29)
    m = eval_oracle.motion(pose)
    whole = ev.read_smoothness(ev.evaluate_smoothness(torch.from_numpy(m).cuda()))
    np.testing.assert_allclose(whole["smoothness"], eval_oracle.smoothness(m), rtol=1e-6)
    np.testing.assert_allclose(whole["jerk"], eval_oracle.jerk(m), rtol=1e-6)
    acc = ev.new_smoothness()
    for lo, hi in ((0, 5), (5, 6), (6, 37)):
        ev.evaluate_smoothness(torch.from_numpy(m[lo:hi]).cuda(), accum=acc)
    parts = ev.read_smoothness(acc)
    assert parts["n_accel"] == whole["n_accel"] and parts["n_jerk"] == whole["n_jerk"]
    np.testing.assert_allclose(parts["sum_accel_norm"], whole["sum_accel_norm"], rtol=1e-12)
    np.testing.assert_allclose(parts["sum_jerk_norm"], whole["sum_jerk_norm"], rtol=1e-12)
    g = torch.Generator().manual_seed(5)
    odd = torch.randn(3, 11, 7, generator=g)
    got = ev.read_smoothness(ev.evaluate_smoothness(odd.cuda()))
    np.testing.assert_allclose(got["smoothness"], eval_oracle.smoothness(odd.numpy()), rtol=1e-6)
    np.testing.assert_allclose(got["jerk"], eval_oracle.jerk(odd.numpy()), rtol=1e-6)
    two = ev.read_smoothness(ev.evaluate_smoothness(odd[:, :2].cuda()))             # 2 velocities: 1 accel, no jerk
    assert two["n_accel"] == 3 and two["n_jerk"] == 0 and np.isnan(two["jerk"])
    one = ev.read_smoothness(ev.evaluate_smoothness(odd[:, :1].cuda()))
    assert one["n_accel"] == 0 and np.isnan(one["smoothness"])
    with pytest.raises(ValueError):
        ev.evaluate_smoothness(torch.zeros(4, 5).cuda())
    with pytest.raises(Exception):
        ev.evaluate_smoothness(torch.zeros(2, 5, 200).cuda())
