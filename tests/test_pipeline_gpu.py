"""GPU test of the composed hot path (pipeline.AudioToPosePipeline: mel -> adapter D2 -> SelfAttention_G ->
L1/PCK partials) against the oracle composition, and of the stream-lane pipelining (results must not depend on
the number of lanes or on how clips are batched)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import eval_oracle, mel_oracle, model_oracle, synth, weights

pytestmark = pytest.mark.gpu
PKG = "audio-to-motion-generation_b200"


@pytest.fixture(scope="module")
def setup(pkg):
    mods = pkg.install_dropin()
    pipeline = importlib.import_module(PKG + ".pipeline")
    sd = weights.make_state_dict(0, "stress")
    model = mods["real_motion_model"].SelfAttention_G().cuda().eval()
    model.load_state_dict(sd)
    return pipeline, model, sd


def run(pipeline, model, lanes, batches, graphs=False):
    pipe = pipeline.AudioToPosePipeline(model, lanes=lanes, graphs=graphs)
    poses = []
    for w, g in batches:
        p = pipe.step(torch.from_numpy(w).cuda(), torch.from_numpy(g).cuda())
        if graphs:
            pipe.sync_lanes()                   # graph lanes hand back their static output buffer: copy it out now
        poses.append(p.clone() if graphs else p)
    out = pipe.finish()
    return out, [p.cpu() for p in poses]


def test_composed_path_matches_oracle(setup):
    pipeline, model, sd = setup
    wav, gt = synth.wav_batch(0, 3), synth.gt_pose_batch(0, 3)
    out, poses = run(pipeline, model, 1, [(wav, gt)])
    mel = mel_oracle.log_mel_batch(wav)
    x = torch.from_numpy(synth.adapter(mel).astype(np.float32))
    ref_pose, _ = model_oracle.generator_forward(sd, x)
    rel = ((poses[0] - ref_pose).abs().sum() / ref_pose.abs().sum()).item()
    assert rel <= 1e-2, rel
    # the metrics of the GPU poses, evaluated by the oracle: hit counts bit-exact, sums to fp64 round-off
    ref = eval_oracle.metric_partials(poses[0].numpy(), gt)
    assert out["pck_hits"] == ref["pck_hits"] and out["n_keypoints"] == ref["n_keypoints"] == 3 * 64 * 52
    np.testing.assert_allclose(out["abs_pose"], ref["abs_pose"], rtol=1e-9)
    np.testing.assert_allclose(out["abs_motion"], ref["abs_motion"], rtol=1e-9)
    fin = eval_oracle.finalize(ref)
    np.testing.assert_allclose([out["pck"], out["l1_pose"], out["l1_motion"]],
                               [fin["pck"], fin["l1_pose"], fin["l1_motion"]], rtol=1e-9)


def test_lanes_and_batching_do_not_change_results(setup):
    pipeline, model, _ = setup
    wav, gt = synth.wav_batch(10, 8), synth.gt_pose_batch(10, 8)
    whole, p1 = run(pipeline, model, 1, [(wav, gt)])
    split = [(wav[i:i + 2], gt[i:i + 2]) for i in range(0, 8, 2)]
    for lanes in (1, 2, 3):
        out, poses = run(pipeline, model, lanes, split)
        assert torch.equal(torch.cat(poses), p1[0])
        assert out["pck_hits"] == whole["pck_hits"] and out["n_frames"] == whole["n_frames"] == 8 * 64
        np.testing.assert_allclose(out["abs_pose"], whole["abs_pose"], rtol=1e-12)
        np.testing.assert_allclose(out["abs_motion"], whole["abs_motion"], rtol=1e-12)


def test_cuda_graph_replay_equals_eager(setup):
    pipeline, model, _ = setup
    wav, gt = synth.wav_batch(30, 8), synth.gt_pose_batch(30, 8)
    split = [(wav[i:i + 2], gt[i:i + 2]) for i in range(0, 8, 2)]
    eager, p_eager = run(pipeline, model, 1, split)
    for lanes in (1, 2):
        out, poses = run(pipeline, model, lanes, split, graphs=True)
        assert torch.equal(torch.cat(poses), torch.cat(p_eager))
        assert out["pck_hits"] == eager["pck_hits"] and out["n_frames"] == eager["n_frames"]
        np.testing.assert_allclose(out["abs_pose"], eager["abs_pose"], rtol=1e-12)


def test_host_batches_end_to_end(setup):
    pipeline, model, _ = setup
    wav, gt = synth.wav_batch(20, 4), synth.gt_pose_batch(20, 4)
    ref, _ = run(pipeline, model, 1, [(wav, gt)])
    pipe = pipeline.AudioToPosePipeline(model, lanes=2)
    host = [(torch.from_numpy(wav[i:i + 1]).pin_memory(), torch.from_numpy(gt[i:i + 1]).pin_memory()) for i in range(4)]
    assert pipe.run_host_batches(host) == 4
    out = pipe.finish()
    assert out["pck_hits"] == ref["pck_hits"]
    np.testing.assert_allclose(out["abs_pose"], ref["abs_pose"], rtol=1e-12)


def test_long_form_sliding_windows(setup):
    """BASELINE config 4 (streaming mel + sliding-window generation) at a reduced length: the windows are strided
    views of one log-mel per clip; results must equal the oracle composition window by window."""
    pipeline, model, sd = setup
    n = 16000 * 6 + 123                                     # 6 s -> 598 frames -> 8 windows (hop 30 frames)
    wav = np.stack([synth.wav_clip(40 + i, n) for i in range(2)])
    pipe = pipeline.AudioToPosePipeline(model, lanes=1)
    poses = pipe.generate_long(torch.from_numpy(wav).cuda()).cpu()
    mel = mel_oracle.log_mel_batch(wav)
    starts = pipeline.window_starts(mel.shape[1])
    assert starts == list(range(0, mel.shape[1] - 384, 30)) and poses.shape == (2, len(starts), 64, 104)
    for b in range(2):
        x = np.stack([mel[b, s:s + 384:6] for s in starts]).astype(np.float32)
        ref, _ = model_oracle.generator_forward(sd, torch.from_numpy(x))
        rel = ((poses[b] - ref).abs().sum() / ref.abs().sum()).item()
        assert rel <= 1e-2, rel
    # the strided view must give exactly what a gathered copy gives
    lm = pipeline.audio_repr.log_mel_spectograms(torch.from_numpy(wav[:1]).cuda())[0]
    view = pipeline.sliding_windows(lm)
    gathered = torch.stack([lm[s:s + 384:6] for s in starts]).contiguous()
    assert torch.equal(view, gathered)
    assert torch.equal(model(view)[0], model(gathered)[0])
    # one launch program over all streams x windows gives, clip for clip, what a forward per clip gives
    assert torch.equal(poses[0], model(view)[0].cpu())


def test_long_form_full_size_shape(setup):
    """60 s streams: 960 000 samples -> 5 998 frames -> 188 windows per clip."""
    pipeline, model, _ = setup
    assert len(pipeline.window_starts(5998)) == 188
    wav = 0.1 * torch.randn(2, 960000, device="cuda")
    poses = pipeline.AudioToPosePipeline(model, lanes=1).generate_long(wav)
    assert poses.shape == (2, 188, 64, 104) and torch.isfinite(poses).all()


def test_pipeline_smoothness_metrics(setup):
    """Optional validation-loop metrics (version5_model_train.py:456-459) of the generated poses, accumulated over
    batches and lanes: equal to the oracle's smoothness / jerk of the same GPU poses."""
    pipeline, model, _ = setup
    wav, gt = synth.wav_batch(30, 6), synth.gt_pose_batch(30, 6)
    pipe = pipeline.AudioToPosePipeline(model, lanes=2, smoothness=True)
    poses = [pipe.step(torch.from_numpy(wav[i:i + 2]).cuda(), torch.from_numpy(gt[i:i + 2]).cuda()) for i in (0, 2, 4)]
    out = pipe.finish()
    m = eval_oracle.motion(torch.cat(poses).cpu().numpy())
    assert out["n_accel"] == 6 * 62 and out["n_jerk"] == 6 * 61
    np.testing.assert_allclose(out["smoothness"], eval_oracle.smoothness(m), rtol=1e-6)
    np.testing.assert_allclose(out["jerk"], eval_oracle.jerk(m), rtol=1e-6)
    pipe.reset()
    assert int(pipe.smooth.abs().sum()) == 0


def test_pats_front_end_f128_matches_oracle_composition(setup):
    """The shipped configuration's features: log_mel_512 (128 bands, pats/data_loading/audio.py:58-79) -> stride-6
    adapter -> SelfAttention_G with F = 128.  Poses within the bf16 bar of the oracle composition."""
    from oracle import pats_oracle
    pipeline, model, sd = setup
    n = 380 * 512                                                        # 381 frames >= the adapter's 379
    wav = np.stack([synth.wav_clip(70 + i, n) for i in range(2)])
    gt = synth.gt_pose_batch(70, 2)
    pipe = pipeline.AudioToPosePipeline(model, lanes=1, front_end="log_mel_512", sample_rate=44100)
    pose = pipe.step(torch.from_numpy(wav).cuda(), torch.from_numpy(gt).cuda())
    out = pipe.finish()
    mel = np.stack([pats_oracle.log_mel_512(w, 44100) for w in wav])
    assert mel.shape == (2, 381, 128)
    x = torch.from_numpy(synth.adapter(mel).astype(np.float32))
    assert x.shape == (2, 64, 128)
    ref_pose, _ = model_oracle.generator_forward(sd, x)
    rel = ((pose.cpu() - ref_pose).abs().sum() / ref_pose.abs().sum()).item()
    assert rel <= 1e-2, rel
    ref = eval_oracle.metric_partials(pose.cpu().numpy(), gt)
    assert out["pck_hits"] == ref["pck_hits"] and out["n_frames"] == 128
    with pytest.raises(ValueError):
        pipeline.AudioToPosePipeline(model, front_end="mfcc")


def test_adapter_frames_only_is_bit_identical(setup):
    """Computing only the 64 log-mel frames the adapter picks (hop 960 samples) gives the same frames, poses and
    metrics as computing all 425 and slicing."""
    pipeline, model, _ = setup
    wav, gt = synth.wav_batch(80, 4), synth.gt_pose_batch(80, 4)
    mods = importlib.import_module(PKG).install_dropin()
    lm = mods["pose_video.audio_repr"].log_mel_spectograms
    w = torch.from_numpy(wav).cuda()
    dense = lm(w)[:, 0:384:6, :]
    sparse = lm(w, hop_length_secs=0.06)[:, :64, :]
    assert torch.equal(dense, sparse)
    full, p_full = run(pipeline, model, 1, [(wav, gt)])
    pipe = pipeline.AudioToPosePipeline(model, lanes=1, adapter_frames_only=True)
    p_sparse = pipe.step(w, torch.from_numpy(gt).cuda())
    out = pipe.finish()
    assert torch.equal(p_sparse.cpu(), p_full[0])
    assert out["pck_hits"] == full["pck_hits"] and out["abs_pose"] == full["abs_pose"]


def test_lanes_follow_later_changes_of_the_model(pkg):
    """The stream lanes are native handles of ONE module: load_state_dict, an in-place weight edit and
    set_output_denorm made after the pipeline was built reach every lane (alternating batches would otherwise run on
    stale weights)."""
    mods = pkg.install_dropin()
    pipeline = importlib.import_module(PKG + ".pipeline")
    model = mods["real_motion_model"].SelfAttention_G().cuda().eval()
    model.load_state_dict(weights.make_state_dict(0, "stress"))
    pipe = pipeline.AudioToPosePipeline(model, lanes=2)
    wav = torch.from_numpy(synth.wav_batch(60, 2)).cuda()
    gt = torch.from_numpy(synth.gt_pose_batch(60, 2)).cuda()

    def two_lanes(p):
        x, y = p.step(wav, gt), p.step(wav, gt)                          # lane 0, lane 1
        p.finish()                                                       # the poses are complete after the lanes are joined
        return x.clone(), y.clone()

    a0, a1 = two_lanes(pipe)
    assert torch.equal(a0, a1)
    model.load_state_dict(weights.make_state_dict(5, "stress"))          # new weights after construction
    b0, b1 = two_lanes(pipe)
    assert torch.equal(b0, b1) and not torch.equal(b0, a0)
    fresh = mods["real_motion_model"].SelfAttention_G().cuda().eval()
    fresh.load_state_dict(weights.make_state_dict(5, "stress"))
    assert torch.equal(b0, two_lanes(pipeline.AudioToPosePipeline(fresh, lanes=1))[0])
    with torch.no_grad():
        model.body_logits.bias.add_(1.0)                                 # in-place edit
    c0, c1 = two_lanes(pipe)
    assert torch.equal(c0, c1) and torch.allclose(c0[..., :20], b0[..., :20] + 1.0, atol=1e-5)
    assert torch.equal(c0[..., 20:], b0[..., 20:])
    mean, std = torch.full((104,), 3.0), torch.full((104,), 2.0)
    model.set_output_denorm(mean, std)
    d0, d1 = two_lanes(pipe)
    assert torch.equal(d0, d1) and torch.equal(d0, c0 * 2.0 + 3.0)
    model.set_output_denorm()


def test_int16_pcm_batches_end_to_end(setup):
    """Pinned int16 host batches (half the H2D bytes) give the metrics of the same samples fed as fp32."""
    pipeline, model, _ = setup
    pcm = np.stack([synth.wav_clip(70 + i, synth.CLIP_SAMPLES, "int16").astype(np.int16) for i in range(4)])
    gt = synth.gt_pose_batch(70, 4)
    ref, p_ref = run(pipeline, model, 1, [(pcm.astype(np.float32), gt)])
    pipe = pipeline.AudioToPosePipeline(model, lanes=2)
    host = [(torch.from_numpy(pcm[i:i + 2]).pin_memory(), torch.from_numpy(gt[i:i + 2]).pin_memory()) for i in (0, 2)]
    assert host[0][0].dtype == torch.int16
    assert pipe.run_host_batches(host) == 4
    out = pipe.finish()
    assert out["n_frames"] == ref["n_frames"]
    assert abs(out["pck_hits"] - ref["pck_hits"]) <= 2                   # the two mel paths agree to fp32 rounding
    np.testing.assert_allclose(out["abs_pose"], ref["abs_pose"], rtol=1e-3)
