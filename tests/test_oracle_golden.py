"""CPU: pin the oracle restatements against vectors produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container).  Not a product test."""
import numpy as np
import pytest
import torch

from oracle import mel_oracle, eval_oracle, model_oracle, synth, weights
from oracle.make_golden import BLOCK_CASES, MEL_CASES, MEL_NFFT_CASES, MODEL_CASES, model_input, real_pose_input


@pytest.mark.parametrize("name,kind,n,idx", MEL_CASES)
def test_mel_oracle_matches_reference(golden, name, kind, n, idx):
    got = mel_oracle.log_mel_audio_repr(synth.wav_clip(idx, n, kind))
    ref = golden["mel"][name]
    assert got.shape == ref.shape and got.dtype == np.float64
    np.testing.assert_array_equal(got, ref)          # same fp64 operation sequence -> bit equal


def test_mel_oracle_pieces(golden):
    g = golden["mel"]
    np.testing.assert_array_equal(mel_oracle.hann(400), g["hann_400"])
    w = mel_oracle.mel_matrix(64, 257, 16000, 125, 7500)
    r, c = np.nonzero(w)
    assert len(r) == 461 and r.min() == 5 and r.max() == 239
    np.testing.assert_array_equal(r, g["melw_rows"])
    np.testing.assert_array_equal(c, g["melw_cols"])
    np.testing.assert_array_equal(w[r, c], g["melw_vals"])
    np.testing.assert_array_equal(mel_oracle.stft_mag(synth.wav_clip(9, 2000), 512, 160, 400), g["stft_mag_2000"])
    assert tuple(g["frames_shape_1000"]) == mel_oracle.frames(synth.wav_clip(4, 1000), 400, 160).shape
    np.testing.assert_array_equal(mel_oracle.log_mel(synth.wav_clip(10, 4000), log_offset=1e-3),
                                  g["default_params_4000"])
    np.testing.assert_array_equal(mel_oracle.stft_mag(synth.wav_clip(35, 5000), 1024, 300, 700),
                                  golden["melnfft"]["stft_mag_1024"])


@pytest.mark.parametrize("name,idx,n,kind,kw", MEL_NFFT_CASES)
def test_mel_oracle_other_fft_lengths(golden, name, idx, n, kind, kw):
    """fft lengths 128 / 256 / 1024 / 2048 (mel_features.py:212-214 at other sample rates)."""
    wav = synth.wav_clip(idx, n, kind)
    if kind == "int16":
        wav = wav.astype(np.int16)
    np.testing.assert_array_equal(mel_oracle.log_mel(wav, **kw), golden["melnfft"][name])


def test_mel_self_checks():
    # SURVEY.md section 8c known answers that come from the reference code itself
    z = mel_oracle.log_mel_audio_repr(np.zeros(1000, np.float32))
    assert np.all(z == np.log(0.01))
    h = mel_oracle.hann(400)
    assert h[0] == 0.0 and abs(h[200] - 1.0) < 1e-15 and abs(h[1] - 6.168e-5) < 1e-8
    assert mel_oracle.stft_geometry(16000, 0.025, 0.010) == (400, 160, 512)
    assert mel_oracle.num_frames(synth.CLIP_SAMPLES, 400, 160) == 425
    with pytest.raises(ValueError):
        mel_oracle.mel_matrix(64, 257, 16000, 125, 9000)
    with pytest.raises(ValueError):
        mel_oracle.mel_matrix(64, 257, 16000, 4000, 3000)
    with pytest.raises(ValueError):
        mel_oracle.frames(np.zeros(100), 400, 160)


def test_eval_oracle_matches_reference(golden):
    g = golden["eval"]
    gt, pred = synth.gt_pose_batch(0, 4), synth.noisy_pred_batch(0, 4)
    gtf, prf = eval_oracle.poses_as_frames(gt), eval_oracle.poses_as_frames(pred)
    np.testing.assert_array_equal(eval_oracle.pck(prf, gtf, 0.2), g["pck_alpha02"])
    np.testing.assert_array_equal(eval_oracle.pck(prf, gtf, 0.1), g["pck_alpha01"])
    rad = eval_oracle.pck_radius(gtf, 0.2)
    assert rad.dtype == np.float32 and rad.shape == (256, 52)
    np.testing.assert_array_equal(rad[:, 0], g["radius_alpha02"])
    np.testing.assert_array_equal(eval_oracle.pck(gtf, gtf), g["pck_identity"])
    assert np.all(g["pck_identity"] == 1.0)
    np.testing.assert_array_equal(eval_oracle.pck(prf.astype(np.float64), gtf.astype(np.float64)), g["pck_alpha02_f64"])
    part = eval_oracle.metric_partials(pred, gt)
    fin = eval_oracle.finalize(part)
    assert part["pck_hits"] == int(round(g["pck_alpha02"].sum() * 52))
    assert 0.5 < fin["pck"] < 1.0                     # both branches of the compare are exercised
    np.testing.assert_allclose(fin["l1_pose"], g["l1_pose"], rtol=1e-5)
    np.testing.assert_allclose(fin["l1_motion"], g["l1_motion"], rtol=1e-5)


def test_smoothness_oracle_matches_reference(golden):
    """compute_temporal_smoothness_loss / compute_jerk_loss of the unmodified version5_model_train.py."""
    from oracle.make_golden import SMOOTH_CASES
    g = golden["smooth"]
    for name, first, n in SMOOTH_CASES:
        m = eval_oracle.motion(synth.noisy_pred_batch(first, n))
        np.testing.assert_allclose(eval_oracle.smoothness(m), g[name + "_smoothness"], rtol=1e-6)
        np.testing.assert_allclose(eval_oracle.jerk(m), g[name + "_jerk"], rtol=1e-6)
    m = eval_oracle.motion(synth.noisy_pred_batch(9, 2)[:, :4])
    np.testing.assert_allclose(eval_oracle.smoothness(m), g["short_smoothness"], rtol=1e-6)
    np.testing.assert_allclose(eval_oracle.jerk(m), g["short_jerk"], rtol=1e-6)


def test_weight_contract():
    con = weights.contract()
    sd = weights.make_state_dict(0, "stress")
    assert len(con) == 340 and list(sd.keys()) == [n for n, _, _ in con]
    assert weights.num_parameters(sd) == 45875858
    again = weights.make_state_dict(0, "stress")
    assert all(torch.equal(sd[k], again[k]) for k in sd)
    be, he = weights.edge_templates()
    assert be.shape == (2, 18) and he.shape == (2, 80)
    hand_t, body_t = model_oracle.angle_triples()
    assert len(hand_t) == 30 and len(body_t) == 5


@pytest.mark.parametrize("name,seed,mode,B,T,F,with_pose", MODEL_CASES)
def test_model_oracle_matches_reference(golden, name, seed, mode, B, T, F, with_pose):
    g = golden["model"]
    torch.set_num_threads(max(1, torch.get_num_threads()))
    sd = weights.make_state_dict(seed, mode)
    x = model_input(seed, B, T, F)
    rp = real_pose_input(seed, B, T) if with_pose else None
    pose, losses = model_oracle.generator_forward(sd, x, rp)
    ref = torch.from_numpy(g[name + "_pose"])
    assert pose.shape == ref.shape
    rel = (pose - ref).abs().sum() / ref.abs().sum()
    assert rel < 2e-6, rel
    assert (pose - ref).abs().max() < 1e-4
    np.testing.assert_allclose([l.item() for l in losses], g[name + "_losses"], rtol=2e-5, atol=1e-6)
    if name == "stress_b2":
        enc = model_oracle.audio_encoder(sd, x)
        np.testing.assert_allclose(enc.numpy(), g[name + "_enc"], rtol=1e-4, atol=1e-5)
        un = model_oracle.unet(sd, enc)
        np.testing.assert_allclose(un.numpy(), g[name + "_unet"], rtol=1e-4, atol=2e-5)


def test_model_oracle_rejects_bad_t():
    with pytest.raises(ValueError):
        model_oracle.generator_forward({}, torch.zeros(1, 62, 64))


@pytest.mark.parametrize("name,cls,args,kwargs,cin,T", BLOCK_CASES)
def test_block_oracles_match_reference_classes(name, cls, args, kwargs, cin, T):
    """The functional restatements of the layer classes (model_oracle.conv_norm_act, self_attention, ...) against the
    unmodified reference classes' outputs on the same parameters (tests/golden/blocks_reference.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blocks_reference.npz"))
    sd = {"p." + k[len(name) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(name + "/sd/")}
    x = torch.from_numpy(g[name + "/x"])
    fn = {"ConvNormRelu": lambda: model_oracle.conv_norm_act(sd, "p", x, stride=2 if kwargs.get("downsample") else 1,
                                                             padding=1, leaky=kwargs.get("leaky", False)),
          "ConvTranspose1D": lambda: model_oracle.conv_transpose_block(sd, "p", x),
          "SelfAttention": lambda: model_oracle.self_attention(sd, "p", x),
          "ChannelAttention": lambda: model_oracle.channel_attention(sd, "p", x),
          "ResBlock": lambda: model_oracle.res_block(sd, "p", x)}[cls]
    torch.testing.assert_close(fn(), torch.from_numpy(g[name + "/y"]), rtol=1e-5, atol=1e-5)


def test_discriminator_oracle_matches_reference():
    """model_oracle.discriminator_forward (dense-adjacency GAT, functional convs) against the unmodified SelfAttention_D
    run through the scatter-style stand-ins (tests/golden/disc_reference.npz); fp32 round-off only."""
    import os
    from oracle.make_golden import DISC_CASES
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "disc_reference.npz"))
    assert int(g["n_tensors"]) == len(weights.disc_contract())
    for name, seed, B, T in DISC_CASES:
        sd = weights.make_state_dict(seed, "stress", discriminator=True)
        y = model_oracle.discriminator_forward(sd, real_pose_input(seed, B, T))
        assert tuple(y.shape) == g[name].shape
        np.testing.assert_allclose(y.numpy(), g[name], rtol=1e-3, atol=2e-5)
