"""Oracle (test infrastructure): produce tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
Inputs are regenerated from seeds (oracle/synth.py, oracle/weights.py), so the fixtures hold
only the reference's *outputs*.  tests/test_oracle_golden.py checks the oracle restatements
against these vectors; the GPU tests check the CUDA path against the oracle and the vectors.
"""
import os
import sys
import numpy as np
import torch

from . import synth, weights, ref_shim, mel_oracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

MEL_CASES = [  # (name, kind, n_samples, clip index)
    ("noise_full", "noise", synth.CLIP_SAMPLES, 0),
    ("noise_8000", "noise", 8000, 1),
    ("tone_8000", "tone", 8000, 2),
    ("int16_8000", "int16", 8000, 3),
    ("zeros_1000", "zeros", 1000, 4),
    ("one_frame_400", "noise", 400, 5),
    ("no_frame_399", "noise", 399, 6),
    ("ragged_559", "noise", 559, 7),       # 1 frame, 159-sample tail dropped
    ("two_frames_560", "noise", 560, 8),
]

MODEL_CASES = [  # (name, weight seed, weight mode, B, T, F, with real_pose)
    ("stress_b2", 0, "stress", 2, 64, 64, True),
    ("default_b1", 1, "default", 1, 64, 64, False),
    ("stress_t32_f128", 2, "stress", 1, 32, 128, False),
]


def model_input(case_seed, B, T, F):
    """Seeded mel-like input: N(-1.5, 1.5^2), roughly the spread of log-mel of 0.1*noise."""
    g = torch.Generator(device="cpu")
    g.manual_seed(777 + case_seed)
    return -1.5 + 1.5 * torch.randn(B, T, F, generator=g, dtype=torch.float32)


def real_pose_input(case_seed, B, T):
    g = torch.Generator(device="cpu")
    g.manual_seed(888 + case_seed)
    return torch.randn(B, T, synth.POSE_FEATS, generator=g, dtype=torch.float32)


def reference_functions(path, names):
    """The UNMODIFIED source of top-level functions of a reference script that cannot be imported (it starts data
    loaders and training at import time): parsed with ast, compiled from the original text, nothing rewritten."""
    import ast
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return {n: ns[n] for n in names}


def smoothness_golden():
    """tests/golden/smooth_reference.npz: compute_temporal_smoothness_loss / compute_jerk_loss / pos_to_motion of
    version5_model_train.py on seeded pose batches."""
    fns = reference_functions(os.path.join(ref_shim.REFERENCE_ROOT, "version5_model_train.py"),
                              ["pos_to_motion", "compute_temporal_smoothness_loss", "compute_jerk_loss"])
    out = {}
    for name, first, n in SMOOTH_CASES:
        pose = torch.from_numpy(synth.noisy_pred_batch(first, n))
        motion = fns["pos_to_motion"](pose)
        out[name + "_smoothness"] = np.array(fns["compute_temporal_smoothness_loss"](motion).item())
        out[name + "_jerk"] = np.array(fns["compute_jerk_loss"](motion).item())
    short = torch.from_numpy(synth.noisy_pred_batch(9, 2)[:, :4])           # 4 poses -> 3 velocities -> 2 accel -> 1 jerk
    m = fns["pos_to_motion"](short)
    out["short_smoothness"] = np.array(fns["compute_temporal_smoothness_loss"](m).item())
    out["short_jerk"] = np.array(fns["compute_jerk_loss"](m).item())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "smooth_reference.npz"), **out)
    print("smoothness goldens:", {k: float(v) for k, v in out.items()})


SMOOTH_CASES = [("b4", 0, 4), ("b1", 7, 1), ("b16", 20, 16)]     # (name, first clip, clips)

# log_mel_spectrogram at sample rates whose 25 ms window needs another fft length (mel_features.py:212-214):
# (name, clip index, samples, kind, kwargs)
MEL_NFFT_CASES = [
    ("sr4000_nfft128", 30, 3000, "noise", dict(audio_sample_rate=4000, log_offset=1e-3, num_mel_bins=16,
                                               lower_edge_hertz=60.0, upper_edge_hertz=1900.0)),
    ("sr8000_nfft256", 31, 6000, "noise", dict(audio_sample_rate=8000, log_offset=0.01)),
    ("sr22050_nfft1024", 32, 12000, "noise", dict(audio_sample_rate=22050, log_offset=0.01, num_mel_bins=64,
                                                  lower_edge_hertz=125.0, upper_edge_hertz=7500.0)),
    ("sr44100_nfft2048", 33, 20000, "tone", dict(audio_sample_rate=44100, log_offset=0.01, num_mel_bins=128,
                                                 lower_edge_hertz=20.0, upper_edge_hertz=20000.0)),
    ("sr48000_nfft2048_i16", 34, 20000, "int16", dict(audio_sample_rate=48000, log_offset=1.0, num_mel_bins=80,
                                                      lower_edge_hertz=50.0, upper_edge_hertz=12000.0)),
]


def mel_nfft_golden():
    """tests/golden/melnfft_reference.npz: the UNMODIFIED mel_features.log_mel_spectrogram on seeded clips at the
    sample rates above (fft lengths 128 / 256 / 1024 / 2048), plus one |STFT| of a 1024-point transform."""
    mf = ref_shim.import_reference()["mel_features"]
    out = {}
    for name, idx, n, kind, kw in MEL_NFFT_CASES:
        wav = synth.wav_clip(idx, n, kind)
        if kind == "int16":
            wav = wav.astype(np.int16)
        out[name] = mf.log_mel_spectrogram(wav, **kw)
    out["stft_mag_1024"] = mf.stft_magnitude(synth.wav_clip(35, 5000), fft_length=1024, hop_length=300, window_length=700)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "melnfft_reference.npz"), **out)
    print("mel nfft goldens:", {k: v.shape for k, v in out.items()})


# stand-alone building blocks of model_layers.py: (name, class, ctor args, ctor kwargs, C_in, T)
BLOCK_CASES = [
    ("conv_k3_leaky", "ConvNormRelu", (64, 128), dict(type="1d", leaky=True), 64, 24),
    ("conv_k3_relu", "ConvNormRelu", (128, 64), dict(type="1d", leaky=False), 128, 20),
    ("conv_down", "ConvNormRelu", (128, 128), dict(type="1d", leaky=True, downsample=True), 128, 32),
    ("conv_transpose", "ConvTranspose1D", (128, 64), dict(), 128, 16),
    ("self_attention", "SelfAttention", (128,), dict(), 128, 24),
    ("self_attention_long", "SelfAttention", (64,), dict(), 64, 100),
    ("channel_attention", "ChannelAttention", (128,), dict(), 128, 32),
    ("res_block", "ResBlock", (64,), dict(type="1d", p=0.1), 64, 32),
]


def randomize_block(module, seed):
    """Non-trivial BatchNorm statistics / affine parameters and attention gates (the defaults make BatchNorm and
    gamma = 0 attention almost the identity)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if name.endswith("running_var"):
                t.copy_(0.5 + torch.rand(t.shape, generator=g))
            elif name.endswith("running_mean"):
                t.copy_(0.2 * torch.randn(t.shape, generator=g))
            elif name.endswith("norm.weight") or name.endswith("bn.weight"):
                t.copy_(0.8 + 0.4 * torch.rand(t.shape, generator=g))
            elif name.endswith("norm.bias") or name.endswith("bn.bias"):
                t.copy_(0.1 * torch.randn(t.shape, generator=g))
            elif name.endswith("gamma"):
                t.fill_(0.5)


def block_golden():
    """tests/golden/blocks_reference.npz: the UNMODIFIED layer classes of model_layers.py (eval mode) on seeded inputs;
    per case the module's state_dict (small channel counts keep the fixture small), the input and the output."""
    ml = ref_shim.import_reference()["model_layers"]
    out = {}
    for i, (name, cls, args, kwargs, cin, T) in enumerate(BLOCK_CASES):
        torch.manual_seed(100 + i)
        mod = getattr(ml, cls)(*args, **kwargs).eval()
        randomize_block(mod, 200 + i)
        g = torch.Generator().manual_seed(300 + i)
        x = torch.randn(3, cin, T, generator=g)
        with torch.no_grad():
            y = mod(x)
        for k, v in mod.state_dict().items():
            if not k.endswith("num_batches_tracked"):
                out["%s/sd/%s" % (name, k)] = v.numpy()
        out[name + "/x"] = x.numpy()
        out[name + "/y"] = y.numpy()
        print(name, tuple(x.shape), "->", tuple(y.shape), float(y.abs().mean()))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "blocks_reference.npz"), **out)


DISC_CASES = [("disc_b3_t64", 11, 3, 64), ("disc_b2_t40", 12, 2, 40)]      # (name, weight seed, B, T)


def disc_golden():
    """tests/golden/disc_reference.npz: the UNMODIFIED SelfAttention_D (eval mode, torch_geometric stand-ins of
    oracle/ref_shim.py) on seeded poses; weights from oracle/weights.make_state_dict(seed, discriminator=True)."""
    rm = ref_shim.import_reference()["real_motion_model"]
    torch.manual_seed(0)
    model = rm.SelfAttention_D().eval()
    ref_sd = model.state_dict()
    con = weights.disc_contract()
    assert [n for n, _, _ in con] == list(ref_sd.keys()), "discriminator contract differs from the reference state_dict"
    for n, shape, _ in con:
        assert tuple(ref_sd[n].shape) == tuple(shape), (n, ref_sd[n].shape, shape)
    out = {"n_tensors": np.array(len(con))}
    for name, seed, B, T in DISC_CASES:
        model.load_state_dict(weights.make_state_dict(seed, "stress", discriminator=True), strict=True)
        model.eval()
        pose = real_pose_input(seed, B, T)
        with torch.no_grad():
            y, losses = model(pose)
        assert losses == []
        out[name] = y.numpy()
        print(name, tuple(pose.shape), "->", tuple(y.shape), y.numpy().ravel()[:4])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "disc_reference.npz"), **out)


def main():
    if "--disc-only" in sys.argv:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        return disc_golden()
    if "--blocks-only" in sys.argv:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        return block_golden()
    if "--smoothness-only" in sys.argv:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        return smoothness_golden()
    if "--mel-nfft-only" in sys.argv:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        return mel_nfft_golden()
    ref = ref_shim.import_reference()
    mf, me, rm = ref["mel_features"], ref["motion_evaluation"], ref["real_motion_model"]
    os.makedirs(GOLDEN_DIR, exist_ok=True)

    # ---- mel: the reference log_mel_spectrogram with audio_repr's parameters -------------------
    kw = mel_oracle.AUDIO_REPR_KW
    out = {}
    for name, kind, n, idx in MEL_CASES:
        out[name] = mf.log_mel_spectrogram(synth.wav_clip(idx, n, kind), **kw)
    w = mf.spectrogram_to_mel_matrix(num_mel_bins=64, num_spectrogram_bins=257, audio_sample_rate=16000,
                                     lower_edge_hertz=125, upper_edge_hertz=7500)
    r, c = np.nonzero(w)
    out["melw_rows"], out["melw_cols"], out["melw_vals"] = r.astype(np.int32), c.astype(np.int32), w[r, c]
    out["hann_400"] = mf.periodic_hann(400)
    out["stft_mag_2000"] = mf.stft_magnitude(synth.wav_clip(9, 2000), fft_length=512, hop_length=160, window_length=400)
    out["frames_shape_1000"] = np.array(mf.frame(synth.wav_clip(4, 1000), 400, 160).shape)
    # default-parameter call (8 kHz, 20 mel bins, log_offset 0 is -inf prone -> use 1e-3)
    out["default_params_4000"] = mf.log_mel_spectrogram(synth.wav_clip(10, 4000), log_offset=1e-3)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "mel_reference.npz"), **out)
    print("mel goldens:", {k: v.shape for k, v in out.items()})

    # ---- eval: compute_pck / compute_pck_radius / L1Loss ---------------------------------------
    gt = synth.gt_pose_batch(0, 4)
    pred = synth.noisy_pred_batch(0, 4)
    gtf, prf = gt.reshape(-1, 2, 52), pred.reshape(-1, 2, 52)
    ev = {"pck_alpha02": me.compute_pck(prf, gtf, 0.2), "pck_alpha01": me.compute_pck(prf, gtf, 0.1),
          "radius_alpha02": me.compute_pck_radius(gtf, 0.2)[:, 0],
          "pck_identity": me.compute_pck(gtf, gtf)}
    tp, tg_ = torch.from_numpy(pred), torch.from_numpy(gt)
    ev["l1_pose"] = np.array(torch.nn.L1Loss()(tp, tg_).item())
    ev["l1_motion"] = np.array(torch.nn.L1Loss()(torch.diff(tp, n=1, dim=1), torch.diff(tg_, n=1, dim=1)).item())
    # fp64 inputs keep fp64 arithmetic in the reference
    ev["pck_alpha02_f64"] = me.compute_pck(prf.astype(np.float64), gtf.astype(np.float64), 0.2)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "eval_reference.npz"), **ev)
    print("eval goldens:", {k: v.shape for k, v in ev.items()})

    # ---- model: the reference SelfAttention_G (shim + D1), eval mode ---------------------------
    torch.manual_seed(0)
    model = rm.SelfAttention_G().eval()
    ref_sd = model.state_dict()
    con = weights.contract()
    assert [n for n, _, _ in con] == list(ref_sd.keys()), "contract key order differs from the reference state_dict"
    for n, shape, _ in con:
        assert tuple(ref_sd[n].shape) == tuple(shape), (n, ref_sd[n].shape, shape)
    n_params = sum(p.numel() for p in model.parameters())
    print("contract ok:", len(con), "tensors,", n_params, "parameters")
    mo = {"n_tensors": np.array(len(con)), "n_params": np.array(n_params)}
    captured = {}
    model.audio_encoder.register_forward_hook(lambda m, i, o: captured.__setitem__("enc", o.detach().clone()))
    model.unet.register_forward_hook(lambda m, i, o: captured.__setitem__("unet", o.detach().clone()))
    for name, seed, mode, B, T, F, with_pose in MODEL_CASES:
        sd = weights.make_state_dict(seed, mode)
        model.load_state_dict(sd, strict=True)
        model.eval()
        x = model_input(seed, B, T, F)
        rp = real_pose_input(seed, B, T) if with_pose else None
        with torch.no_grad():
            pose, losses = model(x, real_pose=rp)
        mo[name + "_pose"] = pose.numpy()
        mo[name + "_losses"] = np.array([l.item() for l in losses], dtype=np.float64)
        if name == "stress_b2":
            mo[name + "_enc"] = captured["enc"].numpy()
            mo[name + "_unet"] = captured["unet"].numpy()
        print(name, pose.shape, mo[name + "_losses"], float(pose.abs().mean()))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "model_reference.npz"), **mo)
    smoothness_golden()
    mel_nfft_golden()
    block_golden()
    disc_golden()
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    sys.exit(main())
