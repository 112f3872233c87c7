"""GPU parity of the tcgen05 tap-offset implicit-GEMM operator (a2m_gemm_taps) for every layer class
of the generator, against torch fp32 convolutions on the same bf16-rounded operands (the oracle's
conv_norm_act arithmetic).  bf16 outputs: |err| <= 1e-2 * max|ref| (one bf16 rounding of the result);
fp32 outputs: <= 2e-4 * max|ref| (accumulation order only)."""
import pytest
import torch
import torch.nn.functional as F

import gemm_cases as gc

pytestmark = pytest.mark.gpu


def check(got, ref, bf16_out=True):
    got = got.float().cpu()
    assert got.shape == ref.shape
    tol = (1e-2 if bf16_out else 2e-4) * ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= tol, (err, tol)
    assert (got - ref).abs().sum() / ref.abs().sum() <= (4e-3 if bf16_out else 1e-4)


@pytest.mark.parametrize("M,K,N,act", [(300, 256, 640, "none"), (128, 64, 64, "relu"), (1000, 2688, 256, "leaky"),
                                        (64, 256, 320, "none")])
def test_linear(M, K, N, act):
    x = gc.bf16_round(gc.gen((M, K), 1))
    w = gc.gen((N, K), 2, K ** -0.5)
    b = gc.gen((N,), 3, 0.1)
    ref = gc.act_ref(x @ gc.bf16_round(w).t() + b, act)
    xd = x.to(torch.bfloat16).cuda()
    out = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    kw = dict(sources=[(xd, [K, M], [1, K])], box=[128, 1, 1, 1], m_extent=[M, 1, 1, 1], out_stride=[N, 0, 0, 0],
              taps=[(0, (0, 0, 0, 0), K, 0)], N=N, act=act)
    gc.run(kw, w.cuda(), K, 1, None, b.cuda(), out)
    check(out, gc.bf16_round(ref))


@pytest.mark.parametrize("B,T,C,N,act", [(3, 64, 256, 512, "leaky"), (9, 16, 128, 256, "leaky"), (2, 32, 64, 64, "none"),
                                          (1, 24, 64, 128, "relu")])
def test_conv1d_k3_with_folded_batchnorm(B, T, C, N, act):
    x = gc.bf16_round(gc.gen((B, C, T), 4))
    w = gc.gen((N, C, 3), 5, (3 * C) ** -0.5)
    scale = 0.5 + torch.rand(N, generator=torch.Generator().manual_seed(6))
    b = gc.gen((N,), 7, 0.1)
    w_eff = gc.bf16_round(w * scale[:, None, None])
    ref = gc.act_ref(F.conv1d(x, w_eff, b, padding=1), act).permute(0, 2, 1).contiguous()       # [B,T,N]
    xd = x.permute(0, 2, 1).contiguous().to(torch.bfloat16).cuda()                               # [B,T,C]
    out = torch.full((B, T, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    box = gc.rows_box(T, B)
    kw = dict(sources=[(xd, [C, T, B], [1, C, T * C])], box=box, m_extent=[T, B, 1, 1],
              out_stride=[N, T * N, 0, 0], taps=[(0, (j - 1, 0, 0, 0), C, j) for j in range(3)], N=N, act=act)
    gc.run(kw, w.cuda(), C * 3, 3, scale.cuda(), b.cuda(), out)
    check(out, gc.bf16_round(ref))


@pytest.mark.parametrize("B,T,C,N", [(5, 32, 128, 128), (2, 64, 512, 512), (17, 16, 64, 192)])
def test_conv1d_k4_stride2(B, T, C, N):
    x = gc.bf16_round(gc.gen((B, C, T), 8))
    w = gc.gen((N, C, 4), 9, (4 * C) ** -0.5)
    b = gc.gen((N,), 10, 0.1)
    ref = F.leaky_relu(F.conv1d(x, gc.bf16_round(w), b, stride=2, padding=1), 0.2).permute(0, 2, 1).contiguous()
    To = T // 2
    xd = x.permute(0, 2, 1).contiguous().to(torch.bfloat16).cuda()
    out = torch.full((B, To, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    # view [B, T/2, 2, C]: dims (C, parity, T/2, B); tap j reads input 2t + j - 1
    taps = [(0, (1, -1, 0, 0), C, 0), (0, (0, 0, 0, 0), C, 1), (0, (1, 0, 0, 0), C, 2), (0, (0, 1, 0, 0), C, 3)]
    box = [1] + gc.rows_box(To, B)[:3]
    kw = dict(sources=[(xd, [C, 2, To, B], [1, C, 2 * C, T * C])], box=box, m_extent=[1, To, B, 1],
              out_stride=[0, N, To * N, 0], taps=taps, N=N, act="leaky")
    gc.run(kw, w.cuda(), C * 4, 4, None, b.cuda(), out)
    check(out, gc.bf16_round(ref))


def test_conv1d_two_source_skip_concat():
    B, T, C0, C1, N = 4, 32, 128, 192, 256
    a = gc.bf16_round(gc.gen((B, C0, T), 11))
    s = gc.bf16_round(gc.gen((B, C1, T), 12))
    w = gc.gen((N, C0 + C1, 3), 13, (3 * (C0 + C1)) ** -0.5)
    b = gc.gen((N,), 14, 0.1)
    ref = F.leaky_relu(F.conv1d(torch.cat([a, s], 1), gc.bf16_round(w), b, padding=1), 0.2).permute(0, 2, 1).contiguous()
    ad = a.permute(0, 2, 1).contiguous().to(torch.bfloat16).cuda()
    sd = s.permute(0, 2, 1).contiguous().to(torch.bfloat16).cuda()
    out = torch.full((B, T, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    taps = [(0, (j - 1, 0, 0, 0), C0, j) for j in range(3)] + [(1, (j - 1, 0, 0, 0), C1, C0 * 3 + j) for j in range(3)]
    kw = dict(sources=[(ad, [C0, T, B], [1, C0, T * C0]), (sd, [C1, T, B], [1, C1, T * C1])],
              box=gc.rows_box(T, B), m_extent=[T, B, 1, 1], out_stride=[N, T * N, 0, 0], taps=taps, N=N, act="leaky")
    gc.run(kw, w.cuda(), (C0 + C1) * 3, 3, None, b.cuda(), out)
    check(out, gc.bf16_round(ref))


def test_conv_transpose_k3_s2_as_two_parity_gemms():
    B, T, C, N = 3, 16, 256, 128
    x = gc.bf16_round(gc.gen((B, C, T), 15))
    w = gc.gen((C, N, 3), 16, (3 * C) ** -0.5)              # ConvTranspose1d weight [C_in, C_out, k]
    b = gc.gen((N,), 17, 0.1)
    ref = F.relu(F.conv_transpose1d(x, gc.bf16_round(w), b, stride=2, padding=1, output_padding=1)).permute(0, 2, 1).contiguous()
    xd = x.permute(0, 2, 1).contiguous().to(torch.bfloat16).cuda()
    out = torch.full((B, 2 * T, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    src = [(xd, [C, T, B], [1, C, T * C])]
    common = dict(sources=src, box=gc.rows_box(T, B), m_extent=[T, B, 1, 1], out_stride=[2 * N, 2 * T * N, 0, 0], N=N,
                  act="relu")
    # out[2j] = W[:, :, 1] x[j];  out[2j+1] = W[:, :, 2] x[j] + W[:, :, 0] x[j+1]
    gc.run(dict(common, taps=[(0, (0, 0, 0, 0), C, 1)], out_base=0), w.cuda(), 3, N * 3, None, b.cuda(), out)
    gc.run(dict(common, taps=[(0, (0, 0, 0, 0), C, 2), (0, (1, 0, 0, 0), C, 0)], out_base=N), w.cuda(), 3, N * 3, None,
           b.cuda(), out)
    check(out, gc.bf16_round(ref))


@pytest.mark.parametrize("B,H,W,C,N", [(3, 32, 32, 64, 128), (5, 16, 16, 128, 256)])
def test_conv2d_k4_stride2(B, H, W, C, N):
    x = gc.bf16_round(gc.gen((B, C, H, W), 18))
    w = gc.gen((N, C, 4, 4), 19, (16 * C) ** -0.5)
    b = gc.gen((N,), 20, 0.1)
    ref = F.leaky_relu(F.conv2d(x, gc.bf16_round(w), b, stride=2, padding=1), 0.2).permute(0, 2, 3, 1).contiguous()
    Ho, Wo = H // 2, W // 2
    xd = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()                 # [B,H,W,C]
    out = torch.full((B, Ho, Wo, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    # two maps (row parity folded into the base pointer); dims (C, pw, W/2, H/2, B)
    dims, strides = [C, 2, Wo, Ho, B], [1, C, 2 * C, 2 * W * C, H * W * C]
    srcs = [(xd, dims, strides), (xd.view(-1)[W * C:], dims, strides)]
    par = {0: (1, -1), 1: (0, 0), 2: (1, 0), 3: (0, 1)}                               # k index -> (parity, shift)
    taps = []
    for i in range(4):
        for j in range(4):
            (ph, dh), (pw, dw) = par[i], par[j]
            taps.append((ph, (pw, dw, dh, 0), C, i * 4 + j))
    wo_b = min(Wo, 128); ho_b = min(Ho, 128 // wo_b); b_b = 128 // (wo_b * ho_b)
    kw = dict(sources=srcs, box=[1, wo_b, ho_b, b_b], m_extent=[1, Wo, Ho, B],
              out_stride=[0, N, Wo * N, Ho * Wo * N], taps=taps, N=N, act="leaky")
    # the second map starts one row down: its last row-pair reads past the tensor -> shrink its H/2 extent
    srcs[1] = (srcs[1][0], [C, 2, Wo, Ho, B], strides)
    gc.run(kw, w.cuda(), C * 16, 16, None, b.cuda(), out)
    check(out, gc.bf16_round(ref))


def test_conv2d_k3_and_centre_column_k3x8():
    B, H, W, C, N = 5, 8, 8, 128, 256
    x = gc.bf16_round(gc.gen((B, C, H, W), 21))
    xd = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    w = gc.gen((N, C, 3, 3), 22, (9 * C) ** -0.5)
    b = gc.gen((N,), 23, 0.1)
    ref = F.leaky_relu(F.conv2d(x, gc.bf16_round(w), b, padding=1), 0.2).permute(0, 2, 3, 1).contiguous()
    out = torch.full((B, H, W, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    taps = [(0, (j - 1, i - 1, 0, 0), C, i * 3 + j) for i in range(3) for j in range(3)]
    kw = dict(sources=[(xd, [C, W, H, B], [1, C, W * C, H * W * C])], box=[8, 8, 2, 1], m_extent=[W, H, B, 1],
              out_stride=[N, W * N, H * W * N, 0], taps=taps, N=N, act="leaky")
    gc.run(kw, w.cuda(), C * 9, 9, None, b.cuda(), out)
    check(out, gc.bf16_round(ref))
    # AudioEncoder conv 4: kernel (3,8), padding (1,3) -> only the centre output column (w_out = 3) is needed
    w2 = gc.gen((N, C, 3, 8), 24, (24 * C) ** -0.5)
    full = F.leaky_relu(F.conv2d(x, gc.bf16_round(w2), b, padding=(1, 3)), 0.2)        # [B,N,8,7]
    ref2 = full[:, :, :, 3].permute(0, 2, 1).contiguous()                               # [B,8,N]
    out2 = torch.full((B, H, N), float("nan"), dtype=torch.float32, device="cuda")
    taps2 = [(0, (j, i - 1, 0, 0), C, i * 8 + j) for i in range(3) for j in range(8)]
    kw2 = dict(sources=[(xd, [C, W, H, B], [1, C, W * C, H * W * C])], box=[1, 8, 16, 1], m_extent=[1, H, B, 1],
               out_stride=[0, N, H * N, 0], taps=taps2, N=N, act="leaky")
    gc.run(kw2, w2.cuda(), C * 24, 24, None, b.cuda(), out2)
    check(out2, ref2, bf16_out=False)


@pytest.mark.parametrize("N,col", [(20, 0), (84, 20)])
def test_logits_fp32_into_pose_layout(N, col):
    B, T, C = 3, 64, 256
    x = gc.bf16_round(gc.gen((B * T, C), 25))
    w = gc.gen((N, C), 26, C ** -0.5)
    b = gc.gen((N,), 27, 0.1)
    ref = x @ gc.bf16_round(w).t() + b
    out = torch.zeros((B * T, 104), dtype=torch.float32, device="cuda")
    kw = dict(sources=[(x.to(torch.bfloat16).cuda(), [C, B * T], [1, C])], box=[128, 1, 1, 1], m_extent=[B * T, 1, 1, 1],
              out_stride=[104, 0, 0, 0], taps=[(0, (0, 0, 0, 0), C, 0)], N=N, out_base=col)
    gc.run(kw, w.cuda(), C, 1, None, b.cuda(), out)
    check(out[:, col:col + N], ref, bf16_out=False)
    mask = torch.ones(104, dtype=torch.bool); mask[col:col + N] = False
    assert (out[:, mask] == 0).all()                       # nothing written outside the column block


def test_full_size_layer_and_linearity():
    """The largest UNet layer at BASELINE config-2 size (B=256: cat 1024+1024 -> 1024, T=32, K=6144):
    a random sample of outputs is checked against the fp32 reference, and doubling A doubles D exactly."""
    B, T, C, N = 256, 32, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(B, T, C, device="cuda", generator=g).to(torch.bfloat16)
    s = torch.randn(B, T, C, device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(N, 2 * C, 3, device="cuda", generator=g) * (6 * C) ** -0.5
    out = torch.empty(B, T, N, dtype=torch.bfloat16, device="cuda")
    taps = [(0, (j - 1, 0, 0, 0), C, j) for j in range(3)] + [(1, (j - 1, 0, 0, 0), C, C * 3 + j) for j in range(3)]
    kw = dict(sources=[(a, [C, T, B], [1, C, T * C]), (s, [C, T, B], [1, C, T * C])], box=[32, 4, 1, 1],
              m_extent=[T, B, 1, 1], out_stride=[N, T * N, 0, 0], taps=taps, N=N, act="none")
    gc.run(kw, w, 2 * C * 3, 3, None, None, out)
    for b_i in (0, 17, 255):
        x = torch.cat([a[b_i], s[b_i]], 1).float().t().unsqueeze(0).cpu()             # [1, 2C, T]
        ref = F.conv1d(x, gc.bf16_round(w.cpu()), None, padding=1)[0].t()
        check(out[b_i], gc.bf16_round(ref))
    out2 = torch.empty_like(out)
    kw["sources"] = [(a * 2, [C, T, B], [1, C, T * C]), (s * 2, [C, T, B], [1, C, T * C])]
    gc.run(kw, w, 2 * C * 3, 3, None, None, out2)
    assert torch.equal(out2.float(), 2 * out.float())


@pytest.mark.parametrize("hint", [256, 512])
@pytest.mark.parametrize("B,T,C,N,act", [(8, 64, 256, 256, "leaky"), (5, 64, 128, 512, "none"), (3, 32, 320, 768, "relu"),
                                          (1, 128, 64, 256, "leaky")])
def test_wide_tiles_conv1d_k3(B, T, C, N, act, hint):
    """The 128 x 256 variant (hint 256) and the CTA-pair variant (hint 512: two neighbouring M tiles and 256 output
    channels are one cta_group::2 MMA tile; B = 5, 3 give odd M-tile counts, i.e. a pair whose second half is idle)."""
    x = gc.bf16_round(gc.gen((B, C, T), 40))
    w = gc.gen((N, C, 3), 41, (3 * C) ** -0.5)
    b = gc.gen((N,), 42, 0.1)
    ref = gc.act_ref(F.conv1d(x, gc.bf16_round(w), b, padding=1), act).permute(0, 2, 1).contiguous()
    xd = x.permute(0, 2, 1).contiguous().to(torch.bfloat16).cuda()
    out = torch.full((B, T, N), float("nan"), dtype=torch.bfloat16, device="cuda")
    kw = dict(sources=[(xd, [C, T, B], [1, C, T * C])], box=gc.rows_box(T, B), m_extent=[T, B, 1, 1],
              out_stride=[N, T * N, 0, 0], taps=[(0, (j - 1, 0, 0, 0), C, j) for j in range(3)], N=N, act=act, tile_hint=hint)
    gc.run(kw, w.cuda(), C * 3, 3, None, b.cuda(), out)
    check(out, gc.bf16_round(ref))


def test_pair_tiles_match_single_cta_tiles_bit_for_bit():
    """Same K order, same fp32 accumulation per output element: the CTA-pair kernel must reproduce the 128 x 256 kernel
    exactly (a large-K two-source layer, the shape of unet.up1)."""
    B, T, C0, C1, N = 6, 32, 256, 256, 512
    a = gc.gen((B, T, C0), 50).to(torch.bfloat16).cuda()
    s = gc.gen((B, T, C1), 51).to(torch.bfloat16).cuda()
    w = gc.gen((N, C0 + C1, 3), 52, (3 * (C0 + C1)) ** -0.5).cuda()
    b = gc.gen((N,), 53, 0.1).cuda()
    outs = []
    for hint in (256, 512):
        out = torch.full((B, T, N), float("nan"), dtype=torch.bfloat16, device="cuda")
        taps = [(0, (j - 1, 0, 0, 0), C0, j) for j in range(3)] + [(1, (j - 1, 0, 0, 0), C1, C0 * 3 + j) for j in range(3)]
        kw = dict(sources=[(a, [C0, T, B], [1, C0, T * C0]), (s, [C1, T, B], [1, C1, T * C1])], box=gc.rows_box(T, B),
                  m_extent=[T, B, 1, 1], out_stride=[N, T * N, 0, 0], taps=taps, N=N, act="leaky", tile_hint=hint)
        gc.run(kw, w, (C0 + C1) * 3, 3, None, b, out)
        outs.append(out)
    assert not torch.isnan(outs[0].float()).any()
    assert torch.equal(outs[0], outs[1])
