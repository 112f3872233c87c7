"""Oracle (test infrastructure): the generator's state_dict contract and a seeded weight factory.

The contract (names, shapes, dtypes) restates what ``SelfAttention_G.__init__`` registers
(``/root/reference/real_motion_model.py:22-129`` on top of ``model_layers.py:53-118,125-131,
155-165,179-183,198-209,249-261,303-339``); ``oracle/make_golden.py`` asserts it equals the
state_dict of the *unmodified* reference class key for key (340 tensors,
45 875 858 parameters).

``make_state_dict(seed, mode)`` builds identical weights on any machine (CPU torch generator),
so the GPU box can rebuild exactly the weights the goldens were produced with.
"""
from collections import OrderedDict
import math
import torch

# pats/data_loading/skeleton.py:94-110 -- 52-joint tree (index 0 = Neck)
PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 7, 6,
           10, 11, 12, 13, 10, 15, 16, 17, 10, 19, 20, 21, 10, 23, 24, 25, 10, 27, 28, 29,
           3,
           31, 32, 33, 34, 31, 36, 37, 38, 31, 40, 41, 42, 31, 44, 45, 46, 31, 48, 49, 50]
N_BODY, N_HAND, JOINT_FEAT, GAT_HEADS = 10, 42, 64, 4   # real_motion_model.py:33-35,78


def edge_templates():
    """Directed edge lists [2,E] (row0 = source, row1 = target), real_motion_model.py:43-60."""
    body, hand = [], []
    for i, par in enumerate(PARENTS[:N_BODY]):
        par = par if par < N_BODY else -1
        if par != -1:
            body += [[par, i], [i, par]]
    for i, par in enumerate(PARENTS[10:10 + N_HAND]):
        par = par - 10 if par >= 10 else -1
        if par != -1:
            hand += [[par, i], [i, par]]
    return (torch.tensor(body, dtype=torch.long).t().contiguous(),
            torch.tensor(hand, dtype=torch.long).t().contiguous())


def _cnr(prefix, c_out, c_in, *k, bn="norm", conv="conv", transpose=False):
    """ConvNormRelu / ConvTranspose1D entries: conv weight+bias and the 5 BatchNorm tensors."""
    wshape = (c_in, c_out) + k if transpose else (c_out, c_in) + k
    return [(f"{prefix}.{conv}.weight", wshape, "w"), (f"{prefix}.{conv}.bias", (c_out,), "b"),
            (f"{prefix}.{bn}.weight", (c_out,), "bn_w"), (f"{prefix}.{bn}.bias", (c_out,), "bn_b"),
            (f"{prefix}.{bn}.running_mean", (c_out,), "bn_m"),
            (f"{prefix}.{bn}.running_var", (c_out,), "bn_v"),
            (f"{prefix}.{bn}.num_batches_tracked", (), "bn_n")]


def _attn(prefix, c):
    return [(f"{prefix}.gamma", (1,), "gamma"),
            (f"{prefix}.query_conv.weight", (c // 8, c, 1), "w"), (f"{prefix}.query_conv.bias", (c // 8,), "b"),
            (f"{prefix}.key_conv.weight", (c // 8, c, 1), "w"), (f"{prefix}.key_conv.bias", (c // 8,), "b"),
            (f"{prefix}.value_conv.weight", (c, c, 1), "w"), (f"{prefix}.value_conv.bias", (c,), "b")]


def _chan(prefix, c, r=8):
    return [(f"{prefix}.fc.0.weight", (c // r, c), "w"), (f"{prefix}.fc.0.bias", (c // r,), "b"),
            (f"{prefix}.fc.2.weight", (c, c // r), "w"), (f"{prefix}.fc.2.bias", (c,), "b")]


def _res(prefix, c):
    return _cnr(f"{prefix}.conv1", c, c, 3) + _cnr(f"{prefix}.conv2", c, c, 3) + _attn(f"{prefix}.attention", c)


def _gat(prefix, f=JOINT_FEAT, h=GAT_HEADS):
    return [(f"{prefix}.att_src", (1, h, f), "att"), (f"{prefix}.att_dst", (1, h, f), "att"),
            (f"{prefix}.bias", (f,), "b"), (f"{prefix}.lin.weight", (h * f, f), "w")]


def _gconv(prefix, f=JOINT_FEAT):
    return [(f"{prefix}.lin_rel.weight", (f, f), "w"), (f"{prefix}.lin_rel.bias", (f,), "b"),
            (f"{prefix}.lin_root.weight", (f, f), "w")]


def contract(in_channels=256, out_channels=256, out_feats=104):
    """Ordered [(name, shape, kind)] of every state_dict entry of SelfAttention_G."""
    c, oc = in_channels, out_channels
    e = [("body_edge_index_template", (2, 18), "edge_body"),
         ("hand_edge_index_template", (2, 80), "edge_hand")]
    enc = [(64, 1, 4, 4), (128, 64, 4, 4), (256, 128, 4, 4), (512, 256, 3, 3), (256, 512, 3, 8)]
    for i, (co, ci, kh, kw) in enumerate(enc):
        e += _cnr(f"audio_encoder.conv.{i}", co, ci, kh, kw)
    ds = [(2 * c, c, 3), (2 * c, 2 * c, 4), (4 * c, 2 * c, 3), (4 * c, 4 * c, 4)]
    for i, (co, ci, k) in enumerate(ds):
        e += _cnr(f"unet.downsample_layers.{i}", co, ci, k)
    e += _cnr("unet.upsample_layers.0", 4 * c, 8 * c, 3, bn="bn", conv="conv_transpose", transpose=True)
    e += _cnr("unet.upsample_layers.1", 4 * c, 8 * c, 3)
    e += _cnr("unet.upsample_layers.2", 2 * c, 4 * c, 3, bn="bn", conv="conv_transpose", transpose=True)
    e += _cnr("unet.upsample_layers.3", 2 * c, 4 * c, 3)
    e += _cnr("unet.bottleneck", 8 * c, 4 * c, 3)
    e += [("unet.final_conv.weight", (oc, 2 * c, 1), "w"), ("unet.final_conv.bias", (oc,), "b")]
    e += _attn("unet.bottleneck_attention", 8 * c) + _attn("unet.up_attention", 4 * c)
    body_f = 20
    for part, nj, nf in (("body", N_BODY, body_f), ("hand", N_HAND, out_feats - body_f)):
        pre = f"{part}_decoder_pre"
        e += _res(f"{pre}.0", oc) + _cnr(f"{pre}.1", oc, oc, 3)
        if part == "body":        # real_motion_model.py:70-75 vs :96-101 (order differs)
            e += _chan(f"{pre}.2", oc) + _attn(f"{pre}.3", oc)
        else:
            e += _attn(f"{pre}.2", oc) + _chan(f"{pre}.3", oc)
        e += [(f"{part}_proj_in.weight", (nj * JOINT_FEAT, oc), "w"), (f"{part}_proj_in.bias", (nj * JOINT_FEAT,), "b")]
        for li in range(1, 6):
            e += _gat(f"{part}_gcn{li}") if li % 2 == 1 else _gconv(f"{part}_gcn{li}")
        for li in range(5):
            e += [(f"{part}_layer_norms.{li}.weight", (JOINT_FEAT,), "ln_w"),
                  (f"{part}_layer_norms.{li}.bias", (JOINT_FEAT,), "ln_b")]
        e += [(f"{part}_proj_out.weight", (oc, nj * JOINT_FEAT), "w"), (f"{part}_proj_out.bias", (oc,), "b"),
              (f"{part}_norm.weight", (oc,), "ln_w"), (f"{part}_norm.bias", (oc,), "ln_b")]
        post = f"{part}_decoder_post"
        e += _res(f"{post}.0", oc) + _cnr(f"{post}.1", oc, oc, 3) + _attn(f"{post}.2", oc)
        if part == "hand":
            e += _chan(f"{post}.3", oc)
        e += [(f"{part}_logits.weight", (nf, oc, 1), "w"), (f"{part}_logits.bias", (nf,), "b")]
    return e


def _conv_bn(prefix, i, c_out, c_in, k):
    """nn.Sequential entries i (Conv1d) and i + 1 (BatchNorm1d) of the discriminator's conv stacks."""
    return [(f"{prefix}.{i}.weight", (c_out, c_in, k), "w"), (f"{prefix}.{i}.bias", (c_out,), "b"),
            (f"{prefix}.{i + 1}.weight", (c_out,), "bn_w"), (f"{prefix}.{i + 1}.bias", (c_out,), "bn_b"),
            (f"{prefix}.{i + 1}.running_mean", (c_out,), "bn_m"), (f"{prefix}.{i + 1}.running_var", (c_out,), "bn_v"),
            (f"{prefix}.{i + 1}.num_batches_tracked", (), "bn_n")]


def disc_contract(in_channels=104, out_channels=64, n_downsampling=2, aux_classes=10):
    """Ordered [(name, shape, kind)] of every state_dict entry of SelfAttention_D with groups = 1
    (real_motion_model.py:464-577)."""
    e = [("body_edge_index_template", (2, 18), "edge_body"), ("hand_edge_index_template", (2, 80), "edge_hand")]
    c = out_channels
    e += _conv_bn("conv1", 0, c, in_channels, 4) + _conv_bn("conv1", 4, c, c, 4)
    for n in range(1, n_downsampling + 1):
        mul = min(2 ** n, 16)
        e += _conv_bn(f"conv2.{n - 1}", 0, c * mul, c, 4) + _conv_bn(f"conv2.{n - 1}", 4, c * mul, c * mul, 4)
        c *= mul
    e += _conv_bn("conv3", 0, 2 * c, c, 4) + _conv_bn("conv3", 4, 4 * c, 2 * c, 4) + _attn("conv3.8", 4 * c)
    e += _conv_bn("conv3", 9, 4 * c, 4 * c, 3)
    e += [("body_proj.weight", (N_BODY * JOINT_FEAT, 2 * c), "w"), ("body_proj.bias", (N_BODY * JOINT_FEAT,), "b"),
          ("hand_proj.weight", (N_HAND * JOINT_FEAT, 2 * c), "w"), ("hand_proj.bias", (N_HAND * JOINT_FEAT,), "b")]
    e += _gat("body_gat") + _gat("hand_gat")
    e += [("body_graph_out.weight", (2 * c, N_BODY * JOINT_FEAT), "w"), ("body_graph_out.bias", (2 * c,), "b"),
          ("hand_graph_out.weight", (2 * c, N_HAND * JOINT_FEAT), "w"), ("hand_graph_out.bias", (2 * c,), "b"),
          ("audio_fusion.weight", (4 * c, 256, 1), "w"), ("audio_fusion.bias", (4 * c,), "b"),
          ("logits.weight", (1, 8 * c, 3), "w"), ("logits.bias", (1,), "b"),
          ("aux_classifier.0.weight", (512, 4 * c), "w"), ("aux_classifier.0.bias", (512,), "b"),
          ("aux_classifier.3.weight", (aux_classes, 512), "w"), ("aux_classifier.3.bias", (aux_classes,), "b")]
    return e


def _fan_in(shape, name):
    if "conv_transpose" in name:             # weight is [C_in, C_out, k]
        return shape[1] * math.prod(shape[2:])
    return math.prod(shape[1:]) if len(shape) > 1 else shape[0]


def make_state_dict(seed=0, mode="stress", **kw):
    """Seeded state_dict.

    mode="default": torch-style U(-1/sqrt(fan_in), 1/sqrt(fan_in)) conv/linear weights and biases,
                    identity BatchNorm/LayerNorm, gamma = 0  (what a fresh reference model looks like:
                    every SelfAttention is the identity, model_layers.py:130).
    mode="stress" : same weights, but BatchNorm running stats / affine and LayerNorm affine are
                    randomised and every gamma = 0.5 so BN folding, D1 and the attention paths are
                    observable.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    body_e, hand_e = edge_templates()
    sd = OrderedDict()
    stress = mode == "stress"
    entries = disc_contract() if kw.pop("discriminator", False) else contract(**kw)
    for name, shape, kind in entries:
        if kind == "edge_body":
            t = body_e.clone()
        elif kind == "edge_hand":
            t = hand_e.clone()
        elif kind == "bn_n":
            t = torch.tensor(0, dtype=torch.long)
        elif kind in ("w", "att"):
            bound = 1.0 / math.sqrt(_fan_in(shape, name))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "b":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind in ("bn_w", "ln_w"):
            t = 0.8 + 0.4 * torch.rand(shape, generator=g) if stress else torch.ones(shape)
        elif kind in ("bn_b", "ln_b"):
            t = 0.1 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        elif kind == "bn_m":
            t = 0.1 * torch.randn(shape, generator=g) if stress else torch.zeros(shape)
        elif kind == "bn_v":
            t = 0.5 + torch.rand(shape, generator=g) if stress else torch.ones(shape)
        elif kind == "gamma":
            t = torch.full(shape, 0.5 if stress else 0.0)
        else:
            raise KeyError(kind)
        sd[name] = t
    return sd


def num_parameters(sd):
    return sum(v.numel() for k, v in sd.items()
               if v.dtype.is_floating_point and "running_" not in k)
