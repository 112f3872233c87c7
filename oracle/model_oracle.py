"""Oracle (test infrastructure): torch-fp32 functional restatement of ``SelfAttention_G.forward``
in eval mode (decision D3), driven by a plain state_dict (see oracle/weights.py).

Reference followed (all under /root/reference):
  conv_norm_act       <- ConvNormRelu.forward          model_layers.py:112-118 (+ ctor :60-110)
  self_attention      <- SelfAttention.forward         model_layers.py:133-146
  channel_attention   <- ChannelAttention.forward      model_layers.py:167-174
  res_block           <- ResBlock.forward              model_layers.py:185-190
  conv_transpose_block<- ConvTranspose1D.forward       model_layers.py:211-215
  audio_encoder       <- AudioEncoder.forward          model_layers.py:267-280
  unet                <- UNet1D.forward                model_layers.py:341-374  (with D1, below)
  gat / graph_conv    <- torch_geometric GATConv / GraphConv as called at
                         real_motion_model.py:78-82,104-108,172-201,224-253
                         (third-party, absent, version unpinned -> restated from PyG's documented
                         semantics; PARITY UNPINNED at this boundary)
  decoder / generator <- SelfAttention_G.forward       real_motion_model.py:154-278
  bone / angle losses <- real_motion_model.py:307-461

D1: ``up_attention`` (declared 1024 channels, model_layers.py:339) is applied to the
ConvTranspose output *before* the skip concat; the shipped order (:364-365) raises.
"""
import math
import torch
import torch.nn.functional as F

from .weights import PARENTS, N_BODY, N_HAND, JOINT_FEAT, GAT_HEADS, edge_templates

BN_EPS = 1e-5
LN_EPS = 1e-5
SLOPE = 0.2


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], training=False, eps=BN_EPS)


def conv_norm_act(sd, p, x, stride=1, padding=1, leaky=True):
    w, b = sd[p + ".conv.weight"], sd[p + ".conv.bias"]
    y = F.conv2d(x, w, b, stride, padding) if w.dim() == 4 else F.conv1d(x, w, b, stride, padding)
    y = _bn(sd, p + ".norm", y)                 # dropout is the identity in eval mode
    return F.leaky_relu(y, SLOPE) if leaky else F.relu(y)


def self_attention(sd, p, x):
    q = F.conv1d(x, sd[p + ".query_conv.weight"], sd[p + ".query_conv.bias"]).transpose(1, 2)  # [B,T,C/8]
    k = F.conv1d(x, sd[p + ".key_conv.weight"], sd[p + ".key_conv.bias"])                      # [B,C/8,T]
    v = F.conv1d(x, sd[p + ".value_conv.weight"], sd[p + ".value_conv.bias"])                  # [B,C,T]
    att = torch.softmax(torch.bmm(q, k), dim=-1)        # no 1/sqrt(d) scaling in the reference
    out = torch.bmm(att, v.transpose(1, 2)).transpose(1, 2)
    return sd[p + ".gamma"] * out + x


def channel_attention(sd, p, x):
    def mlp(z):
        z = F.relu(F.linear(z, sd[p + ".fc.0.weight"], sd[p + ".fc.0.bias"]))
        return torch.sigmoid(F.linear(z, sd[p + ".fc.2.weight"], sd[p + ".fc.2.bias"]))
    att = mlp(x.mean(dim=-1)) + mlp(x.amax(dim=-1))     # sigmoid on each branch, then add
    return x * att.unsqueeze(-1)


def res_block(sd, p, x):
    y = conv_norm_act(sd, p + ".conv1", x)
    y = conv_norm_act(sd, p + ".conv2", y)
    return self_attention(sd, p + ".attention", y) + x


def conv_transpose_block(sd, p, x):
    y = F.conv_transpose1d(x, sd[p + ".conv_transpose.weight"], sd[p + ".conv_transpose.bias"],
                           stride=2, padding=1, output_padding=1)
    return F.relu(_bn(sd, p + ".bn", y))


def audio_encoder(sd, mel, time_steps=None, p="audio_encoder"):
    """mel [B,T,F] -> [B,256,time_steps]."""
    if time_steps is None:
        time_steps = mel.shape[-2]
    x = mel.unsqueeze(1)
    for i, (s, pad) in enumerate([(2, 1), (2, 1), (2, 1), (1, 1), (1, (1, 3))]):
        x = conv_norm_act(sd, f"{p}.conv.{i}", x, stride=s, padding=pad)
    x = F.interpolate(x, size=(time_steps, 1), mode="bilinear")
    return x.squeeze(-1)


def unet(sd, x, p="unet"):
    d = p + ".downsample_layers"
    u = p + ".upsample_layers"
    s0 = conv_norm_act(sd, d + ".0", x)
    x = conv_norm_act(sd, d + ".1", s0, stride=2, padding=1)
    s1 = conv_norm_act(sd, d + ".2", x)
    x = conv_norm_act(sd, d + ".3", s1, stride=2, padding=1)
    x = conv_norm_act(sd, p + ".bottleneck", x)
    x = self_attention(sd, p + ".bottleneck_attention", x)
    x = conv_transpose_block(sd, u + ".0", x)
    x = self_attention(sd, p + ".up_attention", x)          # D1: before the concat
    x = conv_norm_act(sd, u + ".1", torch.cat([x, s1], dim=1))
    x = conv_transpose_block(sd, u + ".2", x)
    x = conv_norm_act(sd, u + ".3", torch.cat([x, s0], dim=1))
    return F.conv1d(x, sd[p + ".final_conv.weight"], sd[p + ".final_conv.bias"])


def dense_adjacency(edge_index, n):
    """adj[i, j] = 1 iff there is an edge j -> i (row0 = source, row1 = target)."""
    adj = torch.zeros(n, n)
    adj[edge_index[1], edge_index[0]] = 1.0
    return adj


def gat(sd, p, x, adj):
    """GATConv(64, 64, heads=4, concat=False) on a batch of identical graphs, x [G,J,64].
    PyG defaults: shared linear (no bias), self loops added, LeakyReLU(0.2) on the logits,
    softmax over incoming edges, mean over heads, + bias."""
    G, J, Fd = x.shape
    h = F.linear(x, sd[p + ".lin.weight"]).view(G, J, GAT_HEADS, Fd)
    a_src = (h * sd[p + ".att_src"]).sum(-1)                 # [G,J,H]
    a_dst = (h * sd[p + ".att_dst"]).sum(-1)
    mask = ((adj + torch.eye(J)) > 0)                        # self loops
    e = F.leaky_relu(a_dst.unsqueeze(2) + a_src.unsqueeze(1), SLOPE)     # [G,i,j,H]
    e = e.masked_fill(~mask[None, :, :, None], float("-inf"))
    alpha = torch.softmax(e, dim=2)
    out = torch.einsum("gijh,gjhf->gihf", alpha, h).mean(dim=2)
    return out + sd[p + ".bias"]


def graph_conv(sd, p, x, adj):
    """GraphConv(64, 64), aggr='add': lin_rel(sum_{j->i} x_j) + lin_root(x_i)."""
    agg = torch.einsum("ij,gjf->gif", adj, x)
    return F.linear(agg, sd[p + ".lin_rel.weight"], sd[p + ".lin_rel.bias"]) + \
        F.linear(x, sd[p + ".lin_root.weight"])


def gnn_stack(sd, part, x):
    """The five graph layers of one branch (real_motion_model.py:172-201 / :224-253): GAT, GraphConv, GAT,
    GraphConv, GAT, each followed by LayerNorm(64) -> LeakyReLU(0.2) -> + residual.  x [G, J, 64] -> same."""
    nj = x.shape[1]
    adj = dense_adjacency(sd[f"{part}_edge_index_template"], nj)
    for li in range(1, 6):
        layer = gat if li % 2 == 1 else graph_conv
        y = layer(sd, f"{part}_gcn{li}", x, adj)
        y = F.layer_norm(y, (JOINT_FEAT,), sd[f"{part}_layer_norms.{li - 1}.weight"],
                         sd[f"{part}_layer_norms.{li - 1}.bias"], LN_EPS)
        x = F.leaky_relu(y, SLOPE) + x
    return x


def decoder(sd, part, feats):
    """One of the two branches of real_motion_model.py:160-262; feats [B,256,T] -> [B,n_feat,T]."""
    nj = N_BODY if part == "body" else N_HAND
    pre, post = f"{part}_decoder_pre", f"{part}_decoder_post"
    x = res_block(sd, pre + ".0", feats)
    x = conv_norm_act(sd, pre + ".1", x)
    if part == "body":
        x = self_attention(sd, pre + ".3", channel_attention(sd, pre + ".2", x))
    else:
        x = channel_attention(sd, pre + ".3", self_attention(sd, pre + ".2", x))
    B, C, T = x.shape
    x = F.linear(x.permute(0, 2, 1), sd[f"{part}_proj_in.weight"], sd[f"{part}_proj_in.bias"])
    x = x.reshape(B * T, nj, JOINT_FEAT)
    x = gnn_stack(sd, part, x)
    x = x.reshape(B, T, nj * JOINT_FEAT)
    x = F.linear(x, sd[f"{part}_proj_out.weight"], sd[f"{part}_proj_out.bias"])
    x = F.layer_norm(x, (C,), sd[f"{part}_norm.weight"], sd[f"{part}_norm.bias"], LN_EPS)
    x = x.permute(0, 2, 1)
    x = res_block(sd, post + ".0", x)
    x = conv_norm_act(sd, post + ".1", x)
    x = self_attention(sd, post + ".2", x)
    if part == "hand":
        x = channel_attention(sd, post + ".3", x)
    return F.conv1d(x, sd[f"{part}_logits.weight"], sd[f"{part}_logits.bias"])


def angle_triples():
    """(hand, body) triples (parent, joint, first child), real_motion_model.py:280-304."""
    hand, body = [], []
    for i in range(N_HAND):
        par = PARENTS[i + 10] - 10 if PARENTS[i + 10] >= 10 else -1
        if par != -1:
            for j in range(i + 1, N_HAND):
                if PARENTS[j + 10] - 10 == i:
                    hand.append((par, i, j))
                    break
    for i in range(N_BODY):
        par = PARENTS[i] if PARENTS[i] < N_BODY else -1
        if par != -1:
            for j in range(i + 1, N_BODY):
                if PARENTS[j] == i:
                    body.append((par, i, j))
                    break
    return hand, body


def _angles(joints, triples):
    out = []
    for p, j, c in triples:
        a = joints[:, :, j, :] - joints[:, :, p, :]
        b = joints[:, :, c, :] - joints[:, :, j, :]
        dot = (a * b).sum(-1)
        cross = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
        out.append(torch.atan2(cross, dot))
    return torch.stack(out, dim=-1)


def angle_loss(pose):
    """0.7 * hand + 0.3 * body range penalties on the *interleaved* [B,T,52,2] view
    (real_motion_model.py:359-461; the xy interleave is the reference's own)."""
    B, T, _ = pose.shape
    j = pose.reshape(B, T, len(PARENTS), 2)
    hand_t, body_t = angle_triples()
    th = _angles(j[:, :, 10:52, :], hand_t)
    hand = (torch.relu(0.0 - th) + torch.relu(th - math.pi)).mean()
    tb = _angles(j[:, :, :10, :], body_t)
    body = (torch.relu(-math.pi / 2 - tb) + torch.relu(tb - math.pi)).mean()
    return 0.7 * hand + 0.3 * body


def bone_loss(real_pose, gen_pose):
    """MSE between time-averaged bone lengths (real_motion_model.py:307-347); joint_subset is
    all 52 joints (:124), so the parent remap is the identity."""
    B, T, _ = real_pose.shape

    def lengths(p):
        p = p.reshape(B, T, len(PARENTS), 2)
        segs = [torch.norm(p[:, :, i, :] - p[:, :, par, :], dim=-1)
                for i, par in enumerate(PARENTS) if par != -1]
        return torch.stack(segs, dim=-1).mean(dim=1)
    return F.mse_loss(lengths(gen_pose), lengths(real_pose))


@torch.no_grad()
def generator_forward(sd, mel, real_pose=None):
    """mel [B,T,F] fp32 -> (pose [B,T,104] fp32, [losses]) == SelfAttention_G(...).eval()(mel)."""
    if mel.shape[1] % 4 != 0:
        raise ValueError("time steps must be a multiple of 4 (UNet1D skip concat)")
    if "body_edge_index_template" not in sd:
        sd = dict(sd)
        sd["body_edge_index_template"], sd["hand_edge_index_template"] = edge_templates()
    feats = unet(sd, audio_encoder(sd, mel))
    out = torch.cat([decoder(sd, "body", feats), decoder(sd, "hand", feats)], dim=1).transpose(1, 2)
    out = out.contiguous()
    losses = []
    if real_pose is not None:
        losses.append(bone_loss(real_pose, out))
    losses.append(angle_loss(out))
    return out, losses


# ---------------------------------------------------------------------------------------------------------------
# discriminator
# ---------------------------------------------------------------------------------------------------------------
def _conv_bn_act(sd, p, i, x, stride, padding):
    """nn.Sequential entries i (Conv1d), i + 1 (BatchNorm1d), LeakyReLU(0.2), Dropout (identity in eval mode)."""
    y = F.conv1d(x, sd[f"{p}.{i}.weight"], sd[f"{p}.{i}.bias"], stride, padding)
    return F.leaky_relu(_bn(sd, f"{p}.{i + 1}", y), SLOPE)


def discriminator_forward(sd, pose, n_downsampling=2):
    """SelfAttention_D.forward(x) with audio = None, aux_labels = None (real_motion_model.py:580-642; the two optional
    arguments cannot work as shipped: with audio the concat has 6144 channels but `logits` takes 4096, and the aux
    classifier is fed a [B] tensor).  pose [B, T, 104] -> scores [B, T']."""
    x = pose.transpose(-1, -2)                                               # :582
    if x.size(2) < 4:
        x = F.pad(x, (0, 4 - x.size(2) % 4))                                 # :583-584
    x = _conv_bn_act(sd, "conv1", 0, x, 2, 1)                                # :586, ctor :504-513
    x = _conv_bn_act(sd, "conv1", 4, x, 1, 1)
    for n in range(n_downsampling):                                          # :587-588, ctor :518-532
        x = _conv_bn_act(sd, f"conv2.{n}", 0, x, 2, 1)
        x = _conv_bn_act(sd, f"conv2.{n}", 4, x, 1, 1)
    x = _conv_bn_act(sd, "conv3", 0, x, 1, 1)                                # :590, ctor :535-550
    x = _conv_bn_act(sd, "conv3", 4, x, 1, 1)
    x = self_attention(sd, "conv3.8", x)
    x = _conv_bn_act(sd, "conv3", 9, x, 1, 1)
    B, C, T = x.shape
    body_e, hand_e = edge_templates()
    outs = []
    for part, nj, e, feats in (("body", N_BODY, body_e, x[:, :C // 2]), ("hand", N_HAND, hand_e, x[:, C // 2:])):
        h = F.linear(feats.mean(dim=2), sd[f"{part}_proj.weight"], sd[f"{part}_proj.bias"])            # :599-600
        h = gat(sd, f"{part}_gat", h.view(B, nj, JOINT_FEAT), dense_adjacency(e, nj))                  # :601-604
        outs.append(F.linear(h.reshape(B, -1), sd[f"{part}_graph_out.weight"], sd[f"{part}_graph_out.bias"]))
    xg = torch.cat(outs, dim=1).unsqueeze(2).repeat(1, 1, T)                                            # :619-621
    x = torch.cat([x, xg], dim=1)
    x = F.conv1d(x, sd["logits.weight"], sd["logits.bias"], 1, 1)                                       # :630
    return x.transpose(-1, -2).squeeze(dim=-1)
