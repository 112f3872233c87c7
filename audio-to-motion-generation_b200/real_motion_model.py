"""B200 drop-in for ``real_motion_model.py`` of the reference: the ``SelfAttention_G`` generator and the
``SelfAttention_D`` discriminator (inference forward).

Same constructor and ``forward(audio, real_pose=None) -> (pose [B, T, 104], [losses])`` contract
(real_motion_model.py:22,154-278) and the same 340 state_dict keys, so reference
checkpoints load with ``load_state_dict``.  The forward is one native launch program (csrc/model.cu):
tcgen05 implicit-GEMM convolutions / linears in bf16 with fp32 accumulation, fused attention, the
static-skeleton GAT / GraphConv layers (restated from torch_geometric's documented semantics -- PyG is
not a dependency here), and the angle / bone losses.

Differences from the reference, all deliberate: inference (eval) semantics only (D3);
``UNet1D.up_attention`` before the skip concat (D1); no Skeleton2D / CSV side effects -- the skeleton is
the constant parent list of pats/data_loading/skeleton.py:94-110 (D5).
"""
import math

import torch
import torch.nn as nn

from . import _cabi
from ._native_module import NativeModule, as_input
from .model_layers import (AudioEncoder, UNet1D, ResBlock, ConvNormRelu, ChannelAttention, SelfAttention,
                           _Block)

# pats/data_loading/skeleton.py:94-110
SKELETON_PARENTS = [-1, 0, 1, 2, 0, 4, 5, 0, 7, 7, 6,
                    10, 11, 12, 13, 10, 15, 16, 17, 10, 19, 20, 21, 10, 23, 24, 25, 10, 27, 28, 29,
                    3,
                    31, 32, 33, 34, 31, 36, 37, 38, 31, 40, 41, 42, 31, 44, 45, 46, 31, 48, 49, 50]
SKELETON_JOINT_NAMES = (
    ['Neck', 'RShoulder', 'RElbow', 'RWrist', 'LShoulder', 'LElbow', 'LWrist', 'Nose', 'REye', 'LEye'] +
    [side + 'Hand' + part for side in 'LR'
     for part in ['Root'] + [f + str(i) for f in ('Thumb', 'Index', 'Middle', 'Ring', 'Little') for i in range(1, 5)]])


class _Lin(nn.Module):
    def __init__(self, cin, cout, bias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        nn.init.xavier_uniform_(self.weight)
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None


class GATConv(_Block):
    """Parameter container with torch_geometric's GATConv key names (att_src, att_dst, bias, lin.weight);
    semantics implemented natively: heads averaged (concat=False), self loops, LeakyReLU(0.2) logits."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True):
        super().__init__()
        if concat or (in_channels, out_channels, heads) != (64, 64, 4):
            raise NotImplementedError("the native path implements GATConv(64, 64, heads=4, concat=False)")
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin = _Lin(in_channels, heads * out_channels, bias=False)
        nn.init.xavier_uniform_(self.att_src)
        nn.init.xavier_uniform_(self.att_dst)


class GraphConv(_Block):
    """torch_geometric GraphConv key names (lin_rel.{weight,bias}, lin_root.weight), aggr='add'."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_rel = _Lin(in_channels, out_channels, bias=True)
        self.lin_root = _Lin(in_channels, out_channels, bias=False)


def _edge_index(parents, lo, count):
    """Undirected parent/child edges among joints [lo, lo+count) re-indexed from 0
    (real_motion_model.py:43-60); row 0 = source, row 1 = target."""
    edges = []
    for i, par in enumerate(parents[lo:lo + count]):
        par -= lo
        if 0 <= par < count:
            edges += [[par, i], [i, par]]
    return torch.tensor(edges, dtype=torch.long).t().contiguous()


class SelfAttention_G(NativeModule):
    '''
    input_shape:  (N, time, frequency)
    output_shape: (N, time, pose_feats)
    '''

    def __init__(self, time_steps=64, in_channels=256, out_channels=256, out_feats=104, p=0.2):
        super().__init__()
        if (in_channels, out_channels, out_feats) != (256, 256, 104):
            raise NotImplementedError("the native generator implements in/out channels 256 and 104 pose features")
        self.audio_encoder = AudioEncoder(output_feats=time_steps, p=p)
        self.unet = UNet1D(input_channels=in_channels, output_channels=out_channels, p=p)
        self.body_feats, self.hand_feats = 20, out_feats - 20
        self.num_body_joints, self.num_hand_joints, self.joint_feat_dim = 10, 42, 64
        self.parents, self.joint_names = list(SKELETON_PARENTS), list(SKELETON_JOINT_NAMES)
        self.joint_subset = list(range(out_feats // 2))
        self.body_edge_index = _edge_index(self.parents, 0, self.num_body_joints)
        self.hand_edge_index = _edge_index(self.parents, 10, self.num_hand_joints)
        self.register_buffer('body_edge_index_template', self.body_edge_index)
        self.register_buffer('hand_edge_index_template', self.hand_edge_index)
        for part, joints, feats in (("body", self.num_body_joints, self.body_feats),
                                    ("hand", self.num_hand_joints, self.hand_feats)):
            pre = [ResBlock(out_channels, type='1d', p=p),
                   ConvNormRelu(out_channels, out_channels, type='1d', leaky=True, downsample=False, p=p)]
            pre += [ChannelAttention(out_channels), SelfAttention(out_channels)] if part == "body" else \
                [SelfAttention(out_channels), ChannelAttention(out_channels)]
            setattr(self, part + "_decoder_pre", nn.Sequential(*pre))
            setattr(self, part + "_proj_in", nn.Linear(out_channels, joints * self.joint_feat_dim))
            for li in range(1, 6):
                layer = GATConv(64, 64, heads=4, concat=False) if li % 2 else GraphConv(64, 64)
                setattr(self, "%s_gcn%d" % (part, li), layer)
            setattr(self, part + "_layer_norms", nn.ModuleList([nn.LayerNorm(64) for _ in range(5)]))
            setattr(self, part + "_relu", nn.LeakyReLU(0.2))
            setattr(self, part + "_dropout", nn.Dropout(p=p))
            setattr(self, part + "_proj_out", nn.Linear(joints * self.joint_feat_dim, out_channels))
            setattr(self, part + "_norm", nn.LayerNorm(out_channels))
            post = [ResBlock(out_channels, type='1d', p=p),
                    ConvNormRelu(out_channels, out_channels, type='1d', leaky=True, downsample=False, p=p),
                    SelfAttention(out_channels)]
            if part == "hand":
                post.append(ChannelAttention(out_channels))
            setattr(self, part + "_decoder_post", nn.Sequential(*post))
            setattr(self, part + "_logits", nn.Conv1d(out_channels, feats, kernel_size=1, stride=1))
        self.hand_triples = self._triples(10, self.num_hand_joints)
        self.body_triples = self._triples(0, self.num_body_joints)

    def _triples(self, lo, count):
        """(parent, joint, first child) per joint that has both (real_motion_model.py:280-304)."""
        local = [p - lo if lo <= p < lo + count else -1 for p in self.parents[lo:lo + count]]
        out = []
        for i, par in enumerate(local):
            if par == -1:
                continue
            child = next((j for j in range(i + 1, count) if local[j] == i), None)
            if child is not None:
                out.append((par, i, child))
        return out

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # torch_geometric renamed GATConv's linear over versions (lin / lin_src+lin_dst / lin_l+lin_r)
        for part in ("body", "hand"):
            for li in (1, 3, 5):
                base = "%s%s_gcn%d." % (prefix, part, li)
                for alt in ("lin_src.weight", "lin_l.weight"):
                    if base + alt in state_dict and base + "lin.weight" not in state_dict:
                        state_dict[base + "lin.weight"] = state_dict.pop(base + alt)
                for dup in ("lin_dst.weight", "lin_r.weight"):
                    state_dict.pop(base + dup, None)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def graph_stack(self, part, x):
        """The five fused graph layers of one branch on their own (real_motion_model.py:172-201 / :224-253):
        x [n_graphs, J, 64] fp32 -> same shape.  part: "body" (J = 10) or "hand" (J = 42).  Diagnostic / test
        surface of csrc/gnn_fused.cu (a2m_model_gnn_forward)."""
        self._require_eval()
        h = self.native()
        joints = {"body": 10, "hand": 42}[part]
        x = torch.as_tensor(x).to(device=h.device, dtype=torch.float32).contiguous()
        if x.dim() != 3 or x.shape[1] != joints or x.shape[2] != 64:
            raise ValueError("graph_stack(%r) expects [n_graphs, %d, 64], got %s" % (part, joints, tuple(x.shape)))
        out = torch.empty_like(x)
        with torch.cuda.device(h.device):
            _cabi.check(_cabi.lib().a2m_model_gnn_forward(h.ptr, 0 if part == "body" else 1, _cabi.ptr(x), x.shape[0],
                                                          _cabi.ptr(out), _cabi.stream_ptr(h.device)))
        return out

    def set_output_denorm(self, pose_mean=None, pose_std=None):
        """Fuse the de-normalisation that follows the generator in real use (generate_motion_video.py:259-260,
        ``pose * std + mean``) into the forward's output pass; ``set_output_denorm()`` switches it off.  The
        returned poses then equal ``normalization_tools.denormalize_pose(forward(x)[0], mean, std)`` bit for bit;
        the internal losses are still those of the normalised output."""
        if (pose_mean is None) != (pose_std is None):
            raise ValueError("give both pose_mean and pose_std, or neither")
        if pose_mean is None:
            self.__dict__["_denorm"] = None
        else:
            mean = torch.as_tensor(pose_mean).detach().to(dtype=torch.float32).reshape(-1).clone()
            std = torch.as_tensor(pose_std).detach().to(dtype=torch.float32).reshape(-1).clone()
            if mean.numel() != 104 or std.numel() != 104:
                raise ValueError("pose_mean and pose_std must have 104 elements")
            self.__dict__["_denorm"] = (mean, std)
        self.__dict__["_denorm_token"] = object()     # a handle carries the token of the setting it was given

    def _sync_denorm(self, h):
        token = self.__dict__.get("_denorm_token")
        if token is None or getattr(h, "denorm_token", None) is token:
            return
        d = self.__dict__["_denorm"]
        with torch.cuda.device(h.device):
            if d is None:
                _cabi.check(_cabi.lib().a2m_model_set_output_denorm(h.ptr, None, None, _cabi.stream_ptr(h.device)))
            else:
                mean, std = d[0].to(h.device), d[1].to(h.device)
                _cabi.check(_cabi.lib().a2m_model_set_output_denorm(h.ptr, _cabi.ptr(mean), _cabi.ptr(std),
                                                                    _cabi.stream_ptr(h.device)))
                mean.record_stream(torch.cuda.current_stream(h.device))
                std.record_stream(torch.cuda.current_stream(h.device))
        h.denorm_token = token

    def forward_windows(self, logmel, n_windows, frames=64, stride=6, window_hop=5, lane=0):
        """Sliding windows over long streams in ONE launch program: logmel [S, n_frames, F] (unit inner stride) -> pose
        [S, n_windows, frames, 104]; window w of stream s reads log-mel rows w * window_hop * stride + t * stride in place
        (the reference's window arithmetic, dataUtils.py:585-620,648-654; S * n_windows <= 65535)."""
        self._require_eval()
        h = self.native(lane)
        self._sync_denorm(h)
        x = as_input(logmel, h.device, "forward_windows expects log-mel [S, n_frames, F], got %s", keep_strides=True)
        S, n_frames, F = x.shape
        last = (n_windows - 1) * window_hop * stride + (frames - 1) * stride
        if n_windows < 1 or last >= n_frames:
            raise ValueError("%d windows of %d frames (stride %d, hop %d) need %d log-mel frames, got %d"
                             % (n_windows, frames, stride, window_hop, last + 1, n_frames))
        pose = torch.empty((S, n_windows, frames, 104), dtype=torch.float32, device=h.device)
        losses = torch.empty(2, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _cabi.check(_cabi.lib().a2m_model_forward_windows(
                h.ptr, _cabi.ptr(x), x.stride(0), window_hop * stride * x.stride(1), stride * x.stride(1), S, n_windows, frames, F,
                _cabi.ptr(pose), _cabi.ptr(losses), _cabi.stream_ptr(h.device)))
        return pose

    def forward(self, audio, real_pose=None, lane=0):
        """audio [B, T, F] (log-mel) -> (pose [B, T, 104] fp32, [angle_loss]) or
        (pose, [bone_loss, angle_loss]) when real_pose [B, T, 104] is given.  `lane` (an extension) selects one of
        the module's native handles, so that forwards enqueued on different CUDA streams do not share an arena."""
        self._require_eval()
        h = self.native(lane)
        self._sync_denorm(h)
        x = as_input(audio, h.device, "SelfAttention_G expects audio [B, T, F], got %s", keep_strides=True)
        B, T, F = x.shape
        if T % 4 != 0:
            raise ValueError("SelfAttention_G needs T %% 4 == 0 (UNet1D skip concats), got T = %d" % T)
        rp = None
        if real_pose is not None:
            rp = as_input(real_pose, h.device, "real_pose must be [B, T, 104], got %s")
            if tuple(rp.shape) != (B, T, 104):
                raise ValueError("real_pose must be [%d, %d, 104], got %s" % (B, T, tuple(rp.shape)))
        pose = torch.empty((B, T, 104), dtype=torch.float32, device=h.device)
        losses = torch.empty(2, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _cabi.check(_cabi.lib().a2m_model_forward(h.ptr, _cabi.ptr(x), x.stride(0), x.stride(1), B, T, F,
                                                      _cabi.ptr(pose), _cabi.ptr(losses), _cabi.ptr(rp),
                                                      _cabi.stream_ptr(h.device)))
        internal = [losses[1], losses[0]] if rp is not None else [losses[0]]
        return pose, internal



class SelfAttention_D(NativeModule):
    """The discriminator of the reference (real_motion_model.py:464-642), eval-mode forward on the B200 path:
    ``forward(x [B, T, 104]) -> (scores [B, T'], [])`` with T' = a2m_disc_out_length(T) (4 for T = 64).  Same constructor
    and state_dict keys (conv1.0.weight ... aux_classifier.3.bias), so reference checkpoints load.

    ``audio`` and ``aux_labels`` cannot work in the reference as shipped (with audio the concatenated tensor has 6144
    channels where ``logits`` takes 4096, :625-630; the auxiliary classifier is handed a [B] tensor, :637-638); here they
    raise NotImplementedError instead of a shape error.  ``audio_fusion`` / ``aux_classifier`` parameters are kept for
    checkpoint compatibility only."""

    def __init__(self, in_channels=104, out_channels=64, n_downsampling=2, p=0.3, groups=1, aux_classes=10, **kwargs):
        super().__init__()
        if (in_channels, out_channels, groups) != (104, 64, 1) or not 0 <= n_downsampling <= 2 or kwargs.get("out_shape", 1) != 1:
            raise NotImplementedError("the native discriminator implements in_channels 104, out_channels 64, groups 1, "
                                      "n_downsampling <= 2, out_shape 1")
        self.n_downsampling, self.groups, self.p = n_downsampling, groups, p
        self.num_body_joints, self.num_hand_joints, self.joint_feat_dim = 10, 42, 64
        self.body_edge_index = _edge_index(SKELETON_PARENTS, 0, 10)
        self.hand_edge_index = _edge_index(SKELETON_PARENTS, 10, 42)
        self.register_buffer('body_edge_index_template', self.body_edge_index)
        self.register_buffer('hand_edge_index_template', self.hand_edge_index)

        def stack(cin, cout, first_stride):
            return [nn.Conv1d(cin, cout, kernel_size=4, stride=first_stride, padding=1), nn.BatchNorm1d(cout),
                    nn.LeakyReLU(negative_slope=0.2), nn.Dropout(p=p)]
        c = out_channels
        self.conv1 = nn.Sequential(*(stack(in_channels, c, 2) + stack(c, c, 1)))
        self.conv2 = nn.ModuleList()
        for n in range(1, n_downsampling + 1):
            mul = min(2 ** n, 16)
            self.conv2.append(nn.Sequential(*(stack(c, c * mul, 2) + stack(c * mul, c * mul, 1))))
            c *= mul
        self.conv3 = nn.Sequential(*(stack(c, 2 * c, 1) + stack(2 * c, 4 * c, 1) + [SelfAttention(4 * c)] +
                                     [nn.Conv1d(4 * c, 4 * c, kernel_size=3, stride=1, padding=1), nn.BatchNorm1d(4 * c),
                                      nn.LeakyReLU(negative_slope=0.2), nn.Dropout(p=p)]))
        self.body_proj = nn.Linear(2 * c, 10 * 64)
        self.hand_proj = nn.Linear(2 * c, 42 * 64)
        self.body_gat = GATConv(64, 64, heads=4, concat=False)
        self.hand_gat = GATConv(64, 64, heads=4, concat=False)
        self.body_graph_out = nn.Linear(10 * 64, 2 * c)
        self.hand_graph_out = nn.Linear(42 * 64, 2 * c)
        self.audio_fusion = nn.Conv1d(256, 4 * c, kernel_size=1)
        self.logits = nn.Conv1d(8 * c, 1, kernel_size=3, stride=1, padding=1)
        self.aux_classifier = nn.Sequential(nn.Linear(4 * c, 512), nn.LeakyReLU(0.2), nn.Dropout(p), nn.Linear(512, aux_classes))
        self.aux_loss_fn = nn.CrossEntropyLoss()

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for part in ("body", "hand"):                                # torch_geometric version-dependent key names
            base = "%s%s_gat." % (prefix, part)
            for alt in ("lin_src.weight", "lin_l.weight"):
                if base + alt in state_dict and base + "lin.weight" not in state_dict:
                    state_dict[base + "lin.weight"] = state_dict.pop(base + alt)
            for dup in ("lin_dst.weight", "lin_r.weight"):
                state_dict.pop(base + dup, None)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def native(self, lane=0):
        _cabi.require_cuda(type(self).__name__)
        fp = self._fingerprint()
        handles = self.__dict__.setdefault("_handles", {})
        entry = handles.get(lane)
        if entry is None or entry[1] != fp:
            dev = next(self.parameters()).device
            if dev.type != "cuda":
                raise RuntimeError("SelfAttention_D: parameters are on %s; move the module to a CUDA device (.cuda()) -- "
                                   "there is no CPU fallback" % dev)
            entry = handles[lane] = (_DiscHandle(self._native_state(), dev, self.n_downsampling), fp)
        return entry[0]

    def forward(self, x, audio=None, aux_labels=None):
        if audio is not None or aux_labels is not None:
            raise NotImplementedError("SelfAttention_D: the audio / aux_labels branches raise shape errors in the reference "
                                      "as shipped (real_motion_model.py:625-639) and are not implemented")
        self._require_eval()
        h = self.native()
        x = as_input(x, h.device, "SelfAttention_D expects poses [B, T, 104], got %s")
        B, T, feats = x.shape
        if feats != 104:
            raise ValueError("SelfAttention_D expects 104 pose features, got %d" % feats)
        t_out = int(_cabi.lib().a2m_disc_out_length(T, self.n_downsampling))
        if t_out < 1:
            raise ValueError("SelfAttention_D: %d steps are too few for the convolution stack" % T)
        out = torch.empty((B, t_out), dtype=torch.float32, device=h.device)
        if B:
            with torch.cuda.device(h.device):
                _cabi.check(_cabi.lib().a2m_disc_forward(h.ptr, _cabi.ptr(x), B, T, _cabi.ptr(out), _cabi.stream_ptr(h.device)))
        return out, []


class _DiscHandle:
    """Owns the a2m_model* of a discriminator (a2m_disc_create)."""

    def __init__(self, state, device, n_downsampling):
        import ctypes
        descs = (_cabi.TensorDesc * len(state))()
        keep = []
        for i, (name, t) in enumerate(state.items()):
            dtype = 1 if t.dtype == torch.int64 else 0
            if not dtype and t.dtype != torch.float32:
                t = t.to(torch.float32)
            t = t.detach().to(device).contiguous()
            keep.append(t)
            descs[i].name, descs[i].data, descs[i].dtype, descs[i].ndim = name.encode(), t.data_ptr(), dtype, t.dim()
            for k, sdim in enumerate(t.shape):
                descs[i].shape[k] = sdim
        out = ctypes.c_void_p()
        with torch.cuda.device(device):
            _cabi.check(_cabi.lib().a2m_disc_create(descs, len(state), int(n_downsampling), device.index, ctypes.byref(out)))
        self.ptr, self.device = out, device

    def __del__(self):
        ptr, self.ptr = getattr(self, "ptr", None), None
        if ptr:
            try:
                _cabi.lib().a2m_model_destroy(ptr)
            except Exception:
                pass
