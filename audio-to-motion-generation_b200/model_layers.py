"""B200 drop-in for the hot-path subset of the reference's ``model_layers.py``.

The classes keep the reference's constructor signatures, attribute names and state_dict keys
(model_layers.py:53-118 ConvNormRelu, :125-131 SelfAttention, :155-165 ChannelAttention, :179-183
ResBlock, :198-209 ConvTranspose1D, :249-261 AudioEncoder, :303-339 UNet1D), so checkpoints written
by the reference load unchanged.  The arithmetic of a whole ``AudioEncoder`` / ``UNet1D`` / ``SelfAttention_G`` forward runs as one native launch program
in liba2m_b200.so (tcgen05 implicit-GEMM convolutions with folded BatchNorm + fused activations, csrc/conv_gemm.cu, and
the kernels of csrc/layers.cu).  The small building blocks (ConvNormRelu, SelfAttention, ChannelAttention, ResBlock,
ConvTranspose1D) also run on their own -- ``forward(x[B, C, T]) -> [B, C', T']`` like the reference's, eval semantics
(a2m_block_forward) -- on the same kernels, for the 1-D geometries the generator uses; anything else raises
NotImplementedError, never a PyTorch fallback.

Decision D1: ``UNet1D.up_attention`` (declared with 4*C channels, :339) is applied to the
ConvTranspose output *before* the skip concat; the reference's own forward (:364-365) raises.
"""
import torch
import torch.nn as nn

from . import _cabi
from ._native_module import NativeModule, as_input


BLOCK_CONV_K3, BLOCK_CONV_K4S2, BLOCK_CONV_TRANSPOSE, BLOCK_SELF_ATTENTION, BLOCK_CHANNEL_ATTENTION, BLOCK_RESBLOCK = range(1, 7)


class _Block(NativeModule):
    """A layer class with the reference's parameters and a native stand-alone forward.  Subclasses set
    ``self._block = (kind, in_channels, out_channels, leaky)`` when their geometry is one the kernels implement,
    or leave it None (then forward raises NotImplementedError; the module still works as a parameter container
    inside AudioEncoder / UNet1D / SelfAttention_G)."""

    _state_prefix = "blk."

    def forward(self, x, **kwargs):
        if self._block is None:
            raise NotImplementedError(
                "%s: this geometry has no stand-alone forward on the B200 path (1-D blocks with channel counts that are "
                "multiples of 64 do); run it through AudioEncoder, UNet1D or SelfAttention_G" % type(self).__name__)
        self._require_eval()
        h = self.native()
        kind, cin, cout, _ = self._block
        x = as_input(x, h.device, type(self).__name__ + " expects [B, C, T], got %s")
        B, C, T = x.shape
        if C != cin:
            raise ValueError("%s expects %d input channels, got %d" % (type(self).__name__, cin, C))
        if kind == BLOCK_CONV_K4S2 and T % 2:
            raise NotImplementedError("the stride-2 block needs an even number of steps on the native path, got %d" % T)
        t_out = T // 2 if kind == BLOCK_CONV_K4S2 else 2 * T if kind == BLOCK_CONV_TRANSPOSE else T
        out = torch.empty((B, cout, t_out), dtype=torch.float32, device=h.device)
        if B == 0 or T == 0:
            return out
        with torch.cuda.device(h.device):
            _cabi.check(_cabi.lib().a2m_block_forward(h.ptr, _cabi.ptr(x), B, T, _cabi.ptr(out), _cabi.stream_ptr(h.device)))
        return out


def _channels_ok(*cs):
    return all(c >= 64 and c % 64 == 0 for c in cs)


def _auto_padding(kernel_size, stride):
    """int((k - s) / 2) with the reference's per-dim rules (model_layers.py:68-82)."""
    if isinstance(kernel_size, int) and isinstance(stride, tuple):
        return tuple(int((kernel_size - st) / 2) for st in stride)
    if isinstance(kernel_size, tuple) and isinstance(stride, int):
        return tuple(int((ks - stride) / 2) for ks in kernel_size)
    if isinstance(kernel_size, tuple) and isinstance(stride, tuple):
        assert len(kernel_size) == len(stride), \
            'dims in kernel_size are {} and stride are {}. They must be the same'.format(len(kernel_size), len(stride))
        return tuple(int((ks - st) / 2) for ks, st in zip(kernel_size, kernel_size))
    return int((kernel_size - stride) / 2)


class ConvNormRelu(_Block):
    """conv -> dropout -> BatchNorm -> (Leaky)ReLU block (model_layers.py:51-118).  Stand-alone forward for the 1-D
    k3 s1 p1 and k4 s2 p1 (downsample) geometries."""

    def __init__(self, in_channels, out_channels, type='1d', leaky=False, downsample=False,
                 kernel_size=None, stride=None, padding=None, p=0, groups=1):
        super().__init__()
        if kernel_size is None and stride is None:
            kernel_size, stride = (4, 2) if downsample else (3, 1)
        if padding is None:
            padding = _auto_padding(kernel_size, stride)
        if groups != 1:
            raise NotImplementedError("grouped convolutions are not on the hot path (groups=%d)" % groups)
        conv, norm, drop = (nn.Conv1d, nn.BatchNorm1d, nn.Dropout) if type == '1d' else \
            (nn.Conv2d, nn.BatchNorm2d, nn.Dropout2d)
        self.conv = conv(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding)
        self.norm = norm(out_channels)
        self.dropout = drop(p=p)
        self.relu = nn.LeakyReLU(negative_slope=0.2) if leaky else nn.ReLU()
        if type == '1d' and _channels_ok(in_channels) and out_channels % 8 == 0:
            if (kernel_size, stride, padding) == (3, 1, 1):
                self._block = (BLOCK_CONV_K3, in_channels, out_channels, bool(leaky))
            elif (kernel_size, stride, padding) == (4, 2, 1):
                self._block = (BLOCK_CONV_K4S2, in_channels, out_channels, bool(leaky))


class SelfAttention(_Block):
    """q/k (C/8) and v (C) 1x1 convs, zero-initialised residual gate gamma."""

    def __init__(self, in_channels):
        super().__init__()
        self.query_conv = nn.Conv1d(in_channels, in_channels // 8, kernel_size=1)
        self.key_conv = nn.Conv1d(in_channels, in_channels // 8, kernel_size=1)
        self.value_conv = nn.Conv1d(in_channels, in_channels, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)
        if _channels_ok(in_channels):
            self._block = (BLOCK_SELF_ATTENTION, in_channels, in_channels, False)


class ChannelAttention(_Block):
    def __init__(self, channel, reduction=8):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool1d(1)
        self.max_pool = nn.AdaptiveMaxPool1d(1)
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel), nn.Sigmoid())
        if _channels_ok(channel) and channel <= 1024 and reduction == 8:
            self._block = (BLOCK_CHANNEL_ATTENTION, channel, channel, False)


class ResBlock(_Block):
    def __init__(self, channels, type='1d', p=0.1):
        super().__init__()
        self.conv1 = ConvNormRelu(channels, channels, type=type, leaky=True, p=p)
        self.conv2 = ConvNormRelu(channels, channels, type=type, leaky=True, p=p)
        self.attention = SelfAttention(channels)
        if type == '1d' and _channels_ok(channels):
            self._block = (BLOCK_RESBLOCK, channels, channels, True)


class ConvTranspose1D(_Block):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=2, padding=1, output_padding=1):
        super().__init__()
        if (kernel_size, stride, padding, output_padding) != (3, 2, 1, 1):
            raise NotImplementedError("the native path implements ConvTranspose1d(k=3, s=2, p=1, op=1) only")
        self.conv_transpose = nn.ConvTranspose1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                                 padding=padding, output_padding=output_padding)
        self.bn = nn.BatchNorm1d(out_channels)
        self.relu = nn.ReLU(inplace=True)
        if _channels_ok(in_channels) and out_channels % 8 == 0:
            self._block = (BLOCK_CONV_TRANSPOSE, in_channels, out_channels, False)


class AudioEncoder(NativeModule):
    """[B, T, F] log-mel -> [B, 256, time_steps]: five Conv2d blocks, then a bilinear resize to
    (time_steps, 1).  Only the centre frequency column of the last conv survives that resize, so the
    native path computes just that column."""

    _state_prefix = "audio_encoder."

    def __init__(self, output_feats=64, input_channels=1, kernel_size=None, stride=None, p=0, groups=1):
        super().__init__()
        if input_channels != 1 or kernel_size is not None or stride is not None:
            raise NotImplementedError("the native AudioEncoder implements the default geometry only")
        mk = dict(type='2d', leaky=True, p=p, groups=groups)
        self.conv = nn.ModuleList([
            ConvNormRelu(input_channels, 64, downsample=True, **mk),
            ConvNormRelu(64, 128, downsample=True, **mk),
            ConvNormRelu(128, 256, downsample=True, **mk),
            ConvNormRelu(256, 512, downsample=False, **mk),
            ConvNormRelu(512, 256, downsample=False, kernel_size=(3, 8), stride=1, **{k: v for k, v in mk.items()}),
        ])

    def forward(self, x, time_steps=None):
        self._require_eval()
        h = self.native()
        x = as_input(x, h.device, "AudioEncoder expects [B, T, F], got %s")
        B, T, F = x.shape
        steps = T if time_steps is None else int(time_steps)        # model_layers.py:268-269
        if steps < 1:
            raise ValueError("time_steps must be positive, got %d" % steps)
        out = torch.empty((B, 256, steps), dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _cabi.check(_cabi.lib().a2m_model_encoder_forward_ex(h.ptr, _cabi.ptr(x), B, T, F, steps, _cabi.ptr(out),
                                                                 _cabi.stream_ptr(h.device)))
        return out


class UNet1D(NativeModule):
    """1-D UNet, channels C -> 2C -> 2C(/2) -> 4C -> 4C(/2) -> 8C -> attention -> up x2 with two skip
    concats -> 1x1 conv; [B, C, T] -> [B, output_channels, T] (C = 256 on the native path)."""

    _state_prefix = "unet."

    def __init__(self, input_channels, output_channels, max_depth=5, kernel_size=None, stride=None, p=0, groups=1):
        super().__init__()
        if input_channels != 256 or output_channels != 256 or kernel_size is not None or stride is not None:
            raise NotImplementedError("the native UNet1D implements input_channels = output_channels = 256")
        c = input_channels
        mk = dict(type='1d', leaky=True, p=p, groups=groups)
        self.downsample_layers = nn.ModuleList([
            ConvNormRelu(c, 2 * c, downsample=False, **mk), ConvNormRelu(2 * c, 2 * c, downsample=True, **mk),
            ConvNormRelu(2 * c, 4 * c, downsample=False, **mk), ConvNormRelu(4 * c, 4 * c, downsample=True, **mk)])
        self.upsample_layers = nn.ModuleList([
            ConvTranspose1D(8 * c, 4 * c, stride=2, output_padding=1), ConvNormRelu(8 * c, 4 * c, downsample=False, **mk),
            ConvTranspose1D(4 * c, 2 * c, stride=2, output_padding=1), ConvNormRelu(4 * c, 2 * c, downsample=False, **mk)])
        self.bottleneck = ConvNormRelu(4 * c, 8 * c, downsample=False, **mk)
        self.max_depth = max_depth
        self.final_conv = nn.Conv1d(2 * c, output_channels, kernel_size=1)
        self.bottleneck_attention = SelfAttention(8 * c)
        self.up_attention = SelfAttention(4 * c)

    def forward(self, x):
        self._require_eval()
        h = self.native()
        x = as_input(x, h.device, "UNet1D expects [B, C, T], got %s")
        B, C, T = x.shape
        if C != 256:
            raise ValueError("UNet1D expects 256 input channels, got %d" % C)
        if T % 4 != 0:
            raise ValueError("UNet1D needs T %% 4 == 0 for its skip concats, got T = %d" % T)
        out = torch.empty((B, 256, T), dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _cabi.check(_cabi.lib().a2m_model_unet_forward(h.ptr, _cabi.ptr(x), B, T, _cabi.ptr(out),
                                                           _cabi.stream_ptr(h.device)))
        return out
