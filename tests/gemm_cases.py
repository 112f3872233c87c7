"""Test helper: build a2m_gemm_desc records for each layer class of the generator and the matching
CPU reference (torch fp32 conv on the same bf16-rounded operands)."""
import ctypes
import importlib

import torch
import torch.nn.functional as F

from conftest import PKG_NAME

ACT = {"none": 0, "leaky": 1, "relu": 2}


def cabi():
    return importlib.import_module(PKG_NAME + "._cabi")


def _pow2_at_least(x):
    p = 1
    while p < x:
        p *= 2
    return p


def rows_box(*extents):
    """Split 128 tile rows over up to 4 dims, innermost first (power-of-two padded)."""
    box, left = [], 128
    for e in extents:
        b = min(left, _pow2_at_least(e))
        box.append(b)
        left //= b
    while len(box) < 4:
        box.append(1)
    box[len(extents) - 1] *= left          # whatever is left goes to the outermost real dim
    return box


def run(desc_kw, w, w_stride_n, w_stride_c, scale, bias, out):
    c = cabi()
    d = c.GemmDesc()
    srcs = desc_kw["sources"]              # list of (tensor, dims, strides)
    d.n_src = len(srcs)
    for s, (t, dims, strides) in enumerate(srcs):
        d.a_ptr[s] = t.data_ptr()
        d.a_rank[s] = len(dims)
        for i, (dm, st) in enumerate(zip(dims, strides)):
            d.a_dims[s][i] = dm
            d.a_strides[s][i] = st
    for i in range(4):
        d.box[i] = desc_kw["box"][i]
        d.m_extent[i] = desc_kw["m_extent"][i]
        d.out_stride[i] = desc_kw["out_stride"][i]
    taps = desc_kw["taps"]                 # list of (src, (o1,o2,o3,o4), channels, w_off)
    d.n_taps = len(taps)
    for t, (src, off, ch, w_off) in enumerate(taps):
        d.tap_src[t] = src
        d.tap_channels[t] = ch
        d.tap_w_off[t] = w_off
        for i in range(4):
            d.tap_off[t][i] = off[i]
    d.N = desc_kw["N"]
    d.act = ACT[desc_kw.get("act", "none")]
    d.out_type = 1 if out.dtype == torch.float32 else 0
    d.out_base = desc_kw.get("out_base", 0)
    d.tile_hint = desc_kw.get("tile_hint", 0)
    c.check(c.lib().a2m_gemm_taps(ctypes.byref(d), c.ptr(w), w_stride_n, w_stride_c, c.ptr(scale), c.ptr(bias),
                                  c.ptr(out), c.stream_ptr()))
    torch.cuda.synchronize()
    return out


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def act_ref(y, act):
    return F.leaky_relu(y, 0.2) if act == "leaky" else F.relu(y) if act == "relu" else y


def gen(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(shape, generator=g)
