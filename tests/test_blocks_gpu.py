"""GPU parity of the stand-alone building blocks (the layer classes of model_layers.py through the drop-in ->
a2m_block_forward) against vectors of the UNMODIFIED reference classes (tests/golden/blocks_reference.npz, made by
oracle/make_golden.py) and against the oracle restatement at the generator's own sizes.

Tolerance: the kernels take bf16 operands with fp32 accumulation, so sum|a - b| / sum|b| <= 4e-3 per block (the whole
generator's bar is 1e-2), and max|a - b| <= 3e-2 * max|b|."""
import os

import numpy as np
import pytest
import torch

from oracle import model_oracle
from oracle.make_golden import BLOCK_CASES, randomize_block

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ml(pkg):
    return pkg.install_dropin()["model_layers"]


@pytest.fixture(scope="module")
def blocks():
    return np.load(os.path.join(ROOT, "tests", "golden", "blocks_reference.npz"))


def close(got, ref, rel=4e-3):
    got, ref = got.detach().cpu().double(), torch.as_tensor(ref).double()
    assert got.shape == ref.shape
    err = (got - ref).abs()
    assert (err.sum() / ref.abs().sum()).item() <= rel, (err.sum() / ref.abs().sum()).item()
    assert err.max().item() <= 3e-2 * ref.abs().max().item(), (err.max().item(), ref.abs().max().item())


@pytest.mark.parametrize("name,cls,args,kwargs,cin,T", BLOCK_CASES)
def test_block_matches_reference_golden(ml, blocks, name, cls, args, kwargs, cin, T):
    mod = getattr(ml, cls)(*args, **kwargs)
    sd = {k[len(name) + 4:]: torch.from_numpy(blocks[k]) for k in blocks.files if k.startswith(name + "/sd/")}
    missing, unexpected = mod.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing), (missing, unexpected)
    mod = mod.cuda().eval()
    y = mod(torch.from_numpy(blocks[name + "/x"]).cuda())
    assert y.is_cuda and y.dtype == torch.float32
    close(y, blocks[name + "/y"])


def _sd(mod):
    return {k: v.detach().cpu() for k, v in mod.state_dict().items()}


def test_blocks_at_generator_sizes_match_oracle(ml):
    """The shapes SelfAttention_G instantiates: 256-channel decoder blocks at T = 64, the 1024- and 2048-channel UNet
    attentions at T = 32 / 16 (tensor-core attention core), and odd lengths (generic kernels)."""
    g = torch.Generator().manual_seed(7)

    def run(mod, x, oracle_fn, seed):
        torch.manual_seed(seed)
        randomize_block(mod, seed)
        mod = mod.cuda().eval()
        sd = {"p." + k: v for k, v in _sd(mod).items()}
        close(mod(x.cuda()), oracle_fn(sd, "p", x))

    x = torch.randn(4, 256, 64, generator=g)
    run(ml.ConvNormRelu(256, 256, type="1d", leaky=True), x, lambda sd, p, v: model_oracle.conv_norm_act(sd, p, v), 1)
    run(ml.ConvNormRelu(256, 512, type="1d", leaky=True, downsample=True), x,
        lambda sd, p, v: model_oracle.conv_norm_act(sd, p, v, stride=2, padding=1), 2)
    run(ml.ResBlock(256, type="1d"), x, model_oracle.res_block, 3)
    run(ml.SelfAttention(256), x, model_oracle.self_attention, 4)                    # fused projection + attention
    run(ml.ChannelAttention(256), x, model_oracle.channel_attention, 5)
    run(ml.SelfAttention(256), torch.randn(3, 256, 20, generator=g), model_oracle.self_attention, 6)   # 20 does not divide 128
    run(ml.SelfAttention(1024), torch.randn(2, 1024, 32, generator=g), model_oracle.self_attention, 7)
    run(ml.SelfAttention(2048), torch.randn(2, 2048, 16, generator=g), model_oracle.self_attention, 8)
    run(ml.SelfAttention(256), torch.randn(2, 256, 200, generator=g), model_oracle.self_attention, 9)  # long sequence
    run(ml.ConvTranspose1D(2048, 1024), torch.randn(2, 2048, 16, generator=g), model_oracle.conv_transpose_block, 10)
    run(ml.ConvNormRelu(512, 512, type="1d", leaky=True), torch.randn(2, 512, 300, generator=g),
        lambda sd, p, v: model_oracle.conv_norm_act(sd, p, v), 11)                    # more than one 128-row tile per clip


def test_block_contracts(ml):
    with pytest.raises(NotImplementedError):
        ml.ConvNormRelu(48, 64, type="1d").cuda().eval()(torch.zeros(1, 48, 8).cuda())           # channels not a multiple of 64
    with pytest.raises(NotImplementedError):
        ml.ConvNormRelu(1, 64, type="2d", downsample=True).cuda().eval()(torch.zeros(1, 1, 8, 8).cuda())
    conv = ml.ConvNormRelu(64, 64, type="1d").cuda()
    with pytest.raises(RuntimeError):
        conv(torch.zeros(1, 64, 8).cuda())                          # training mode: eval semantics only
    conv.eval()
    with pytest.raises(ValueError):
        conv(torch.zeros(1, 128, 8).cuda())
    assert conv(torch.zeros(0, 64, 8).cuda()).shape == (0, 64, 8)
    with pytest.raises(RuntimeError):
        ml.ConvNormRelu(64, 64, type="1d").eval()(torch.zeros(1, 64, 8))                          # parameters on the CPU
