"""not-gpu: analytic known-answer tests pinning the restated VAE leaves of oracle/vae_oracle.py.

The reference holds NO artefact for this path (no tests, no fixtures; diffusers itself is absent from
the image), so the VAE oracle stays "parity unpinned" against the reference; what can be pinned is
the published structure and arithmetic of diffusers' AutoencoderKL with the FLUX.1-dev vae config:
parameter count and state-dict names, GroupNorm groups/eps, the asymmetric down-sampling pad, nearest
up-sampling, single-head attention scale, posterior / scaling / shift formulas and the uint8
post-process.  `test_against_real_diffusers` cross-checks the whole restatement as soon as a real
diffusers is importable (SURVEY §8c plan item 4); it is skipped in this image."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import flux_oracle as fo
from oracle import vae_oracle as vo


def test_parameter_count_and_state_dict_names():
    m = vo.AutoencoderKL()
    n = sum(p.numel() for p in m.parameters())
    # SD-1.x AutoencoderKL (4 latent channels, with quant convs) has 83 653 863 parameters; the 16-channel
    # FLUX/SD3 variant changes encoder.conv_out (512 -> 32), decoder.conv_in (16 -> 512) and drops the
    # two 1x1 quant convs: + 110 616 + 55 296 - 92
    assert n == 83_653_863 + 110_616 + 55_296 - 92 == 83_819_683
    sd = m.state_dict()
    for k, shape in {"encoder.conv_in.weight": (128, 3, 3, 3),
                     "encoder.down_blocks.1.resnets.0.conv_shortcut.weight": (256, 128, 1, 1),
                     "encoder.down_blocks.2.downsamplers.0.conv.weight": (512, 512, 3, 3),
                     "encoder.mid_block.attentions.0.to_q.weight": (512, 512),
                     "encoder.mid_block.attentions.0.to_out.0.bias": (512,),
                     "encoder.conv_out.weight": (32, 512, 3, 3),
                     "decoder.conv_in.weight": (512, 16, 3, 3),
                     "decoder.mid_block.resnets.1.norm2.weight": (512,),
                     "decoder.up_blocks.2.resnets.0.conv_shortcut.weight": (256, 512, 1, 1),
                     "decoder.up_blocks.3.resnets.2.conv2.weight": (128, 128, 3, 3),
                     "decoder.up_blocks.2.upsamplers.0.conv.bias": (256,),
                     "decoder.conv_norm_out.weight": (128,),
                     "decoder.conv_out.weight": (3, 128, 3, 3)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in sd  # last block does not downsample
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in sd
    assert not any("quant_conv" in k for k in sd)                        # use_quant_conv = False
    assert len(m.decoder.up_blocks[0].resnets) == 3 and len(m.encoder.down_blocks[0].resnets) == 2


def test_spatial_factors():
    m = vo.AutoencoderKL().float()
    with torch.no_grad():
        assert m.decode(torch.zeros(1, 16, 4, 6)).shape == (1, 3, 32, 48)
        mean, logvar = m.encode_moments(torch.zeros(1, 3, 32, 48))
    assert mean.shape == logvar.shape == (1, 16, 4, 6)


def test_groupnorm_is_32_groups_eps_1e6():
    r = vo.ResnetBlock2D(64, 64)
    assert r.norm1.num_groups == 32 and r.norm1.eps == 1e-6 and r.norm1.affine
    # a group = 2 channels here: values {a, b} per pixel-constant channel -> +-1/sqrt(1 + eps/var)
    x = torch.zeros(1, 64, 2, 2)
    x[:, 0::2] = 3.0
    x[:, 1::2] = 1.0
    y = F.group_norm(x, 32, eps=1e-6)
    assert y[0, 0, 0, 0].item() == pytest.approx(1.0 / math.sqrt(1 + 1e-6), rel=1e-6)
    assert y[0, 1, 0, 0].item() == pytest.approx(-1.0 / math.sqrt(1 + 1e-6), rel=1e-6)


def test_downsample_pads_right_and_bottom_only():
    d = vo.Downsample2D(1)
    with torch.no_grad():
        d.conv.bias.zero_()
        d.conv.weight.zero_()
        d.conv.weight[0, 0, 0, 0] = 1.0          # top-left tap: picks x[2i, 2j] -> no padding on the left/top
        x = torch.arange(36.0).view(1, 1, 6, 6)
        assert torch.equal(d(x)[0, 0], x[0, 0, ::2, ::2])
        d.conv.weight.zero_()
        d.conv.weight[0, 0, 2, 2] = 1.0          # bottom-right tap: x[2i+2, 2j+2], zero past the edge
        y = d(x)[0, 0]
    assert y.shape == (3, 3)
    assert torch.equal(y[:2, :2], x[0, 0, 2::2, 2::2]) and torch.all(y[2] == 0) and torch.all(y[:, 2] == 0)


def test_upsample_is_nearest_then_conv():
    u = vo.Upsample2D(1)
    with torch.no_grad():
        u.conv.bias.zero_()
        u.conv.weight.zero_()
        u.conv.weight[0, 0, 1, 1] = 1.0
        x = torch.arange(6.0).view(1, 1, 2, 3)
        y = u(x)[0, 0]
    assert torch.equal(y, x[0, 0].repeat_interleave(2, 0).repeat_interleave(2, 1))


def test_attention_single_head_scale_and_residual():
    a = vo.VaeAttention(64).float()
    with torch.no_grad():
        for lin in (a.to_q, a.to_k, a.to_v, a.to_out[0]):
            lin.bias.zero_()
        a.group_norm.weight.fill_(1.0)
        a.group_norm.bias.zero_()
        eye = torch.eye(64)
        a.to_v.weight.copy_(eye)
        a.to_out[0].weight.copy_(eye)
        a.to_q.weight.zero_()                      # all-zero scores -> uniform softmax -> token mean of v
        a.to_k.weight.zero_()
        x = torch.randn(1, 64, 3, 5)
        y = a(x)
        t = F.group_norm(x.view(1, 64, 15), 32, eps=1e-6)
        want = x + t.mean(dim=2).view(1, 64, 1, 1)
        assert torch.allclose(y, want, atol=1e-5)
        # scale = 1/sqrt(dim_head) with ONE head of the full width
        a.to_q.weight.copy_(eye)
        a.to_k.weight.copy_(eye)
        y2 = a(x)
        tt = t.transpose(1, 2)
        p = torch.softmax(tt @ tt.transpose(1, 2) / math.sqrt(64.0), dim=-1)
        want2 = x + (p @ tt).transpose(1, 2).reshape(1, 64, 3, 5)
        assert torch.allclose(y2, want2, atol=1e-4)


def test_scaling_shift_posterior_and_pack():
    assert vo.SCALING_FACTOR == 0.3611 and vo.SHIFT_FACTOR == 0.1159

    class Ident:
        def decode(self, z):
            return z

        def encode_moments(self, x):
            return torch.full((1, 16, 2, 2), 0.5), torch.full((1, 16, 2, 2), math.log(4.0))
    z = torch.randn(1, 16, 4, 4)
    packed = fo.pack_latents(z, 1, 16, 4, 4)
    out = vo.decode_latents(Ident(), packed, 32, 32)
    assert torch.allclose(out, z / 0.3611 + 0.1159)
    img = torch.zeros(16, 16, 3, dtype=torch.uint8)
    eps = torch.ones(16, 2, 2)
    enc = vo.encode_images(Ident(), img, eps, dtype=torch.float32)   # mean .5, std exp(.5 ln 4) = 2
    assert enc.shape == (1, 1, 64) and torch.allclose(enc, torch.full_like(enc, (0.5 + 2.0 - 0.1159) * 0.3611))
    assert torch.allclose(vo.encode_images(Ident(), img, None, dtype=torch.float32),
                          torch.full((1, 1, 64), (0.5 - 0.1159) * 0.3611))   # eps=None -> mode
    m = vo.AutoencoderKL().float()
    with torch.no_grad():
        m.encoder.conv_out.weight.zero_()
        m.encoder.conv_out.bias[:16] = 0.0
        m.encoder.conv_out.bias[16:] = 50.0
        _, logvar = m.encode_moments(torch.zeros(1, 3, 16, 16))
    assert torch.all(logvar == 20.0)                                  # clamp(-30, 20)


def test_postprocess_rounding():
    x = torch.tensor([-1.0, -0.5, 0.0, 0.5, 1.0, 3.0, -3.0, 1 / 255.0]).view(1, 1, 1, 8).expand(1, 3, 1, 8)
    assert vo.postprocess_uint8(x)[0, 0, :, 0].tolist() == [0, 64, 128, 191, 255, 255, 0, 128]
    # 63.75 -> 64, 127.5 -> 128 (half to even), 191.25 -> 191


def test_against_real_diffusers():
    diffusers = pytest.importorskip("diffusers", reason="diffusers is not installed in this image")
    if "shims" in (getattr(diffusers, "__file__", "") or "") or not hasattr(diffusers, "AutoencoderKL"):
        pytest.skip("only the oracle's import shim named `diffusers` is on the path, not the real package")
    ref = diffusers.AutoencoderKL(in_channels=3, out_channels=3, latent_channels=16,
                                  down_block_types=("DownEncoderBlock2D",) * 4,
                                  up_block_types=("UpDecoderBlock2D",) * 4,
                                  block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                                  use_quant_conv=False, use_post_quant_conv=False, mid_block_add_attention=True)
    ours = vo.AutoencoderKL()
    vo.init_weights_(ours, seed=0, dtype=torch.float32)
    ref.load_state_dict(ours.state_dict(), strict=True)
    z = torch.randn(1, 16, 8, 8)
    with torch.no_grad():
        assert torch.allclose(ref.decode(z).sample, ours.decode(z), atol=1e-4)
        x = torch.randn(1, 3, 64, 64)
        assert torch.allclose(ref.encode(x).latent_dist.mean, ours.encode_moments(x)[0], atol=1e-4)
