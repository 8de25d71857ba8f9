"""CPU oracle of the FLUX VAE (diffusers AutoencoderKL with the FLUX.1-dev vae/config.json):
decoder (`generate.py:302-307`: latents / scaling + shift -> vae.decode -> postprocess) and encoder
(`pipeline_tools.py:7-30`).  TEST INFRASTRUCTURE ONLY (see oracle/flux_oracle.py header).

Restated from diffusers.models.autoencoders.{autoencoder_kl,vae} and
diffusers.models.{resnet,attention_processor,upsampling,downsampling} (not vendored in the reference,
not installed here): block_out_channels (128, 256, 512, 512), layers_per_block 2, latent_channels 16,
norm_num_groups 32, act silu, no quant / post-quant conv, mid-block attention with 1 head of dim 512,
scaling_factor 0.3611, shift_factor 0.1159, force_upcast irrelevant in bf16/fp32 here.
Parameter names follow the diffusers state dict (decoder.up_blocks.0.resnets.0.conv1.weight, ...)."""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

SCALING_FACTOR = 0.3611
SHIFT_FACTOR = 0.1159


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6, output_scale_factor=1)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))  # dropout(0) omitted
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h  # output_scale_factor == 1


class VaeAttention(nn.Module):
    """diffusers Attention(512, heads=1, dim_head=512, norm_num_groups=32, residual_connection=True,
    bias=True, _from_deprecated_attn_block=True) with AttnProcessor2_0 (SDPA)."""

    def __init__(self, c: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, c, eps=1e-6)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        res = x
        t = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)  # [b, hw, c]
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o)
        o = o.transpose(1, 2).reshape(b, c, h, w)
        return o + res  # rescale_output_factor == 1


class UNetMidBlock2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.attentions = nn.ModuleList([VaeAttention(c)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Upsample2D(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class Downsample2D(nn.Module):
    """diffusers Downsample2D(padding=0): asymmetric pad (0,1,0,1) then conv3x3 stride 2."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, add_upsample: bool, n: int = 3):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(n)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, add_downsample: bool, n: int = 2):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(n)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class Decoder(nn.Module):
    def __init__(self, latent: int = 16, out_ch: int = 3, chans=(128, 256, 512, 512)):
        super().__init__()
        rev = list(reversed(chans))  # 512, 512, 256, 128
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = UNetMidBlock2D(rev[0])
        ups: List[nn.Module] = []
        prev = rev[0]
        for i, c in enumerate(rev):
            ups.append(UpDecoderBlock2D(prev, c, add_upsample=i != len(rev) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(32, chans[0], eps=1e-6)
        self.conv_out = nn.Conv2d(chans[0], out_ch, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for u in self.up_blocks:
            x = u(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Encoder(nn.Module):
    def __init__(self, in_ch: int = 3, latent: int = 16, chans=(128, 256, 512, 512)):
        super().__init__()
        self.conv_in = nn.Conv2d(in_ch, chans[0], 3, padding=1)
        downs, prev = [], chans[0]
        for i, c in enumerate(chans):
            downs.append(DownEncoderBlock2D(prev, c, add_downsample=i != len(chans) - 1))
            prev = c
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = UNetMidBlock2D(chans[-1])
        self.conv_norm_out = nn.GroupNorm(32, chans[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(chans[-1], 2 * latent, 3, padding=1)  # mean | logvar

    def forward(self, x):
        x = self.conv_in(x)
        for d in self.down_blocks:
            x = d(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    def __init__(self, chans=(128, 256, 512, 512), latent: int = 16):
        super().__init__()
        self.encoder = Encoder(3, latent, chans)
        self.decoder = Decoder(latent, 3, chans)
        self.latent = latent

    @torch.no_grad()
    def decode(self, z):
        return self.decoder(z)

    @torch.no_grad()
    def encode_moments(self, x):
        m = self.encoder(x)
        mean, logvar = m.chunk(2, dim=1)
        return mean, torch.clamp(logvar, -30.0, 20.0)


def _seed(name: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in (name + f"#vae{seed}").encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFF


def init_weights_(model: nn.Module, seed: int = 0, dtype=torch.bfloat16) -> nn.Module:
    """Seeded synthetic weights (pure function of parameter name): convs ~ N(0, 1/fan_in)."""
    with torch.no_grad():
        for name, p in model.named_parameters():
            g = torch.Generator().manual_seed(_seed(name, seed))
            if p.ndim >= 2:
                fan_in = p[0].numel()
                w = torch.randn(p.shape, generator=g) / math.sqrt(fan_in)
            elif name.endswith("weight"):  # GroupNorm gamma
                w = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
            else:
                w = 0.05 * torch.randn(p.shape, generator=g)
            p.data = w.to(dtype)
    return model


def postprocess_uint8(image: torch.Tensor) -> torch.Tensor:
    """diffusers VaeImageProcessor.postprocess(..., 'pil'/'np') up to the uint8 array:
    (x / 2 + 0.5).clamp(0, 1) -> NHWC float32 -> (x * 255).round().astype(uint8)."""
    x = (image / 2 + 0.5).clamp(0, 1)
    x = x.cpu().permute(0, 2, 3, 1).float()
    return (x * 255).round().to(torch.uint8)


def decode_latents(vae: AutoencoderKL, packed_latents, height: int, width: int):
    """generate.py:302-307: unpack -> / scaling + shift -> vae.decode."""
    from .flux_oracle import unpack_latents
    z = unpack_latents(packed_latents, height, width)
    z = (z / SCALING_FACTOR) + SHIFT_FACTOR
    return vae.decode(z)


def encode_images(vae: AutoencoderKL, image_u8_hwc: torch.Tensor, eps: torch.Tensor = None,
                  dtype=torch.bfloat16):
    """train_flux/flux/pipeline_tools.py:7-30 with the posterior noise made explicit:
    VaeImageProcessor.preprocess (uint8 -> [0,1] fp32 -> 2x-1) -> .to(dtype) -> vae.encode ->
    DiagonalGaussianDistribution.sample (mean + std * eps; eps=None -> mode) -> (z - shift) * scale
    -> FluxPipeline._pack_latents.  image_u8_hwc: [H, W, 3] uint8."""
    from .flux_oracle import pack_latents
    x = image_u8_hwc.permute(2, 0, 1)[None].float() / 255.0
    x = (2.0 * x - 1.0).to(dtype)
    mean, logvar = vae.encode_moments(x)
    if eps is None:
        z = mean
    else:
        std = torch.exp(0.5 * logvar)
        z = mean + std * eps.to(dtype)[None]
    z = (z - SHIFT_FACTOR) * SCALING_FACTOR
    b, c, h, w = z.shape
    return pack_latents(z, b, c, h, w)
