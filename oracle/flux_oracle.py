"""CPU oracle for the FLUX.1-dev DiT hot path of Diffusion-CoT/ReflectionFlow.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (reflectionflow_b200/) imports this file;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and
only as the checker or the CPU comparator.

What it is: a plain-PyTorch restatement of
  * the reference's own orchestration of the path — train_flux/flux/block.py (attn_forward :7-170,
    block_forward :173-272, single_block_forward :275-333), train_flux/flux/transformer.py
    (tranformer_forward :47-252), train_flux/flux/generate.py (denoise loop :193-276),
    train_flux/flux/lora_controller.py (enable_lora :5-42), tts/utils.py (get_noises :131-155,
    prepare_latents_for_flux :71-87) — and
  * the third-party leaf modules that code calls.  Those live in `diffusers` (UNPINNED in the
    reference's requirements.txt:1; contemporaneous release 0.32/0.33), `peft` (LoRA) — neither is
    vendored under the reference tree nor installed here, so their published algorithms are
    restated below (class docstrings name the diffusers module each one follows).

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
("parity unpinned" by the reference itself).  What IS pinned, by oracle/make_golden.py run in the
build container: the restated orchestration here is checked bit-for-bit against the reference's
unmodified block.py / transformer.py / lora_controller.py imported from /root/reference over the
same leaf modules (oracle/shims), and the resulting input/output vectors are committed under
tests/golden/.  The leaf modules themselves can only be pinned by analytic known-answer tests
(tests/test_oracle_kat.py) until a real diffusers install is available.

Rounding model: every op rounds to the tensor dtype exactly where eager PyTorch would (each
Linear output, LayerNorm output, every elementwise op), because that is what the reference
executes.  Run it in float32 for a "true value" comparator or in bfloat16 for the reference's
actual arithmetic.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------
# configuration (black-forest-labs/FLUX.1-dev transformer/config.json values)
# ------------------------------------------------------------------------------------------


@dataclass
class FluxConfig:
    num_layers: int = 19
    num_single_layers: int = 38
    num_attention_heads: int = 24
    attention_head_dim: int = 128
    in_channels: int = 64
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)
    patch_size: int = 1

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @staticmethod
    def dev() -> "FluxConfig":
        return FluxConfig()

    @staticmethod
    def tiny_depth(double: int = 2, single: int = 2) -> "FluxConfig":
        """Full width (every GEMM shape real), few layers."""
        return FluxConfig(num_layers=double, num_single_layers=single)


# ------------------------------------------------------------------------------------------
# leaf modules (restated diffusers)
# ------------------------------------------------------------------------------------------


class RMSNorm(nn.Module):
    """diffusers.models.normalization.RMSNorm(dim, eps, elementwise_affine=True), as used for
    attn.norm_q / norm_k / norm_added_q / norm_added_k (call sites block.py:38-41,60-67,92-95)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        in_dtype = x.dtype
        variance = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(variance + self.eps)  # promotes to fp32
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            x = x.to(self.weight.dtype)
        x = x * self.weight
        return x if x.dtype == in_dtype or self.weight.dtype != torch.float32 else x.to(in_dtype)


class AdaLayerNormZero(nn.Module):
    """diffusers.models.normalization.AdaLayerNormZero (num_embeddings=None variant);
    call sites block.py:186,191,201."""

    def __init__(self, dim: int):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb):
        e = self.linear(self.silu(emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = e.chunk(6, dim=1)
        x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa, shift_mlp, scale_mlp, gate_mlp


class AdaLayerNormZeroSingle(nn.Module):
    """diffusers AdaLayerNormZeroSingle; call sites block.py:295,299."""

    def __init__(self, dim: int):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(dim, 3 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb):
        e = self.linear(self.silu(emb))
        shift, scale, gate = e.chunk(3, dim=1)
        x = self.norm(x) * (1 + scale[:, None]) + shift[:, None]
        return x, gate


class AdaLayerNormContinuous(nn.Module):
    """diffusers AdaLayerNormContinuous(dim, dim, elementwise_affine=False, eps=1e-6);
    call site transformer.py:243.  NOTE the chunk order: scale first, then shift."""

    def __init__(self, dim: int):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(dim, 2 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, conditioning):
        emb = self.linear(self.silu(conditioning).to(x.dtype))
        scale, shift = emb.chunk(2, dim=1)
        return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]


class GELUProj(nn.Module):
    """diffusers.models.activations.GELU(dim_in, dim_out, approximate='tanh'): proj then gelu."""

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    """diffusers.models.attention.FeedForward(dim, dim_out=dim, activation_fn='gelu-approximate'):
    net = [GELU(dim -> 4 dim), Dropout(0), Linear(4 dim -> dim)]; call sites block.py:252-259."""

    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([GELUProj(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class Attention(nn.Module):
    """Parameter container with the attribute names of diffusers Attention that block.py touches
    (:24-67,146-155).  The math itself is in joint_attention() below, as it is in the reference's
    attn_forward (diffusers' own processor is bypassed there too)."""

    def __init__(self, dim: int, heads: int, head_dim: int, added_kv: bool, pre_only: bool):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim)
        self.to_k = nn.Linear(dim, dim)
        self.to_v = nn.Linear(dim, dim)
        self.norm_q = RMSNorm(head_dim)
        self.norm_k = RMSNorm(head_dim)
        if added_kv:
            self.add_q_proj = nn.Linear(dim, dim)
            self.add_k_proj = nn.Linear(dim, dim)
            self.add_v_proj = nn.Linear(dim, dim)
            self.norm_added_q = RMSNorm(head_dim)
            self.norm_added_k = RMSNorm(head_dim)
            self.to_add_out = nn.Linear(dim, dim)
        else:
            self.norm_added_q = None
            self.norm_added_k = None
        if not pre_only:
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])


class FluxTransformerBlock(nn.Module):
    """diffusers FluxTransformerBlock parameter layout (double-stream block)."""

    def __init__(self, dim: int, heads: int, head_dim: int):
        super().__init__()
        self.norm1 = AdaLayerNormZero(dim)
        self.norm1_context = AdaLayerNormZero(dim)
        self.attn = Attention(dim, heads, head_dim, added_kv=True, pre_only=False)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(dim)
        self.norm2_context = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff_context = FeedForward(dim)


class FluxSingleTransformerBlock(nn.Module):
    """diffusers FluxSingleTransformerBlock parameter layout (mlp_ratio 4)."""

    def __init__(self, dim: int, heads: int, head_dim: int):
        super().__init__()
        self.mlp_hidden_dim = 4 * dim
        self.norm = AdaLayerNormZeroSingle(dim)
        self.proj_mlp = nn.Linear(dim, self.mlp_hidden_dim)
        self.act_mlp = nn.GELU(approximate="tanh")
        self.proj_out = nn.Linear(dim + self.mlp_hidden_dim, dim)
        self.attn = Attention(dim, heads, head_dim, added_kv=False, pre_only=True)


def get_timestep_embedding(timesteps, embedding_dim=256, flip_sin_to_cos=True,
                           downscale_freq_shift=0.0, scale=1.0, max_period=10000):
    """diffusers.models.embeddings.get_timestep_embedding."""
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32,
                                                    device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels=256):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, t):
        return get_timestep_embedding(t, self.num_channels)


class TwoLayerSiluMLP(nn.Module):
    """TimestepEmbedding(in,dim) and PixArtAlphaTextProjection(in,dim,act_fn='silu') share this
    shape: linear_1 -> SiLU -> linear_2."""

    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
    """diffusers.models.embeddings.CombinedTimestepGuidanceTextProjEmbeddings; call sites
    transformer.py:102-114."""

    def __init__(self, dim: int, pooled_dim: int):
        super().__init__()
        self.time_proj = Timesteps(256)
        self.timestep_embedder = TwoLayerSiluMLP(256, dim)
        self.guidance_embedder = TwoLayerSiluMLP(256, dim)
        self.text_embedder = TwoLayerSiluMLP(pooled_dim, dim)

    def forward(self, timestep, guidance, pooled):
        t_emb = self.timestep_embedder(self.time_proj(timestep).to(dtype=pooled.dtype))
        g_emb = self.guidance_embedder(self.time_proj(guidance).to(dtype=pooled.dtype))
        tg = t_emb + g_emb
        return tg + self.text_embedder(pooled)


class CombinedTimestepTextProjEmbeddings(nn.Module):
    """guidance_embeds=False variant (FLUX.1-schnell); forward(timestep, pooled)."""

    def __init__(self, dim: int, pooled_dim: int):
        super().__init__()
        self.time_proj = Timesteps(256)
        self.timestep_embedder = TwoLayerSiluMLP(256, dim)
        self.text_embedder = TwoLayerSiluMLP(pooled_dim, dim)

    def forward(self, timestep, pooled):
        t_emb = self.timestep_embedder(self.time_proj(timestep).to(dtype=pooled.dtype))
        return t_emb + self.text_embedder(pooled)


def get_1d_rotary_pos_embed(dim: int, pos: torch.Tensor, theta: float = 10000.0):
    """diffusers get_1d_rotary_pos_embed(use_real=True, repeat_interleave_real=True,
    freqs_dtype=float64)."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64, device=pos.device)[: dim // 2] / dim))
    freqs = torch.outer(pos, freqs)
    cos = freqs.cos().repeat_interleave(2, dim=1).float()
    sin = freqs.sin().repeat_interleave(2, dim=1).float()
    return cos, sin


class FluxPosEmbed(nn.Module):
    """diffusers.models.embeddings.FluxPosEmbed(theta=10000, axes_dim); call sites
    transformer.py:131,134."""

    def __init__(self, theta: int, axes_dim):
        super().__init__()
        self.theta = theta
        self.axes_dim = list(axes_dim)

    def forward(self, ids: torch.Tensor):
        pos = ids.float()
        cos_out, sin_out = [], []
        for i in range(ids.shape[-1]):
            c, s = get_1d_rotary_pos_embed(self.axes_dim[i], pos[:, i], self.theta)
            cos_out.append(c)
            sin_out.append(s)
        return torch.cat(cos_out, dim=-1).to(ids.device), torch.cat(sin_out, dim=-1).to(ids.device)


def apply_rotary_emb(x, freqs_cis):
    """diffusers.models.embeddings.apply_rotary_emb(use_real=True, use_real_unbind_dim=-1):
    interleaved (GPT-J) pairs, fp32 math, one rounding back to x.dtype."""
    cos, sin = freqs_cis
    cos = cos[None, None].to(x.device)
    sin = sin[None, None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class FluxTransformer2DModel(nn.Module):
    """Parameter tree with diffusers' FluxTransformer2DModel state-dict key names."""

    def __init__(self, cfg: FluxConfig):
        super().__init__()
        d = cfg.inner_dim
        self.cfg = cfg
        self.config = _Cfg(in_channels=cfg.in_channels, guidance_embeds=cfg.guidance_embeds,
                           num_layers=cfg.num_layers, num_single_layers=cfg.num_single_layers,
                           num_attention_heads=cfg.num_attention_heads,
                           attention_head_dim=cfg.attention_head_dim,
                           joint_attention_dim=cfg.joint_attention_dim,
                           pooled_projection_dim=cfg.pooled_projection_dim)
        self.pos_embed = FluxPosEmbed(10000, cfg.axes_dims_rope)
        if cfg.guidance_embeds:
            self.time_text_embed = CombinedTimestepGuidanceTextProjEmbeddings(
                d, cfg.pooled_projection_dim)
        else:
            self.time_text_embed = CombinedTimestepTextProjEmbeddings(d, cfg.pooled_projection_dim)
        self.context_embedder = nn.Linear(cfg.joint_attention_dim, d)
        self.x_embedder = nn.Linear(cfg.in_channels, d)
        self.transformer_blocks = nn.ModuleList(
            [FluxTransformerBlock(d, cfg.num_attention_heads, cfg.attention_head_dim)
             for _ in range(cfg.num_layers)])
        self.single_transformer_blocks = nn.ModuleList(
            [FluxSingleTransformerBlock(d, cfg.num_attention_heads, cfg.attention_head_dim)
             for _ in range(cfg.num_single_layers)])
        self.norm_out = AdaLayerNormContinuous(d)
        self.proj_out = nn.Linear(d, cfg.patch_size * cfg.patch_size * cfg.in_channels)
        self.gradient_checkpointing = False


# ------------------------------------------------------------------------------------------
# deterministic synthetic weights (no FLUX checkpoint exists offline)
# ------------------------------------------------------------------------------------------


def _name_seed(name: str, seed: int) -> int:
    h = 1469598103934665603
    for ch in (name + f"#{seed}").encode():
        h = ((h ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h & 0x7FFFFFFF


def init_param(name: str, shape, seed: int = 0, device="cpu") -> torch.Tensor:
    """Seeded fp32 values for one parameter, a pure function of (name, shape, seed): the same
    numbers on every machine.  Scales keep activations O(1) through 57 blocks."""
    g = torch.Generator(device="cpu").manual_seed(_name_seed(name, seed))
    shape = tuple(shape)
    if name.endswith(".weight") and len(shape) == 2:
        fan_in = shape[1]
        std = 1.0 / math.sqrt(fan_in)
        if ".norm1.linear" in name or ".norm1_context.linear" in name or ".norm.linear" in name \
                or name.startswith("norm_out.linear"):
            std *= 0.5  # modulation stays moderate
        w = torch.randn(shape, generator=g) * std
    elif name.endswith(".weight") and len(shape) == 1:  # RMSNorm scale
        w = 1.0 + 0.1 * torch.randn(shape, generator=g)
    else:  # biases
        w = 0.05 * torch.randn(shape, generator=g)
    return w.to(device)


def init_weights_(model: nn.Module, seed: int = 0, dtype=torch.bfloat16) -> nn.Module:
    with torch.no_grad():
        for name, p in model.named_parameters():
            p.data = init_param(name, p.shape, seed).to(dtype)
    return model


LORA_TARGET_SUFFIXES_DOUBLE = ["norm1.linear", "attn.to_q", "attn.to_k", "attn.to_v",
                               "attn.to_out.0", "ff.net.2"]
LORA_TARGET_SUFFIXES_SINGLE = ["norm.linear", "proj_mlp", "proj_out", "attn.to_q", "attn.to_k",
                               "attn.to_v"]


def lora_target_modules(cfg: FluxConfig) -> List[str]:
    """The peft target list of train_flux/config.yaml:53 expanded to module paths."""
    out = ["x_embedder"]
    for i in range(cfg.num_layers):
        out += [f"transformer_blocks.{i}.{s}" for s in LORA_TARGET_SUFFIXES_DOUBLE]
    for i in range(cfg.num_single_layers):
        out += [f"single_transformer_blocks.{i}.{s}" for s in LORA_TARGET_SUFFIXES_SINGLE]
    return out


def make_lora_weights(model: nn.Module, cfg: FluxConfig, rank: int = 32, seed: int = 1,
                      dtype=torch.bfloat16) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """Seeded random LoRA factors {module: (A [r,in], B [out,r])}.  (peft initialises B to zero;
    a trained adapter has both non-zero, which is what matters for parity.)"""
    mods = dict(model.named_modules())
    out = {}
    for name in lora_target_modules(cfg):
        lin = mods[name]
        g = torch.Generator().manual_seed(_name_seed(name + ".lora", seed))
        A = torch.randn(rank, lin.in_features, generator=g) / math.sqrt(lin.in_features)
        B = torch.randn(lin.out_features, rank, generator=g) * (0.5 / math.sqrt(rank))
        out[name] = (A.to(dtype), B.to(dtype))
    return out


# ------------------------------------------------------------------------------------------
# the reference's orchestration, restated
# ------------------------------------------------------------------------------------------


class LoraSet:
    """peft semantics the reference relies on (lora_controller.py:5-42):
    y = W x + b + scaling * B(A(x)); inside `enable_lora(mods, activated=False)` the scaling of
    `mods` is 0.  `latent_lora` == activated."""

    def __init__(self, weights: Optional[Dict[str, Tuple[torch.Tensor, torch.Tensor]]] = None,
                 scaling: float = 1.0):
        self.weights = weights or {}
        self.scaling = scaling

    def linear(self, name: str, lin: nn.Linear, x: torch.Tensor, lora_on: bool) -> torch.Tensor:
        y = lin(x)
        if name in self.weights:
            A, B = self.weights[name]
            s = self.scaling if lora_on else 0.0
            # peft Linear.forward: result = result + lora_B(lora_A(dropout(x))) * scaling
            y = y + F.linear(F.linear(x, A.to(x.dtype)), B.to(x.dtype)) * s
        return y


def _heads(t: torch.Tensor, heads: int) -> torch.Tensor:
    b, n, d = t.shape
    return t.view(b, n, heads, d // heads).transpose(1, 2)


def joint_attention(attn: Attention, prefix: str, lora: LoraSet, latent_lora: bool,
                    img: torch.Tensor, txt: Optional[torch.Tensor], cond: Optional[torch.Tensor],
                    rope, cond_rope, model_config: dict):
    """block.py:7-170 (attn_forward).  Returns per-stream outputs in the reference's order."""
    h = attn.heads
    q = _heads(lora.linear(prefix + "to_q", attn.to_q, img, latent_lora), h)
    k = _heads(lora.linear(prefix + "to_k", attn.to_k, img, latent_lora), h)
    v = _heads(lora.linear(prefix + "to_v", attn.to_v, img, latent_lora), h)
    q = attn.norm_q(q)
    k = attn.norm_k(k)
    if txt is not None:
        tq = attn.norm_added_q(_heads(attn.add_q_proj(txt), h))
        tk = attn.norm_added_k(_heads(attn.add_k_proj(txt), h))
        tv = _heads(attn.add_v_proj(txt), h)
        q = torch.cat([tq, q], dim=2)
        k = torch.cat([tk, k], dim=2)
        v = torch.cat([tv, v], dim=2)
    if rope is not None:
        q = apply_rotary_emb(q, rope)
        k = apply_rotary_emb(k, rope)
    n_cond = 0
    if cond is not None:
        cq = attn.norm_q(_heads(lora.linear(prefix + "to_q", attn.to_q, cond, True), h))
        ck = attn.norm_k(_heads(lora.linear(prefix + "to_k", attn.to_k, cond, True), h))
        cv = _heads(lora.linear(prefix + "to_v", attn.to_v, cond, True), h)
        if cond_rope is not None:
            cq = apply_rotary_emb(cq, cond_rope)
            ck = apply_rotary_emb(ck, cond_rope)
        q = torch.cat([q, cq], dim=2)
        k = torch.cat([k, ck], dim=2)
        v = torch.cat([v, cv], dim=2)
        n_cond = cond.shape[1]
    mask = None
    if not model_config.get("union_cond_attn", True):
        mask = torch.ones(q.shape[2], k.shape[2], dtype=torch.bool, device=q.device)
        mask[-n_cond:, :-n_cond] = False
        mask[:-n_cond, -n_cond:] = False
    c_factor = model_config.get("_c_factor")  # generate(condition_scale != 1), generate.py:86-90
    if c_factor is not None:
        mask = torch.zeros(q.shape[2], k.shape[2], dtype=q.dtype, device=q.device)
        bias = torch.log(torch.ones(1, 1) * c_factor)[0]
        mask[-n_cond:, :-n_cond] = bias
        mask[:-n_cond, -n_cond:] = bias
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False, attn_mask=mask)
    b = o.shape[0]
    o = o.transpose(1, 2).reshape(b, -1, h * o.shape[-1]).to(q.dtype)
    if txt is not None:
        n_txt = txt.shape[1]
        o_txt = o[:, :n_txt]
        o_img = o[:, n_txt: o.shape[1] - n_cond]
        o_cond = o[:, o.shape[1] - n_cond:] if cond is not None else None
        o_img = lora.linear(prefix + "to_out.0", attn.to_out[0], o_img, latent_lora)
        o_txt = attn.to_add_out(o_txt)
        if o_cond is not None:
            o_cond = lora.linear(prefix + "to_out.0", attn.to_out[0], o_cond, True)
        return o_img, o_txt, o_cond
    if cond is not None:
        return o[:, : o.shape[1] - n_cond], None, o[:, o.shape[1] - n_cond:]
    return o, None, None


def double_block(blk: FluxTransformerBlock, prefix: str, lora: LoraSet, img, txt, cond, temb,
                 cond_temb, rope, cond_rope, model_config: dict):
    """block.py:173-272 (block_forward) -> (txt, img, cond)."""
    ll = model_config.get("latent_lora", False)

    def ada(x, emb, lora_on):
        n1 = blk.norm1
        e = lora.linear(prefix + "norm1.linear", n1.linear, n1.silu(emb), lora_on)
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = e.chunk(6, dim=1)
        return n1.norm(x) * (1 + sc_a[:, None]) + sh_a[:, None], g_a, sh_m, sc_m, g_m

    n_img, g_img, shm_img, scm_img, gm_img = ada(img, temb, ll)
    n_txt, g_txt, shm_txt, scm_txt, gm_txt = blk.norm1_context(txt, temb)
    use_cond = cond is not None
    if use_cond:
        n_cond, g_cond, shm_cond, scm_cond, gm_cond = ada(cond, cond_temb, True)
    a_img, a_txt, a_cond = joint_attention(blk.attn, prefix + "attn.", lora, ll, n_img, n_txt,
                                           n_cond if use_cond else None, rope,
                                           cond_rope if use_cond else None, model_config)
    a_img = g_img.unsqueeze(1) * a_img
    img = img + a_img
    a_txt = g_txt.unsqueeze(1) * a_txt
    txt = txt + a_txt
    if use_cond:
        a_cond = g_cond.unsqueeze(1) * a_cond
        cond = cond + a_cond
        if model_config.get("add_cond_attn", False):
            img = img + a_cond

    def mlp_in(x, norm, scale, shift):
        return norm(x) * (1 + scale[:, None]) + shift[:, None]

    def ff_lora(x, lora_on):  # ff.net[0] has no LoRA target; ff.net[2] does (config.yaml:53)
        hdn = blk.ff.net[0](x)
        return lora.linear(prefix + "ff.net.2", blk.ff.net[2], hdn, lora_on)

    f_img = gm_img.unsqueeze(1) * ff_lora(mlp_in(img, blk.norm2, scm_img, shm_img), ll)
    f_txt = gm_txt.unsqueeze(1) * blk.ff_context(mlp_in(txt, blk.norm2_context, scm_txt, shm_txt))
    if use_cond:
        f_cond = gm_cond.unsqueeze(1) * ff_lora(mlp_in(cond, blk.norm2, scm_cond, shm_cond), True)
    img = img + f_img
    txt = txt + f_txt
    if use_cond:
        cond = cond + f_cond
    return txt, img, (cond if use_cond else None)


def single_block(blk: FluxSingleTransformerBlock, prefix: str, lora: LoraSet, x, temb, rope,
                 cond, cond_temb, cond_rope, model_config: dict):
    """block.py:275-333 (single_block_forward)."""
    ll = model_config.get("latent_lora", False)

    def pre(t, emb, lora_on):
        e = lora.linear(prefix + "norm.linear", blk.norm.linear, blk.norm.silu(emb), lora_on)
        shift, scale, gate = e.chunk(3, dim=1)
        n = blk.norm.norm(t) * (1 + scale[:, None]) + shift[:, None]
        m = blk.act_mlp(lora.linear(prefix + "proj_mlp", blk.proj_mlp, n, lora_on))
        return n, gate, m

    n_x, gate, m_x = pre(x, temb, ll)
    use_cond = cond is not None
    if use_cond:
        n_c, gate_c, m_c = pre(cond, cond_temb, True)
    a_x, _, a_c = joint_attention(blk.attn, prefix + "attn.", lora, ll, n_x, None,
                                  n_c if use_cond else None, rope,
                                  cond_rope if use_cond else None, model_config)
    y = gate.unsqueeze(1) * lora.linear(prefix + "proj_out", blk.proj_out,
                                        torch.cat([a_x, m_x], dim=2), ll)
    x_out = x + y
    if not use_cond:
        return x_out, None
    yc = gate_c.unsqueeze(1) * lora.linear(prefix + "proj_out", blk.proj_out,
                                           torch.cat([a_c, m_c], dim=2), True)
    return x_out, cond + yc


@torch.no_grad()
def transformer_forward(model: FluxTransformer2DModel, hidden_states, encoder_hidden_states,
                        pooled_projections, timestep, img_ids, txt_ids, guidance=None,
                        condition_latents=None, condition_ids=None, model_config=None,
                        lora: Optional[LoraSet] = None, c_t: float = 0.0, scalar_dtype=None):
    """transformer.py:47-252 (tranformer_forward); with condition_latents=None it is also the
    stock diffusers FluxTransformer2DModel.forward that FluxPipeline.__call__ uses (entry A)."""
    model_config = model_config or {}
    lora = lora or LoraSet()
    ll = model_config.get("latent_lora", False)
    use_cond = condition_latents is not None
    x = lora.linear("x_embedder", model.x_embedder, hidden_states, ll)
    cond = lora.linear("x_embedder", model.x_embedder, condition_latents, True) if use_cond else None
    # `scalar_dtype`: dtype in which the reference scales timestep / guidance by 1000
    # (transformer.py:95,98 — the model dtype, i.e. bf16: 3.5 -> 3504).  An fp32 "true value" run
    # passes torch.bfloat16 here so that it evaluates the SAME function of the inputs.
    sd = scalar_dtype or x.dtype
    timestep = (timestep.to(sd) * 1000).to(x.dtype)
    if guidance is not None:
        guidance = (guidance.to(sd) * 1000).to(x.dtype)
        temb = model.time_text_embed(timestep, guidance, pooled_projections)
        cond_temb = model.time_text_embed(torch.ones_like(timestep) * c_t * 1000,
                                          torch.ones_like(guidance) * 1000, pooled_projections)
    else:
        temb = model.time_text_embed(timestep, pooled_projections)
        cond_temb = model.time_text_embed(torch.ones_like(timestep) * c_t * 1000,
                                          pooled_projections)
    txt = model.context_embedder(encoder_hidden_states)
    rope = model.pos_embed(torch.cat((txt_ids, img_ids), dim=0))
    cond_rope = model.pos_embed(condition_ids) if use_cond else None
    for i, blk in enumerate(model.transformer_blocks):
        txt, x, cond = double_block(blk, f"transformer_blocks.{i}.", lora, x, txt, cond, temb,
                                    cond_temb if use_cond else None, rope, cond_rope,
                                    model_config)
    n_txt = txt.shape[1]
    x = torch.cat([txt, x], dim=1)
    for i, blk in enumerate(model.single_transformer_blocks):
        x, cond = single_block(blk, f"single_transformer_blocks.{i}.", lora, x, temb, rope, cond,
                               cond_temb if use_cond else None, cond_rope, model_config)
    x = x[:, n_txt:, ...]
    x = model.norm_out(x, temb)
    return model.proj_out(x)


# ------------------------------------------------------------------------------------------
# scheduler / pipeline helpers (restated diffusers; call sites generate.py:193-213,276,302-307)
# ------------------------------------------------------------------------------------------


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5,
                    max_shift=1.15):
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def flow_match_sigmas(num_inference_steps: int, image_seq_len: int):
    """generate.py:193-209 + FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=..., mu=...)
    with the FLUX.1-dev scheduler config (use_dynamic_shifting=True).  Returns
    (timesteps fp32 [N], sigmas fp32 [N+1])."""
    sig = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps).astype(np.float32)
    mu = calculate_shift(image_seq_len)
    sig = math.exp(mu) / (math.exp(mu) + (1 / sig - 1) ** 1.0)
    sigmas = torch.from_numpy(np.asarray(sig, dtype=np.float32)).to(torch.float32)
    timesteps = sigmas * 1000
    sigmas = torch.cat([sigmas, torch.zeros(1, dtype=torch.float32)])
    return timesteps, sigmas


def euler_step(sample, model_output, sigma, sigma_next):
    """FlowMatchEulerDiscreteScheduler.step."""
    prev = sample.to(torch.float32) + (sigma_next - sigma) * model_output
    return prev.to(model_output.dtype)


def pack_latents(latents, batch, ch, h, w):
    latents = latents.view(batch, ch, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5)
    return latents.reshape(batch, (h // 2) * (w // 2), ch * 4)


def unpack_latents(latents, height, width, vae_scale_factor=8):
    b, n, c = latents.shape
    h = 2 * (int(height) // (vae_scale_factor * 2))
    w = 2 * (int(width) // (vae_scale_factor * 2))
    latents = latents.view(b, h // 2, w // 2, c // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
    return latents.reshape(b, c // 4, h, w)


def prepare_latent_image_ids(h2: int, w2: int, dtype=torch.bfloat16):
    ids = torch.zeros(h2, w2, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(h2)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w2)[None, :]
    return ids.reshape(h2 * w2, 3).to(dtype)


def condition_ids(cond_size: int, position_delta=(0, None), dtype=torch.bfloat16):
    """condition.py:96-132 + tts_reflectionflow.py:278: ids of the (cond/16)^2 token grid shifted
    by position_delta = [0, -cond_size // 16]."""
    g = cond_size // 16
    ids = prepare_latent_image_ids(g, g, dtype)
    d0 = position_delta[0]
    d1 = position_delta[1] if position_delta[1] is not None else -g
    ids[:, 1] += d0
    ids[:, 2] += d1
    return ids


def prepare_latents_for_flux(height: int, width: int, seed: int, dtype=torch.bfloat16):
    """tts/utils.py:71-87 driven the way get_noises does (:131-155): the draw happens on the
    GLOBAL CPU generator reseeded with `seed`, directly in `dtype`."""
    h = 2 * (int(height) // 16)
    w = 2 * (int(width) // 16)
    g = torch.manual_seed(int(seed))
    latents = torch.randn((1, 16, h, w), generator=g, dtype=dtype)
    return pack_latents(latents, 1, 16, h, w)


def get_noises(max_seed: int, num_samples: int, height: int, width: int, dtype=torch.bfloat16):
    seeds = torch.randint(0, high=max_seed, size=(num_samples,))
    return {int(s): prepare_latents_for_flux(height, width, int(s), dtype) for s in seeds}


@torch.no_grad()
def denoise(model, latents, prompt_embeds, pooled, num_inference_steps: int, guidance_scale=3.5,
            img_ids=None, txt_ids=None, condition_latents=None, cond_ids=None, model_config=None,
            lora=None, return_trajectory=False, scalar_dtype=None):
    """generate.py:193-276 (== FluxPipeline.__call__ steps 5-6 when condition_latents is None)."""
    dtype = latents.dtype
    timesteps, sigmas = flow_match_sigmas(num_inference_steps, latents.shape[1])
    if txt_ids is None:
        txt_ids = torch.zeros(prompt_embeds.shape[1], 3, dtype=dtype)
    traj = []
    sd = scalar_dtype or dtype
    for i, t in enumerate(timesteps):
        timestep = t.expand(latents.shape[0]).to(sd)  # generate.py:222 rounds t to the model dtype
        guidance = None
        if model.config.guidance_embeds:
            guidance = torch.tensor([guidance_scale]).expand(latents.shape[0])
        v = transformer_forward(model, latents, prompt_embeds, pooled, timestep / 1000, img_ids,
                                txt_ids, guidance, condition_latents, cond_ids, model_config, lora,
                                scalar_dtype=sd)
        latents = euler_step(latents, v, sigmas[i], sigmas[i + 1])
        if return_trajectory:
            traj.append((v, latents))
    return (latents, traj) if return_trajectory else latents
