"""Model hyper-parameters of the FLUX DiT (black-forest-labs/FLUX.1-dev transformer/config.json)."""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass
class FluxDiTConfig:
    num_layers: int = 19
    num_single_layers: int = 38
    num_attention_heads: int = 24
    attention_head_dim: int = 128
    in_channels: int = 64
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    axes_dims_rope: tuple = (16, 56, 56)

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    def to_dict(self):
        return asdict(self)

    @classmethod
    def from_any(cls, cfg) -> "FluxDiTConfig":
        """Accept another dataclass / dict / diffusers-style config object with the same fields."""
        if isinstance(cfg, cls):
            return cfg
        get = (lambda k, d: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d: getattr(cfg, k, d))
        base = cls()
        return cls(**{k: get(k, getattr(base, k)) for k in asdict(base)})

    def param_count(self) -> int:
        d = self.inner_dim
        dbl = 2 * (6 * d * d + 6 * d) + 2 * 4 * (d * d + d) + 4 * 128 + 2 * (8 * d * d + 5 * d)
        sgl = (3 * d * d + 3 * d) + 3 * (d * d + d) + 2 * 128 + (4 * d * d + 4 * d) + (5 * d * d + d)
        emb = (self.in_channels * d + d) + (self.joint_attention_dim * d + d) \
            + (2 if self.guidance_embeds else 1) * (256 * d + d + d * d + d) \
            + (self.pooled_projection_dim * d + d + d * d + d) + (2 * d * d + 2 * d) \
            + (d * self.in_channels + self.in_channels)
        return self.num_layers * dbl + self.num_single_layers * sgl + emb
