"""Symbols transformer.py:7-14 imports."""
import logging as _logging
from dataclasses import dataclass

import torch
from oracle.flux_oracle import FluxTransformer2DModel  # noqa: F401

logger = _logging.getLogger("diffusers.shim")
USE_PEFT_BACKEND = True


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor


def scale_lora_layers(model, weight):
    """diffusers.utils.peft_utils.scale_lora_layers: no-op for weight == 1.0."""
    if weight == 1.0:
        return
    for m in model.modules():
        if hasattr(m, "scale_layer") and hasattr(m, "lora_A"):
            m.scale_layer(weight)


def unscale_lora_layers(model, weight=None):
    if weight is None or weight == 1.0:
        return
    for m in model.modules():
        if hasattr(m, "unscale_layer"):
            m.unscale_layer(weight)
