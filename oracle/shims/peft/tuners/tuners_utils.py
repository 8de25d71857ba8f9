"""peft.tuners.tuners_utils.BaseTunerLayer + a LoRA Linear with peft's forward semantics
(lora_controller.py:1)."""
import torch
import torch.nn as nn


class BaseTunerLayer:
    active_adapters = ["default"]

    def scale_layer(self, scale):
        if scale == 1:
            return
        for a in self.active_adapters:
            if a in self.scaling:
                self.scaling[a] *= scale


class LoraLinear(nn.Module, BaseTunerLayer):
    """peft.tuners.lora.Linear: y = base(x) + lora_B(lora_A(dropout(x))) * scaling."""

    def __init__(self, base: nn.Linear, A: torch.Tensor, B: torch.Tensor, alpha: float, adapter="default"):
        super().__init__()
        self.base_layer = base
        r = A.shape[0]
        self.lora_A = nn.ModuleDict({adapter: nn.Linear(A.shape[1], r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter: nn.Linear(r, B.shape[0], bias=False)})
        self.lora_A[adapter].weight.data = A.clone()
        self.lora_B[adapter].weight.data = B.clone()
        self.scaling = {adapter: alpha / r}
        self.active_adapters = [adapter]
        self.in_features, self.out_features = base.in_features, base.out_features

    def forward(self, x):
        result = self.base_layer(x)
        for a in self.active_adapters:
            x_ = x.to(self.lora_A[a].weight.dtype)
            result = result + self.lora_B[a](self.lora_A[a](x_)) * self.scaling[a]
        return result
