"""SetCriterion of the reference (model/univtg.py:157-351) behind the same interface; filled in by the loss kernels."""
import torch
from torch import nn


class SetCriterion(nn.Module):
    def __init__(self, weight_dict, eos_coef, losses, temperature, span_loss_type, max_v_l, saliency_margin=1):
        super().__init__()
        self.weight_dict = weight_dict
        self.losses = losses
        self.span_loss_type = span_loss_type
        self.max_v_l = max_v_l
        self.saliency_margin = saliency_margin
        self.temperature = 0.07  # the reference overrides the argument (model/univtg.py:185)
        self.eos_coef = eos_coef
        empty_weight = torch.ones(2)
        empty_weight[-1] = self.eos_coef
        self.register_buffer("empty_weight", empty_weight)

    def forward(self, outputs, targets, hl_only=False):
        from .losses import criterion_forward

        return criterion_forward(self, outputs, targets)
