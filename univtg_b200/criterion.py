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

    def weighted_total(self, loss_dict):
        """`sum(loss_dict[k] * weight_dict[k] for k in loss_dict if k in weight_dict)` - the scalar the reference's training
        loops back-propagate (main/train_mr.py:56-58, main/train_vlp_ddp.py:57-59) - as one dot product on the loss vector
        instead of ~25 single-element kernels (5 selects, 5 multiplies, 4 adds and their backward)."""
        from .losses import LOSS_NAMES

        vec = getattr(loss_dict, "vector", None)
        if vec is None:
            return sum(loss_dict[k] * self.weight_dict[k] for k in loss_dict.keys() if k in self.weight_dict)
        key = (vec.device, tuple(sorted(loss_dict.keys())), tuple(sorted(self.weight_dict.items())))
        cache = self.__dict__.setdefault("_wvec_cache", {})
        w = cache.get(key)
        if w is None:
            w = torch.tensor([float(self.weight_dict[k]) if (k in loss_dict and k in self.weight_dict) else 0.0
                              for k in LOSS_NAMES], dtype=torch.float32, device=vec.device)
            cache.clear()
            cache[key] = w
        return torch.dot(vec, w)
