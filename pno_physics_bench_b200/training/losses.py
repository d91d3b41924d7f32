"""Loss mirror of the reference's training/losses.py (PNOLoss :17-149, ELBOLoss :152-183): same constructor arguments,
same returned dict keys and the same numbers, including the reference's aliasing quirk (the in-place `+=` on
`total_loss` makes the reported 'nll' / 'reconstruction' entry equal 'total' on the ELBO path, losses.py:172-177).
Host-side torch ops: the loss epilogue is row N1 of SURVEY.md section 8(f), not part of the spectral hot path."""
import math
from typing import Dict

import torch
import torch.nn as nn


def gaussian_nll(mean, std, target):
    """losses.py:97-117."""
    std = torch.clamp(std, min=1e-6)
    return (0.5 * (math.log(2 * math.pi) + 2 * torch.log(std) + (target - mean) ** 2 / std ** 2)).mean()


def coverage_penalty(mean, std, target, levels=(0.5, 0.9, 0.95)):
    """losses.py:119-149: mean over levels of (empirical central coverage - level)^2 (piecewise constant in the inputs)."""
    resid = (target - mean).abs()
    total = 0.0
    for level in levels:
        z = torch.distributions.Normal(0, 1).icdf(torch.tensor(0.5 + level / 2)).to(std.device)
        total = total + ((resid <= z * std).float().mean() - level) ** 2
    return total / len(levels)


class PNOLoss(nn.Module):
    def __init__(self, reconstruction_weight: float = 1.0, kl_weight: float = 1e-4, calibration_weight: float = 0.1):
        super().__init__()
        self.reconstruction_weight = reconstruction_weight
        self.kl_weight = kl_weight
        self.calibration_weight = calibration_weight

    def _parts(self, predictions, targets, model) -> Dict[str, torch.Tensor]:
        parts = {}
        std = None
        if isinstance(predictions, tuple):
            if len(predictions) == 3:
                mean, log_var, parts["kl"] = predictions
            elif len(predictions) == 2:
                mean, log_var = predictions
            else:
                raise ValueError(f"Invalid prediction format: {len(predictions)} elements")
            std = torch.exp(0.5 * log_var)
        else:
            mean = predictions
        parts["reconstruction"] = torch.mean((mean - targets) ** 2)
        if std is not None:
            parts["nll"] = gaussian_nll(mean, std, targets)
            if self.calibration_weight > 0:
                parts["calibration"] = coverage_penalty(mean, std, targets)
        if "kl" not in parts and model is not None:
            if hasattr(model, "compute_kl_divergence"):
                parts["kl"] = model.compute_kl_divergence()
            else:
                parts["kl"] = torch.tensor(0.0, device=mean.device)
        return parts

    def forward(self, predictions, targets, model=None) -> Dict[str, torch.Tensor]:
        parts = self._parts(predictions, targets, model)
        total = parts["nll"] if "nll" in parts else self.reconstruction_weight * parts["reconstruction"]
        total = total + 0.0                       # new tensor: PNOLoss itself does not alias (losses.py:80)
        if "kl" in parts:
            total = total + self.kl_weight * parts["kl"]
        if "calibration" in parts:
            total = total + self.calibration_weight * parts["calibration"]
        parts["total"] = total
        return parts


class ELBOLoss(PNOLoss):
    def __init__(self, beta: float = 1.0, num_samples: int = 1, **kwargs):
        super().__init__(**kwargs)
        self.beta = beta
        self.num_samples = num_samples            # stored and never read, like the reference (losses.py:163)

    def forward(self, predictions, targets, model=None) -> Dict[str, torch.Tensor]:
        parts = super().forward(predictions, targets, model)
        if "kl" in parts:
            parts["kl"] = self.beta * parts["kl"]
            base_key = "nll" if "nll" in parts else "reconstruction"
            total = parts[base_key] + self.kl_weight * parts["kl"]
            if "calibration" in parts:
                total = total + self.calibration_weight * parts["calibration"]
            parts[base_key] = total               # the reference's in-place update of the aliased tensor
            parts["total"] = total
        return parts
