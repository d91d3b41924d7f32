"""Loss mirror of the reference's training/losses.py (PNOLoss :17-149, ELBOLoss :152-183): same constructor arguments,
same returned dict keys and the same numbers, including the reference's aliasing quirk (the in-place `+=` on
`total_loss` makes the reported 'nll' / 'reconstruction' entry equal 'total' on the ELBO path, losses.py:172-177).
Host-side torch ops: the loss epilogue is row N1 of SURVEY.md section 8(f), not part of the spectral hot path."""
import math
from typing import Dict

import torch
import torch.nn as nn

from .. import _lib

_LEVELS = (0.5, 0.9, 0.95)
_Z_LEVELS = None


class _LossStats(torch.autograd.Function):
    """Sufficient statistics of the data terms in one pass on the device (csrc/pno_uq.cu): returns the float64 vector
    [n, sum|e|, sum e^2, sum nll, sum std, sum std^2, sum std|e|, count_50, count_90, count_95]; backward is one elementwise
    kernel for d/d mean and d/d log_var of the two differentiable sums (the coverage counts are piecewise constant)."""

    @staticmethod
    def forward(ctx, mean, log_var, target):
        from ..metrics import uq_stats
        global _Z_LEVELS
        if _Z_LEVELS is None:                                                          # losses.py:136 (float32 icdf)
            _Z_LEVELS = [float(torch.distributions.Normal(0, 1).icdf(torch.tensor(0.5 + lv / 2))) for lv in _LEVELS]
        z = _Z_LEVELS
        has_lv = log_var is not None
        stats = uq_stats(mean, log_var if has_lv else torch.zeros_like(mean), target, spread_is_log_var=True,
                         z_levels=z if has_lv else ())
        ctx.save_for_backward(mean, log_var if has_lv else mean, target)
        ctx.has_lv = has_lv
        return stats

    @staticmethod
    def backward(ctx, gstats):
        mean, log_var, target = ctx.saved_tensors
        lib = _lib.load()
        g = gstats[2:4].to(torch.float64).contiguous()
        mean_c, target_c = mean.contiguous(), target.contiguous()
        lv_c = log_var.contiguous() if ctx.has_lv else None
        dmean = torch.empty_like(mean_c)
        dlv = torch.empty_like(mean_c) if ctx.has_lv else None
        _lib.check(lib.pno_loss_data_bwd(_lib.ptr(mean_c), _lib.ptr(lv_c) if ctx.has_lv else None, _lib.ptr(target_c),
                                         mean_c.numel(), _lib.ptr(g), _lib.ptr(dmean), _lib.ptr(dlv) if ctx.has_lv else None,
                                         _lib.stream_ptr()), "pno_loss_data_bwd")
        return dmean, dlv, None


def _on_device(*tensors):
    return all(t.is_cuda and t.dtype == torch.float32 for t in tensors)


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
        log_var = None
        if isinstance(predictions, tuple):
            if len(predictions) == 3:
                mean, log_var, parts["kl"] = predictions
            elif len(predictions) == 2:
                mean, log_var = predictions
            else:
                raise ValueError(f"Invalid prediction format: {len(predictions)} elements")
        else:
            mean = predictions
        if _on_device(mean, targets) and (log_var is None or _on_device(log_var)) and mean.shape == targets.shape:
            # data terms from one pass on the device: same numbers, no intermediate [B,1,H,W] tensors
            st = _LossStats.apply(mean, log_var, targets)
            n = st[0]
            parts["reconstruction"] = (st[2] / n).float()
            if log_var is not None:
                parts["nll"] = (st[3] / n).float()
                if self.calibration_weight > 0:
                    cov = (st[7:10] / n).float()
                    parts["calibration"] = ((cov[0] - _LEVELS[0]) ** 2 + (cov[1] - _LEVELS[1]) ** 2 + (cov[2] - _LEVELS[2]) ** 2) / 3
            return self._with_kl(parts, model, mean)
        std = torch.exp(0.5 * log_var) if log_var is not None else None
        parts["reconstruction"] = torch.mean((mean - targets) ** 2)
        if std is not None:
            parts["nll"] = gaussian_nll(mean, std, targets)
            if self.calibration_weight > 0:
                parts["calibration"] = coverage_penalty(mean, std, targets)
        return self._with_kl(parts, model, mean)

    @staticmethod
    def _with_kl(parts, model, mean):
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
