"""CalibrationMetrics mirror (reference metrics.py:21-392): same class / method names, arguments and returned numbers.

Every metric is a function of one vector of fp64 sufficient statistics that `pno_uq_stats` (csrc/pno_uq.cu) accumulates in ONE
pass over (prediction, uncertainty, target) on the device.  The statistics are additive, so `UQAccumulator` lets an evaluation
loop keep its predictions on the GPU and read ~100 doubles at the end, where the reference copies every batch to the host and
re-reads the concatenated tensors once per metric (training/trainer.py:540-566, metrics.py:349-381).
"""
import math
from typing import Dict, List, Sequence

import torch

from . import _lib

_STAT_SLOTS = 7


def norm_ppf(p: float) -> float:
    """Standard normal quantile (what scipy.stats.norm.ppf gives in metrics.py:118, 148): Acklam's rational approximation
    plus one Halley step; |error| < 1e-15."""
    a = [-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02,
         -3.066479806614716e+01, 2.506628277459239e+00]
    b = [-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01,
         -1.328068155288572e+01]
    c = [-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00,
         4.374664141464968e+00, 2.938163982698783e+00]
    d = [7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00]
    if not 0.0 < p < 1.0:
        return math.copysign(math.inf, p - 0.5) if p in (0.0, 1.0) else math.nan
    if p < 0.02425:
        q = math.sqrt(-2 * math.log(p))
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / \
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1)
    elif p <= 1 - 0.02425:
        q = p - 0.5
        r = q * q
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / \
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1)
    else:
        q = math.sqrt(-2 * math.log(1 - p))
        x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / \
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1)
    e = 0.5 * math.erfc(-x / math.sqrt(2)) - p
    u = e * math.sqrt(2 * math.pi) * math.exp(x * x / 2)
    return x - u / (1 + x * u / 2)


def uq_stats(mean, spread, target, *, spread_is_log_var=False, z_levels: Sequence[float] = (), alphas: Sequence[float] = (),
             z_alphas: Sequence[float] = (), bin_edges=None, acc=None):
    """Accumulate the sufficient statistics of one batch into `acc` (float64, CUDA; created zeroed when None).
    Layout: include/pno_b200.h (`pno_uq_stats`)."""
    import ctypes
    lib = _lib.load()
    for name, t in (("mean", mean), ("spread", spread), ("target", target)):
        _lib.require_cuda(t, name)
        if t.dtype != torch.float32:
            raise RuntimeError(f"uq_stats: {name} must be float32, got {t.dtype}")
    if mean.shape != spread.shape or mean.shape != target.shape:
        raise RuntimeError(f"uq_stats: shapes differ: {tuple(mean.shape)}, {tuple(spread.shape)}, {tuple(target.shape)}")
    mean, spread, target = mean.contiguous(), spread.contiguous(), target.contiguous()
    n_bins = 0 if bin_edges is None else bin_edges.numel() - 1
    slots = _STAT_SLOTS + len(z_levels) + len(alphas) + 3 * n_bins
    if acc is None:
        acc = torch.zeros(slots, dtype=torch.float64, device=mean.device)
    elif acc.numel() != slots or acc.dtype != torch.float64:
        raise RuntimeError(f"uq_stats: accumulator must hold {slots} float64 values")
    if mean.numel() == 0:
        return acc
    zl = (ctypes.c_float * max(len(z_levels), 1))(*z_levels)
    al = (ctypes.c_float * max(len(alphas), 1))(*alphas)
    za = (ctypes.c_float * max(len(z_alphas), 1))(*z_alphas)
    _lib.check(lib.pno_uq_stats(_lib.ptr(mean), _lib.ptr(spread), _lib.ptr(target), mean.numel(), int(spread_is_log_var),
                                ctypes.cast(zl, ctypes.c_void_p), len(z_levels), ctypes.cast(al, ctypes.c_void_p),
                                ctypes.cast(za, ctypes.c_void_p), len(alphas),
                                _lib.ptr(bin_edges) if n_bins else None, n_bins, _lib.ptr(acc), _lib.stream_ptr()),
               "pno_uq_stats")
    return acc


class UQAccumulator:
    """Streaming form of `CalibrationMetrics.compute_all_metrics`: `update` once per batch (one kernel, no host sync),
    `compute` once at the end (one device->host read of the statistics)."""

    def __init__(self, confidence_levels: Sequence[float] = (0.5, 0.8, 0.9, 0.95, 0.99), alphas: Sequence[float] = (0.1, 0.05, 0.01),
                 num_bins: int = 20, device="cuda"):
        self.levels, self.alphas, self.num_bins = list(confidence_levels), list(alphas), int(num_bins)
        self.z_levels = [norm_ppf(0.5 + lv / 2) for lv in self.levels]                      # metrics.py:118
        self.z_alphas = [norm_ppf(1 - a / 2) for a in self.alphas]                          # metrics.py:148
        self.device = torch.device(device)
        self.edges = torch.linspace(0, 1, self.num_bins + 1).to(self.device) if self.num_bins else None   # metrics.py:66
        self.acc = None

    def update(self, predictions, uncertainties, targets):
        self.acc = uq_stats(predictions, uncertainties, targets, z_levels=self.z_levels, alphas=self.alphas,
                            z_alphas=self.z_alphas, bin_edges=self.edges, acc=self.acc)
        return self

    def stats(self) -> List[float]:
        if self.acc is None:
            raise RuntimeError("UQAccumulator.compute() before any update()")
        return self.acc.cpu().tolist()

    def ece(self, s, norm="l1") -> float:
        n, base = s[0], _STAT_SLOTS + len(self.levels) + len(self.alphas)
        ece = 0.0
        for b in range(self.num_bins):                                                       # metrics.py:71-92
            cnt, sconf, hit = s[base + 3 * b: base + 3 * b + 3]
            if cnt > 0:
                gap = abs(hit / cnt - sconf / cnt)
                ece += (gap * gap if norm == "l2" else gap) * (cnt / n)
        return math.sqrt(ece) if norm == "l2" else ece

    def compute(self) -> Dict[str, float]:
        s = self.stats()
        n = s[0]
        m = {"mae": s[1] / n, "rmse": math.sqrt(s[2] / n)}                                    # metrics.py:347-349
        if self.num_bins:
            m["ece"] = self.ece(s)
        m["nll"] = s[3] / n
        m["sharpness"] = s[4] / n
        for k, lv in enumerate(self.levels):
            m[f"coverage_{int(lv * 100)}"] = s[_STAT_SLOTS + k] / n
        for k, a in enumerate(self.alphas):
            m[f"interval_score_{int((1 - a) * 100)}"] = s[_STAT_SLOTS + len(self.levels) + k] / n
        cov = s[6] / n - (s[4] / n) * (s[1] / n)                                             # metrics.py:370-381 (Pearson r)
        var_u, var_e = s[5] / n - (s[4] / n) ** 2, s[2] / n - (s[1] / n) ** 2
        m["uncertainty_error_correlation"] = cov / math.sqrt(var_u * var_e) if var_u > 0 and var_e > 0 else 0.0
        return m


class CalibrationMetrics:
    """metrics.py:21: every method is one `pno_uq_stats` launch + a read of a few doubles."""

    def __init__(self, device=None):
        self.device = device or "cuda"

    def _one(self, predictions, uncertainties, targets, **kw):
        acc = UQAccumulator(device=predictions.device, **kw)
        return acc, acc.update(predictions, uncertainties, targets).stats()

    def expected_calibration_error(self, predictions, uncertainties, targets, num_bins: int = 20, norm: str = "l1") -> float:
        acc, s = self._one(predictions, uncertainties, targets, confidence_levels=(), alphas=(), num_bins=num_bins)
        return acc.ece(s, norm)

    def coverage_at_confidence(self, predictions, uncertainties, targets, confidence: float = 0.9) -> float:
        _, s = self._one(predictions, uncertainties, targets, confidence_levels=(confidence,), alphas=(), num_bins=0)
        return s[_STAT_SLOTS] / s[0]

    def interval_score(self, predictions, uncertainties, targets, alpha: float = 0.1) -> float:
        _, s = self._one(predictions, uncertainties, targets, confidence_levels=(), alphas=(alpha,), num_bins=0)
        return s[_STAT_SLOTS] / s[0]

    def sharpness(self, uncertainties) -> float:
        _, s = self._one(uncertainties, uncertainties, uncertainties, confidence_levels=(), alphas=(), num_bins=0)
        return s[4] / s[0]

    def negative_log_likelihood(self, predictions, uncertainties, targets) -> float:
        _, s = self._one(predictions, uncertainties, targets, confidence_levels=(), alphas=(), num_bins=0)
        return s[3] / s[0]

    def compute_all_metrics(self, predictions, uncertainties, targets,
                            confidence_levels: Sequence[float] = (0.5, 0.8, 0.9, 0.95, 0.99)) -> Dict[str, float]:
        acc = UQAccumulator(confidence_levels=confidence_levels, device=predictions.device)
        return acc.update(predictions, uncertainties, targets).compute()
