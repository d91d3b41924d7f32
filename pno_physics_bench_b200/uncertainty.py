"""The reference's other Monte-Carlo loops on the fused device-side MC driver:

* `UncertaintyDecomposer.decompose`  (reference uncertainty.py:18-94)
* `VariationalEnsemble.predict`      (reference uncertainty/ensemble_methods.py:298-381)

Same names, arguments and returned quantities.  The reference stacks `[S, B, 1, H, W]` sample tensors and reduces them with
`var` / `mean` / `std`; here every pass adds into fp64 moment buffers on the device (`pno_mc_accumulate`,
`pno_mc_accumulate_decompose`), the passes are sharded over the ranks of a process group by GLOBAL sample index like
`predict_with_uncertainty`, and one all-reduce of the moment buffers makes the result identical on every rank.

Reference behaviour kept: both classes put the model in `eval()` and call `model(x, sample=True)`; in eval mode the
reference's sampling gate (models.py:65, 86, 264-267) makes every pass deterministic, so for a PNO the reference returns zero
epistemic uncertainty.  `sample_weights=True` (additive keyword) runs the passes under `train()` + `no_grad()` exactly like
`predict_with_uncertainty` (models.py:352-357) so that the posterior is actually sampled.
"""
from typing import Dict, List, Optional

import torch

from . import _lib
from . import functional as F_
from .distributed import allreduce_moments, mc_shard, shared_seed

_SITE_DECOMPOSE = 0xFFFE        # Philox site of the aleatoric draw of uncertainty.py:70-71


def _accumulate_decompose(mean, log_var, seed, sample_idx, s1, s2, sa, want_pred):
    lib = _lib.load()
    mean = mean.contiguous()
    lv = log_var.contiguous() if log_var is not None else None
    pred = torch.empty_like(mean) if want_pred else None
    _lib.check(lib.pno_mc_accumulate_decompose(_lib.ptr(mean), _lib.ptr(lv) if lv is not None else None, mean.numel(), seed,
                                               sample_idx, _SITE_DECOMPOSE, _lib.ptr(s1), _lib.ptr(s2), _lib.ptr(sa),
                                               _lib.ptr(pred) if want_pred else None, _lib.stream_ptr()),
               "pno_mc_accumulate_decompose")
    return pred


class _Sampling:
    """eval() like the reference, or train() for the duration of the passes when the posterior is to be sampled."""

    def __init__(self, model, sample_weights):
        self.model, self.sample_weights = model, sample_weights

    def __enter__(self):
        self.was_training = self.model.training
        self.model.train(self.sample_weights)
        return self

    def __exit__(self, *exc):
        # the reference leaves the model in eval(); a sampled run restores the previous mode like models.py:377-379
        self.model.train(self.was_training if self.sample_weights else False)


class UncertaintyDecomposer:
    """uncertainty.py:18."""

    def __init__(self, device: Optional[str] = None):
        self.device = device or "cuda"

    def decompose(self, model: torch.nn.Module, input_data: torch.Tensor, num_forward_passes: int = 100,
                  return_samples: bool = False, *, sample_weights: bool = False, seed: Optional[int] = None, group=None):
        """uncertainty.py:29-94: returns (aleatoric, epistemic[, samples]) with
        epistemic = unbiased std over the passes of `prediction` (mean + sqrt(exp(log_var) + 1e-8) eps when the model returns
        (mean, log_var, ...), else the model output) and aleatoric = sqrt(mean over the passes of exp(log_var)) (zeros when
        the model has no variance output).  With `return_samples` the passes of THIS rank are returned stacked."""
        if num_forward_passes <= 0:
            raise ValueError(f"num_forward_passes must be positive, got {num_forward_passes}")
        x = input_data.to(self.device)
        samples = []
        with _Sampling(model, sample_weights), torch.no_grad():
            seed = shared_seed(seed, x.device, group)
            first, count = mc_shard(num_forward_passes, group)
            s1 = s2 = sa = None
            for s in range(first, first + count):
                with F_.noise_scope(seed, s, 0):
                    out = model(x, sample=True) if hasattr(model, "forward") else model(x)
                mean, log_var = (out[0], out[1]) if isinstance(out, tuple) and len(out) >= 2 else (out, None)
                if s1 is None:
                    s1 = torch.zeros(mean.shape, dtype=torch.float64, device=mean.device)
                    s2, sa = torch.zeros_like(s1), torch.zeros_like(s1)
                pred = _accumulate_decompose(mean, log_var, seed, s, s1, s2, sa, return_samples)
                if return_samples:
                    samples.append(pred)
            if s1 is None:            # this rank drew no passes: contribute zeros of the output's shape
                probe = model(x[:1], sample=False) if hasattr(model, "forward") else model(x[:1])
                probe = probe[0] if isinstance(probe, tuple) else probe
                s1 = torch.zeros((x.shape[0],) + tuple(probe.shape[1:]), dtype=torch.float64, device=x.device)
                s2, sa = torch.zeros_like(s1), torch.zeros_like(s1)
            allreduce_moments(s1, s2, group)
            allreduce_moments(sa, torch.zeros_like(sa), group)
            _, epistemic = F_.mc_finalize(s1, s2, num_forward_passes)           # predictions.var(0).sqrt()  (unbiased)
            aleatoric = torch.sqrt(sa / num_forward_passes).float()             # aleatoric_variances.mean(0).sqrt()
        if return_samples:
            return aleatoric, epistemic, torch.stack(samples, dim=0) if samples else torch.empty(0, device=x.device)
        return aleatoric, epistemic


class VariationalEnsemble:
    """uncertainty/ensemble_methods.py:298-381 (`EnsembleBase.__init__` :26-40: models moved to the device, eval mode)."""

    def __init__(self, models: List[torch.nn.Module], num_samples_per_model: int = 10, device: Optional[str] = None):
        self.models = models
        self.device = device or "cuda"
        self.num_models = len(models)
        self.num_samples_per_model = num_samples_per_model
        for model in self.models:
            model.to(self.device)
            model.eval()

    def predict(self, x: torch.Tensor, decompose_uncertainty: bool = True, *, sample_weights: bool = False,
                seed: Optional[int] = None):
        """ensemble_methods.py:326-381: (ensemble mean, total uncertainty, {'aleatoric', 'epistemic', 'total'} or None).
        aleatoric = mean over models of the within-model (unbiased) variance, epistemic = unbiased variance of the model means;
        without decomposition the total is the unbiased std over all M*S samples.  A model that returns a tuple contributes its
        first element (the reference's `torch.stack` cannot stack tuples)."""
        x = x.to(self.device)
        S, Mn = self.num_samples_per_model, self.num_models
        seed = shared_seed(seed, x.device, None)
        all1 = all2 = mm1 = mm2 = within = None
        with torch.no_grad():
            for mi, model in enumerate(self.models):
                with _Sampling(model, sample_weights):
                    s1 = s2 = None
                    for s in range(S):
                        with F_.noise_scope(seed, mi * S + s, 0):
                            y = model(x, sample=True)
                        y = y[0] if isinstance(y, tuple) else y
                        if s1 is None:
                            s1 = torch.zeros(y.shape, dtype=torch.float64, device=y.device)
                            s2 = torch.zeros_like(s1)
                        F_.mc_accumulate(y, s1, s2)
                model_mean, model_std = F_.mc_finalize(s1, s2, S)                    # :344-346 (unbiased std; NaN when S == 1)
                if all1 is None:
                    all1, all2 = torch.zeros_like(s1), torch.zeros_like(s1)
                    mm1, mm2, within = torch.zeros_like(s1), torch.zeros_like(s1), torch.zeros_like(s1)
                all1 += s1
                all2 += s2
                F_.mc_accumulate(model_mean, mm1, mm2)
                within += model_std.double() ** 2
        ensemble_mean, total_all = F_.mc_finalize(all1, all2, S * Mn)                # :357, :379
        if not decompose_uncertainty:
            return ensemble_mean, total_all, None
        aleatoric = (within / Mn).float()                                            # :361
        _, model_mean_std = F_.mc_finalize(mm1, mm2, Mn)                              # :364 torch.var (unbiased)
        epistemic = model_mean_std ** 2
        total = torch.sqrt(aleatoric + epistemic)
        decomposition: Dict[str, torch.Tensor] = {"aleatoric": torch.sqrt(aleatoric), "epistemic": torch.sqrt(epistemic),
                                                  "total": total}
        return ensemble_mean, total, decomposition
