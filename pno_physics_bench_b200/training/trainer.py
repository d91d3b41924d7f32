"""PNOTrainer mirror (reference training/trainer.py:34-575): same constructor signature, `fit` / `evaluate` /
`save_checkpoint` / `load_checkpoint`, same checkpoint dict keys, same step order (zero_grad -> sampled forward ->
loss -> backward -> clip_grad_norm_ -> optimizer.step).  Additions: `train_step` is public, the step's Philox seed is
shared across data-parallel ranks, gradients are averaged with bucketed NCCL all-reduces, and per-step `.item()` host
syncs are deferred to the end of the epoch."""
import contextlib
import logging
import os
import time
from pathlib import Path
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import _lib
from .. import functional as F_
from ..distributed import GradientSynchronizer, shared_seed, world
from .losses import ELBOLoss
from .optim import FusedAdamW

logger = logging.getLogger(__name__)


class PNOTrainer:
    def __init__(self, model: nn.Module, loss_fn: Optional[nn.Module] = None, optimizer=None, scheduler=None,
                 device: Optional[str] = None, gradient_clipping: Optional[float] = 1.0, mixed_precision: bool = False,
                 num_samples: int = 5, kl_weight: float = 1e-4, log_interval: int = 10, use_wandb: bool = False,
                 wandb_project: Optional[str] = None, callbacks: Optional[List] = None, seed: Optional[int] = None,
                 group=None):
        self.device = torch.device(device if device is not None else "cuda")
        self.model = model.to(self.device)
        self.loss_fn = loss_fn if loss_fn is not None else ELBOLoss(kl_weight=kl_weight, num_samples=num_samples)
        # default optimiser: the reference's AdamW(lr 1e-3, weight decay 1e-4) (trainer.py:71-75) as the device-side fused step,
        # with the gradient clipping of trainer.py:283-288 folded in; a user-supplied optimiser takes the reference's path
        self.optimizer = optimizer if optimizer is not None else FusedAdamW(
            self.model.parameters(), lr=1e-3, weight_decay=1e-4, max_grad_norm=gradient_clipping)
        self.scheduler = scheduler
        self.gradient_clipping = gradient_clipping
        self.mixed_precision = mixed_precision      # accepted for signature compatibility; the kernels choose their
        self.num_samples = num_samples              # arithmetic through model.set_engine(), not autocast
        self.log_interval = log_interval
        self.use_wandb = False
        self.callbacks = callbacks or []
        self.group = group
        self.current_epoch = 0
        self.global_step = 0
        self.best_val_loss = float("inf")
        self.training_history = {"train_loss": [], "val_loss": [], "learning_rate": []}
        self.seed = shared_seed(seed, self.device, group)
        _, ws = world(group)
        # data parallel: gradient all-reduces are launched from autograd hooks while backward is still running
        self.grad_sync = GradientSynchronizer(self.model.parameters(), group) if ws > 1 else None
        if isinstance(self.optimizer, FusedAdamW):
            self.optimizer.shard_group = group          # norm of mode-sharded parameters (mode_parallel.py) is summed over the group

    # ------------------------------------------------------------------------------------------------------------------
    def train_step(self, inputs, targets, batch_offset: int = 0) -> Dict[str, torch.Tensor]:
        """One optimisation step (trainer.py:242-295).  Returns the loss dict as device tensors (no host sync)."""
        self.model.train()
        inputs = inputs.to(self.device, non_blocking=True)
        targets = targets.to(self.device, non_blocking=True)
        self.optimizer.zero_grad()
        with F_.noise_scope(self.seed, self.global_step, batch_offset):
            outputs = self.model(inputs, sample=True, return_kl=True)
            kl_leaf = None
            if isinstance(outputs, tuple) and len(outputs) == 3 and hasattr(self.model, "kl_params"):
                # the KL term enters the loss as a leaf: its gradient w.r.t. the weights is analytic and is added straight
                # into the .grad buffers below, scaled by d total / d kl (whatever the loss does with it)
                kl_leaf = outputs[2].detach().requires_grad_(True)
                outputs = (outputs[0], outputs[1], kl_leaf)
            losses = self.loss_fn(outputs, targets, self.model)
            with self._sm_reserve_for_allreduce():
                losses["total"].backward()
            if kl_leaf is not None and kl_leaf.grad is not None:
                if isinstance(self.optimizer, FusedAdamW):      # identical on every rank: added after the gradient all-reduce
                    self.optimizer.set_analytic_kl(self.model.kl_params(), kl_leaf.grad)     # added inside the optimiser kernels
                else:
                    F_.kl_grad_accumulate_(self.model.kl_params(), kl_leaf.grad)
        if self.grad_sync is not None:
            self.grad_sync.finish()
        fused_clip = isinstance(self.optimizer, FusedAdamW) and self.optimizer.max_grad_norm == self.gradient_clipping
        if self.gradient_clipping and not fused_clip:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.gradient_clipping)
        self.optimizer.step()
        self.global_step += 1
        return losses

    @contextlib.contextmanager
    def _sm_reserve_for_allreduce(self):
        """Data parallel: the gradient all-reduces are launched from autograd hooks while backward is still running, but every
        kernel of this library is persistent and owns all 148 SMs, so the NCCL kernels would only run in the gaps.  During backward
        the library leaves `PNO_DP_SM_RESERVE` SMs (default 8) to them (`pno_set_sm_reserve`)."""
        n = int(os.environ.get("PNO_DP_SM_RESERVE", "8")) if (self.grad_sync is not None and self.grad_sync.has_large()) else 0
        if n <= 0:
            yield
            return
        lib = _lib.load()
        _lib.check(lib.pno_set_sm_reserve(n), "pno_set_sm_reserve")
        try:
            yield
        finally:
            _lib.check(lib.pno_set_sm_reserve(0), "pno_set_sm_reserve")

    def _train_epoch(self, loader) -> Dict[str, float]:
        sums, count = {}, 0
        rank, ws = world(self.group)
        for batch_idx, (inputs, targets) in enumerate(loader):
            for cb in self.callbacks:
                cb.on_batch_begin(batch_idx, inputs, targets, self)
            bs = inputs.shape[0]
            losses = self.train_step(inputs, targets, batch_offset=rank * bs)
            for k, v in losses.items():
                sums[k] = sums.get(k, 0.0) + v.detach() * bs
            count += bs
            for cb in self.callbacks:
                cb.on_batch_end(batch_idx, losses, self)
        return {k: float(v) / max(count, 1) for k, v in sums.items()}

    def _allreduce_sums(self, t):
        """Sum an additive statistics vector over the data-parallel group (no-op on one rank)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _validate_epoch(self, loader) -> Dict[str, float]:
        """trainer.py:321-397: predict_with_uncertainty(S=num_samples) -> (mean, 2 log std) -> loss, plus `ece` / `coverage_90`.

        Data-parallel runs shard the validation DATA like the training data, so every rank draws all S samples of its own batches
        (shard_samples=False) and the additive sums (losses x batch, sample count, calibration statistics) are all-reduced once
        at the end of the epoch."""
        from ..metrics import UQAccumulator
        self.model.eval()
        keys, sums, count = None, None, 0
        acc = UQAccumulator(confidence_levels=(0.9,), alphas=(), device=self.device)
        with torch.no_grad():
            for inputs, targets in loader:
                inputs, targets = inputs.to(self.device), targets.to(self.device)
                mean, std = self.model.predict_with_uncertainty(inputs, num_samples=self.num_samples, group=self.group,
                                                                shard_samples=False)
                losses = self.loss_fn((mean, 2 * torch.log(std)), targets, self.model)          # trainer.py:346
                bs = inputs.shape[0]
                if keys is None:
                    keys = sorted(losses)
                    sums = torch.zeros(len(keys) + 1, dtype=torch.float64, device=self.device)
                sums[:-1] += torch.stack([losses[k].detach().double().reshape(()) for k in keys]) * bs
                count += bs
                acc.update(mean, std, targets)
        if keys is None:
            return {}
        sums[-1] = count
        self._allreduce_sums(sums)
        if acc.acc is not None:
            self._allreduce_sums(acc.acc)
        host = sums.cpu().tolist()
        out = {k: host[i] / max(host[-1], 1.0) for i, k in enumerate(keys)}
        try:
            m = acc.compute()
            out["ece"] = m["ece"]                                                            # trainer.py:383-391
            out["coverage_90"] = m["coverage_90"]
        except Exception as e:      # noqa: BLE001  (the reference logs and carries on)
            logger.warning("Failed to compute uncertainty metrics: %s", e)
        return out

    def fit(self, train_loader, val_loader=None, epochs: int = 100, save_dir: Optional[str] = None,
            resume_from: Optional[str] = None) -> Dict[str, List[float]]:
        if resume_from:
            self.load_checkpoint(resume_from)
        for cb in self.callbacks:
            cb.on_train_begin(self)
        t0 = time.time()
        for epoch in range(self.current_epoch, epochs):
            self.current_epoch = epoch
            for cb in self.callbacks:
                cb.on_epoch_begin(epoch, self)
            train = self._train_epoch(train_loader)
            val = self._validate_epoch(val_loader) if val_loader is not None else {}
            if self.scheduler is not None:
                self.scheduler.step()
            self.training_history["train_loss"].append(train.get("total", float("nan")))
            self.training_history["val_loss"].append(val.get("total", float("nan")))
            self.training_history["learning_rate"].append(self.optimizer.param_groups[0]["lr"])
            if val and val["total"] < self.best_val_loss:
                self.best_val_loss = val["total"]
                if save_dir:
                    self.save_checkpoint(Path(save_dir) / "best_model.pt")
            if (epoch + 1) % self.log_interval == 0:
                logger.info("epoch %d train %.4e val %.4e", epoch + 1, train.get("total", float("nan")),
                            val.get("total", float("nan")))
            stop = False
            for cb in self.callbacks:
                if cb.on_epoch_end(epoch, {"train": train, "val": val}, self):
                    stop = True
            if stop:
                break
        self.current_epoch = epochs
        logger.info("training finished in %.1f s", time.time() - t0)
        for cb in self.callbacks:
            cb.on_train_end(self)
        return self.training_history

    def evaluate(self, loader, num_samples: int = 100) -> Dict[str, float]:
        """trainer.py:499-575: MC predictive mean/std with S=100 per batch, the loss on (mean, 2 log(std + 1e-8)) and the
        CalibrationMetrics report.  Predictions stay on the device: every batch adds its sufficient statistics to one fp64
        vector (`metrics.UQAccumulator`), read once at the end -- the reference copies every batch to the host and
        concatenates (trainer.py:548-556).  Returns the reference's keys plus 'mse' / 'nll'."""
        from ..metrics import UQAccumulator
        self.model.eval()
        acc = UQAccumulator(device=self.device)
        total_loss = None
        total_samples = 0
        with torch.no_grad():
            for inputs, targets in loader:
                inputs, targets = inputs.to(self.device), targets.to(self.device)
                mean, std = self.model.predict_with_uncertainty(inputs, num_samples=num_samples, group=self.group,
                                                                shard_samples=False)
                losses = self.loss_fn((mean, 2 * torch.log(std + 1e-8)), targets, self.model)      # trainer.py:533-538
                batch = inputs.shape[0]
                total_loss = losses["total"].detach() * batch if total_loss is None else total_loss + losses["total"].detach() * batch
                total_samples += batch
                acc.update(mean, std, targets)
        if total_samples == 0:
            return {}
        # data-parallel evaluation: each rank saw its own shard of the loader; the statistics are additive
        tot = self._allreduce_sums(torch.stack([total_loss.double().reshape(()),
                                                torch.tensor(float(total_samples), dtype=torch.float64, device=self.device)]))
        self._allreduce_sums(acc.acc)
        total_loss, total_samples = float(tot[0]), float(tot[1])
        m = acc.compute()
        out = {"test_loss": float(total_loss) / total_samples, "rmse": m["rmse"], "mae": m["mae"]}
        out.update(m)
        out["mse"] = m["rmse"] ** 2
        return out

    # ------------------------------------------------------------------------------------------------------------------
    def save_checkpoint(self, path, metadata: Optional[dict] = None):
        """Same dict layout as trainer.py:440-463; tensors are saved in the reference's logical shapes."""
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        ckpt = {
            "epoch": self.current_epoch,
            "model_state_dict": {k: v.contiguous() for k, v in self.model.state_dict().items()},
            "best_val_loss": self.best_val_loss,
            "training_history": self.training_history,
            "optimizer_state_dict": self.optimizer.state_dict(),
            "metadata": metadata or {},
            # additive keys (the reference loader ignores them): the noise of step k is keyed by (seed, k), so a resumed run must
            # continue the step count under the same seed instead of replaying the keys of steps 0..k
            "global_step": self.global_step,
            "noise_seed": self.seed,
        }
        if self.scheduler is not None:
            ckpt["scheduler_state_dict"] = self.scheduler.state_dict()
        torch.save(ckpt, path)

    def load_checkpoint(self, path):
        ckpt = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(ckpt["model_state_dict"])
        if "optimizer_state_dict" in ckpt:
            self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if self.scheduler is not None and "scheduler_state_dict" in ckpt:
            self.scheduler.load_state_dict(ckpt["scheduler_state_dict"])
        self.global_step = ckpt.get("global_step", self.global_step)
        self.seed = ckpt.get("noise_seed", self.seed)
        self.current_epoch = ckpt.get("epoch", 0)
        self.best_val_loss = ckpt.get("best_val_loss", float("inf"))
        self.training_history = ckpt.get("training_history", self.training_history)
        return ckpt
