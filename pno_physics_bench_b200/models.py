"""Drop-in mirror of the reference's model API (src/pno_physics_bench/models.py) whose tensor work is done by
libpno_b200.so (hand-written sm_100a CUDA, C ABI in include/pno_b200.h).

Same class names, constructor signatures, attribute / submodule names, `state_dict` keys, shapes and dtypes, the same
train/eval sampling gate and the same exception types as the reference; citations give the reference lines mirrored.
Differences that are deliberate:
  * spectral parameters keep the reference's logical shape [Ci,Co,m1,m2] but are strided views of kernel-native
    [m1,m2,Ci,Co] storage, so a CTA streams one mode's weight matrix with unit stride (checkpoints are unaffected);
  * noise comes from a counter-based Philox stream (seed drawn from torch's global generator once per forward, or fixed
    by `noise_scope`), not from torch.randn_like, so results do not depend on GPU count or batch sharding;
  * additive API: `forward(..., return_kl=)`, `compute_kl_divergence`, `engine=` / `set_engine()`, `PNOBlock`.
There is no CPU path: tensors must live on a B200.
"""
import warnings
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from . import functional as F_

__all__ = ["SpectralConv2d", "SpectralConv2d_Probabilistic", "PNOBlock", "FourierNeuralOperator",
           "ProbabilisticNeuralOperator", "noise_scope"]

noise_scope = F_.noise_scope

_ACT_CODES = {"gelu": _lib.ACT_GELU, "relu": _lib.ACT_RELU}

# Default arithmetic engine of every module: "auto" = the fp32-parity tensor-core engine (error-compensated 3xTF32, rel-L2 <= 1e-5
# against the reference's fp32 path) for every stage whose shape its tiling covers, the fp32 CUDA-core engine for the others (both
# are reference-grade fp32, so the dispatch is invisible in the results; `_lib.stage_counts()` shows which ran).
# Explicit engines are STRICT: "tf32x3" / "tf32" raise when a stage's shape is not tiled by the tensor-core kernels instead of
# silently computing it elsewhere ("tf32x3+fallback" / "tf32+fallback" opt into the fallback).  "tf32" is the reduced-precision
# mode (rel-L2 ~1e-3, the bf16-compute class of BASELINE.json); "fp32" forces CUDA cores.
DEFAULT_ENGINE = "auto"


def _engine_code(engine, allow_fallback=False):
    if isinstance(engine, int):
        return engine | (_lib.ENGINE_ALLOW_FALLBACK if allow_fallback and (engine & 0xFF) != _lib.ENGINE_FP32 else 0)
    try:
        code = _lib.ENGINES[engine]
    except KeyError:
        raise ValueError(f"engine must be one of {sorted(_lib.ENGINES)}, got {engine!r}") from None
    if allow_fallback and code != _lib.ENGINE_FP32:
        code |= _lib.ENGINE_ALLOW_FALLBACK
    return code


def _spectral_param(values):
    """nn.Parameter with the logical shape of `values` on kernel-native [m1,m2,Ci,Co] storage."""
    ci, co, m1, m2 = values.shape
    view = F_.alloc_spectral(ci, co, m1, m2, values.dtype)
    view.copy_(values)
    return nn.Parameter(view)


class SpectralConv2d(nn.Module):
    """Deterministic spectral convolution (reference models.py:18-50)."""

    def __init__(self, in_channels: int, out_channels: int, modes1: int, modes2: int, engine=DEFAULT_ENGINE):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2 = modes1, modes2
        self.scale = 1 / (in_channels * out_channels)
        shape = (in_channels, out_channels, modes1, modes2)
        # same draws, in the same order, as the reference initialiser (models.py:30-31)
        self.weights1 = _spectral_param(self.scale * torch.rand(*shape, dtype=torch.cfloat))
        self.weights2 = _spectral_param(self.scale * torch.rand(*shape, dtype=torch.cfloat))
        self.engine = _engine_code(engine)

    def compl_mul2d(self, input: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        """models.py:33-35. Utility kept for API compatibility; forward() does not use it."""
        return torch.einsum("bixy,ioxy->boxy", input, weights)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F_.spectral_conv(x, self.weights1, self.weights2, engine=self.engine)


class SpectralConv2d_Probabilistic(SpectralConv2d):
    """Spectral convolution with a diagonal Gaussian posterior over its weights (reference models.py:53-104)."""

    def __init__(self, in_channels: int, out_channels: int, modes1: int, modes2: int, engine=DEFAULT_ENGINE):
        super().__init__(in_channels, out_channels, modes1, modes2, engine)
        shape = (in_channels, out_channels, modes1, modes2)
        init = torch.log(torch.full(shape, 0.01))                       # sigma = 0.1 (models.py:59-60)
        self.weights1_log_var = _spectral_param(init)
        self.weights2_log_var = _spectral_param(init.clone())
        self.noise_site = 0

    def sample_weights(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """models.py:63-80. Materialises one draw (API compatibility / inspection); forward() samples in-kernel."""
        if not self.training:
            return self.weights1, self.weights2
        sc = F_.scope_or_fresh()
        re, im = F_.materialize_spectral_noise(sc.seed, sc.sample_idx, self.noise_site, self.in_channels,
                                               self.out_channels, self.modes1, self.modes2, self.weights1.device)
        out = []
        for blk, (w, lv) in enumerate(((self.weights1, self.weights1_log_var), (self.weights2, self.weights2_log_var))):
            std = torch.exp(0.5 * lv)
            out.append(torch.complex(w.real + std * re[blk], w.imag + std * im[blk]))
        return tuple(out)

    def forward(self, x: torch.Tensor, sample: bool = True) -> torch.Tensor:
        if sample and self.training:                                    # the reference's gate (models.py:65, 86)
            sc = F_.scope_or_fresh()
            return F_.spectral_conv(x, self.weights1, self.weights2, self.weights1_log_var, self.weights2_log_var,
                                    seed=sc.seed, sample_idx=sc.sample_idx, site=self.noise_site, engine=self.engine)
        return F_.spectral_conv(x, self.weights1, self.weights2, engine=self.engine)

    def kl_divergence(self) -> torch.Tensor:
        """models.py:100-104."""
        return F_.kl_divergence(self.weights1, self.weights1_log_var, self.weights2, self.weights2_log_var)


class PNOBlock(nn.Module):
    """The PNO block.

    local_noise="weight" reproduces the README listing (reference README.md:153-187; intended test contract
    tests/unit/test_models_comprehensive.py:99-147): spectral(x) + conv1x1 with reparameterised WEIGHTS, optional
    activation.  local_noise="activation" reproduces the block the reference model actually runs
    (models.py:293-306): act(spectral(x) + reparameterize(w_mean(x), w_log_var(x))).
    """

    def __init__(self, in_channels, out_channels, modes1, modes2, activation: Optional[str] = None,
                 local_noise: str = "weight", engine=DEFAULT_ENGINE):
        super().__init__()
        if activation not in (None, "gelu", "relu", "tanh"):
            raise ValueError(f"Unknown activation: {activation}")
        if local_noise not in ("weight", "activation"):
            raise ValueError(f"local_noise must be 'weight' or 'activation', got {local_noise}")
        self.activation_name, self.local_noise = activation, local_noise
        self.engine = _engine_code(engine)
        self.conv = SpectralConv2d_Probabilistic(in_channels, out_channels, modes1, modes2, engine)
        self.w_mean = nn.Conv2d(in_channels, out_channels, 1)
        self.w_log_var = nn.Conv2d(in_channels, out_channels, 1)
        self.site_conv, self.site_local = 1, 3

    def forward(self, x, sample=True):
        stochastic = sample and self.training
        sc = F_.scope_or_fresh() if stochastic else None
        kw = dict(seed=sc.seed, sample_idx=sc.sample_idx) if stochastic else {}
        act = _ACT_CODES.get(self.activation_name, _lib.ACT_NONE)
        if self.local_noise == "activation":
            local = F_.local_branch(x, self.w_mean.weight, self.w_mean.bias,
                                    self.w_log_var.weight if stochastic else None,
                                    self.w_log_var.bias if stochastic else None, site=self.site_local,
                                    batch_offset=sc.batch_offset if stochastic else 0, engine=self.engine, **kw)
        else:
            weight = self.w_mean.weight
            if stochastic:
                eps = F_.materialize_activation_noise(sc.seed, sc.sample_idx, self.site_local, weight.shape,
                                                      device=weight.device)
                weight = weight + torch.exp(0.5 * self.w_log_var.weight) * eps
            local = F_.local_branch(x, weight, self.w_mean.bias, engine=self.engine)
        lv = (self.conv.weights1_log_var, self.conv.weights2_log_var) if stochastic else (None, None)
        out = F_.spectral_conv(x, self.conv.weights1, self.conv.weights2, *lv, addend=local, act=act,
                               site=self.site_conv, engine=self.engine, **kw)
        return torch.tanh(out) if self.activation_name == "tanh" else out

    def kl_divergence(self):
        """README.md:180-187: KL of the 1x1 weights only."""
        lv = self.w_log_var.weight
        return -0.5 * torch.sum(1 + lv - self.w_mean.weight.pow(2) - lv.exp())


class FourierNeuralOperator(nn.Module):
    """Deterministic FNO baseline (reference models.py:121-183)."""

    def __init__(self, input_dim: int = 3, hidden_dim: int = 256, num_layers: int = 4, modes: int = 20,
                 activation: str = "gelu", engine=DEFAULT_ENGINE):
        super().__init__()
        self.input_dim, self.hidden_dim, self.num_layers, self.modes = input_dim, hidden_dim, num_layers, modes
        self.activation_name = "relu" if activation == "relu" else "gelu"      # models.py:140-145
        self.engine = _engine_code(engine)
        self.lift = nn.Conv2d(input_dim, hidden_dim, 1)
        self.fourier_layers = nn.ModuleList(SpectralConv2d(hidden_dim, hidden_dim, modes, modes, engine)
                                            for _ in range(num_layers))
        self.linear_layers = nn.ModuleList(nn.Conv2d(hidden_dim, hidden_dim, 1) for _ in range(num_layers))
        self.project = nn.Conv2d(hidden_dim, 1, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        act = _ACT_CODES[self.activation_name]
        h = F_.local_branch(x, self.lift.weight, self.lift.bias, engine=self.engine)
        for spec, lin in zip(self.fourier_layers, self.linear_layers):
            local = F_.local_branch(h, lin.weight, lin.bias, engine=self.engine)
            h = F_.spectral_conv(h, spec.weights1, spec.weights2, addend=local, act=act, engine=self.engine)
        return F_.local_branch(h, self.project.weight, self.project.bias, engine=self.engine)

    def predict_with_uncertainty(self, x: torch.Tensor, num_samples: int = 100):
        """models.py:178-183: no native uncertainty."""
        with torch.no_grad():
            mean = self.forward(x)
        return mean, torch.zeros_like(mean)


class ProbabilisticNeuralOperator(nn.Module):
    """PNO (reference models.py:186-389)."""

    def __init__(self, input_dim: int = 3, hidden_dim: int = 256, num_layers: int = 4, modes: int = 20,
                 uncertainty_type: str = "full", posterior: str = "variational", activation: str = "gelu",
                 engine=DEFAULT_ENGINE):
        # the reference validates before nn.Module.__init__ (models.py:199-213)
        for name, val in (("input_dim", input_dim), ("hidden_dim", hidden_dim), ("num_layers", num_layers),
                          ("modes", modes)):
            if val <= 0:
                raise ValueError(f"{name} must be positive, got {val}")
        for name, val, allowed in (("uncertainty_type", uncertainty_type, ["full", "diagonal", "scalar"]),
                                   ("posterior", posterior, ["variational", "ensemble", "mc_dropout"]),
                                   ("activation", activation, ["gelu", "relu", "tanh", "swish"])):
            if val not in allowed:
                raise ValueError(f"{name} must be one of {allowed}, got {val}")
        super().__init__()
        self.input_dim, self.hidden_dim, self.num_layers, self.modes = input_dim, hidden_dim, num_layers, modes
        self.uncertainty_type, self.posterior = uncertainty_type, posterior
        self.activation_name = "relu" if activation == "relu" else "gelu"      # tanh / swish fall back to gelu (:224-229)
        self.engine = _engine_code(engine)

        # construction order follows the reference so that a seeded default init draws the same numbers (models.py:231-250)
        self.lift_mean = nn.Conv2d(input_dim, hidden_dim, 1)
        self.lift_log_var = nn.Conv2d(input_dim, hidden_dim, 1)
        self.fourier_layers = nn.ModuleList(SpectralConv2d_Probabilistic(hidden_dim, hidden_dim, modes, modes, engine)
                                            for _ in range(num_layers))
        self.linear_mean = nn.ModuleList(nn.Conv2d(hidden_dim, hidden_dim, 1) for _ in range(num_layers))
        self.linear_log_var = nn.ModuleList(nn.Conv2d(hidden_dim, hidden_dim, 1) for _ in range(num_layers))
        self.project_mean = nn.Conv2d(hidden_dim, 1, 1)
        self.project_log_var = nn.Conv2d(hidden_dim, 1, 1)
        for conv in [*self.linear_log_var, self.lift_log_var, self.project_log_var]:     # models.py:253-262
            nn.init.constant_(conv.weight, -3.0)
            nn.init.constant_(conv.bias, -3.0)
        for l, layer in enumerate(self.fourier_layers):
            layer.noise_site = 1 + 3 * l

    # -- configuration ------------------------------------------------------------------------------------------------
    def set_engine(self, engine, allow_fallback=False):
        """'auto' (default), 'fp32' (CUDA cores, any shape), 'tf32x3' (tcgen05, fp32-grade) or 'tf32' (tcgen05, reduced
        precision).  The explicit tensor-core engines raise for shapes their tilings do not cover unless allow_fallback=True."""
        self.engine = _engine_code(engine, allow_fallback)
        for layer in self.fourier_layers:
            layer.engine = self.engine
        from .graphs import drop_graphs
        drop_graphs(self)
        return self

    # -- pieces -------------------------------------------------------------------------------------------------------
    def reparameterize(self, mean: torch.Tensor, log_var: torch.Tensor) -> torch.Tensor:
        """models.py:264-275 for user code that calls it directly; the fused kernels do this in their epilogue."""
        if not self.training:
            return mean
        sc = F_.scope_or_fresh()
        std = torch.exp(0.5 * torch.clamp(log_var, min=-10.0, max=10.0)).clamp(min=1e-6)
        eps = F_.materialize_activation_noise(sc.seed, sc.sample_idx, 0xFFFF, mean.shape, sc.batch_offset, mean.device)
        return mean + eps * std

    def _validate(self, x):
        if not isinstance(x, torch.Tensor):
            raise TypeError(f"Expected torch.Tensor, got {type(x)}")
        if x.dim() != 4:
            raise ValueError(f"Expected 4D tensor (batch, channels, height, width), got {x.dim()}D tensor")
        if x.size(1) != self.input_dim:
            raise ValueError(f"Expected {self.input_dim} input channels, got {x.size(1)}")

    def _trunk(self, x, stochastic, sample_weights, sc):
        """lift + L blocks (models.py:288-306); returns the hidden features."""
        act = _ACT_CODES[self.activation_name]
        eng = self.engine
        kw = dict(seed=sc.seed, sample_idx=sc.sample_idx) if sc is not None else {}
        boff = sc.batch_offset if sc is not None else 0
        h = F_.local_branch(x, self.lift_mean.weight, self.lift_mean.bias,
                            self.lift_log_var.weight if stochastic else None,
                            self.lift_log_var.bias if stochastic else None, site=0, batch_offset=boff, engine=eng, **kw)
        for l, layer in enumerate(self.fourier_layers):
            local = F_.local_branch(h, self.linear_mean[l].weight, self.linear_mean[l].bias,
                                    self.linear_log_var[l].weight if stochastic else None,
                                    self.linear_log_var[l].bias if stochastic else None, site=2 + 3 * l,
                                    batch_offset=boff, engine=eng, **kw)
            lv = (layer.weights1_log_var, layer.weights2_log_var) if sample_weights else (None, None)
            h = F_.spectral_conv(h, layer.weights1, layer.weights2, *lv, addend=local, act=act, site=1 + 3 * l,
                                 engine=eng, **kw)
        return h

    def forward(self, x: torch.Tensor, sample: bool = True, return_kl: bool = False):
        """models.py:277-316.  return_kl=True (additive; what training/trainer.py:257,279 asks for) returns
        (mean, log_var, kl) with mean = project_mean(h), log_var = clamp(project_log_var(h), -10, 10)."""
        self._validate(x)
        stochastic = sample and self.training
        sc = F_.scope_or_fresh() if stochastic else None
        h = self._trunk(x, stochastic, stochastic, sc)
        if return_kl:
            mean = F_.local_branch(h, self.project_mean.weight, self.project_mean.bias, engine=self.engine)
            log_var = F_.local_branch(h, self.project_log_var.weight, self.project_log_var.bias, engine=self.engine)
            return mean, torch.clamp(log_var, -10.0, 10.0), self.kl_divergence()
        kw = dict(seed=sc.seed, sample_idx=sc.sample_idx, batch_offset=sc.batch_offset) if stochastic else {}
        return F_.local_branch(h, self.project_mean.weight, self.project_mean.bias,
                               self.project_log_var.weight if stochastic else None,
                               self.project_log_var.bias if stochastic else None, site=1 + 3 * self.num_layers,
                               engine=self.engine, **kw)

    def get_final_features(self, x: torch.Tensor) -> torch.Tensor:
        """models.py:331-342."""
        return self._trunk(x, False, False, None)

    def predict_distributional(self, x: torch.Tensor):
        """models.py:318-329 (switches the model to eval() and leaves it there, like the reference)."""
        self.eval()
        with torch.no_grad():
            mean = self.forward(x, sample=False)
            feats = self.get_final_features(x)
            log_var = F_.local_branch(feats, self.project_log_var.weight, self.project_log_var.bias, engine=self.engine)
            std = torch.exp(0.5 * torch.clamp(log_var, min=-10.0, max=10.0)).clamp(min=1e-6)
        return mean, std

    def predict_with_uncertainty(self, x: torch.Tensor, num_samples: int = 100, seed: Optional[int] = None,
                                 group=None, shard_samples: bool = True, use_graph: Optional[bool] = None):
        """models.py:344-379: S sampled forwards under train()+no_grad; mean and unbiased std over the samples.

        The moments are accumulated on device (no [S,B,1,H,W] stack).  The sample loop replays ONE captured CUDA graph of the
        forward (graphs.MCForwardGraph; `use_graph=False` runs the eager loop, None = graph when this rank draws >= 4 samples).
        With torch.distributed initialised (or `group` given) and shard_samples=True the S samples are sharded across ranks by
        GLOBAL sample index and the two moment buffers are all-reduced over NCCL, so every rank returns the same (mean, std)
        for any number of GPUs -- this requires the SAME x on every rank; pass shard_samples=False
        when each rank holds its own batch (data-parallel evaluation)."""
        if not isinstance(x, torch.Tensor):
            raise TypeError(f"Expected torch.Tensor, got {type(x)}")
        if num_samples <= 0:
            raise ValueError(f"num_samples must be positive, got {num_samples}")
        if num_samples > 1000:
            warnings.warn(f"Large num_samples ({num_samples}) may be slow and memory-intensive")
        from .distributed import mc_shard, allreduce_moments, shared_seed
        was_training = self.training
        self.train()
        try:
            with torch.no_grad():
                if shard_samples:
                    seed = shared_seed(seed, x.device, group)
                    first, count = mc_shard(num_samples, group)
                else:
                    seed = F_.fresh_seed() if seed is None else int(seed)
                    first, count = 0, num_samples
                graph = count >= 4 if use_graph is None else bool(use_graph)
                if graph and count > 0:
                    self._validate(x)
                    s1, s2 = self._mc_graph_for(x).run(x, seed, first, count)
                else:
                    s1 = s2 = None
                    for s in range(first, first + count):
                        with F_.noise_scope(seed, s, 0):
                            y = self.forward(x, sample=True)
                        if s1 is None:
                            moments = torch.zeros((2,) + tuple(y.shape), dtype=torch.float64, device=y.device)
                            s1, s2 = moments[0], moments[1]
                        F_.mc_accumulate(y, s1, s2)
                    if s1 is None:      # this rank drew no samples: contribute zeros of the right shape
                        moments = torch.zeros((2, x.shape[0], 1) + tuple(x.shape[2:]), dtype=torch.float64, device=x.device)
                        s1, s2 = moments[0], moments[1]
                if shard_samples:
                    allreduce_moments(s1, s2, group)
                mean, std = F_.mc_finalize(s1, s2, num_samples)
        except torch.cuda.OutOfMemoryError as e:
            raise RuntimeError(f"CUDA out of memory during uncertainty estimation. "
                               f"Try reducing num_samples or batch_size. Original error: {e}")
        except Exception as e:
            raise RuntimeError(f"Error during uncertainty estimation: {e}")
        finally:
            self.train(was_training)
        return mean, std

    def _mc_graph_for(self, x):
        from .graphs import mc_graph_for
        return mc_graph_for(self, x)

    def kl_divergence(self) -> torch.Tensor:
        """models.py:381-389: fourier layers only; one fused reduction over all blocks."""
        params = []
        for layer in self.fourier_layers:
            params += [layer.weights1, layer.weights1_log_var, layer.weights2, layer.weights2_log_var]
        return F_.kl_divergence(*params)

    compute_kl_divergence = kl_divergence          # the name training/losses.py:75 looks for

    def kl_params(self):
        params = []
        for layer in self.fourier_layers:
            params += [layer.weights1, layer.weights1_log_var, layer.weights2, layer.weights2_log_var]
        return params
