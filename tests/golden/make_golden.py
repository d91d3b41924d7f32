#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (`/root/reference/src/pno_physics_bench/models.py`,
`training/losses.py`) in the build container on seeded inputs, with this repo's Philox noise stream injected through
`torch.randn_like` in the reference's own draw order (models.py:67-76, 290, 300-301, 311-312).

The reference has no golden vectors of its own, and /root/reference does not exist on the GPU box, so
these fixtures are what pins the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz   (needs /root/reference)
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

for name in ("matplotlib", "matplotlib.pyplot", "h5py", "seaborn", "wandb"):       # absent optional deps of the package
    sys.modules.setdefault(name, mock.MagicMock())

from oracle import philox  # noqa: E402
from oracle import pno_oracle as orc  # noqa: E402
import pno_physics_bench.models as ref  # noqa: E402
from pno_physics_bench.training import losses as ref_losses  # noqa: E402

SEED = 1234


class EpsQueue:
    """Context manager: torch.randn_like returns the queued tensors, in order, inside the reference."""

    def __init__(self, tensors):
        self.q = list(tensors)
        self.n = 0

    def __enter__(self):
        self._orig = torch.randn_like

        def fake(t, **kw):
            e = self.q[self.n]
            self.n += 1
            assert tuple(e.shape) == tuple(t.shape), (e.shape, t.shape)
            return e.to(t.dtype)

        torch.randn_like = fake
        return self

    def __exit__(self, *a):
        torch.randn_like = self._orig
        assert self.n == len(self.q), f"consumed {self.n} of {len(self.q)} noise tensors"


def layer_queue(noise, site, ci, co, m1, m2):
    re, im = noise.spectral(site, ci, co, m1, m2)
    return [re[0], re[1], im[0], im[1]]                       # models.py:67,68,74,76


def model_queue(noise, model, x):
    b, _, h, w = x.shape
    c, nl, m = model.hidden_dim, model.num_layers, model.modes
    q = [noise.activation(orc.SITE_LIFT, (b, c, h, w))]
    for l in range(nl):
        q += layer_queue(noise, orc.site_spectral(l), c, c, m, m)
        q.append(noise.activation(orc.site_linear(l), (b, c, h, w)))
    q.append(noise.activation(orc.site_project(nl), (b, 1, h, w)))
    return q


def trained_like_(layer, gen, lv_lo=-6.0, lv_hi=-2.0):
    """mu ~ CN(0, 1/Ci), log_var ~ U(lo, hi): makes both the mean path and the noise visible."""
    with torch.no_grad():
        for w, lv in ((layer.weights1, layer.weights1_log_var), (layer.weights2, layer.weights2_log_var)):
            s = (1.0 / w.shape[0]) ** 0.5
            w.copy_(torch.complex(torch.randn(w.shape, generator=gen) * s, torch.randn(w.shape, generator=gen) * s))
            lv.copy_(torch.rand(lv.shape, generator=gen) * (lv_hi - lv_lo) + lv_lo)


def npify(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


LAYER_CASES = {          # name: (B, Ci, Co, H, W, m1, m2, trained_like)
    "even": (2, 3, 4, 16, 12, 4, 5, True),
    "oddw": (2, 3, 4, 16, 13, 4, 7, True),
    "nyquist": (2, 3, 4, 16, 12, 4, 7, True),
    "fullrows": (1, 2, 2, 8, 8, 4, 5, True),
    "overlap": (1, 2, 2, 8, 8, 6, 3, True),
    "default_init": (2, 4, 4, 16, 16, 5, 5, False),
    "oddco": (3, 5, 3, 12, 20, 3, 4, True),
}


def gen_layers():
    out = {}
    for name, (b, ci, co, h, w, m1, m2, tl) in LAYER_CASES.items():
        gen = torch.Generator().manual_seed(hash(name) % 2 ** 31 if False else sum(map(ord, name)))
        torch.manual_seed(sum(map(ord, name)))
        layer = ref.SpectralConv2d_Probabilistic(ci, co, m1, m2)
        if tl:
            trained_like_(layer, gen)
        x = torch.randn(b, ci, h, w, generator=gen, requires_grad=True)
        g = torch.randn(b, co, h, w, generator=gen)
        pre = name + "."
        out[pre + "shape"] = np.array([b, ci, co, h, w, m1, m2])
        for k, v in layer.state_dict().items():
            out[pre + "p." + k] = v.numpy().copy()
        out[pre + "x"], out[pre + "g"] = x.detach().numpy().copy(), g.numpy()
        # deterministic forward (sample=False) + grads
        layer.train()
        y = layer(x, sample=False)
        y.backward(g)
        out[pre + "det.y"] = y.detach().numpy()
        out[pre + "det.dx"] = x.grad.numpy().copy()
        out[pre + "det.dw1"], out[pre + "det.dw2"] = layer.weights1.grad.numpy().copy(), layer.weights2.grad.numpy().copy()
        # eval gate: sample=True in eval mode must equal sample=False
        layer.eval()
        with torch.no_grad():
            out[pre + "evalgate.y"] = layer(x, sample=True).numpy()
        # sampled forward in train mode + grads, site 7, sample_idx 3
        layer.train()
        layer.zero_grad()
        x.grad = None
        noise = orc.PhiloxNoise(SEED, 3)
        with EpsQueue(layer_queue(noise, 7, ci, co, m1, m2)):
            y = layer(x, sample=True)
        y.backward(g)
        out[pre + "smp.y"] = y.detach().numpy()
        out[pre + "smp.dx"] = x.grad.numpy().copy()
        for k in ("weights1", "weights2", "weights1_log_var", "weights2_log_var"):
            out[pre + "smp.d" + k] = getattr(layer, k).grad.numpy().copy()
        # KL + grads
        layer.zero_grad()
        kl = layer.kl_divergence()
        kl.backward()
        out[pre + "kl"] = kl.detach().numpy()
        for k in ("weights1", "weights2", "weights1_log_var", "weights2_log_var"):
            out[pre + "kl.d" + k] = getattr(layer, k).grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "layer_cases.npz"), **out)
    print("layer_cases.npz", len(out), "arrays")


def gen_model():
    out = {}
    torch.manual_seed(7)
    gen = torch.Generator().manual_seed(7)
    # "gelu"/"relu": well-conditioned activation-space log-var (mostly unclamped, some at -10): parity to 1e-5.
    # "stress": wide spread so that BOTH clamp edges of models.py:268 occur (sigma up to e^5); the
    # reference's own fp32 gradients are then ~1e-4 from the fp64 truth, so tests use a looser bound for this case.
    for case, act, wsc, bsc, boff in (("gelu", "gelu", 0.5, 1.0, -3.0), ("relu", "relu", 0.5, 1.0, -3.0),
                                      ("stress", "gelu", 3.0, 4.0, -2.0)):
        model = ref.ProbabilisticNeuralOperator(input_dim=3, hidden_dim=8, num_layers=2, modes=4, activation=act)
        for layer in model.fourier_layers:
            trained_like_(layer, gen)
        with torch.no_grad():
            for conv in [model.lift_log_var, model.project_log_var, *model.linear_log_var]:
                conv.weight.copy_(torch.randn(conv.weight.shape, generator=gen) * wsc)
                conv.bias.copy_(torch.randn(conv.bias.shape, generator=gen) * bsc + boff)
        x = torch.randn(2, 3, 16, 16, generator=gen)
        tgt = torch.randn(2, 1, 16, 16, generator=gen)
        pre = case + "."
        for k, v in model.state_dict().items():
            out[pre + "p." + k] = v.numpy().copy()
        out[pre + "x"], out[pre + "target"] = x.numpy(), tgt.numpy()
        # train-mode sampled forward, sample_idx 5; loss = mse + 1e-4*KL (debug_gradient_nan.py:48-57 style)
        model.train()
        with EpsQueue(model_queue(orc.PhiloxNoise(SEED, 5), model, x)):
            y = model(x, sample=True)
        loss = torch.mean((y - tgt) ** 2) + 1e-4 * model.kl_divergence()
        loss.backward()
        out[pre + "train.y"], out[pre + "train.loss"] = y.detach().numpy(), loss.detach().numpy()
        for k, p_ in model.named_parameters():
            out[pre + "train.grad." + k] = (p_.grad if p_.grad is not None else torch.zeros_like(p_)).numpy().copy()
        out[pre + "kl"] = model.kl_divergence().detach().numpy()
        # gating rows of SURVEY.md A.1
        with torch.no_grad():
            out[pre + "train_nosample.y"] = model(x, sample=False).numpy()
            model.eval()
            out[pre + "eval_sample.y"] = model(x, sample=True).numpy()
            out[pre + "features"] = model.get_final_features(x).numpy()
            mean_d, std_d = model.predict_distributional(x)
            out[pre + "distributional.mean"], out[pre + "distributional.std"] = mean_d.numpy(), std_d.numpy()
        # MC loop: S = 3, sample_idx 0..2
        q = []
        for s in range(3):
            q += model_queue(orc.PhiloxNoise(SEED, s), model, x)
        with EpsQueue(q):
            mean, std = model.predict_with_uncertainty(x, num_samples=3)
        assert not model.training
        out[pre + "mc.mean"], out[pre + "mc.std"] = mean.numpy(), std.numpy()
    # FNO
    torch.manual_seed(11)
    fno = ref.FourierNeuralOperator(input_dim=3, hidden_dim=6, num_layers=2, modes=3)
    with torch.no_grad():
        for layer in fno.fourier_layers:
            for w in (layer.weights1, layer.weights2):
                w.copy_(torch.complex(torch.randn(w.shape), torch.randn(w.shape)) * 0.3)
    x = torch.randn(2, 3, 12, 12)
    for k, v in fno.state_dict().items():
        out["fno.p." + k] = v.numpy().copy()
    out["fno.x"] = x.numpy()
    with torch.no_grad():
        out["fno.y"] = fno(x).numpy()
    np.savez_compressed(os.path.join(HERE, "model_cases.npz"), **out)
    print("model_cases.npz", len(out), "arrays")


def gen_losses():
    out = {}
    gen = torch.Generator().manual_seed(3)
    mean = torch.randn(2, 1, 8, 8, generator=gen)
    lv = torch.randn(2, 1, 8, 8, generator=gen) - 1.0
    tgt = torch.randn(2, 1, 8, 8, generator=gen)
    kl = torch.tensor(123.5)
    loss = ref_losses.ELBOLoss(kl_weight=1e-4, num_samples=5)
    res_t = loss((mean.clone(), lv.clone(), kl.clone()), tgt)
    model_stub = types.SimpleNamespace()              # no compute_kl_divergence -> KL silently 0
    res_m = loss(mean.clone(), tgt, model_stub)
    out["mean"], out["log_var"], out["target"], out["kl_in"] = mean.numpy(), lv.numpy(), tgt.numpy(), kl.numpy()
    for k, v in res_t.items():
        out["tuple." + k] = v.detach().numpy()
    for k, v in res_m.items():
        out["tensor." + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "loss_cases.npz"), **out)
    print("loss_cases.npz", {k: float(v) for k, v in out.items() if np.ndim(v) == 0})


def gen_metrics():
    """CalibrationMetrics of the reference (metrics.py:21-392) on seeded (prediction, std, target) fields."""
    from pno_physics_bench.metrics import CalibrationMetrics
    cm = CalibrationMetrics(device="cpu")
    g = torch.Generator().manual_seed(SEED + 7)
    out = {}
    cases = {"calibrated": (1.0, 1.0), "overconfident": (0.4, 1.0), "underconfident": (2.5, 1.0), "tiny_std": (1e-7, 1e-3)}
    for name, (std_scale, err_scale) in cases.items():
        target = torch.randn(3, 1, 16, 16, generator=g)
        true_std = 0.2 + 0.3 * torch.rand(3, 1, 16, 16, generator=g)
        pred = target + err_scale * true_std * torch.randn(3, 1, 16, 16, generator=g)
        std = std_scale * true_std
        m = cm.compute_all_metrics(pred, std, target)
        out[name + ".pred"], out[name + ".std"], out[name + ".target"] = pred.numpy(), std.numpy(), target.numpy()
        for k, v in m.items():
            out[name + ".m." + k] = np.float64(v)
        out[name + ".m.coverage_at_0.68"] = np.float64(cm.coverage_at_confidence(pred, std, target, 0.68))
        out[name + ".m.ece_10bins"] = np.float64(cm.expected_calibration_error(pred, std, target, num_bins=10))
    np.savez_compressed(os.path.join(HERE, "metrics_cases.npz"), **out)
    print("metrics_cases.npz", {k: float(v) for k, v in out.items() if ".m." in k and k.startswith("calibrated")})


if __name__ == "__main__":
    gen_layers()
    gen_model()
    gen_losses()
    gen_metrics()
