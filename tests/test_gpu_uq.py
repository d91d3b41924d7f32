"""Loss / calibration epilogue on the device: `pno_uq_stats`, `pno_loss_data_bwd` through the Python mirror
(`pno_physics_bench_b200.metrics`, `.training.losses`, `PNOTrainer.evaluate`) against the oracle and the reference's golden values."""
import math

import numpy as np
import pytest
import torch

from oracle import pno_oracle as orc
import pno_physics_bench_b200 as pno
from pno_physics_bench_b200 import metrics as M
from pno_physics_bench_b200.training import ELBOLoss, PNOLoss, PNOTrainer

pytestmark = pytest.mark.gpu
DEV = "cuda"
CASES = ["calibrated", "overconfident", "underconfident", "tiny_std"]


def close(a, b, rel=2e-6):
    return abs(a - b) <= rel * max(1.0, abs(b))


@pytest.mark.parametrize("name", CASES)
def test_calibration_metrics_match_reference_goldens(golden, name):
    g = golden("metrics_cases.npz")
    pred, std, tgt = (torch.from_numpy(g[f"{name}.{k}"]).to(DEV) for k in ("pred", "std", "target"))
    cm = M.CalibrationMetrics()
    got = cm.compute_all_metrics(pred, std, tgt)
    for k, v in got.items():
        assert close(v, float(g[f"{name}.m.{k}"]), 5e-6), (k, v, float(g[f"{name}.m.{k}"]))
    assert close(cm.coverage_at_confidence(pred, std, tgt, 0.68), float(g[name + ".m.coverage_at_0.68"]))
    assert close(cm.expected_calibration_error(pred, std, tgt, num_bins=10), float(g[name + ".m.ece_10bins"]), 5e-6)
    assert close(cm.negative_log_likelihood(pred, std, tgt), float(g[name + ".m.nll"]), 5e-6)
    assert close(cm.sharpness(std), float(g[name + ".m.sharpness"]))
    assert close(cm.interval_score(pred, std, tgt, 0.05), float(g[name + ".m.interval_score_95"]), 5e-6)


def test_stats_vector_matches_oracle_on_a_large_ragged_input():
    gen = torch.Generator().manual_seed(5)
    n = (1 << 20) + 3
    tgt = torch.randn(n, generator=gen)
    std = 0.05 + torch.rand(n, generator=gen)
    pred = tgt + std * torch.randn(n, generator=gen) * 1.3
    levels, alphas = (0.5, 0.9, 0.99), (0.1, 0.01)
    z = [M.norm_ppf(0.5 + lv / 2) for lv in levels]
    za = [M.norm_ppf(1 - a / 2) for a in alphas]
    edges = torch.linspace(0, 1, 21)
    got = M.uq_stats(pred.to(DEV), std.to(DEV), tgt.to(DEV), z_levels=z, alphas=alphas, z_alphas=za, bin_edges=edges.to(DEV)).cpu().numpy()
    want = orc.uq_stats(pred.numpy(), std.numpy(), tgt.numpy(), z, alphas, 20)
    assert got.shape == want.shape
    assert got[0] == n
    L = len(levels)
    for i, (a, b) in enumerate(zip(got, want)):
        is_count = 7 <= i < 7 + L or (i >= 7 + L + len(alphas) and (i - 7 - L - len(alphas)) % 3 != 1)
        if is_count:            # borderline elements may fall either side (expf / logf differ by an ulp from numpy)
            assert abs(a - b) <= 4, (i, a, b)
        else:
            assert abs(a - b) <= 2e-6 * max(1.0, abs(b)), (i, a, b)
    # log-variance form and batch additivity
    lv = 2 * torch.log(std)
    acc = M.uq_stats(pred[:1000].to(DEV), lv[:1000].to(DEV), tgt[:1000].to(DEV), spread_is_log_var=True, z_levels=z)
    acc = M.uq_stats(pred[1000:].to(DEV), lv[1000:].to(DEV), tgt[1000:].to(DEV), spread_is_log_var=True, z_levels=z, acc=acc)
    one = M.uq_stats(pred.to(DEV), std.to(DEV), tgt.to(DEV), z_levels=z).cpu().numpy()
    two = acc.cpu().numpy()
    assert np.all(np.abs(one[:7] - two[:7]) <= 3e-6 * np.maximum(1.0, np.abs(one[:7])))
    assert np.all(np.abs(one[7:] - two[7:]) <= 4)


def test_empty_and_invalid_inputs():
    e = torch.empty(0, device=DEV)
    acc = M.uq_stats(e, e, e)
    assert float(acc.abs().sum()) == 0.0
    with pytest.raises(RuntimeError):
        M.uq_stats(torch.zeros(4, device=DEV), torch.zeros(5, device=DEV), torch.zeros(4, device=DEV))
    with pytest.raises(RuntimeError):
        M.uq_stats(torch.zeros(4, device=DEV, dtype=torch.float64), torch.zeros(4, device=DEV), torch.zeros(4, device=DEV))
    with pytest.raises(RuntimeError):
        M.uq_stats(torch.zeros(4), torch.zeros(4), torch.zeros(4))


def test_losses_on_device_match_reference_goldens_and_autograd(golden):
    g = golden("loss_cases.npz")
    mean, lv, tgt = (torch.from_numpy(g[k]).to(DEV) for k in ("mean", "log_var", "target"))
    kl = torch.tensor(float(g["kl_in"]), device=DEV)
    res = ELBOLoss()((mean, lv, kl), tgt)
    for k in ("kl", "nll", "calibration", "total", "reconstruction"):
        assert close(float(res[k]), float(g["tuple." + k]), 1e-5), k
    res = ELBOLoss()(mean, tgt, None)
    for k in ("total", "reconstruction"):
        assert close(float(res[k]), float(g["tensor." + k]), 1e-5), k
    # gradients of the fused data terms vs torch autograd on the unfused formula (same device, fp32)
    for with_lv in (True, False):
        m1, l1 = mean.clone().requires_grad_(), lv.clone().requires_grad_()
        m2, l2 = mean.clone().requires_grad_(), lv.clone().requires_grad_()
        if with_lv:
            fused = PNOLoss()((m1, l1), tgt)["total"]
            std = torch.exp(0.5 * l2).clamp(min=1e-6)
            plain = (0.5 * (math.log(2 * math.pi) + 2 * torch.log(std) + (tgt - m2) ** 2 / std ** 2)).mean()
        else:
            fused = PNOLoss(reconstruction_weight=0.7)(m1, tgt)["total"]
            plain = 0.7 * torch.mean((m2 - tgt) ** 2)
        fused.backward()
        plain.backward()
        assert torch.allclose(m1.grad, m2.grad, rtol=2e-5, atol=1e-9)
        if with_lv:
            assert torch.allclose(l1.grad, l2.grad, rtol=2e-5, atol=1e-9)
    # a clamped std (log_var below 2 log 1e-6) has zero gradient w.r.t. log_var, like torch.clamp
    lv_small = torch.full_like(lv, -40.0).requires_grad_()
    PNOLoss(calibration_weight=0.0)((mean, lv_small), tgt)["total"].backward()
    assert float(lv_small.grad.abs().max()) == 0.0


def test_trainer_evaluate_reports_the_reference_metrics_without_host_copies():
    torch.manual_seed(0)
    model = pno.ProbabilisticNeuralOperator(3, 32, 2, 8).to(DEV)
    trainer = PNOTrainer(model, ELBOLoss(), device=DEV)
    data = [(torch.randn(4, 3, 32, 32), torch.randn(4, 1, 32, 32)) for _ in range(3)]
    out = trainer.evaluate(data, num_samples=6)
    for k in ("test_loss", "rmse", "mae", "ece", "nll", "sharpness", "coverage_50", "coverage_99", "interval_score_90",
              "uncertainty_error_correlation", "mse"):
        assert k in out and math.isfinite(out[k]), k
    assert 0.0 <= out["coverage_50"] <= out["coverage_99"] <= 1.0
    assert abs(out["mse"] - out["rmse"] ** 2) < 1e-9


# ---- the other MC-loop callers on the device-side MC driver ------------------------------------------
from pno_physics_bench_b200 import functional as F_  # noqa: E402
from pno_physics_bench_b200.uncertainty import UncertaintyDecomposer, VariationalEnsemble  # noqa: E402


class _Heteroscedastic(torch.nn.Module):
    """Returns (mean, log_var) like the models uncertainty.py:63-73 is written for."""

    def forward(self, x, sample=True):
        return x.mean(1, keepdim=True) * 0.5, (x[:, :1] * 0.3 - 1.0)


def test_decompose_matches_the_reference_formulas_with_injected_noise():
    torch.manual_seed(1)
    x = torch.randn(3, 2, 8, 12, device=DEV)
    S, seed = 7, 4242
    model = _Heteroscedastic()
    alea, epi, samples = UncertaintyDecomposer().decompose(model, x, S, return_samples=True, seed=seed)
    mean, lv = model(x)
    preds = torch.stack([mean + torch.sqrt(torch.exp(lv) + 1e-8) * F_.materialize_activation_noise(seed, s, 0xFFFE, mean.shape)
                         for s in range(S)])                                   # uncertainty.py:66-73
    assert torch.allclose(samples, preds, rtol=1e-6, atol=1e-6)
    assert torch.allclose(epi, preds.var(dim=0).sqrt(), rtol=2e-5, atol=1e-7)   # :84-85
    assert torch.allclose(alea, torch.exp(lv).sqrt(), rtol=1e-6, atol=1e-7)     # :88-89
    a2, e2 = UncertaintyDecomposer().decompose(model, x, S, seed=seed)
    assert torch.equal(a2, alea) and torch.equal(e2, epi)


def test_decompose_on_a_pno_keeps_the_reference_gate_and_can_sample_the_posterior():
    torch.manual_seed(2)
    model = pno.ProbabilisticNeuralOperator(3, 32, 2, 8).to(DEV)
    x = torch.randn(2, 3, 32, 32, device=DEV)
    alea, epi = UncertaintyDecomposer().decompose(model, x, 5, seed=9)
    assert not model.training                      # the reference leaves the model in eval() (uncertainty.py:49)
    assert float(epi.abs().max()) == 0.0 and float(alea.abs().max()) == 0.0      # eval-mode passes are deterministic
    model.train()
    alea, epi = UncertaintyDecomposer().decompose(model, x, 6, seed=9, sample_weights=True)
    assert model.training                          # a sampled run restores the previous mode
    _, std = model.predict_with_uncertainty(x, num_samples=6, seed=9)
    assert torch.allclose(epi, std, rtol=1e-6, atol=1e-7) and float(epi.mean()) > 0
    assert float(alea.abs().max()) == 0.0          # the PNO forward returns a tensor: no variance output (uncertainty.py:75-77)


def test_variational_ensemble_matches_the_reference_formulas():
    torch.manual_seed(3)
    models = [pno.ProbabilisticNeuralOperator(3, 32, 1, 8) for _ in range(3)]
    ens = VariationalEnsemble(models, num_samples_per_model=4)
    x = torch.randn(2, 3, 32, 32, device=DEV)
    mean, total, dec = ens.predict(x, sample_weights=True, seed=77)
    # the same passes, stacked and reduced like ensemble_methods.py:340-375
    per_model = []
    for mi, m in enumerate(models):
        m.train()
        with torch.no_grad():
            ys = []
            for s in range(4):
                with pno.noise_scope(77, mi * 4 + s, 0):
                    ys.append(m(x, sample=True))
        m.eval()
        per_model.append(torch.stack(ys))
    allp = torch.cat(per_model)
    means = torch.stack([p.mean(0) for p in per_model])
    stds = torch.stack([p.std(0) for p in per_model])
    alea, epis = (stds ** 2).mean(0), means.var(0)
    assert torch.allclose(mean, allp.mean(0), rtol=1e-5, atol=1e-6)
    assert torch.allclose(dec["aleatoric"], alea.sqrt(), rtol=2e-4, atol=1e-6)
    assert torch.allclose(dec["epistemic"], epis.sqrt(), rtol=2e-4, atol=1e-6)
    assert torch.allclose(total, (alea + epis).sqrt(), rtol=2e-4, atol=1e-6)
    mean2, total2, none = ens.predict(x, decompose_uncertainty=False, sample_weights=True, seed=77)
    assert none is None and torch.allclose(total2, allp.std(0), rtol=2e-4, atol=1e-6)
    # reference behaviour (eval mode): deterministic passes, so no within-model spread; the models still differ
    with torch.no_grad():
        det = torch.stack([m(x, sample=False) for m in models])
    _, t0, d0 = ens.predict(x, seed=77)
    assert float(d0["aleatoric"].abs().max()) == 0.0
    assert torch.allclose(t0, det.var(0).sqrt(), rtol=2e-4, atol=1e-6)


# ---- optimiser step on device (trainer.py:283-291) ----------------------------------------------------------------------
from pno_physics_bench_b200.training import FusedAdamW  # noqa: E402


@pytest.mark.parametrize("max_norm", [0.05, 1e9, None])
def test_fused_adamw_matches_torch_adamw_with_clipping(max_norm):
    torch.manual_seed(4)

    def make():
        torch.manual_seed(11)
        native = torch.randn(4, 6, 5, 7, dtype=torch.cfloat, device=DEV)              # kernel-native storage [m1,m2,Ci,Co]
        return [torch.nn.Parameter(native.permute(2, 3, 0, 1)),                        # logical [Ci,Co,m1,m2] view, like the model
                torch.nn.Parameter(torch.randn(13, 9, device=DEV)),                    # ragged (117 elements)
                torch.nn.Parameter(torch.randn(1027, device=DEV))]

    ours, ref = make(), make()
    opt_o = FusedAdamW(ours, lr=3e-3, weight_decay=1e-2, max_grad_norm=max_norm)
    opt_r = torch.optim.AdamW(ref, lr=3e-3, weight_decay=1e-2)
    for step in range(4):
        torch.manual_seed(100 + step)
        grads = [torch.randn_like(p) * (10.0 if step == 1 else 0.1) for p in ours]
        for p, q, g in zip(ours, ref, grads):
            p.grad = g.clone().contiguous() if step % 2 else g.clone()                 # also a gradient with other strides
            q.grad = g.clone()
        if max_norm is not None:
            want_norm = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        opt_o.step()
        opt_r.step()
        if max_norm is not None:
            assert torch.allclose(opt_o.grad_norm(), want_norm, rtol=1e-6)
        for p, q in zip(ours, ref):
            assert torch.allclose(torch.view_as_real(p) if p.is_complex() else p, torch.view_as_real(q) if q.is_complex() else q,
                                  rtol=2e-6, atol=2e-7), step
    for p, q in zip(ours, ref):
        so, sr = opt_o.state[p], opt_r.state[q]
        assert int(so["step"]) == int(sr["step"]) == 4
        assert torch.allclose(torch.view_as_real(so["exp_avg"]) if p.is_complex() else so["exp_avg"],
                              torch.view_as_real(sr["exp_avg"]) if p.is_complex() else sr["exp_avg"], rtol=1e-5, atol=1e-8)


def test_fused_adamw_adds_the_analytic_kl_gradient_in_kernel():
    """set_analytic_kl: same update as materialising kl_weight * dKL/d(mu, lv) into .grad first (models.py:100-104)."""
    def make():
        torch.manual_seed(21)
        mu = torch.nn.Parameter(torch.randn(3, 5, 4, 6, dtype=torch.cfloat, device=DEV).permute(2, 3, 0, 1))
        lv = torch.nn.Parameter((torch.rand(3, 5, 4, 6, device=DEV) * 4 - 6).permute(2, 3, 0, 1))
        other = torch.nn.Parameter(torch.randn(33, device=DEV))
        return [mu, lv, other]

    ours, ref = make(), make()
    opt_o = FusedAdamW(ours, lr=1e-2, weight_decay=1e-3, max_grad_norm=0.5)
    opt_r = torch.optim.AdamW(ref, lr=1e-2, weight_decay=1e-3)
    gs = torch.tensor(3e-2, device=DEV)
    for step in range(3):
        torch.manual_seed(50 + step)
        grads = [torch.randn_like(p) for p in ours]
        for p, q, g in zip(ours, ref, grads):
            p.grad, q.grad = g.clone(), g.clone()
        with torch.no_grad():
            ref[0].grad += gs * ref[0]                                          # d KL / d mu = mu (re and im parts)
            ref[1].grad += gs * (-0.5 * (1 - torch.exp(ref[1])))                # d KL / d lv
        want_norm = torch.nn.utils.clip_grad_norm_(ref, 0.5)
        opt_o.set_analytic_kl(ours[:2], gs)
        opt_o.step()
        opt_r.step()
        assert torch.allclose(opt_o.grad_norm(), want_norm, rtol=2e-6)
        for p, q in zip(ours, ref):
            assert torch.allclose(torch.view_as_real(p) if p.is_complex() else p, torch.view_as_real(q) if q.is_complex() else q,
                                  rtol=2e-6, atol=2e-7), step


def test_fused_adamw_continues_from_a_torch_adamw_state_dict():
    """Reference checkpoints carry torch.optim.AdamW state (training/trainer.py:452): moments in logical-shape layout."""
    torch.manual_seed(31)
    native = torch.randn(2, 3, 4, 5, dtype=torch.cfloat, device=DEV)
    ours = [torch.nn.Parameter(native.clone().permute(2, 3, 0, 1)), torch.nn.Parameter(torch.randn(7, 3, device=DEV))]
    ref = [torch.nn.Parameter(native.permute(2, 3, 0, 1).contiguous()), torch.nn.Parameter(ours[1].detach().clone())]
    opt_r = torch.optim.AdamW(ref, lr=2e-3, weight_decay=1e-2)
    grads = [[torch.randn_like(p) for p in ref] for _ in range(4)]
    for step in range(2):
        for q, g in zip(ref, grads[step]):
            q.grad = g.clone()
        opt_r.step()
    with torch.no_grad():
        for p, q in zip(ours, ref):
            p.copy_(q)
    opt_o = FusedAdamW(ours, lr=2e-3, weight_decay=1e-2)
    import copy
    opt_o.load_state_dict(copy.deepcopy(opt_r.state_dict()))      # (a checkpoint from disk: no tensors shared with opt_r)
    for step in range(2, 4):
        for p, q, g in zip(ours, ref, grads[step]):
            p.grad, q.grad = g.clone(), g.clone()
        opt_o.step()
        opt_r.step()
    for p, q in zip(ours, ref):
        assert torch.allclose(torch.view_as_real(p) if p.is_complex() else p, torch.view_as_real(q) if q.is_complex() else q,
                              rtol=2e-6, atol=2e-7)
    assert int(opt_o.state[ours[0]]["step"]) == 4
