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
