"""Pins the oracle (oracle/pno_oracle.py, oracle/philox.py) to the live reference through tests/golden/*.npz
(made by tests/golden/make_golden.py from the unmodified reference models.py / training/losses.py)."""
import numpy as np
import pytest
import torch

from conftest import rel_l2, sub, to_torch
from oracle import philox
from oracle import pno_oracle as orc

TOL = 2e-6          # fp32 oracle vs fp32 reference: same algorithm, only summation order differs
LAYERS = ["even", "oddw", "nyquist", "fullrows", "overlap", "default_init", "oddco"]


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in got) == want


def test_philox_normals_are_standard():
    z = philox.normal_activation(99, 0, 0, (4, 16, 32, 32), dtype=np.float64)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs((z ** 4).mean() - 3) < 0.1
    # batch_offset addresses the same global stream
    z2 = philox.normal_activation(99, 0, 0, (2, 16, 32, 32), batch_offset=2, dtype=np.float64)
    assert np.array_equal(z[2:], z2)


@pytest.mark.parametrize("name", LAYERS)
@pytest.mark.parametrize("dense", [False, True])
def test_layer_against_reference(golden, name, dense):
    g = sub(golden("layer_cases.npz"), name + ".")
    p = to_torch(sub(g, "p."))
    for v in p.values():
        v.requires_grad_(True)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    up = torch.from_numpy(g["g"])
    # deterministic
    y = orc.spectral_layer(x, p, "", sample=False, dense=dense)
    assert rel_l2(y, g["det.y"]) < TOL
    y.backward(up)
    assert rel_l2(x.grad, g["det.dx"]) < TOL
    assert rel_l2(p["weights1"].grad, g["det.dw1"]) < TOL
    assert rel_l2(p["weights2"].grad, g["det.dw2"]) < TOL
    # eval gate
    assert rel_l2(orc.spectral_layer(x, p, "", sample=True, training=False, dense=dense), g["evalgate.y"]) < TOL
    # sampled
    x.grad = None
    for v in p.values():
        v.grad = None
    y = orc.spectral_layer(x, p, "", orc.PhiloxNoise(1234, 3), site=7, sample=True, training=True, dense=dense)
    assert rel_l2(y, g["smp.y"]) < TOL
    y.backward(up)
    assert rel_l2(x.grad, g["smp.dx"]) < TOL
    for k in ("weights1", "weights2", "weights1_log_var", "weights2_log_var"):
        assert rel_l2(p[k].grad, g["smp.d" + k]) < TOL, k
    # KL
    for v in p.values():
        v.grad = None
    kl = orc.kl_layer(p, "")
    assert rel_l2(kl, g["kl"]) < 1e-6
    kl.backward()
    for k in ("weights1", "weights2", "weights1_log_var", "weights2_log_var"):
        assert rel_l2(p[k].grad, g["kl.d" + k]) < 1e-6, k


@pytest.mark.parametrize("case", ["gelu", "relu", "stress"])
def test_model_against_reference(golden, case):
    act = "relu" if case == "relu" else "gelu"
    gtol = 1e-3 if case == "stress" else 2e-5      # see make_golden.py: "stress" is ill-conditioned by design
    g = sub(golden("model_cases.npz"), case + ".")
    p = to_torch(sub(g, "p."))
    for v in p.values():
        v.requires_grad_(True)
    x, tgt = torch.from_numpy(g["x"]), torch.from_numpy(g["target"])
    y = orc.pno_forward(p, x, orc.PhiloxNoise(1234, 5), True, True, act)
    assert rel_l2(y, g["train.y"]) < 5e-6
    loss = torch.mean((y - tgt) ** 2) + 1e-4 * orc.pno_kl(p)
    assert rel_l2(loss, g["train.loss"]) < 5e-6
    loss.backward()
    for k, v in p.items():
        want = g["train.grad." + k]
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        assert rel_l2(got, want) < gtol, k
    assert rel_l2(orc.pno_kl(p), g["kl"]) < 1e-6
    with torch.no_grad():
        assert rel_l2(orc.pno_forward(p, x, None, False, True, act), g["train_nosample.y"]) < 5e-6
        assert rel_l2(orc.pno_forward(p, x, None, True, False, act), g["eval_sample.y"]) < 5e-6
        assert rel_l2(orc.pno_forward(p, x, None, False, False, act, features_only=True), g["features"]) < 5e-6
        mean, std, _ = orc.mc_predict(p, x, 3, 1234, act)
        assert rel_l2(mean, g["mc.mean"]) < 5e-6
        assert rel_l2(std, g["mc.std"]) < 5e-6
        s1, s2 = orc.mc_moments(p, x, range(3), 1234, act)
        m2, sd2 = orc.mc_finalize(s1, s2, 3)
        assert rel_l2(m2, g["mc.mean"]) < 5e-6 and rel_l2(sd2, g["mc.std"]) < 2e-5


def test_fno_against_reference(golden):
    g = sub(golden("model_cases.npz"), "fno.")
    p = to_torch(sub(g, "p."))
    x = torch.from_numpy(g["x"])
    assert rel_l2(orc.fno_forward(p, x), g["y"]) < 5e-6
    assert rel_l2(orc.fno_forward(p, x, dense=True), g["y"]) < 5e-6


def test_losses_against_reference(golden):
    g = golden("loss_cases.npz")
    mean, lv, tgt = (torch.from_numpy(g[k]) for k in ("mean", "log_var", "target"))
    res = orc.elbo_loss((mean, lv), tgt, kl=torch.tensor(float(g["kl_in"])))
    for k in ("kl", "nll", "calibration", "total", "reconstruction"):
        assert abs(float(res[k]) - float(g["tuple." + k])) < 1e-5 * max(1, abs(float(g["tuple." + k]))), k
    res = orc.elbo_loss(mean, tgt, kl=None)
    for k in ("kl", "total", "reconstruction"):
        assert abs(float(res[k]) - float(g["tensor." + k])) < 1e-5 * max(1, abs(float(g["tensor." + k]))), k


def test_dense_matches_fft_in_fp64():
    """The two oracles (pocketfft route and dense fp64 twiddles) agree to fp64 round-off."""
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 10, 14, generator=gen, dtype=torch.float64)
    w1 = torch.randn(3, 2, 4, 6, generator=gen, dtype=torch.complex128)
    w2 = torch.randn(3, 2, 4, 6, generator=gen, dtype=torch.complex128)
    assert rel_l2(orc.spectral_conv_dense(x, w1, w2), orc.spectral_conv_fft(x, w1, w2)) < 1e-13


def test_mode_limits_raise():
    x = torch.randn(1, 2, 8, 8)
    w = torch.randn(2, 2, 9, 3, dtype=torch.complex64)
    with pytest.raises(RuntimeError):
        orc.spectral_conv_fft(x, w, w)
    w = torch.randn(2, 2, 3, 6, dtype=torch.complex64)
    with pytest.raises(RuntimeError):
        orc.spectral_conv_fft(x, w, w)
    with pytest.raises(RuntimeError):
        orc.spectral_conv_dense(x, w, w)


METRIC_CASES = ["calibrated", "overconfident", "underconfident", "tiny_std"]


@pytest.mark.parametrize("name", METRIC_CASES)
def test_calibration_metrics_against_reference(golden, name):
    """oracle uq_stats / uq_metrics vs the reference's CalibrationMetrics (metrics.py:21-392)."""
    g = golden("metrics_cases.npz")
    pred, std, tgt = g[name + ".pred"], g[name + ".std"], g[name + ".target"]
    m = orc.uq_metrics(pred, std, tgt)
    for k, v in m.items():
        ref = float(g[f"{name}.m.{k}"])
        assert abs(v - ref) <= 2e-6 * max(1.0, abs(ref)), (k, v, ref)
    s = orc.uq_stats(pred, std, tgt, z_levels=[orc._ndtri(0.5 + 0.68 / 2)], n_bins=10)
    assert abs(s[7] / s[0] - float(g[name + ".m.coverage_at_0.68"])) < 1e-7
    ece10 = orc.uq_metrics_from_stats(s, levels=(0.68,), alphas=(), n_bins=10)["ece"]
    assert abs(ece10 - float(g[name + ".m.ece_10bins"])) < 2e-6
