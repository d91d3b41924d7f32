"""Parity of the CUDA path (through the Python mirror -> C ABI -> sm_100a kernels) with the reference, on a B200.

* against tests/golden/*.npz (outputs of the unmodified reference with this repo's Philox stream injected), and
* against the oracle (oracle/pno_oracle.py, pinned to the reference by tests/test_oracle_golden.py) on larger seeded
  inputs with the kernels' own noise materialised by `pno_philox_materialize_*` and injected.
Tolerance: rel-L2 <= 1e-5 for the fp32 engine (north_star); the ill-conditioned "stress" model case uses 1e-3 because the
reference's own fp32 gradients are ~1e-4 from the fp64 truth there (tests/golden/make_golden.py).
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2, sub, to_torch
from oracle import philox
from oracle import pno_oracle as orc
import pno_physics_bench_b200 as pno
from pno_physics_bench_b200 import functional as F_
from pno_physics_bench_b200 import _lib

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda"
LAYERS = ["even", "oddw", "nyquist", "fullrows", "overlap", "default_init", "oddco"]


def make_layer(g):
    b, ci, co, h, w, m1, m2 = (int(v) for v in g["shape"])
    layer = pno.SpectralConv2d_Probabilistic(ci, co, m1, m2)
    layer.load_state_dict(to_torch(sub(g, "p.")))
    return layer.to(DEV)


def test_philox_stream_matches_oracle():
    bits = F_.philox_bits(0xDEADBEEF12345678, 3, 9, (1 << 33) + 5, 4096).cpu().numpy().view(np.uint32)
    want = philox.quad_bits(0xDEADBEEF12345678, 3, 9, np.arange(4096, dtype=np.uint64) + np.uint64((1 << 33) + 5))
    assert np.array_equal(bits, want)
    z = F_.materialize_activation_noise(77, 2, 4, (3, 5, 12, 20), batch_offset=6).cpu().numpy()
    zw = philox.normal_activation(77, 2, 4, (3, 5, 12, 20), batch_offset=6, dtype=np.float64)
    assert np.abs(z - zw).max() < 5e-6
    z = F_.materialize_activation_noise(77, 2, 4, (2, 3, 7, 9)).cpu().numpy()        # unaligned element count
    assert np.abs(z - philox.normal_activation(77, 2, 4, (2, 3, 7, 9), dtype=np.float64)).max() < 5e-6
    re, im = F_.materialize_spectral_noise(5, 1, 2, 3, 5, 4, 6)
    wre, wim = philox.normal_spectral_weights(5, 1, 2, 3, 5, 4, 6, dtype=np.float64)
    assert np.abs(re.cpu().numpy() - wre).max() < 5e-6 and np.abs(im.cpu().numpy() - wim).max() < 5e-6


@pytest.mark.parametrize("name", LAYERS)
def test_layer_against_reference_golden(golden, name):
    g = sub(golden("layer_cases.npz"), name + ".")
    layer = make_layer(g)
    x = torch.from_numpy(g["x"]).to(DEV).requires_grad_(True)
    up = torch.from_numpy(g["g"]).to(DEV)
    layer.train()
    y = layer(x, sample=False)
    assert rel_l2(y, g["det.y"]) < TOL
    y.backward(up)
    assert rel_l2(x.grad, g["det.dx"]) < TOL
    assert rel_l2(layer.weights1.grad, g["det.dw1"]) < TOL
    assert rel_l2(layer.weights2.grad, g["det.dw2"]) < TOL
    assert layer.weights1_log_var.grad is None
    # eval gate (models.py:79-80): bit-identical to sample=False, no noise consumed
    layer.eval()
    with torch.no_grad():
        y_eval = layer(x, sample=True)
        assert torch.equal(y_eval, layer(x, sample=False))
    assert rel_l2(y_eval, g["evalgate.y"]) < TOL
    # sampled, site 7, sample_idx 3 (as in make_golden.py)
    layer.train()
    layer.zero_grad()
    x.grad = None
    layer.noise_site = 7
    with pno.noise_scope(1234, 3):
        y = layer(x, sample=True)
    assert rel_l2(y, g["smp.y"]) < TOL
    y.backward(up)
    assert rel_l2(x.grad, g["smp.dx"]) < TOL
    for k in ("weights1", "weights2", "weights1_log_var", "weights2_log_var"):
        assert rel_l2(getattr(layer, k).grad, g["smp.d" + k]) < TOL, k
    # KL
    layer.zero_grad()
    kl = layer.kl_divergence()
    assert kl.dim() == 0 and rel_l2(kl, g["kl"]) < 1e-6
    kl.backward()
    for k in ("weights1", "weights2", "weights1_log_var", "weights2_log_var"):
        assert rel_l2(getattr(layer, k).grad, g["kl.d" + k]) < 1e-6, k


def trained_like_(layer, gen):
    with torch.no_grad():
        for w, lv in ((layer.weights1, layer.weights1_log_var), (layer.weights2, layer.weights2_log_var)):
            s = (1.0 / w.shape[0]) ** 0.5
            w.copy_(torch.complex(torch.randn(w.shape, generator=gen) * s, torch.randn(w.shape, generator=gen) * s))
            lv.copy_(torch.rand(lv.shape, generator=gen) * 4 - 6)


@pytest.mark.parametrize("shape", [(8, 32, 32, 64, 64, 12, 12),       # BASELINE config 1 layer
                                   (3, 20, 36, 48, 40, 9, 11),        # nothing divisible by the tile sizes
                                   (2, 64, 64, 256, 256, 48, 48),     # config-5 grid and modes (row-chunked transform)
                                   (70, 8, 8, 16, 16, 4, 4)])         # batch larger than one tile
def test_layer_against_oracle_with_injected_noise(shape):
    b, ci, co, h, w, m1, m2 = shape
    gen = torch.Generator().manual_seed(sum(shape))
    layer = pno.SpectralConv2d_Probabilistic(ci, co, m1, m2)
    trained_like_(layer, gen)
    layer = layer.to(DEV).train()
    layer.noise_site = 4
    x = torch.randn(b, ci, h, w, generator=gen)
    up = torch.randn(b, co, h, w, generator=gen)
    xg = x.to(DEV).requires_grad_(True)
    with pno.noise_scope(99, 11):
        y = layer(xg, sample=True)
    y.backward(up.to(DEV))
    re, im = F_.materialize_spectral_noise(99, 11, 4, ci, co, m1, m2)
    p = {k: v.detach().cpu().contiguous().requires_grad_(True) for k, v in layer.state_dict().items()}
    xo = x.clone().requires_grad_(True)
    yo = orc.spectral_layer(xo, p, "", orc.InjectedNoise({4: (re.cpu(), im.cpu())}), site=4)
    yo.backward(up)
    assert rel_l2(y, yo) < TOL
    assert rel_l2(xg.grad, xo.grad) < TOL
    for k in p:
        assert rel_l2(getattr(layer, k).grad, p[k].grad) < TOL, k
    # the sampled weights the module reports are the ones the kernel used
    with pno.noise_scope(99, 11):
        w1, w2 = layer.sample_weights()
    w1o = orc.sample_spectral_weights(p["weights1"], p["weights1_log_var"], re[0].cpu(), im[0].cpu())
    assert rel_l2(w1, w1o) < 1e-6


@pytest.mark.parametrize("case", ["gelu", "relu", "stress"])
def test_model_against_reference_golden(golden, case):
    act = "relu" if case == "relu" else "gelu"
    gtol = 1e-3 if case == "stress" else 2e-5
    g = sub(golden("model_cases.npz"), case + ".")
    model = pno.ProbabilisticNeuralOperator(3, 8, 2, 4, activation=act)
    model.load_state_dict(to_torch(sub(g, "p.")))
    model = model.to(DEV).train()
    x, tgt = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["target"]).to(DEV)
    with pno.noise_scope(1234, 5):
        y = model(x, sample=True)
    assert rel_l2(y, g["train.y"]) < TOL
    loss = torch.mean((y - tgt) ** 2) + 1e-4 * model.kl_divergence()
    assert rel_l2(loss, g["train.loss"]) < TOL
    loss.backward()
    for k, v in model.named_parameters():
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        assert rel_l2(got, g["train.grad." + k]) < gtol, k
    assert rel_l2(model.kl_divergence(), g["kl"]) < 1e-6
    with torch.no_grad():
        assert rel_l2(model(x, sample=False), g["train_nosample.y"]) < TOL
        model.eval()
        assert rel_l2(model(x, sample=True), g["eval_sample.y"]) < TOL
        assert rel_l2(model.get_final_features(x), g["features"]) < TOL
        model.train()
        mean_d, std_d = model.predict_distributional(x)
        assert not model.training                       # the reference leaves the model in eval mode (models.py:320)
        assert rel_l2(mean_d, g["distributional.mean"]) < TOL and rel_l2(std_d, g["distributional.std"]) < TOL
    mean, std = model.predict_with_uncertainty(x, num_samples=3, seed=1234)
    assert not model.training                           # restored (models.py:375-377)
    assert rel_l2(mean, g["mc.mean"]) < TOL and rel_l2(std, g["mc.std"]) < 2e-5


def test_fno_against_reference_golden(golden):
    g = sub(golden("model_cases.npz"), "fno.")
    fno = pno.FourierNeuralOperator(3, 6, 2, 3)
    fno.load_state_dict(to_torch(sub(g, "p.")))
    fno = fno.to(DEV)
    x = torch.from_numpy(g["x"]).to(DEV)
    with torch.no_grad():
        assert rel_l2(fno(x), g["y"]) < TOL
    mean, std = fno.predict_with_uncertainty(x)
    assert rel_l2(mean, g["y"]) < TOL and float(std.abs().max()) == 0.0


def test_model_cfg1_against_oracle_with_injected_noise():
    """BASELINE config 1 (64x64, in 3, hid 32, 4 layers, modes 12, batch 8): fwd + bwd, sample=True."""
    gen = torch.Generator().manual_seed(5)
    torch.manual_seed(5)
    model = pno.ProbabilisticNeuralOperator(3, 32, 4, 12)
    for layer in model.fourier_layers:
        trained_like_(layer, gen)
    with torch.no_grad():
        for conv in [model.lift_log_var, model.project_log_var, *model.linear_log_var]:
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=gen) * 0.2)
            conv.bias.copy_(torch.randn(conv.bias.shape, generator=gen) - 3.0)
    model = model.to(DEV).train()
    x = torch.randn(8, 3, 64, 64, generator=gen)
    tgt = torch.randn(8, 1, 64, 64, generator=gen)
    seed, sidx = 4242, 2
    with pno.noise_scope(seed, sidx):
        y = model(x.to(DEV), sample=True)
    loss = torch.mean((y - tgt.to(DEV)) ** 2) + 1e-4 * model.kl_divergence()
    loss.backward()
    table = {0: F_.materialize_activation_noise(seed, sidx, 0, (8, 32, 64, 64)).cpu(),
             13: F_.materialize_activation_noise(seed, sidx, 13, (8, 1, 64, 64)).cpu()}
    for l in range(4):
        re, im = F_.materialize_spectral_noise(seed, sidx, 1 + 3 * l, 32, 32, 12, 12)
        table[1 + 3 * l] = (re.cpu(), im.cpu())
        table[2 + 3 * l] = F_.materialize_activation_noise(seed, sidx, 2 + 3 * l, (8, 32, 64, 64)).cpu()
    p = {k: v.detach().cpu().contiguous().requires_grad_(True) for k, v in model.state_dict().items()}
    yo = orc.pno_forward(p, x, orc.InjectedNoise(table), True, True, "gelu")
    lo = torch.mean((yo - tgt) ** 2) + 1e-4 * orc.pno_kl(p)
    lo.backward()
    assert rel_l2(y, yo) < TOL and rel_l2(loss, lo) < TOL
    for k, v in model.named_parameters():
        assert rel_l2(v.grad, p[k].grad) < 5e-5, k


def test_mc_moments_and_edge_cases():
    torch.manual_seed(1)
    model = pno.ProbabilisticNeuralOperator(3, 8, 2, 3).to(DEV)
    x = torch.randn(2, 3, 16, 16, device=DEV)
    model.eval()
    mean1, std1 = model.predict_with_uncertainty(x, num_samples=1, seed=3)
    assert torch.isnan(std1).all() and torch.isfinite(mean1).all()          # torch.std of one sample (models.py:368)
    assert not model.training
    a = model.predict_with_uncertainty(x, num_samples=6, seed=3)
    b = model.predict_with_uncertainty(x, num_samples=6, seed=3)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])              # same seed -> same stream
    c = model.predict_with_uncertainty(x, num_samples=6, seed=4)
    assert not torch.equal(a[0], c[0])
    assert (a[1] >= 0).all()
    # explicit moments agree with stacking the samples like the reference does
    model.train()
    ys = []
    with torch.no_grad():
        for s in range(6):
            with pno.noise_scope(3, s):
                ys.append(model(x, sample=True))
    ys = torch.stack(ys)
    assert rel_l2(a[0], ys.mean(0)) < 1e-6 and rel_l2(a[1], ys.std(0)) < 1e-5
    with pytest.warns(UserWarning):
        model.predict_with_uncertainty(x[:1, :, :8, :8], num_samples=1001, seed=0)
    with pytest.raises(RuntimeError):                                       # re-wrapped inside the MC loop (models.py:373-374)
        model.predict_with_uncertainty(torch.randn(2, 5, 16, 16, device=DEV), num_samples=2)


def test_error_behaviour_matches_reference():
    layer = pno.SpectralConv2d_Probabilistic(2, 2, 9, 3).to(DEV)
    with pytest.raises(RuntimeError):                   # m1 > H
        layer(torch.randn(1, 2, 8, 8, device=DEV))
    layer = pno.SpectralConv2d_Probabilistic(2, 2, 3, 6).to(DEV)
    with pytest.raises(RuntimeError):                   # m2 > W/2+1
        layer(torch.randn(1, 2, 8, 8, device=DEV))
    layer = pno.SpectralConv2d_Probabilistic(2, 2, 3, 3).to(DEV)
    for dt in (torch.float64, torch.float16, torch.bfloat16):
        with pytest.raises(RuntimeError):
            layer(torch.randn(1, 2, 8, 8, device=DEV).to(dt))
    with pytest.raises(RuntimeError):                   # wrong channel count (einsum error in the reference)
        layer(torch.randn(1, 3, 8, 8, device=DEV))


def test_readme_block_against_oracle():
    gen = torch.Generator().manual_seed(8)
    torch.manual_seed(8)
    blk = pno.PNOBlock(6, 10, 4, 5, activation=None, local_noise="weight")
    trained_like_(blk.conv, gen)
    with torch.no_grad():
        blk.w_log_var.weight.copy_(torch.rand(blk.w_log_var.weight.shape, generator=gen) * 3 - 5)
    blk = blk.to(DEV).train()
    x = torch.randn(3, 6, 20, 24, generator=gen)
    xg = x.to(DEV).requires_grad_(True)
    with pno.noise_scope(31, 1):
        y = blk(xg, sample=True)
    up = torch.randn(y.shape, generator=gen)
    y.backward(up.to(DEV))
    re, im = F_.materialize_spectral_noise(31, 1, 1, 6, 10, 4, 5)
    eps_w = F_.materialize_activation_noise(31, 1, 3, (10, 6, 1, 1))
    p = {k: v.detach().cpu().contiguous().requires_grad_(True) for k, v in blk.state_dict().items()}
    xo = x.clone().requires_grad_(True)
    yo = orc.readme_block(xo, p, orc.InjectedNoise({1: (re.cpu(), im.cpu()), 3: eps_w.cpu()}), 1, 3)
    yo.backward(up)
    assert rel_l2(y, yo) < TOL and rel_l2(xg.grad, xo.grad) < TOL
    for k, v in blk.named_parameters():
        if p[k].grad is not None:
            assert rel_l2(v.grad, p[k].grad) < TOL, k
    assert rel_l2(blk.kl_divergence(), orc.readme_block_kl(p)) < 1e-6
    blk.eval()
    with torch.no_grad():
        ye = blk(xg, sample=True)
        yeo = orc.readme_block(x, {k: v.detach() for k, v in p.items()}, None, sample=True, training=False)
    assert rel_l2(ye, yeo) < TOL
    # the block the model runs (activation-space noise + activation)
    blk2 = pno.PNOBlock(6, 6, 4, 5, activation="gelu", local_noise="activation").to(DEV).train()
    with pno.noise_scope(1, 0):
        assert torch.isfinite(blk2(xg.detach())).all()


def test_trainer_step_matches_oracle():
    """P10: three AdamW steps through PNOTrainer.train_step vs the same steps on the oracle (CPU autograd)."""
    from pno_physics_bench_b200.training import PNOTrainer, ELBOLoss
    gen = torch.Generator().manual_seed(21)
    torch.manual_seed(21)
    model = pno.ProbabilisticNeuralOperator(3, 8, 2, 4)
    for layer in model.fourier_layers:
        trained_like_(layer, gen)
    with torch.no_grad():
        for conv in [model.lift_log_var, model.project_log_var, *model.linear_log_var]:
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=gen) * 0.3)
            conv.bias.copy_(torch.randn(conv.bias.shape, generator=gen) - 3.0)
    p = {k: v.detach().clone().contiguous().requires_grad_(True) for k, v in model.state_dict().items()}
    trainer = PNOTrainer(model, device=DEV, seed=77, kl_weight=1e-4)
    opt = torch.optim.AdamW(p.values(), lr=1e-3, weight_decay=1e-4)
    x = torch.randn(4, 3, 16, 16, generator=gen)
    tgt = torch.randn(4, 1, 16, 16, generator=gen)
    for step in range(3):
        losses = trainer.train_step(x, tgt)
        # oracle: same adapter semantics (return_kl=True -> (mean, clamp(log_var), kl) -> ELBOLoss)
        opt.zero_grad()
        noise = orc.PhiloxNoise(77, step)
        h = orc.pno_forward(p, x, noise, True, True, "gelu", features_only=True)
        mean = orc.conv1x1(h, p["project_mean.weight"], p["project_mean.bias"])
        lv = torch.clamp(orc.conv1x1(h, p["project_log_var.weight"], p["project_log_var.bias"]), -10, 10)
        ref = orc.elbo_loss((mean, lv), tgt, kl=orc.pno_kl(p), kl_weight=1e-4)
        ref["total"].backward()
        torch.nn.utils.clip_grad_norm_(p.values(), 1.0)
        opt.step()
        assert rel_l2(losses["total"], ref["total"]) < 2e-5, step
        assert rel_l2(losses["kl"], ref["kl"]) < 1e-6
    for k, v in model.state_dict().items():
        assert rel_l2(v, p[k]) < 2e-5, k


def test_mc_sample_sharding_is_invariant():
    """SURVEY.md 8(e) / P9 on one GPU: the moments of the sample ranges two ranks would draw (GLOBAL sample indices), merged
    the way the all-reduce merges them, give the unsharded predict_with_uncertainty result."""
    from pno_physics_bench_b200 import distributed as D
    torch.manual_seed(3)
    model = pno.ProbabilisticNeuralOperator(input_dim=3, hidden_dim=16, num_layers=2, modes=4).to(DEV)
    x = torch.randn(3, 3, 16, 16, device=DEV)
    S, seed = 7, 4242
    mean, std = model.predict_with_uncertainty(x, num_samples=S, seed=seed)
    model.train()
    s1 = torch.zeros(3, 1, 16, 16, dtype=torch.float64, device=DEV)
    s2 = torch.zeros_like(s1)
    with torch.no_grad():
        for rank in range(2):
            r1, r2 = torch.zeros_like(s1), torch.zeros_like(s1)
            first, count = D.shard_range(S, rank, 2)
            for s in range(first, first + count):
                with pno.noise_scope(seed, s, 0):
                    F_.mc_accumulate(model(x, sample=True), r1, r2)
            s1 += r1
            s2 += r2
    mean2, std2 = F_.mc_finalize(s1, s2, S)
    assert rel_l2(mean2, mean) < 1e-6 and rel_l2(std2, std) < 1e-6


def test_batch_sharding_draws_the_same_activation_noise():
    """Data-parallel invariance (P11): a batch shard run with its GLOBAL batch offset reproduces the rows of the full batch."""
    torch.manual_seed(5)
    model = pno.ProbabilisticNeuralOperator(input_dim=3, hidden_dim=16, num_layers=2, modes=4).to(DEV).train()
    x = torch.randn(4, 3, 16, 16, device=DEV)
    with torch.no_grad():
        with pno.noise_scope(11, 2, 0):
            full = model(x, sample=True)
        with pno.noise_scope(11, 2, 2):
            tail = model(x[2:].contiguous(), sample=True)
    assert rel_l2(tail, full[2:]) < 1e-6


def test_mc_statistics_match_reference_with_independent_noise():
    """P14 (statistical): with INDEPENDENT noise streams (ours: in-kernel Philox; oracle: torch.randn) the MC predictive mean
    and std of a small model agree within sampling error."""
    torch.manual_seed(8)
    model = pno.ProbabilisticNeuralOperator(input_dim=3, hidden_dim=8, num_layers=2, modes=4).to(DEV)
    with torch.no_grad():                       # keep activation noise in the unclamped regime so that statistics are well behaved
        for conv in [*model.linear_log_var, model.lift_log_var, model.project_log_var]:
            conv.weight.mul_(0.02)
            conv.bias.fill_(-2.0)
    x = torch.randn(2, 3, 16, 16, device=DEV)
    S = 768
    mean, std = model.predict_with_uncertainty(x, num_samples=S, seed=31337)
    p = {k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}

    class TorchNoise:
        def __init__(self):
            self.g = torch.Generator().manual_seed(99)

        def activation(self, site, shape, dtype=torch.float32):
            return torch.randn(tuple(shape), generator=self.g, dtype=dtype)

        def spectral(self, site, ci, co, m1, m2, dtype=torch.float32):
            return (torch.randn(2, ci, co, m1, m2, generator=self.g, dtype=dtype),
                    torch.randn(2, ci, co, m1, m2, generator=self.g, dtype=dtype))

    noise = TorchNoise()
    with torch.no_grad():
        ys = torch.stack([orc.pno_forward(p, x.cpu(), noise, True, True, "gelu") for _ in range(S)])
    m_ref, s_ref = ys.mean(0), ys.std(0)
    se = (s_ref ** 2 / S + std.cpu() ** 2 / S).sqrt()            # standard error of the difference of two independent MC means
    z = ((mean.cpu() - m_ref) / se).abs()
    assert float(z.max()) < 5.5 and float((z > 3).float().mean()) < 0.02
    ratio = std.cpu() / s_ref
    assert float((ratio - 1).abs().mean()) < 0.05 and float((ratio - 1).abs().max()) < 0.25
