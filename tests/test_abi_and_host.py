"""CPU-side checks: the C-ABI library loads and exports every symbol include/pno_b200.h declares; the Python mirror has the
reference's constructor / state_dict contract; host logic (sharding, losses, strided parameters) behaves."""
import copy
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2, sub, to_torch
import pno_physics_bench_b200 as pno
from pno_physics_bench_b200 import _lib
from pno_physics_bench_b200.distributed import shard_range
from pno_physics_bench_b200.training import ELBOLoss, PNOLoss


def header_functions():
    text = open(os.path.join(ROOT, "include", "pno_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pno_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.pno_version() == 100
    assert sorted(_lib.SIGNATURES) == names


def test_no_cpu_fallback():
    layer = pno.SpectralConv2d_Probabilistic(2, 2, 2, 2)
    with pytest.raises(_lib.PnoLibraryError):
        layer(torch.randn(1, 2, 8, 8))
    model = pno.ProbabilisticNeuralOperator(3, 4, 1, 2)
    with pytest.raises(_lib.PnoLibraryError):
        model(torch.randn(1, 3, 8, 8))


def test_constructor_contract_and_errors():
    for kw in (dict(input_dim=0), dict(hidden_dim=-1), dict(num_layers=0), dict(modes=0), dict(uncertainty_type="x"),
               dict(posterior="x"), dict(activation="elu")):
        with pytest.raises(ValueError):
            pno.ProbabilisticNeuralOperator(**{**dict(input_dim=3, hidden_dim=4, num_layers=1, modes=2), **kw})
    m = pno.ProbabilisticNeuralOperator(3, 4, 1, 2, activation="swish")
    assert m.activation_name == "gelu"
    with pytest.raises(TypeError):
        m("not a tensor")
    with pytest.raises(ValueError):
        m(torch.randn(3, 8, 8))
    with pytest.raises(ValueError):
        m(torch.randn(1, 2, 8, 8))
    with pytest.raises(TypeError):
        m.predict_with_uncertainty([1, 2])
    with pytest.raises(ValueError):
        m.predict_with_uncertainty(torch.randn(1, 3, 8, 8), num_samples=0)
    with pytest.raises(ValueError):
        pno.PNOBlock(2, 2, 2, 2, activation="nope")
    assert hasattr(m, "compute_kl_divergence")


def test_state_dict_matches_reference_layout(golden):
    """Keys, shapes and dtypes equal the reference's (golden p.* arrays are reference state_dicts); values round-trip."""
    g = sub(golden("model_cases.npz"), "gelu.")
    ref_sd = to_torch(sub(g, "p."))
    m = pno.ProbabilisticNeuralOperator(3, 8, 2, 4)
    sd = m.state_dict()
    assert list(sd) == list(ref_sd)
    for k in sd:
        assert sd[k].shape == ref_sd[k].shape and sd[k].dtype == ref_sd[k].dtype, k
    m.load_state_dict(ref_sd)
    for k, v in m.state_dict().items():
        assert torch.equal(v, ref_sd[k]), k
    w = m.fourier_layers[0].weights1
    assert w.dtype == torch.complex64 and w.permute(2, 3, 0, 1).is_contiguous()
    assert copy.deepcopy(m).fourier_layers[1].weights2_log_var.permute(2, 3, 0, 1).is_contiguous()
    f = pno.FourierNeuralOperator(3, 6, 2, 3)
    assert list(f.state_dict()) == list(sub(golden("model_cases.npz"), "fno.p."))


def test_default_init_draws_like_reference():
    """Spectral weights ~ scale*U[0,1) complex, log_var = log(0.01), log-var convs = -3 (models.py:30-31, 59-60, 256-262)."""
    torch.manual_seed(3)
    layer = pno.SpectralConv2d_Probabilistic(4, 5, 3, 2)
    torch.manual_seed(3)
    want1 = (1 / 20) * torch.rand(4, 5, 3, 2, dtype=torch.cfloat)
    want2 = (1 / 20) * torch.rand(4, 5, 3, 2, dtype=torch.cfloat)
    assert torch.equal(layer.weights1.detach(), want1) and torch.equal(layer.weights2.detach(), want2)
    assert torch.allclose(layer.weights1_log_var.detach(), torch.full((4, 5, 3, 2), float(np.log(np.float32(0.01)))))
    m = pno.ProbabilisticNeuralOperator(3, 4, 2, 2)
    for conv in [m.lift_log_var, m.project_log_var, *m.linear_log_var]:
        assert (conv.weight == -3).all() and (conv.bias == -3).all()


def test_strided_parameters_work_with_adamw_and_clip():
    layer = pno.SpectralConv2d_Probabilistic(3, 4, 2, 2)
    ref = {k: v.detach().clone().contiguous().requires_grad_(True) for k, v in layer.named_parameters()}
    opt_a = torch.optim.AdamW(layer.parameters(), lr=1e-2, weight_decay=1e-2)
    opt_b = torch.optim.AdamW(ref.values(), lr=1e-2, weight_decay=1e-2)
    gen = torch.Generator().manual_seed(0)
    for _ in range(3):
        for (k, p), q in zip(layer.named_parameters(), ref.values()):
            gq = torch.randn(q.shape, generator=gen, dtype=q.dtype)
            q.grad = gq.clone()
            p.grad = gq.permute(2, 3, 0, 1).contiguous().permute(2, 3, 0, 1)      # kernel-native strides
        torch.nn.utils.clip_grad_norm_(layer.parameters(), 1.0)
        torch.nn.utils.clip_grad_norm_(ref.values(), 1.0)
        opt_a.step()
        opt_b.step()
    for (k, p), q in zip(layer.named_parameters(), ref.values()):
        assert rel_l2(p, q) < 1e-6, k
        assert p.permute(2, 3, 0, 1).is_contiguous()


def test_shard_range_partitions():
    for n in (1, 5, 100, 101):
        for ws in (1, 2, 3, 8):
            got = [shard_range(n, r, ws) for r in range(ws)]
            assert sum(c for _, c in got) == n
            nxt = 0
            for first, c in got:
                assert first == nxt
                nxt += c
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def test_losses_match_reference(golden):
    g = golden("loss_cases.npz")
    mean, lv, tgt = (torch.from_numpy(g[k]) for k in ("mean", "log_var", "target"))
    res = ELBOLoss(kl_weight=1e-4, num_samples=5)((mean, lv, torch.tensor(float(g["kl_in"]))), tgt)
    for k in ("kl", "nll", "calibration", "total", "reconstruction"):
        assert abs(float(res[k]) - float(g["tuple." + k])) < 1e-6 * max(1.0, abs(float(g["tuple." + k]))), k

    class NoKL:
        pass

    res = ELBOLoss(kl_weight=1e-4)(mean, tgt, NoKL())
    for k in ("kl", "total", "reconstruction"):
        assert abs(float(res[k]) - float(g["tensor." + k])) < 1e-6, k
    with pytest.raises(ValueError):
        PNOLoss()((mean,), tgt)


def test_metrics_host_logic_matches_the_oracle(golden):
    """Host side of the calibration report (metrics.py mirror): quantiles and the metrics computed from the statistics vector."""
    import numpy as np
    import torch
    from oracle import pno_oracle as orc
    from pno_physics_bench_b200 import metrics as M
    from scipy import stats
    for p in (1e-6, 0.01, 0.3, 0.5, 0.75, 0.975, 0.995, 0.99995):       # the levels in use are 0.75 .. 0.995
        assert abs(M.norm_ppf(p) - stats.norm.ppf(p)) < 1e-11 * max(1.0, abs(stats.norm.ppf(p)))
    g = golden("metrics_cases.npz")
    acc = M.UQAccumulator(device="cpu")
    z = [orc._ndtri(0.5 + lv / 2) for lv in acc.levels]
    stats_vec = orc.uq_stats(g["calibrated.pred"], g["calibrated.std"], g["calibrated.target"], z, acc.alphas, acc.num_bins)
    acc.acc = torch.from_numpy(stats_vec)
    got = acc.compute()
    for k, v in got.items():
        ref = float(g["calibrated.m." + k])
        assert abs(v - ref) <= 2e-6 * max(1.0, abs(ref)), (k, v, ref)
    assert abs(acc.ece(acc.stats(), "l2") ** 2 - sum(
        (abs(h / c - s / c) ** 2) * (c / stats_vec[0])
        for c, s, h in np.asarray(stats_vec[7 + len(acc.levels) + len(acc.alphas):]).reshape(-1, 3) if c > 0)) < 1e-12


# ----------------------------------------------------------------------------------------------------------------------
# which stages of which BASELINE config the tensor-core engines tile (host-only query; no GPU needed)
# ----------------------------------------------------------------------------------------------------------------------
BASELINE_SHAPES = {          # (H, W, m1, m2, B per GPU, Ci, Co)
    "cfg1": (64, 64, 12, 12, 8, 32, 32),
    "cfg2": (64, 64, 20, 20, 32, 256, 256),
    "cfg3": (64, 64, 20, 20, 64, 256, 256),
    "cfg4": (128, 128, 32, 32, 64, 128, 128),
    "cfg5": (256, 256, 48, 48, 2, 256, 256),
}
# stages of BASELINE configs 4 / 5 that still run on CUDA cores: none (the transforms of 256-row images and the fp32-parity
# transforms at 128x128 / 32 modes are the two-pass tensor-core form, csrc/pno_tc_dft2.cu)
KNOWN_GAPS = set()
FWD_STAGES = ("dft_fwd", "mix_fwd", "dft_inv", "local_fwd")
BWD_STAGES = ("mix_bwd_dx", "mix_bwd_dw", "local_bwd")


def test_tensor_core_tilings_cover_the_baseline_configs():
    from pno_physics_bench_b200 import _lib
    sup = {(c, e, st): _lib.engine_supports(st, e, *BASELINE_SHAPES[c]) for c in BASELINE_SHAPES for e in ("tf32x3", "tf32")
           for st in _lib.STAGES}
    # the headline inference config and the trainer config: every stage, both engines
    for e in ("tf32x3", "tf32"):
        for st in FWD_STAGES:
            assert sup[("cfg3", e, st)], (e, st)
        for st in FWD_STAGES + BWD_STAGES:
            assert sup[("cfg2", e, st)], (e, st)
    # cfg 4 (deterministic FNO inference, 128x128, modes 32) and cfg 5 (256x256, modes 48, a training config): every stage, both
    # engines -- except the stages listed in KNOWN_GAPS (kept in step with DESIGN.md section 5)
    for c in ("cfg4", "cfg5"):
        for e in ("tf32x3", "tf32"):
            for st in FWD_STAGES + (BWD_STAGES if c == "cfg5" else ()):
                assert sup[(c, e, st)] != ((c, e, st) in KNOWN_GAPS), (c, e, st)
    # cfg 1 (hidden 32): the transforms are tiled; its 32-channel GEMMs are below the 128-row MMA tile and run on CUDA cores
    for e in ("tf32x3", "tf32"):
        assert sup[("cfg1", e, "dft_fwd")] and sup[("cfg1", e, "dft_inv")]
        assert not sup[("cfg1", e, "mix_fwd")] and not sup[("cfg1", e, "local_fwd")]
    assert not _lib.engine_supports("dft_fwd", "fp32", *BASELINE_SHAPES["cfg3"])
    assert not _lib.engine_supports("dft_fwd", "tf32", 8, 8, 9, 3, 1, 2, 2)           # m1 > H


def test_engine_names():
    from pno_physics_bench_b200 import _lib, models
    assert models._engine_code("auto") == _lib.ENGINE_TC_TF32X3 | _lib.ENGINE_ALLOW_FALLBACK
    assert models._engine_code("tf32") == _lib.ENGINE_TC_TF32
    assert models._engine_code("tf32", allow_fallback=True) == _lib.ENGINE_TC_TF32 | _lib.ENGINE_ALLOW_FALLBACK
    assert models._engine_code("fp32", allow_fallback=True) == _lib.ENGINE_FP32
    with pytest.raises(ValueError):
        models._engine_code("bf16")
