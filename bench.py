#!/usr/bin/env python
"""Headline benchmark: MC-sample fields/s of `predict_with_uncertainty` at 64x64, hidden 256, modes 20 (BASELINE.json
config 3: batch 64, S MC samples), on N B200s of one node, plus the spectral-conv roofline, the other BASELINE configs, the
reference's own eager path on the same B200 and on the host CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--engine tf32x3|tf32|fp32] [--mc-samples S]
    python bench.py --scaling strong --mc-samples-total 100     # config 3 as written: S = 100 in total, sharded over the ranks
    python bench.py --workload train_cfg2                       # data-parallel PNOTrainer step (batch 32 per GPU) as the timed step
    python bench.py --impl reference ...                        # the UNMODIFIED reference (baseline/_ref) on the host CPU cores

A "step" is one `predict_with_uncertainty(x, num_samples=S*N)` over one batch of 64 synthetic fields: S samples per GPU (weak
scaling: per-GPU work fixed), parameters replicated, one NCCL all-reduce of the two moment buffers per step.
`value` = N*S*64 / (max-over-ranks CUDA-event time per step), inputs resident in HBM.  `e2e` = the same call from pinned host
memory (H2D of x, D2H of mean and std inside the timed region).  One JSON line on stdout (rank 0).

The headline engine is `tf32x3`, the fp32-PARITY engine (error-compensated 3xTF32, rel-L2 <= 1e-5 against the reference's
fp32 path: tests/test_gpu_baseline_shapes.py).  The reduced-precision `tf32` engine (north_star's <= 2e-2 mode) is measured with
the same steps / warm-up and reported under `reduced_precision`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(B=64, C=256, H=64, W=64, modes=20, layers=4, in_dim=3)
METRIC = "MC-sample fields/sec at 64^2 hid256 modes20 (predict_with_uncertainty, batch 64)"
UNIT = "fields/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.marks = {}

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self, name):
        self.marks[name] = time.time()

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, start="start", end="end"):
        t0, t1 = self.marks.get(start), self.marks.get(end)
        rows = [r for t, r in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.1)]
        if not rows:                       # region shorter than one sampling period: fall back to the nearest samples
            rows = [r for t, r in self.rows if t1 is None or t <= t1 + 0.3][-3:]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].startswith("Active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_model(engine, device):
    import pno_physics_bench_b200 as pno
    torch.manual_seed(0)
    model = pno.ProbabilisticNeuralOperator(CFG["in_dim"], CFG["C"], CFG["layers"], CFG["modes"], engine=engine)
    return model.to(device)


def event_time(fn, iters, warmup=2, flush=None):
    """mean CUDA-event milliseconds of fn() on torch's current stream (the stream the library launches on)."""
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def block_kernel_times(engine, device, iters=5):
    """CUDA-event time of each of the four kernels of one PNO block at the cfg-3 shape, L2 flushed before every launch, with each
    kernel's ALGORITHMIC HBM bytes:
      local   : x read + local written (+ the two 1x1 weight matrices)
      dft_fwd : x read + truncated spectrum written
      mix     : spectrum read + spectrum written + mu (8 B) + log_var (4 B) per weight     <- SURVEY.md 8(d) parameter term
      dft_inv : spectrum read + local read + y written
    """
    from pno_physics_bench_b200 import _lib
    lib = _lib.load()
    B, C, H, W, m = CFG["B"], CFG["C"], CFG["H"], CFG["W"], CFG["modes"]
    M = 2 * m * m
    plan = _lib.plan(H, W, m, m, device)
    g = torch.Generator(device=device).manual_seed(0)
    x = torch.randn(B, C, H, W, device=device, generator=g)
    loc, y = torch.empty_like(x), torch.empty_like(x)
    xr, xi, yr, yi = (torch.randn(M, B, C, device=device, generator=g) for _ in range(4))
    mu1, mu2 = (torch.randn(m, m, C, C, 2, device=device, generator=g) / C for _ in range(2))
    lv1, lv2 = (torch.full((m, m, C, C), -4.6, device=device) for _ in range(2))
    wm, wl = torch.randn(C, C, device=device, generator=g) / 16, torch.randn(C, C, device=device, generator=g) / 64
    bm, bl = torch.zeros(C, device=device), torch.full((C,), -3.0, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)      # > 126 MB L2
    st, eng = _lib.stream_ptr(), _lib.ENGINES[engine]
    it = [0]

    def call(rc):
        _lib.check(rc, "kernel")
        it[0] += 1

    calls = {
        "local": lambda: call(lib.pno_local_fwd(eng, x.data_ptr(), B, C, C, H * W, wm.data_ptr(), bm.data_ptr(), wl.data_ptr(),
                                                 bl.data_ptr(), 1234, it[0], 2, 0, loc.data_ptr(), 0, st)),
        "dft_fwd": lambda: call(lib.pno_dft_fwd(plan, eng, x.data_ptr(), B, C, 0, xr.data_ptr(), xi.data_ptr(), st)),
        "mix": lambda: call(lib.pno_mix_fwd(plan, eng, xr.data_ptr(), xi.data_ptr(), B, C, C, mu1.data_ptr(), mu2.data_ptr(),
                                            lv1.data_ptr(), lv2.data_ptr(), 1234, it[0], 1, yr.data_ptr(), yi.data_ptr(), st)),
        "dft_inv": lambda: call(lib.pno_dft_inv(plan, eng, yr.data_ptr(), yi.data_ptr(), B, C, 0, loc.data_ptr(), _lib.ACT_GELU,
                                                y.data_ptr(), 0, st)),
    }
    act = 4.0 * B * C * H * W            # one fp32 activation tensor
    spec = 4.0 * M * B * C * 2           # one truncated spectrum (re + im planes)
    nbytes = {"local": 2 * act + 8.0 * C * C, "dft_fwd": act + spec, "mix": 2 * spec + 12.0 * C * C * M,
              "dft_inv": spec + 2 * act}
    # algorithmic flops: two 1x1 convs; W-axis + H-axis contractions of the pruned transforms; complex channel mix
    dft = 4.0 * B * C * H * W * m + 8.0 * B * C * H * 2 * m * m
    flops = {"local": 2 * 2.0 * B * H * W * C * C, "dft_fwd": dft, "mix": 8.0 * B * C * C * M, "dft_inv": dft}
    # tensor-core passes per product in units of one kind::tf32 MMA: single pass 1; fp32-parity engine 3 (hi*hi + lo*hi + hi*lo in
    # tf32), except the local branch, which runs its two cross terms as kind::f16 MMAs on bf16 copies at twice the rate (1 + 2/2)
    passes = {n: (1.0 if engine != "tf32x3" else (2.0 if n == "local" else 3.0)) for n in flops}
    out = {}
    for name, fn in calls.items():
        us = event_time(fn, iters, 2, flush) * 1e3
        out[name] = {"us": round(us, 1), "bytes": nbytes[name], "gbs": round(nbytes[name] / us / 1e3, 1), "flops": flops[name],
                     "tf32_passes": passes[name]}
    return out


def ncu_traffic(kernel, engine="tf32x3"):
    """dram bytes (read + write) per launch of `kernel` from the committed `ncu --set full` summaries, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(engine, {}).get(kernel)
    return None


KERNEL_NAMES = {"local": "k_tc_local (1x1 mean + log-var convs + reparameterised noise)",
                "dft_fwd": "k_tc_dft_fwd (pruned forward transform)",
                "mix": "k_tc_mix_fwd (weight sampling + per-mode complex channel mix)",
                "dft_inv": "k_tc_dft_inv (pruned inverse transform + local add + GELU)"}


def layer_roofline(engine, device):
    """Roofline of the dominant kernel of the block (the slowest of the four) + the whole spectral layer / block."""
    k = block_kernel_times(engine, device)
    hbm, tc, src = measured_peaks()
    tf32_peak = tc / 2.0          # TF/s: kind::tf32 runs at half the measured dense bf16 rate (no separate tf32 measurement exists)
    for v in k.values():          # the roof that binds each kernel: HBM time of its algorithmic bytes vs tensor time of its MMA passes
        t_hbm = v["bytes"] / hbm / 1e3
        t_tc = v["flops"] * v["tf32_passes"] / tf32_peak / 1e6
        v["bound"] = "tensor" if t_tc > t_hbm else "hbm"
        v["roof_us"] = round(max(t_hbm, t_tc), 1)
        v["frac"] = round(max(t_hbm, t_tc) / v["us"], 4)
        v["tflops_tf32_equiv"] = round(v["flops"] * v["tf32_passes"] / v["us"] / 1e6, 1)
    top = max(k, key=lambda n: k[n]["us"])
    B, C, H, W, m = CFG["B"], CFG["C"], CFG["H"], CFG["W"], CFG["modes"]
    layer_bytes = 4.0 * B * C * H * W * 2 + 12.0 * C * C * 2 * m * m        # x in + y out + mu + log_var
    layer_us = k["dft_fwd"]["us"] + k["mix"]["us"] + k["dft_inv"]["us"]
    block_us = layer_us + k["local"]["us"]
    t = k[top]
    tensor = t["bound"] == "tensor"
    return {"bound": t["bound"], "achieved": t["tflops_tf32_equiv"] if tensor else t["gbs"], "peak": round(tf32_peak, 1) if tensor else hbm,
            "unit": "TFLOP/s" if tensor else "GB/s", "frac": t["frac"],
            "traffic": ncu_traffic(top, engine), "peak_source": src + ("; tf32 peak = measured bf16 / 2; flops in tf32-MMA equivalents "
                                                                "(1 tf32 + 2 bf16 passes = 2 per product)" if tensor else ""),
            "kernel": KERNEL_NAMES[top], "kernel_us": t["us"],
            "algorithmic_bytes": t["bytes"], "algorithmic_flops": t["flops"], "engine": engine,
            "timed": "each kernel alone, L2 flushed before every launch, right AFTER the timed region (loop clocks: `clocks_kernels`)",
            "kernels": {n: {"us": v["us"], "gbs": v["gbs"], "bound": v["bound"], "roof_us": v["roof_us"], "frac": v["frac"],
                            "frac_of_hbm": round(v["gbs"] / hbm, 4), "traffic": ncu_traffic(n, engine)} for n, v in k.items()},
            "layer": {"us": round(layer_us, 1), "algorithmic_bytes": layer_bytes,
                      "frac_of_hbm_roofline": round(layer_bytes / layer_us / 1e3 / hbm, 4)},
            "block_us": round(block_us, 1)}


# ----------------------------------------------------------------------------------------------------------------------
# the reference itself (baseline/_ref = `pip install --target` of the unmodified /root/reference)
# ----------------------------------------------------------------------------------------------------------------------
def import_reference():
    """The reference's own `pno_physics_bench.models` from baseline/_ref, or None when the install is not in the tree."""
    if not os.path.isdir(os.path.join(REF_DIR, "pno_physics_bench")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import contextlib
    import io
    try:
        with contextlib.redirect_stdout(io.StringIO()):          # the package prints a warning about matplotlib on import
            from pno_physics_bench import models as ref_models
        return ref_models
    except Exception as e:                                       # noqa: BLE001
        log(f"bench: reference import failed: {e}")
        return None


def cpu_baseline(threads=None, reps=1):
    """The reference's own `predict_with_uncertainty` (unmodified models.py from baseline/_ref; the oracle port when the install is
    absent) on the host cores: ONE MC sample of the full cfg-3 batch (weight sampling does not shrink with the batch, so the batch
    is not reduced); the MC loop is strictly linear in S, so fields/s of one sample is fields/s of the loop."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    B = CFG["B"]
    torch.manual_seed(1)
    x = torch.randn(B, CFG["in_dim"], CFG["H"], CFG["W"])
    ref = import_reference()
    best = float("inf")
    if ref is not None:
        torch.manual_seed(0)
        model = ref.ProbabilisticNeuralOperator(CFG["in_dim"], CFG["C"], CFG["layers"], CFG["modes"])
        model.train()
        with torch.no_grad():
            for _ in range(reps):
                t0 = time.perf_counter()
                model(x, sample=True)          # one iteration of the loop body of predict_with_uncertainty (models.py:361-364)
                best = min(best, time.perf_counter() - t0)
        kind, what = "reference", "unmodified reference models.py (baseline/_ref)"
    else:
        from oracle import pno_oracle as orc
        import pno_physics_bench_b200 as pno
        torch.manual_seed(0)
        model = pno.ProbabilisticNeuralOperator(CFG["in_dim"], CFG["C"], CFG["layers"], CFG["modes"])
        p = {k: v.detach().contiguous() for k, v in model.state_dict().items()}

        class TorchNoise:            # the reference draws with torch.randn_like; time that, not the numpy Philox restatement
            def activation(self, site, shape, dtype=torch.float32):
                return torch.randn(tuple(shape), dtype=dtype)

            def spectral(self, site, ci, co, m1, m2, dtype=torch.float32):
                return torch.randn(2, ci, co, m1, m2, dtype=dtype), torch.randn(2, ci, co, m1, m2, dtype=dtype)

        with torch.no_grad():
            for _ in range(reps):
                t0 = time.perf_counter()
                orc.pno_forward(p, x, TorchNoise(), True, True, "gelu")
                best = min(best, time.perf_counter() - t0)
        kind, what = "port", "oracle port of the reference algorithm (baseline/_ref not in the tree)"
    return {"value": round(B / best, 3), "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"1 MC sample x {B} fields of the cfg-3 model ({best:.2f} s, {what}, torch CPU, {threads} threads)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_baseline()
    base = None
    for _ in range(max(1, args.steps)):
        base = cpu_baseline()
        vals.append(base["value"])
    v = len(vals) / sum(1.0 / u for u in vals)          # total fields / total time
    base["value"] = round(v, 3)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(CFG["B"] / v * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3: predict_with_uncertainty 64x64 hid256 modes20 L4, bounded sample of 1 MC "
                                   "sample x 64 fields per step on the host CPU"},
            "cpu_baseline": base, "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0,
                                          "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def ref_cuda(device):
    """The UNMODIFIED reference classes run eagerly on this B200 (cuFFT / cuBLAS, TF32 off = torch default): the real bar to beat
    (BASELINE.md section 4.2).  cfg 3: fields/s of the MC loop; cfg 1: ms of one fwd+bwd."""
    ref = import_reference()
    if ref is None:
        return {"unavailable": "baseline/_ref is not in the tree"}
    out = {"allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32)}
    try:
        torch.manual_seed(0)
        model = ref.ProbabilisticNeuralOperator(CFG["in_dim"], CFG["C"], CFG["layers"], CFG["modes"]).to(device)
        x = torch.randn(CFG["B"], CFG["in_dim"], CFG["H"], CFG["W"], device=device)
        S = 4
        ms = event_time(lambda: model.predict_with_uncertainty(x, num_samples=S), 3, 1)
        out["cfg3"] = {"value": round(S * CFG["B"] / ms * 1e3, 1), "unit": UNIT, "ms_per_mc_sample": round(ms / S, 2),
                       "what": f"reference predict_with_uncertainty(x[64,3,64,64], num_samples={S}) eager on cuda"}
        del model, x
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        model = ref.ProbabilisticNeuralOperator(3, 32, 4, 12).to(device).train()
        x, y = torch.randn(8, 3, 64, 64, device=device), torch.randn(8, 1, 64, 64, device=device)

        def step():
            model.zero_grad(set_to_none=True)
            loss = torch.mean((model(x, sample=True) - y) ** 2) + 1e-4 * model.kl_divergence()
            loss.backward()
        out["cfg1"] = {"ms_per_step": round(event_time(step, 10, 3), 3), "what": "reference fwd+bwd (MSE + 1e-4 KL), batch 8, eager on cuda"}
        del model
        torch.manual_seed(0)
        fno = ref.FourierNeuralOperator(3, 128, 4, 32).to(device).eval()
        x4 = torch.randn(64, 3, 128, 128, device=device)
        with torch.no_grad():
            ms = event_time(lambda: fno(x4), 5, 2)
        out["cfg4"] = {"value": round(64 / ms * 1e3, 1), "unit": UNIT, "ms_per_batch": round(ms, 3),
                       "what": "reference FourierNeuralOperator(3,128,4,32).eval() batch 64 at 128x128, eager on cuda"}
        del fno, x4
        torch.cuda.empty_cache()
    except Exception as e:                                       # noqa: BLE001
        out["error"] = f"{type(e).__name__}: {e}"[:300]
    return out


# ----------------------------------------------------------------------------------------------------------------------
# the other BASELINE configs (step / batch time; parity for these shapes is tests/, this is the measurement)
# ----------------------------------------------------------------------------------------------------------------------
def other_configs(device, engine):
    import pno_physics_bench_b200 as pno
    from pno_physics_bench_b200 import _lib
    from pno_physics_bench_b200.training import ELBOLoss, PNOTrainer
    out = {}

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:                                   # noqa: BLE001
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    def cfg1():
        torch.manual_seed(0)
        model = pno.ProbabilisticNeuralOperator(3, 32, 4, 12).to(device).train()          # engine "auto": hid 32 is below the MMA tile
        x, y = torch.randn(8, 3, 64, 64, device=device), torch.randn(8, 1, 64, 64, device=device)

        def step(xx, yy):
            loss = torch.mean((model(xx, sample=True) - yy) ** 2) + 1e-4 * model.kl_divergence()
            loss.backward()
            return loss

        def eager():
            model.zero_grad(set_to_none=True)
            with pno.noise_scope(1, 0):
                step(x, y)
        c0 = _lib.stage_counts()
        ms_eager = event_time(eager, 10, 3)
        calls = _lib.tensor_core_calls(c0)
        from pno_physics_bench_b200.graphs import TrainStepGraph
        g = TrainStepGraph(model, None, x, y, step)
        k = [0]

        def graphed():
            k[0] += 1
            g.run(x, y, 1, k[0])
        ms_graph = event_time(graphed, 20, 3)
        return {"what": "PNO 64x64 in3 hid32 L4 modes12 batch 8: fwd+bwd, sample=True, MSE + 1e-4 KL", "engine": "auto",
                "ms_per_step": round(ms_graph, 3), "ms_per_step_eager": round(ms_eager, 3), "mode": "CUDA graph of fwd + loss + bwd",
                "tensor_core_stages": [s for s, (a, b) in calls.items() if b > 0]}

    def cfg2():
        torch.manual_seed(0)
        model = pno.ProbabilisticNeuralOperator(3, 256, 4, 20, engine=engine)
        trainer = PNOTrainer(model, ELBOLoss(kl_weight=1e-4), device=str(device), seed=1)
        x, y = torch.randn(32, 3, 64, 64, device=device), torch.randn(32, 1, 64, 64, device=device)
        ms = event_time(lambda: trainer.train_step(x, y), 4, 2)
        return {"what": "PNOTrainer step (sampled fwd, ELBO, bwd, clip 1.0, AdamW) 64x64 hid256 L4 modes20 batch 32", "engine": engine,
                "ms_per_step": round(ms, 3), "fields_per_s": round(32 / ms * 1e3, 1)}

    def cfg4():
        torch.manual_seed(0)
        fno = pno.FourierNeuralOperator(3, 128, 4, 32, engine=engine).to(device).eval()
        x = torch.randn(64, 3, 128, 128, device=device)
        c0 = _lib.stage_counts()
        with torch.no_grad():
            ms = event_time(lambda: fno(x), 5, 2)
        calls = _lib.tensor_core_calls(c0)
        return {"what": "FourierNeuralOperator(3,128,4,32).eval() batch 64 at 128x128 (sample=False path)", "engine": engine,
                "ms_per_batch": round(ms, 3), "fields_per_s": round(64 / ms * 1e3, 1),
                "cuda_core_stages": [s for s, (a, b) in calls.items() if a > 0 and s not in ("local_fwd",)]}

    def cfg5():
        torch.manual_seed(0)
        with torch.device(device):                 # 14.5 GB of parameters: initialise on the GPU
            model = pno.ProbabilisticNeuralOperator(3, 256, 4, 48, engine=engine).train()
        x, y = torch.randn(2, 3, 256, 256, device=device), torch.randn(2, 1, 256, 256, device=device)
        c0 = _lib.stage_counts()
        with torch.no_grad():
            ms_fwd = event_time(lambda: model(x, sample=True), 3, 1)
        calls = _lib.tensor_core_calls(c0)

        def step():
            model.zero_grad(set_to_none=True)
            loss = torch.mean((model(x, sample=True) - y) ** 2) + 1e-4 * model.kl_divergence()
            loss.backward()
        ms_step = event_time(step, 2, 1)
        return {"what": "PNO 256x256 hid256 L4 modes48, batch 2 (the per-GPU share of global batch 16 on 8 GPUs)", "engine": engine,
                "ms_sampled_forward": round(ms_fwd, 2), "ms_fwd_bwd": round(ms_step, 2),
                "cuda_core_stages": [s for s, (a, b) in calls.items() if a > 0 and s not in ("local_fwd",)]}

    guarded("cfg1", cfg1)          # (cfg 2 = the `train_cfg2` legs of the main line)
    guarded("cfg4", cfg4)
    guarded("cfg5", cfg5)
    return out


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default=os.environ.get("PNO_BENCH_ENGINE", "tf32x3"), choices=["fp32", "tf32x3", "tf32"],
                    help="tf32x3 = the fp32-parity engine (headline); tf32 = the reduced-precision (rel-L2 <= 2e-2) mode of BASELINE.json")
    ap.add_argument("--mc-samples", type=int, default=int(os.environ.get("PNO_BENCH_MC_SAMPLES", "20")),
                    help="MC samples per GPU per step (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--mc-samples-total", type=int, default=100, help="total MC samples per step with --scaling strong (config 3: 100)")
    ap.add_argument("--workload", default="mc_cfg3", choices=["mc_cfg3", "train_cfg2", "train_cfg5"])
    ap.add_argument("--parallel", default="dp", choices=["dp", "mp"], help="train_cfg2: data parallel or mode parallel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (other configs, reference on cuda, tf32)")
    ap.add_argument("--eager", action="store_true", help="run the MC loop eagerly instead of replaying the captured CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # NCCL's banner ("NCCL version ...", printed to stdout when NCCL_DEBUG=VERSION/INFO is set in the environment) must not
        # share stdout with the one JSON line of the contract: stdout points at stderr while the communicator is created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    from pno_physics_bench_b200 import _lib, graphs
    lib = _lib.load()

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    if args.workload in ("train_cfg2", "train_cfg5"):
        return run_train(args, device, rank, world, local, timed)

    model = build_model(args.engine, device)
    B = CFG["B"]
    strong = args.scaling == "strong"
    total_samples = args.mc_samples_total if strong else args.mc_samples * world
    torch.manual_seed(1)
    x_host = torch.randn(B, CFG["in_dim"], CFG["H"], CFG["W"]).pin_memory()
    x_dev = x_host.to(device)
    mean_host = torch.empty(B, 1, CFG["H"], CFG["W"]).pin_memory()
    std_host = torch.empty_like(mean_host)
    use_graph = not args.eager

    def mc_steps(mdl, S):
        def resident(i):
            return mdl.predict_with_uncertainty(x_dev, num_samples=S, seed=1000 + i, use_graph=use_graph)

        def e2e(i):
            xd = x_host.to(device, non_blocking=True)
            mean, std = mdl.predict_with_uncertainty(xd, num_samples=S, seed=2000 + i, use_graph=use_graph)
            mean_host.copy_(mean, non_blocking=True)
            std_host.copy_(std, non_blocking=True)
        return resident, e2e

    def launches_now():
        return lib.pno_launch_count() + graphs.replayed_launches()

    step_resident, step_e2e = mc_steps(model, total_samples)
    roof = strong_line = reduced = None
    with ClockSampler(local) as clocks:        # nvidia-smi needs ~1 s to start: launched before the warm-up, rows filtered by time
        for i in range(args.warmup):
            step_resident(i)
            step_e2e(i)
        launches0 = launches_now()
        clocks.mark("start")
        ms_total = timed(step_resident, args.steps)
        launches = launches_now() - launches0
        ms_e2e = timed(step_e2e, args.steps)
        clocks.mark("end")
        # per-kernel roofline right after the timed region: the part is at its loop clocks (power-capped), not cold
        if rank == 0:
            roof = layer_roofline(args.engine, device)
        clocks.mark("kernels_end")
        if world > 1:
            dist.barrier()
        fields = B * total_samples
        # config 3 as written (strong scaling: S = 100 in total) next to the weak-scaling headline, on every N
        if not strong and not args.no_extras:
            sr, se = mc_steps(model, args.mc_samples_total)
            for i in range(2):
                sr(i)
            ms_s = timed(sr, args.steps)
            ms_se = timed(se, args.steps)
            strong_line = {"scaling": "strong", "mc_samples_total": args.mc_samples_total,
                           "samples_on_busiest_rank": -(-args.mc_samples_total // world),
                           "value": round(B * args.mc_samples_total / (ms_s / args.steps * 1e-3), 2), "unit": UNIT,
                           "ms_per_step": round(ms_s / args.steps, 3),
                           "e2e_value": round(B * args.mc_samples_total / (ms_se / args.steps * 1e-3), 2)}
        # the reduced-precision engine with the same steps / warm-up
        if not args.no_extras and args.engine != "tf32":
            model.set_engine("tf32")
            rr, re_ = mc_steps(model, total_samples)
            for i in range(args.warmup):
                rr(i)
                re_(i)
            clocks.mark("tf32_start")
            ms_r = timed(rr, args.steps)
            ms_re = timed(re_, args.steps)
            clocks.mark("tf32_end")
            reduced = {"engine": "tf32", "dtype": "tf32 (single pass; rel-L2 <= 2e-2 tested on mean, std, loss and gradients)",
                       "value": round(fields / (ms_r / args.steps * 1e-3), 2), "unit": UNIT, "ms_per_step": round(ms_r / args.steps, 3),
                       "e2e_value": round(fields / (ms_re / args.steps * 1e-3), 2), "steps": args.steps, "warmup": args.warmup}
            if rank == 0:
                reduced["roofline"] = layer_roofline("tf32", device)
            clocks.mark("tf32_kernels_end")
            reduced["clocks"] = clocks.summary("tf32_start", "tf32_end")
            model.set_engine(args.engine)
            if world > 1:
                dist.barrier()
    ms_step = ms_total / args.steps
    value = fields / (ms_step * 1e-3)
    e2e_value = fields / (ms_e2e / args.steps * 1e-3)
    base = others = refgpu = train = None
    if not args.no_extras:
        del model
        graphs._MC_GRAPHS.clear()
        torch.cuda.empty_cache()
        # BASELINE config 2 on the same N GPUs: data-parallel and mode-parallel trainer steps (weak scaling, batch 32 per GPU)
        train = {}
        for par in ("dp", "mp"):
            # A failure of a secondary leg must not take the headline line with it.  Shape / tiling errors are raised identically on
            # every rank before any collective of the step, so all ranks skip the leg together; anything else shows up as a
            # collective timeout, which is the honest outcome for a broken exchange.
            try:
                train[par] = train_leg(args.engine, par, device, rank, world, timed, args.steps, 2)
            except Exception as e:                                   # noqa: BLE001
                train[par] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()
    def guarded(fn, *a):
        try:
            return fn(*a)
        except Exception as e:                                       # noqa: BLE001
            return {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank == 0 and world == 1 and not args.no_extras:
        others = guarded(other_configs, device, args.engine)
        refgpu = guarded(ref_cuda, device)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base = guarded(cpu_baseline)
    if world > 1:
        try:
            dist.barrier()
        except Exception as e:                                       # noqa: BLE001
            log(f"bench: barrier after the secondary legs failed: {e}")
    if rank == 0:
        # every MC sample streams all parameters once (2.5 GB) and the working set per layer (> 1 GB) exceeds the 126 MB L2
        S_per = f"{args.mc_samples_total} MC samples in total (strong scaling)" if strong else f"{args.mc_samples} MC samples per GPU per step"
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "tf32x3 (error-compensated 3xTF32 = fp32 parity, rel-L2 <= 1e-5)",
                                               "tf32": "tf32"}[args.engine],
                "data": "synthetic",
                "config": {"workload": f"cfg3: predict_with_uncertainty 64x64 in3 hid256 L4 modes20, batch {B} x {S_per}",
                           "engine": args.engine, "mc_loop": "eager launches" if args.eager else "CUDA graph of one sampled forward, replayed per sample",
                           "l2": "inputs larger than L2 (2.5 GB of parameters + >1 GB of activations streamed per sample)",
                           "parallelism": f"mc-sample sharding x{world}, 1 all-reduce of 2x[B,1,H,W] fp64 per step"},
                "clocks": clocks.summary("start", "end"), "gpu_launches": int(launches),
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                        "d2h_bytes_per_step": mean_host.numel() * 8},
                "roofline": roof, "clocks_kernels": clocks.summary("end", "kernels_end"), "cpu_baseline": base,
                "reduced_precision": reduced, "strong_scaling": strong_line, "train_cfg2": train, "other_configs": others, "ref_cuda": refgpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


TRAIN_CFGS = {   # BASELINE.json configs 2 and 5: (hidden, layers, modes, grid, batch per GPU)
    "train_cfg2": (256, 4, 20, 64, 32),
    "train_cfg5": (256, 4, 48, 256, 2),
}


def train_leg(engine, parallel, device, rank, world, timed, steps, warmup, cfg="train_cfg2"):
    """One PNOTrainer step of BASELINE config 2 (batch 32 per GPU) or 5 (256x256, 48 modes, batch 2 per GPU = 16 on 8 GPUs) as the
    timed step.  parallel = "dp": data parallel (replicated weights, per-parameter gradient all-reduce overlapped with backward);
    "mp": mode parallel (spectral weights sharded by modes, spectrum all-to-alls; mode_parallel.py)."""
    import pno_physics_bench_b200 as pno
    from pno_physics_bench_b200 import _lib
    from pno_physics_bench_b200.training import ELBOLoss, PNOTrainer
    lib = _lib.load()
    hid, layers, modes, hw, Bl = TRAIN_CFGS[cfg]
    torch.manual_seed(0)
    with torch.device(device):                 # config 5 has 14.5 GB of parameters: initialise on the GPU (same seed on every rank)
        model = pno.ProbabilisticNeuralOperator(3, hid, layers, modes, engine=engine)
    if parallel == "mp":
        from pno_physics_bench_b200.mode_parallel import ModeParallelPNO
        model = ModeParallelPNO(model, engine=engine)
        torch.cuda.empty_cache()
    trainer = PNOTrainer(model, ELBOLoss(kl_weight=1e-4), device=str(device), seed=1)
    torch.manual_seed(1 + rank)
    x_host, y_host = torch.randn(Bl, 3, hw, hw).pin_memory(), torch.randn(Bl, 1, hw, hw).pin_memory()
    x, y = x_host.to(device), y_host.to(device)
    loss_host = torch.empty(()).pin_memory()

    def resident(i):
        trainer.train_step(x, y, batch_offset=rank * Bl)

    def e2e(i):
        losses = trainer.train_step(x_host, y_host, batch_offset=rank * Bl)
        loss_host.copy_(losses["total"].detach().float(), non_blocking=True)

    for i in range(warmup):
        resident(i)
    l0 = lib.pno_launch_count()
    ms = timed(resident, steps) / steps
    launches = lib.pno_launch_count() - l0
    ms_e = timed(e2e, steps) / steps
    mem = torch.cuda.max_memory_allocated(device) / 2 ** 30
    if trainer.grad_sync is not None:
        trainer.grad_sync.remove()
    del trainer, model
    torch.cuda.empty_cache()
    return {"parallel": parallel, "engine": engine, "value": round(world * Bl / ms * 1e3, 2), "unit": UNIT, "ms_per_step": round(ms, 3),
            "e2e_value": round(world * Bl / ms_e * 1e3, 2), "batch_per_gpu": Bl, "gpu_launches": int(launches),
            "peak_mem_gib": round(mem, 2), "h2d_bytes_per_step": (x_host.numel() + y_host.numel()) * 4, "d2h_bytes_per_step": 4}


def run_train(args, device, rank, world, local, timed):
    """--workload train_cfg2 [--parallel dp|mp]: the data- or mode-parallel trainer step as the headline of the line."""
    import torch.distributed as dist
    with ClockSampler(local) as clocks:
        clocks.mark("start")
        leg = train_leg(args.engine, args.parallel, device, rank, world, timed, args.steps, args.warmup, args.workload)
        clocks.mark("end")
    if rank == 0:
        how = ("data parallel: per-parameter NCCL all-reduce (AVG) of the 2.5 GB of gradients, launched from autograd hooks while backward "
               "continues" if args.parallel == "dp" else
               "mode parallel: spectral weights / gradients / AdamW state sharded by modes, two all-to-alls of the truncated spectrum per "
               "layer and direction")
        hid, layers, modes, hw, Bl = TRAIN_CFGS[args.workload]
        line = {"metric": f"PNOTrainer fields/sec at {hw}^2 hid{hid} modes{modes} (one ELBO step: fwd, bwd, clip, AdamW)", "value": leg["value"],
                "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.engine, "data": "synthetic",
                "config": {"workload": f"{args.workload[6:]}: PNOTrainer.train_step {hw}x{hw} in3 hid{hid} L{layers} modes{modes}, batch {Bl} per GPU (global {Bl * world})",
                           "engine": args.engine, "parallelism": f"x{world}, {how}",
                           "l2": "inputs larger than L2 (10 GB (cfg 2) / 58 GB (cfg 5) of parameters / gradients / AdamW state per step)"},
                "clocks": clocks.summary("start", "end"), "gpu_launches": leg["gpu_launches"], "peak_mem_gib": leg["peak_mem_gib"],
                "e2e": {"value": leg["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": leg["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": leg["d2h_bytes_per_step"]}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
