#!/usr/bin/env python
"""Headline benchmark: MC-sample fields/s of `predict_with_uncertainty` at 64x64, hidden 256, modes 20 (BASELINE.json
config 3: batch 64, S MC samples), on N B200s of one node, plus the spectral-conv roofline and a CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--engine fp32|tf32x3|tf32] [--mc-samples S]
    python bench.py --impl reference ...        # the reference's algorithm (oracle port) on the host CPU cores

A "step" is one `predict_with_uncertainty(x, num_samples=S*N)` over one batch of 64 synthetic fields: S samples per
GPU (weak scaling: per-GPU work fixed), parameters replicated, one NCCL all-reduce of the two moment buffers per step.
`value` = N*S*64 / (max-over-ranks CUDA-event time per step), inputs resident in HBM.  `e2e` = the same call from
pinned host memory (H2D of x, D2H of mean and std inside the timed region).  One JSON line on stdout (rank 0).
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
        self.t0 = self.t1 = None          # wall-clock bounds of the timed region (rows outside are dropped)

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

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = [r for t, r in self.rows if (self.t0 is None or t >= self.t0) and (self.t1 is None or t <= self.t1 + 0.1)]
        if not rows:                       # region shorter than one sampling period: fall back to the nearest samples
            rows = [r for _, r in self.rows[-3:]]
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


def block_kernel_times(engine, device, iters=5):
    """CUDA-event time (on torch's current stream, the stream the kernels are launched on) of each of the four kernels of
    one PNO block at the cfg-3 shape, L2 flushed before every launch, with each kernel's ALGORITHMIC HBM bytes:
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
    calls = {
        "local": lambda it: lib.pno_local_fwd(eng, x.data_ptr(), B, C, C, H * W, wm.data_ptr(), bm.data_ptr(), wl.data_ptr(),
                                              bl.data_ptr(), 1234, it, 2, 0, loc.data_ptr(), 0, st),
        "dft_fwd": lambda it: lib.pno_dft_fwd(plan, eng, x.data_ptr(), B, C, 0, xr.data_ptr(), xi.data_ptr(), st),
        "mix": lambda it: lib.pno_mix_fwd(plan, eng, xr.data_ptr(), xi.data_ptr(), B, C, C, mu1.data_ptr(), mu2.data_ptr(),
                                          lv1.data_ptr(), lv2.data_ptr(), 1234, it, 1, yr.data_ptr(), yi.data_ptr(), st),
        "dft_inv": lambda it: lib.pno_dft_inv(plan, eng, yr.data_ptr(), yi.data_ptr(), B, C, 0, loc.data_ptr(), _lib.ACT_GELU,
                                              y.data_ptr(), 0, st),
    }
    act = 4.0 * B * C * H * W            # one fp32 activation tensor
    spec = 4.0 * M * B * C * 2           # one truncated spectrum (re + im planes)
    nbytes = {"local": 2 * act + 8.0 * C * C, "dft_fwd": act + spec, "mix": 2 * spec + 12.0 * C * C * M,
              "dft_inv": spec + 2 * act}
    out = {}
    for name, fn in calls.items():
        ts = []
        for it in range(iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn(it), name)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        us = sum(ts) / len(ts)
        out[name] = {"us": round(us, 1), "bytes": nbytes[name], "gbs": round(nbytes[name] / us / 1e3, 1)}
    return out


def ncu_traffic(kernel):
    """dram bytes (read + write) per launch of `kernel` from the committed `ncu --set full` summary, or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(kernel)
    return None


def layer_roofline(engine, device):
    """Roofline of the dominant kernel of the block (the slowest of the four) + the whole spectral layer / block."""
    k = block_kernel_times(engine, device)
    hbm, tc, src = measured_peaks()
    top = max(k, key=lambda n: k[n]["us"])
    B, C, H, W, m = CFG["B"], CFG["C"], CFG["H"], CFG["W"], CFG["modes"]
    layer_bytes = 4.0 * B * C * H * W * 2 + 12.0 * C * C * 2 * m * m        # x in + y out + mu + log_var
    layer_us = k["dft_fwd"]["us"] + k["mix"]["us"] + k["dft_inv"]["us"]
    block_us = layer_us + k["local"]["us"]
    names = {"local": "k_tc_local (1x1 mean + log-var convs + reparameterised noise)",
             "dft_fwd": "k_tc_dft_fwd (pruned forward transform)",
             "mix": "k_tc_mix_fwd (weight sampling + per-mode complex channel mix)",
             "dft_inv": "k_tc_dft_inv (pruned inverse transform + local add + GELU)"}
    return {"bound": "hbm", "achieved": k[top]["gbs"], "peak": hbm, "unit": "GB/s", "frac": round(k[top]["gbs"] / hbm, 4),
            "traffic": ncu_traffic(top), "peak_source": src, "kernel": names[top], "kernel_us": k[top]["us"],
            "algorithmic_bytes": k[top]["bytes"],
            "kernels": {n: {"us": v["us"], "gbs": v["gbs"], "frac": round(v["gbs"] / hbm, 4), "traffic": ncu_traffic(n)}
                        for n, v in k.items()},
            "layer": {"us": round(layer_us, 1), "algorithmic_bytes": layer_bytes,
                      "frac_of_hbm_roofline": round(layer_bytes / layer_us / 1e3 / hbm, 4)},
            "block_us": round(block_us, 1)}


def cpu_baseline(sample_b=64, threads=None, reps=1):
    """The reference's algorithm (oracle/pno_oracle.py, pinned to the reference by tests/golden) on the host cores:
    ONE MC sample of the full cfg-3 batch (weight sampling does not shrink with the batch, so the batch is not reduced);
    the MC loop is strictly linear in S, so fields/s of one sample is fields/s of the loop."""
    from oracle import pno_oracle as orc
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    import pno_physics_bench_b200 as pno
    torch.manual_seed(0)
    model = pno.ProbabilisticNeuralOperator(CFG["in_dim"], CFG["C"], CFG["layers"], CFG["modes"])
    p = {k: v.detach().contiguous() for k, v in model.state_dict().items()}
    x = torch.randn(sample_b, CFG["in_dim"], CFG["H"], CFG["W"])

    class TorchNoise:            # the reference draws with torch.randn_like; time that, not the numpy Philox restatement
        def activation(self, site, shape, dtype=torch.float32):
            return torch.randn(tuple(shape), dtype=dtype)

        def spectral(self, site, ci, co, m1, m2, dtype=torch.float32):
            return torch.randn(2, ci, co, m1, m2, dtype=dtype), torch.randn(2, ci, co, m1, m2, dtype=dtype)

    best = float("inf")
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.pno_forward(p, x, TorchNoise(), True, True, "gelu")
            best = min(best, time.perf_counter() - t0)
    return {"value": round(sample_b / best, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"1 MC sample x {sample_b} fields of the cfg-3 model ({best:.2f} s, torch CPU, {threads} threads)"}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default=os.environ.get("PNO_BENCH_ENGINE", "tf32"), choices=["fp32", "tf32x3", "tf32"],
                    help="tf32 = the reduced-precision (bf16-compute class, rel-L2 <= 2e-2) mode of BASELINE.json; tf32x3 = the fp32-parity mode")
    ap.add_argument("--no-strict", action="store_true", help="skip the secondary fp32-parity (tf32x3) measurement")
    ap.add_argument("--mc-samples", type=int, default=int(os.environ.get("PNO_BENCH_MC_SAMPLES", "20")),
                    help="MC samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from pno_physics_bench_b200 import _lib
    lib = _lib.load()
    model = build_model(args.engine, device)
    B, S = CFG["B"], args.mc_samples
    torch.manual_seed(1)
    x_host = torch.randn(B, CFG["in_dim"], CFG["H"], CFG["W"]).pin_memory()
    x_dev = x_host.to(device)
    mean_host = torch.empty(B, 1, CFG["H"], CFG["W"]).pin_memory()
    std_host = torch.empty_like(mean_host)
    total_samples = S * world

    def step_resident(i):
        return model.predict_with_uncertainty(x_dev, num_samples=total_samples, seed=1000 + i)

    def step_e2e(i):
        xd = x_host.to(device, non_blocking=True)
        mean, std = model.predict_with_uncertainty(xd, num_samples=total_samples, seed=2000 + i)
        mean_host.copy_(mean, non_blocking=True)
        std_host.copy_(std, non_blocking=True)

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

    # per-kernel roofline first: each kernel timed alone (L2 flushed), before the long timed region heats the part into its power cap
    roof = None
    if rank == 0:
        roof = layer_roofline(args.engine, device)
    if world > 1:
        dist.barrier()
    with ClockSampler(local) as clocks:        # nvidia-smi needs ~1 s to start: launched before the warm-up, rows filtered by time
        for i in range(args.warmup):
            step_resident(i)
            step_e2e(i)
        launches0 = lib.pno_launch_count()
        clocks.mark_start()
        ms_total = timed(step_resident, args.steps)
        launches = lib.pno_launch_count() - launches0
        ms_e2e = timed(step_e2e, args.steps)
        clocks.mark_end()
    ms_step = ms_total / args.steps
    fields = B * total_samples
    value = fields / (ms_step * 1e-3)
    e2e_value = fields / (ms_e2e / args.steps * 1e-3)
    # secondary measurement in the fp32-parity mode (error-compensated 3xTF32): one warm-up + one timed step
    strict = None
    if not args.no_strict and args.engine != "tf32x3":
        model.set_engine("tf32x3")
        step_resident(0)
        ms_strict = timed(step_resident, 1)
        model.set_engine(args.engine)
        strict = {"engine": "tf32x3", "value": round(fields / (ms_strict * 1e-3), 2), "unit": UNIT,
                  "ms_per_step": round(ms_strict, 3), "parity": "rel-L2 <= 1e-5 vs the fp32 reference (tests/test_gpu_tc.py)"}
    base = None
    if rank == 0:
        if strict is not None:
            ks = block_kernel_times("tf32x3", device, iters=3)
            strict["block_kernels_us"] = {n: v["us"] for n, v in ks.items()}
        if not args.no_cpu_baseline and world == 1:
            base = cpu_baseline()
    if world > 1:
        dist.barrier()
    if rank == 0:
        # every MC sample streams all parameters once (2.5 GB) and the working set per layer (> 1 GB) exceeds the 126 MB L2
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "tf32x3 (fp32-equivalent)", "tf32": "tf32"}[args.engine],
                "data": "synthetic",
                "config": {"workload": f"cfg3: predict_with_uncertainty 64x64 in3 hid256 L4 modes20, batch {B} x {S} MC "
                                       f"samples per GPU per step", "engine": args.engine,
                           "l2": "inputs larger than L2 (2.5 GB of parameters + >1 GB of activations streamed per sample)",
                           "precision_mode": {"tf32": "north_star's reduced-precision (bf16-compute class) mode: tf32 operands, fp32 "
                                                      "accumulate, rel-L2 <= 2e-2 tested; fp32-parity numbers under `strict`",
                                              "tf32x3": "fp32-parity mode (3xTF32, rel-L2 <= 1e-5)",
                                              "fp32": "fp32 CUDA-core engine"}[args.engine],
                           "parallelism": f"mc-sample sharding x{world}, 1 all-reduce of 2x[B,1,H,W] fp64 per step"},
                "clocks": clocks.summary(), "gpu_launches": int(launches),
                "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                        "d2h_bytes_per_step": mean_host.numel() * 8},
                "roofline": roof, "cpu_baseline": base, "strict": strict}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
