"""ctypes binding of libpno_b200.so (C ABI in include/pno_b200.h).

There is no CPU or library fallback behind this module: if the shared library is missing, or the device is not a
Blackwell sm_100 part, every compute entry point raises.
"""
import contextlib
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PNO_LIB_PATH") or os.path.join(_HERE, "libpno_b200.so")      # PNO_LIB_PATH: diagnostics builds

PNO_OK, PNO_E_INVALID, PNO_E_MODES, PNO_E_CUDA, PNO_E_ARCH, PNO_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
ENGINE_FP32, ENGINE_TC_TF32X3, ENGINE_TC_TF32 = 0, 1, 2
ENGINE_ALLOW_FALLBACK = 0x100
# "tf32x3" / "tf32" are STRICT: a stage whose shape the tensor-core tiling does not cover raises instead of silently running on
# CUDA cores.  "auto" (the modules' default) = the fp32-parity tensor-core engine where it tiles the shape, the fp32 CUDA-core
# engine elsewhere (both reference-grade fp32); "<engine>+fallback" opts an explicit engine into the same fallback.
ENGINES = {"fp32": ENGINE_FP32, "tf32x3": ENGINE_TC_TF32X3, "tf32": ENGINE_TC_TF32,
           "auto": ENGINE_TC_TF32X3 | ENGINE_ALLOW_FALLBACK,
           "tf32x3+fallback": ENGINE_TC_TF32X3 | ENGINE_ALLOW_FALLBACK, "tf32+fallback": ENGINE_TC_TF32 | ENGINE_ALLOW_FALLBACK}
STAGES = ("dft_fwd", "dft_inv", "mix_fwd", "mix_bwd_dx", "mix_bwd_dw", "local_fwd", "local_bwd", "local_thin")

_c_float_p = ctypes.c_void_p      # device pointers travel as integers
_u64, _u32, _int, _vp = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p
_f32 = ctypes.c_float

# name -> (restype, argtypes); must list every symbol include/pno_b200.h declares (tests/test_abi.py checks this)
SIGNATURES = {
    "pno_version": (_int, []),
    "pno_last_error": (ctypes.c_char_p, []),
    "pno_check_device": (_int, []),
    "pno_launch_count": (_u64, []),
    "pno_set_sm_reserve": (_int, [_int]),
    "pno_stage_counts": (_int, [_vp]),
    "pno_engine_supports": (_int, [_int] * 9),
    "pno_set_dynamic_noise": (_int, [_vp]),
    "pno_noise_key_write": (_int, [_vp, _u64, _u32, _vp]),
    "pno_noise_key_advance": (_int, [_vp, _u32, _vp]),
    "pno_plan_create": (_int, [_int, _int, _int, _int, ctypes.POINTER(_vp)]),
    "pno_plan_destroy": (_int, [_vp]),
    "pno_spectral_workspace_bytes": (ctypes.c_size_t, [_vp, _int, _int, _int]),
    "pno_philox_normal_activation": (_int, [_u64, _u32, _u32, _u64, _u64, _vp, _vp]),
    "pno_philox_normal_spectral": (_int, [_u64, _u32, _u32, _int, _int, _int, _int, _vp, _vp, _vp]),
    "pno_philox_bits": (_int, [_u64, _u32, _u32, _u64, _u64, _vp, _vp]),
    "pno_dft_fwd": (_int, [_vp, _int, _vp, _int, _int, _int, _vp, _vp, _vp]),
    "pno_dft_inv": (_int, [_vp, _int, _vp, _vp, _int, _int, _int, _vp, _int, _vp, _vp, _vp]),
    "pno_mix_fwd": (_int, [_vp, _int, _vp, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _vp, _vp, _vp]),
    "pno_mix_bwd_dx": (_int, [_vp, _int, _vp, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _vp, _vp,
                              _vp]),
    "pno_mix_bwd_dw": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _int, _int, _int, _vp, _vp, _u64, _u32, _u32, _vp, _vp,
                              _vp, _vp, _vp]),
    "pno_mix_fwd_window": (_int, [_vp, _int, _vp, _vp, _int, _int, _int, _vp, _vp, _u64, _u32, _u32, _int, _int, _vp, _vp, _vp]),
    "pno_mix_bwd_dx_window": (_int, [_vp, _int, _vp, _vp, _int, _int, _int, _vp, _vp, _u64, _u32, _u32, _int, _int, _vp, _vp, _vp]),
    "pno_mix_bwd_dw_window": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _int, _int, _int, _vp, _u64, _u32, _u32, _int, _int, _vp, _vp,
                                     _vp]),
    "pno_spectral_fwd": (_int, [_vp, _int, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _vp, _int, _vp,
                                _vp, _vp, _vp]),
    "pno_spectral_bwd": (_int, [_vp, _int, _vp, _vp, _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _vp,
                                _vp, _vp, _vp, _vp, _vp, _vp]),
    "pno_local_fwd": (_int, [_int, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _u64, _u32, _u32, _u64, _vp, _vp,
                             _vp]),
    "pno_local_bwd_workspace_bytes": (ctypes.c_size_t, [_int, _int, _int, _int, _int]),
    "pno_local_bwd": (_int, [_int, _vp, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pno_act_bwd": (_int, [_vp, _vp, _u64, _int, _vp, _vp]),
    "pno_kl_fwd": (_int, [_vp, _vp, _u64, _vp, _vp]),
    "pno_kl_bwd": (_int, [_vp, _vp, _u64, _vp, _int, _vp, _vp, _vp]),
    "pno_mc_accumulate": (_int, [_vp, _u64, _vp, _vp, _vp]),
    "pno_mc_accumulate_decompose": (_int, [_vp, _vp, _u64, _u64, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    "pno_mc_finalize": (_int, [_vp, _vp, _u64, _u32, _vp, _vp, _vp]),
    "pno_uq_stats": (_int, [_vp, _vp, _vp, _u64, _int, _vp, _int, _vp, _vp, _int, _vp, _int, _vp, _vp]),
    "pno_loss_data_bwd": (_int, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp]),
    "pno_sumsq_accumulate": (_int, [_vp, _u64, _vp, _vp, _int, _vp, _vp]),
    "pno_adamw_step": (_int, [_vp, _vp, _vp, _vp, _u64, _f32, _f32, _f32, _f32, _f32, _u32, _vp, _f32, _int, _int, _vp, _vp]),
    "pno_sumsq_accumulate_multi": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pno_adamw_step_multi": (_int, [_int, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _f32, _u32, _vp, _f32, _int, _vp, _vp]),
    "pno_debug_set_buffer": (_int, [_vp]),
    "pno_selftest_umma": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pno_selftest_umma_mixed": (_int, [_int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()
_plans = {}
launch_count = 0          # number of C-ABI compute calls issued (bench.py reports kernel launches from this)


class PnoLibraryError(RuntimeError):
    pass


def load():
    """Load libpno_b200.so (once). Raises PnoLibraryError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PnoLibraryError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                    f"`make -C {os.path.join(_HERE, 'csrc')}`; pno_physics_bench_b200 has no CPU / eager fallback")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            if lib.pno_version() != 100:
                raise PnoLibraryError(f"{LIB_PATH}: unexpected ABI version {lib.pno_version()}")
            _lib = lib
    return _lib


def last_error():
    return load().pno_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    """Translate a status code into the exception type the reference raises for the same condition."""
    global launch_count
    launch_count += 1
    if rc == PNO_OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == PNO_E_MODES:
        raise RuntimeError(msg)          # reference: einsum shape mismatch -> RuntimeError (models.py:93-94)
    if rc == PNO_E_INVALID:
        raise ValueError(msg)
    raise PnoLibraryError(msg)


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def stage_counts():
    """{stage: (calls that ran on CUDA cores, calls that ran on the tensor cores)} since the process started."""
    buf = (ctypes.c_uint64 * (2 * len(STAGES)))()
    check(load().pno_stage_counts(ctypes.cast(buf, ctypes.c_void_p)), "pno_stage_counts")
    return {name: (int(buf[2 * i]), int(buf[2 * i + 1])) for i, name in enumerate(STAGES)}


def tensor_core_calls(before, after=None):
    """Per-stage (cuda_core, tensor_core) call deltas between two stage_counts() snapshots."""
    after = after if after is not None else stage_counts()
    return {k: (after[k][0] - before[k][0], after[k][1] - before[k][1]) for k in STAGES}


def engine_supports(stage, engine, h, w, m1, m2, b, ci, co):
    """Host-only: does `engine` run `stage` on its tensor-core kernel at this shape?  (works without a GPU)"""
    code = ENGINES[engine] if isinstance(engine, str) else int(engine)
    return bool(load().pno_engine_supports(STAGES.index(stage), code, h, w, m1, m2, b, ci, co))


def ptr(t):
    return 0 if t is None else t.data_ptr()


_NULL_CTX = contextlib.nullcontext()


def on_device(t):
    """Context that makes t's device current for the C-ABI call (kernels, TMA descriptors and torch's current stream all follow
    the current device); free when it already is."""
    if t.device.index is None or t.device.index == torch.cuda.current_device():
        return _NULL_CTX
    return torch.cuda.device(t.device)


def require_param(t, like, name, dtype=torch.float32):
    """A parameter handed to the kernels must be a CUDA tensor on the input's device with the dtype the kernels read; anything
    else would be reinterpreted memory (the reference raises RuntimeError for mixed devices / dtypes, too)."""
    if t is None:
        return
    if not t.is_cuda or t.device != like.device:
        raise RuntimeError(f"{name} is on {t.device} but the input is on {like.device}: move the module with .to(device) "
                           f"(pno_physics_bench_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} has dtype {t.dtype}; the kernels read {dtype} (model.half() / .double() are not supported)")


def require_cuda(t, name="tensor"):
    if not t.is_cuda:
        raise PnoLibraryError(f"{name} is on {t.device}: pno_physics_bench_b200 runs on a B200 (sm_100a) only and has no "
                              f"CPU fallback")


def plan(h, w, m1, m2, device):
    """Twiddle plan for (H, W, m1, m2) on `device` (cached for the life of the process)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), h, w, m1, m2)
    p = _plans.get(key)
    if p is None:
        lib = load()
        handle = _vp()
        with torch.cuda.device(key[0]):
            check(lib.pno_plan_create(h, w, m1, m2, ctypes.byref(handle)), "pno_plan_create")
        p = handle.value
        _plans[key] = p
    return p
