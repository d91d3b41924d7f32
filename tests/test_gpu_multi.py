"""Multi-GPU checks that need >= 2 B200s in the box (skipped on a single-GPU box): data-parallel invariance of one PNOTrainer step
 and sample-sharding invariance of predict_with_uncertainty (P9) over NCCL, each as a 2-rank torchrun job."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _torchrun(script_args, nproc=2, timeout=600):
    port = 29500 + os.getpid() % 2000
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + script_args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_step_equals_the_whole_batch_step():
    r = _torchrun([os.path.join(ROOT, "tools", "dp_train_check.py")])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "DP_CHECK_OK" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sample_sharding_over_nccl_is_invariant():
    r = _torchrun([os.path.join(ROOT, "tools", "mc_shard_check.py")])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MC_SHARD_OK" in r.stdout, r.stdout[-2000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_mode_parallel_step_equals_the_unsharded_step():
    r = _torchrun([os.path.join(ROOT, "tools", "mp_train_check.py")])
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MP_CHECK_OK" in r.stdout, r.stdout[-2000:]


def test_mode_parallel_on_one_rank_equals_the_model():
    """The mode-window kernels on one GPU (two windows = the two weight blocks): same step as the unsharded model."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mp_train_check.py")], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MP_CHECK_OK" in r.stdout, r.stdout[-2000:]


def test_mode_parallel_batch_chunks_on_one_rank():
    """A global batch larger than one channel-mix launch takes (8 ranks x 32 fields = 256 rows) is cut into batch chunks, and the
    weight-gradient kernel walks the batch in K chunks: same step as the unsharded model (batch 40, 16 rows per launch)."""
    env = dict(os.environ, PNO_MP_BATCH="40", PNO_MP_MAX_BATCH="16")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mp_train_check.py")], cwd=ROOT, capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MP_CHECK_OK" in r.stdout, r.stdout[-2000:]
