#!/usr/bin/env python
"""Time the tensor-core forward transform (cfg-3 shape) for the PNO_FWD_* overrides in the environment."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import block_kernel_times  # noqa: E402

r = block_kernel_times(sys.argv[1] if len(sys.argv) > 1 else "tf32", torch.device("cuda", 0), iters=3)
print({k: os.environ.get(k) for k in ("PNO_FWD_STAGES", "PNO_FWD_KPS")}, "dft_fwd", r["dft_fwd"]["us"], "us")
