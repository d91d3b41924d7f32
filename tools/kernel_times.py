#!/usr/bin/env python
"""CUDA-event time of every kernel of one PNO block at the cfg-3 shape (bench.block_kernel_times), per engine.

    python tools/kernel_times.py [engine ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import block_kernel_times  # noqa: E402

if __name__ == "__main__":
    for engine in (sys.argv[1:] or ["tf32", "tf32x3"]):
        r = block_kernel_times(engine, torch.device("cuda", 0))
        tot = sum(v["us"] for v in r.values())
        print(engine, json.dumps(r), f"block total {tot:.0f} us")
