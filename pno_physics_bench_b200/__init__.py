"""pno_physics_bench_b200 -- B200 (sm_100a) implementation of pno-physics-bench's probabilistic spectral-convolution path.

Drop-in classes (same names / signatures / state_dict as the reference's `pno_physics_bench.models`):
SpectralConv2d, SpectralConv2d_Probabilistic, PNOBlock, FourierNeuralOperator, ProbabilisticNeuralOperator; and the
training mirror in `pno_physics_bench_b200.training` (PNOTrainer, PNOLoss, ELBOLoss).
All tensor work goes through libpno_b200.so (C ABI: include/pno_b200.h); there is no CPU or eager fallback.
"""
from .models import (FourierNeuralOperator, PNOBlock, ProbabilisticNeuralOperator, SpectralConv2d,  # noqa: F401
                     SpectralConv2d_Probabilistic, noise_scope)

__version__ = "0.1.0"
