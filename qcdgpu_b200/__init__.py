"""qcdgpu_b200 -- B200-native (sm_100a CUDA) gauge-update hot path of QCDGPU.

    capi.GPU     the GPU_CL::GPU-shaped C ABI (libqcdgpu_b200.so, include/qcdgpu_b200.h)
    prng.PRNG    host side of the PRNG subsystem (PRNG_CL::PRNG)
    model.Model  the run loop (model_CL::model): init.dat -> sweeps -> measurements -> result file
    biglattice   x-slab decomposition over 1/2/4/8 GPUs with NVLink halo exchange

There is no CPU fallback anywhere in this package.
"""
from .capi import GPU, QcdGpuError, library_path, load  # noqa: F401
from .prng import PRNG  # noqa: F401
from .model import Model  # noqa: F401

__version__ = "0.1.0"
