"""eagleeverything_b200 -- B200-native (sm_100a) implementation of the genome-scan hot path of
Eagle / WMAM: no-space ASCII genotype decode, int8 tensor-core M.Mt, FP64 tensor-core a / var(a)
scan.  `api` mirrors the reference's Rcpp exports over the C ABI of libeaglegpu.so
(include/eagle_gpu.h); `device` and `dist` hold the device-resident and multi-GPU plumbing.
There is no CPU fallback anywhere in this package."""
from . import _lib  # noqa: F401

__all__ = ["api", "device", "dist", "synth", "_lib"]
__version__ = "0.1.0"
