"""primestereomatch_b200 -- B200-native (sm_100a) STEREO_GIF hot path: CVC -> CVF -> WTA.

The product is the CUDA library `libprime_stereo_b200.so` (sources in csrc/, C-ABI in
include/prime_stereo_b200.h).  `DispEst` mirrors the reference facade over that C-ABI.
"""
from .dispest import DispEst, device_count, MAX_CPU_THREADS, OCV_DE, OCL_DE  # noqa: F401
from . import capi, synth  # noqa: F401

__all__ = ["DispEst", "device_count", "capi", "synth"]
