"""lustrine_b200 — B200-native (sm_100a) particle simulation step with the Lustrine API.

Only what the hot path needs lives here:
  csrc/   hand-written CUDA kernels + the C ABI of include/lgpu.h  -> lib/liblgpu.so
  host/   C++ mirror of the reference's Simulation / Lustrine / LustrineWrapper interface
  lgpu.py ctypes binding of the C ABI (tests, bench, multi-GPU driver)
"""
from . import lgpu  # noqa: F401

__all__ = ["lgpu"]
