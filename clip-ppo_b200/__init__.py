"""clip_ppo_b200 - B200-native (sm_100a) implementation of CLIP-PPO's observation path.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of
include/clipppo_b200.h), the ctypes binding, and the host-side mirrors of the reference's
operator interfaces.  The drop-in Python surface is the top-level `shared/` package.
"""
from . import _native  # noqa: F401

__all__ = ["_native"]
