"""Oracle (CPU, numpy) for the opt-in in-kernel Gaussian noise of the disturbance kernel.  TEST INFRASTRUCTURE ONLY.

The reference draws its noise with `torch.randn_like(obs)` (shared/disturbances_gpu.py:97-99 -> [tv]
gaussian_noise_image); the default path of this repository reads that tensor.  `apply_disturbances(noise_seed=...)`
(additive) generates N(0,1) draws inside the kernel instead (csrc/disturb.cu: philox4x32_10 / box_muller /
philox_normal4).  This file restates that generator so the in-kernel noise can be checked value by value:

  * Philox4x32-10 as published (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11;
    the Random123 library) - pinned below against Random123's known-answer vectors (tests/test_oracle.py);
  * counter = (quad_lo, quad_hi, offset_lo, offset_hi), key = (seed_lo, seed_hi), quad = index of the 4-element group
    in the logical contiguous [first_image + B, C, H, W] tensor (W % 4 == 0);
  * outputs (a, b, c, d) -> box_muller(a, b), box_muller(c, d):  u = fl32(fl32(a) * 2^-32 + 2^-33),
    r = sqrt(-2 ln u), theta = fl32(int32(b)) * fl32(pi * 2^-31), draws (r cos theta, r sin theta).
The kernel evaluates log2 / sin / cos with the MUFU approximations (abs. error < 3e-6 on a draw); the parity test allows
1e-5 on the disturbed output.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)

# Random123 kat_vectors, philox4x32 with 10 rounds: (counter, key) -> output
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays (scalars broadcast)."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint32) for v in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def box_muller(a, b):
    a32 = a.astype(np.float32).astype(np.float64)
    u = (a32 * 2.0 ** -32 + 2.0 ** -33).astype(np.float32).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(u))
    th = (b.astype(np.int32).astype(np.float32).astype(np.float64) * np.float64(np.float32(np.pi * 2.0 ** -31)))
    th = th.astype(np.float32).astype(np.float64)
    return (r * np.cos(th)).astype(np.float32), (r * np.sin(th)).astype(np.float32)


def normal_noise(seed: int, offset: int, shape, first_image: int = 0) -> np.ndarray:
    """fp32 [B,C,H,W]: the draws the kernel uses for images first_image .. first_image + B - 1."""
    B, C, H, W = shape
    assert W % 4 == 0
    per_image = C * H * W // 4
    quad = np.arange(first_image * per_image, (first_image + B) * per_image, dtype=np.uint64)
    lo, hi = (quad & _MASK).astype(np.uint32), (quad >> np.uint64(32)).astype(np.uint32)
    o = philox4x32_10(lo, hi, np.uint32(offset & 0xFFFFFFFF), np.uint32((offset >> 32) & 0xFFFFFFFF),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    n0, n1 = box_muller(o[0], o[1])
    n2, n3 = box_muller(o[2], o[3])
    return np.stack([n0, n1, n2, n3], axis=1).reshape(B, C, H, W)
