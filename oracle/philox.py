"""numpy restatement of torch's CUDA Philox4x32-10 streams -- TEST INFRASTRUCTURE ONLY.

The reference's sampler draws from torch's *CUDA* generator (R/snn_model/vq_diffusion.py:105 hard-codes 'cuda'):
``torch.rand_like`` (:118) and ``Tensor.exponential_`` inside ``Categorical.sample`` -> ``multinomial`` (:136-138,
TORCH/distributions/categorical.py:147-148).  Both go through
``distribution_elementwise_grid_stride_kernel`` (TORCH/include/ATen/native/cuda/DistributionTemplates.h:64-87):

    thread idx:  curand_init(seed, subsequence=idx, offset) ; per round one curand_uniform4 ->
    element li = idx + tpg*(4*round + ii) gets component ii,   tpg = 256 * grid,
    grid = min(sm_count * (max_threads_per_sm // 256), ceil(numel / 256))               (:50-62)
    generator offset advances by ((numel-1) // (256*grid*4) + 1) * 4 per call            (:60)

Philox4x32-10 itself is the published algorithm (Salmon et al., SC'11) as implemented by cuRAND
(``curand_philox4x32_x.h``: multipliers 0xD2511F53 / 0xCD9E8D57, Weyl constants 0x9E3779B9 / 0xBB67AE85, counter =
(offset/4 lo, offset/4 hi, subsequence lo, subsequence hi), key = seed).

Pin status: there is no GPU in the build container, so this file is pinned on the GPU box by
``tests/test_gpu_sampling.py`` against ``torch.rand`` / ``Tensor.exponential_`` on CUDA with ``torch.manual_seed``;
the Philox core is additionally pinned here against the Random123 known-answer vectors (tests/test_oracle_golden.py).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs are uint32 arrays (counters) and python ints (key)."""
    c0 = c0.astype(np.uint64); c1 = c1.astype(np.uint64); c2 = c2.astype(np.uint64); c3 = c3.astype(np.uint64)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def execution_policy(numel: int, sm_count: int, max_threads_per_sm: int):
    """(threads per grid, generator offset increment) of the torch launch for ``numel`` elements."""
    grid = min(sm_count * (max_threads_per_sm // 256), (numel + 255) // 256)
    grid = max(grid, 1)
    inc = ((numel - 1) // (256 * grid * 4) + 1) * 4
    return grid * 256, inc


def raw_u32(seed: int, offset: int, numel_global: int, sm_count: int, max_threads_per_sm: int,
            index_base: int = 0, numel: int | None = None) -> np.ndarray:
    """The 32-bit draw of elements [index_base, index_base + numel) of a call over ``numel_global`` elements."""
    assert offset % 4 == 0
    numel = numel_global - index_base if numel is None else numel
    tpg, _ = execution_policy(numel_global, sm_count, max_threads_per_sm)
    li = np.arange(index_base, index_base + numel, dtype=np.uint64)
    idx = li % np.uint64(tpg)
    q = li // np.uint64(tpg)
    rnd = q >> np.uint64(2)
    ii = (q & np.uint64(3)).astype(np.int64)
    ctr = np.uint64(offset // 4) + rnd
    r = philox4x32_10((ctr & MASK).astype(np.uint32), (ctr >> np.uint64(32)).astype(np.uint32),
                      (idx & MASK).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    stacked = np.stack(r, axis=0)
    return stacked[ii, np.arange(numel)]


def _curand_uniform(raw: np.ndarray) -> np.ndarray:
    """cuRAND's _curand_uniform: x * 2^-32 + 2^-33 in fp32, range (0, 1]."""
    two32 = np.float32(2.3283064365386963e-10)
    return (raw.astype(np.float32) * two32 + two32 / np.float32(2.0)).astype(np.float32)


def uniform(seed, offset, numel_global, sm_count, max_threads_per_sm, index_base=0, numel=None) -> np.ndarray:
    """torch.rand on CUDA: the (0,1] draw with 1.0 flipped to 0.0 (DistributionTemplates.h:493-503)."""
    v = _curand_uniform(raw_u32(seed, offset, numel_global, sm_count, max_threads_per_sm, index_base, numel))
    v[v == np.float32(1.0)] = np.float32(0.0)
    return v


def exponential(seed, offset, numel_global, sm_count, max_threads_per_sm, index_base=0, numel=None) -> np.ndarray:
    """Tensor.exponential_(1) on CUDA: -log(u), with log := -eps/2 for u >= 1 - eps/2
    (TORCH/include/ATen/core/TransformationHelper.h:129-146).  On the device torch evaluates the log with the fast
    __logf intrinsic (TORCH/include/ATen/NumericUtils.h:150-157), which numpy's log reproduces only to ~1e-6; the GPU
    test states that tolerance, and the CUDA kernel (same intrinsic) is bit-identical to torch."""
    v = _curand_uniform(raw_u32(seed, offset, numel_global, sm_count, max_threads_per_sm, index_base, numel))
    eps = np.float32(1.1920928955078125e-07)
    lg = np.where(v >= np.float32(1.0) - eps / np.float32(2.0), -eps / np.float32(2.0), np.log(v).astype(np.float32))
    return (np.float32(-1.0) * lg).astype(np.float32)
