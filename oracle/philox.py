"""ORACLE (test infrastructure): numpy restatement of the integer half of torch's CUDA RNG contract that the
reference's `torch.randint(0, 1000, (B,), device=cuda)` (`diffusion/models/stable_diffusion.py:177`) relies on.

Sources restated (headers on disk):
  * Philox4x32-10: /usr/local/cuda/include/curand_philox4x32_x.h:88-91,107-137,160-185
  * curand_init(seed, subsequence, offset): curand_kernel.h:1022-1037 (subsequence -> ctr.zw, offset/4 -> ctr.xy)
  * ATen launch policy + `rand32 % range + base`: ATen/native/cuda/DistributionTemplates.h:50-87,282-346 and
    ATen/core/TransformationHelper.h:42-44
Pinned against the Random123 known-answer vectors for philox4x32-10 (tests/test_oracle.py).  The float half
(Box-Muller with device __sincosf/__logf) cannot be restated bit-exactly on a CPU; it is checked on the GPU
against torch.randn_like itself.
"""
import numpy as np

PHILOX_W32_0, PHILOX_W32_1 = 0x9E3779B9, 0xBB67AE85
PHILOX_M4x32_0, PHILOX_M4x32_1 = 0xD2511F53, 0xCD9E8D57
M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32 array-like, key: (..., 2). Returns (..., 4) uint32."""
    c = np.array(ctr, dtype=np.uint64) & M32
    k = np.array(key, dtype=np.uint64) & M32
    c0, c1, c2, c3 = [c[..., i].copy() for i in range(4)]
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    for r in range(10):
        p0 = np.uint64(PHILOX_M4x32_0) * c0
        p1 = np.uint64(PHILOX_M4x32_1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M32, lo1, (hi0 ^ c3 ^ k1) & M32, lo0
        k0 = (k0 + np.uint64(PHILOX_W32_0)) & M32
        k1 = (k1 + np.uint64(PHILOX_W32_1)) & M32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def aten_grid(numel, num_sms=148, max_threads_per_sm=2048, block=256, unroll=4):
    """ATen `calc_execution_policy` (DistributionTemplates.h:50-87): returns (grid, counter_offset)."""
    grid = (numel + block - 1) // block
    blocks_per_sm = max_threads_per_sm // block
    grid = min(num_sms * blocks_per_sm, grid)
    counter_offset = ((numel - 1) // (block * grid * unroll) + 1) * 4
    return grid, counter_offset


def curand4(seed, subsequence, offset):
    """One `curand4()` after `curand_init(seed, subsequence, offset)`; offset must be a multiple of 4."""
    subsequence = np.asarray(subsequence, dtype=np.uint64)
    assert offset % 4 == 0
    lo = np.uint64(offset // 4)
    ctr = np.stack([
        np.broadcast_to(lo & np.uint64(M32), subsequence.shape),
        np.broadcast_to(lo >> np.uint64(32), subsequence.shape), subsequence & np.uint64(M32),
        subsequence >> np.uint64(32)
    ], axis=-1)
    key = np.broadcast_to(np.array([seed & M32, (seed >> 32) & M32], dtype=np.uint64), subsequence.shape + (2,))
    return philox4x32_10(ctr, key)


def randint_cuda(seed, offset, numel, high, low=0, num_sms=148):
    """torch.randint(low, high, (numel,), device='cuda') for ranges < 2**32 -> (int64 values, new offset)."""
    grid, inc = aten_grid(numel, num_sms)
    nthreads = grid * 256
    out = np.zeros(numel, dtype=np.int64)
    # distribution_elementwise_grid_stride_kernel: each loop round draws one curand4 per thread, element
    # idx + k*nthreads takes component k of that round's draw
    rounds = (numel + nthreads * 4 - 1) // (nthreads * 4)
    tid = np.arange(nthreads, dtype=np.uint64)
    for r in range(rounds):
        draws = curand4(seed, tid, offset + 4 * r)
        for k in range(4):
            li = r * nthreads * 4 + k * nthreads + np.arange(nthreads)
            ok = li < numel
            out[li[ok]] = (draws[ok, k].astype(np.int64) % (high - low)) + low
    return out, offset + inc
