"""Philox4x32-10 (Salmon et al., SC'11; Random123) in numpy, bit-exact with the CUDA
kernel's generator, plus the repo's draw layout.

Draw layout (shared by the kernel, see csrc/petmh_kernels.cuh `philox_draw`):
  key     = (seed_lo, seed_hi)
  counter = (coord i, 2*sweep + block, chain_gid_lo, chain_gid_hi)
  chain_gid = tac_gid * n_chains + chain      (independent of the GPU count)
  output words x0..x3:  normal = sqrt(-2 ln u(x0)) cos(2 pi u(x1)),  u(x) = x*2^-32 + 2^-33
                        log-uniform = ln u(x2);  visit key = x3 (rank by (key, coord))
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(counter[..., i], np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], np.uint64)
    k1 = np.asarray(key[..., 1], np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def u01(x):
    """uint32 -> float32 in (0, 1]: x*2^-32 + 2^-33 with fp32 rounding (as the kernel)."""
    return (x.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float32)


def raw_draws(seed, chain_gid, sweep, block, n_coord=48):
    """The 4 raw words for every coordinate of one (chain, sweep, block)."""
    ctr = np.zeros((n_coord, 4), np.uint32)
    ctr[:, 0] = np.arange(n_coord)
    ctr[:, 1] = np.uint32((2 * sweep + block) & 0xFFFFFFFF)
    ctr[:, 2] = np.uint32(chain_gid & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32((chain_gid >> 32) & 0xFFFFFFFF)
    key = np.zeros((n_coord, 2), np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    return philox4x32_10(ctr, key)


def ranks_from_keys(keys):
    """Visit position of each coordinate: rank by (key, coord index) ascending."""
    order = np.lexsort((np.arange(keys.size), keys))
    rank = np.empty(keys.size, np.int64)
    rank[order] = np.arange(keys.size)
    return rank


def draws(seed, chain_gid, sweep, block, n_coord=48):
    """(normal f32 approx, logu f32 approx, rank int) -- the float transforms are
    evaluated in fp64 here; the kernel uses fp32 fast intrinsics, so these agree to
    ~1e-6 only.  Bit-exact parity uses the explicit tape (oracle/mh.py)."""
    x = raw_draws(seed, chain_gid, sweep, block, n_coord)
    u1 = u01(x[:, 0]).astype(np.float64)
    u2 = u01(x[:, 1]).astype(np.float64)
    normal = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    logu = np.log(u01(x[:, 2]).astype(np.float64))
    return normal.astype(np.float32), logu.astype(np.float32), ranks_from_keys(x[:, 3])
