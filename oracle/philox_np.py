"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the counter-based RNG the CUDA kernels use.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.

Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; the Random123
reference constants).  The reference (PEUQSE) has no counter-based RNG: it draws from numpy's legacy
global MT19937 stream (`np.random.multivariate_normal` at InverseProblem.py:2630, `np.random.uniform`
at InverseProblem.py:2650 and :706).  The north star replaces that stream by Philox keyed on
(seed, stream) with a counter of (index, sub-draw), so that every chain/point is index-addressed.
Parity tests feed the *same* draws to the reference by monkeypatching those two numpy calls.

Layout shared with peuqse_b200/csrc/philox.cuh (must stay in sync):
    key     = (seed_lo, seed_hi)
    counter = (idx_lo, idx_hi, block, stream)
where `idx` is the 64-bit chain / point index, `block` numbers successive 128-bit outputs for that
index (for MH: block = 4*step + j; for sweeps: block = j) and `stream` separates uses
(STREAM_MH = 1, STREAM_UNIFORM_SWEEP = 2, STREAM_ESS = 3).
One 128-bit output gives two 53-bit uniforms: u_a from words (0,1), u_b from words (2,3).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_MH = 1
STREAM_UNIFORM_SWEEP = 2
STREAM_ESS = 3

TWO_POW_M53 = 2.0 ** -53


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32 array-like, key: (..., 2) uint32 array-like -> (..., 4) uint32."""
    c = np.asarray(counter, dtype=np.uint64)
    k = np.asarray(key, dtype=np.uint64)
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0, k1 = (np.broadcast_to(k[..., i], c0.shape).copy() for i in range(2))
    for _ in range(10):
        p0 = M0 * c0  # 32x32 -> 64 bit products, no overflow in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + np.uint64(W0)) & MASK32
        k1 = (k1 + np.uint64(W1)) & MASK32
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _u53(hi_word, lo_word):
    """Two uint32 words -> double in [0,1) with 53 random bits."""
    hi = hi_word.astype(np.uint64) >> np.uint64(5)   # 27 bits
    lo = lo_word.astype(np.uint64) >> np.uint64(6)   # 26 bits
    return ((hi << np.uint64(26)) | lo).astype(np.float64) * TWO_POW_M53


def raw_block(seed, stream, idx, block):
    """128-bit Philox output for (seed, stream, idx, block); idx/block broadcast."""
    idx = np.asarray(idx, dtype=np.uint64)
    block = np.asarray(block, dtype=np.uint64)
    idx, block = np.broadcast_arrays(idx, block)
    ctr = np.stack([idx & MASK32, idx >> np.uint64(32), block & MASK32,
                    np.full(idx.shape, stream, dtype=np.uint64)], axis=-1)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint64)
    return philox4x32_10(ctr, key)


def uniform_pair(seed, stream, idx, block):
    """-> (u_a, u_b), both in [0,1)."""
    r = raw_block(seed, stream, idx, block)
    return _u53(r[..., 0], r[..., 1]), _u53(r[..., 2], r[..., 3])


def normal_pair(seed, stream, idx, block):
    """Box-Muller on one 128-bit block: u1 = 1 - u_a in (0,1], angle = 2*pi*u_b."""
    ua, ub = uniform_pair(seed, stream, idx, block)
    r = np.sqrt(-2.0 * np.log(1.0 - ua))
    ang = (2.0 * np.pi) * ub
    return r * np.cos(ang), r * np.sin(ang)


def sweep_uniforms(seed, first, count, k):
    """u[count,k] in [0,1) for the uniform multistart sweep: point g, parameter d uses block d//2,
    half d%2 (InverseProblem.py:706 draws np.random.uniform(-2,2,[N,k]); here index-addressed)."""
    g = (np.uint64(first) + np.arange(count, dtype=np.uint64))[:, None]
    nblk = (k + 1) // 2
    ua, ub = uniform_pair(seed, STREAM_UNIFORM_SWEEP, g, np.arange(nblk, dtype=np.uint64)[None, :])
    u = np.empty((count, 2 * nblk))
    u[:, 0::2] = ua
    u[:, 1::2] = ub
    return u[:, :k]


def mh_draws(seed, chain, step, k):
    """Standard normals z[k] and the accept uniform u for MH (chain, step).
    Blocks 8*step + j, j = 0..3 give normals (2 per block, so k <= 8 per 4 blocks; more
    parameters continue into blocks j = 4..6); block 8*step + 7 word-pair a gives u."""
    chain = np.asarray(chain, dtype=np.uint64)
    step = np.asarray(step, dtype=np.uint64)
    chain, step = np.broadcast_arrays(chain, step)
    nblk = (k + 1) // 2
    assert nblk <= 7
    z = np.empty(chain.shape + (2 * nblk,))
    for j in range(nblk):
        za, zb = normal_pair(seed, STREAM_MH, chain, np.uint64(8) * step + np.uint64(j))
        z[..., 2 * j] = za
        z[..., 2 * j + 1] = zb
    ua, _ = uniform_pair(seed, STREAM_MH, chain, np.uint64(8) * step + np.uint64(7))
    return z[..., :k], ua
