"""CPU ORACLE -- test infrastructure, NOT product code.

numpy restatement of the device-side parallel-tempering swap round
(raytracerfortran_b200/csrc/rt_kernels.cu: philox4x32_10, swap_perm, swap_round_kernel), which in
turn applies TEMPSWP_MH's rule (prjmh_temper_rf.f90:1339-1344) to a counter-derived pairing:

    keys       = Philox4x32-10(counter = (round_lo, round_hi, 'PERC', 0), key = seed)
    perm       = 4-round Feistel network on 2*bits bits (4**bits >= n), cycle-walked into [0, n)
    pair t     = (perm(2t), perm(2t+1)),  u_t = 53 bits of Philox4x32-10((t, round, 'SWAP'), seed)
    accept iff u_t <= exp((beta_j - beta_i) * (logL_i - logL_j));  accepted pairs exchange betas

Philox4x32-10 is pinned by the Random123 known-answer vectors (tests/test_tempering_cpu.py).
"""
import numpy as np

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 (scalars or arrays), key: 2 uint32.  Returns 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & _MASK for x in ctr]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = int(key[0]) & _MASK, int(key[1]) & _MASK
    for _ in range(10):
        p0 = np.uint64(_M0) * c[0]
        p1 = np.uint64(_M1) * c[2]
        n0 = (p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1)
        c = [n0 & _MASK, p1 & _MASK, n2 & _MASK, p0 & _MASK]
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return [x.astype(np.uint32) for x in c]


def u01(a, b):
    """53 random bits in [0, 1) from two uint32 words (the construction numpy uses for doubles)."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0
            + (b >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def _mix(v):
    v = (v * np.uint64(0x9E3779B1)) & _MASK
    v ^= v >> np.uint64(15)
    v = (v * np.uint64(0x85EBCA77)) & _MASK
    v ^= v >> np.uint64(13)
    return v


def perm_params(n, seed, round_index):
    bits = 1
    while (1 << (2 * bits)) < n:
        bits += 1
    keys = philox4x32_10((round_index & _MASK, (round_index >> 32) & _MASK, 0x50455243, 0),
                         (seed & _MASK, (seed >> 32) & _MASK))
    return bits, [int(np.asarray(k).reshape(-1)[0]) for k in keys]


def swap_perm(n, seed, round_index, idx=None):
    """The round's bijection of [0, n) evaluated at idx (default: everywhere)."""
    bits, keys = perm_params(n, seed, round_index)
    mask = np.uint64((1 << bits) - 1)
    x = np.arange(n, dtype=np.uint64) if idx is None else np.asarray(idx, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        L, R = x[todo] >> np.uint64(bits), x[todo] & mask
        for r in range(4):
            F = _mix(R ^ np.uint64(keys[r])) & mask
            L, R = R, L ^ F
        x[todo] = (L << np.uint64(bits)) | R
        todo = x >= n
    return x.astype(np.int64)


def swap_round(logL, beta, seed, round_index):
    """One swap round over all n chains.  Returns (new beta [n], pairs [n//2, 2], accept [n//2] bool,
    u [n//2])."""
    logL = np.asarray(logL, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    n = logL.size
    npair = n // 2
    p = swap_perm(n, seed, round_index)
    i, j = p[0:2 * npair:2], p[1:2 * npair:2]
    t = np.arange(npair, dtype=np.uint64)
    r = philox4x32_10((t, round_index & _MASK, (round_index >> 32) & _MASK, 0x53574150),
                      (seed & _MASK, (seed >> 32) & _MASK))
    u = u01(r[0], r[1])
    logratio = (beta[j] - beta[i]) * (logL[i] - logL[j])            # prjmh_temper_rf.f90:1339-1340
    with np.errstate(over="ignore", invalid="ignore"):
        accept = u <= np.exp(logratio)                               # :1342
    new = beta.copy()
    new[i] = np.where(accept, beta[j], beta[i])
    new[j] = np.where(accept, beta[i], beta[j])
    return new, np.stack([i, j], axis=1), accept, u
