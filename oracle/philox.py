"""Oracle: Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the uniform / Box-Muller maps the
CUDA path applies to it (rllab_b200/csrc/common.cuh: Philox, noise4).  TEST INFRASTRUCTURE ONLY.

The reference draws from NumPy's global MT19937 stream in program order (gaussian_mlp_policy.py:128, env reset());
that order cannot be reproduced across 65k lanes, so the B200 path defines its own counter-based stream and the
parity tests inject identical noise tensors on both sides ("RNG parity").
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over counter arrays (uint64 holding 32-bit values).  Returns 4 uint32 arrays."""
    c = [np.asarray(x, np.uint64) & MASK for x in (c0, c1, c2, c3)]
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(M0) * c[0]
        p1 = np.uint64(M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [(hi1 ^ c[1] ^ k0) & MASK, lo1, (hi0 ^ c[3] ^ k1) & MASK, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return [x.astype(np.uint32) for x in c]


def raw_block(rows, row0, K, N, lane0, seed, it, stream_id):
    """uint32 words laid out like b200rl_fill_noise: [rows][K][N]."""
    out = np.zeros((rows, K, N), np.uint32)
    lanes = (np.arange(N, dtype=np.uint64) + np.uint64(lane0))
    for r in range(rows):
        for c in range((K + 3) // 4):
            w = philox4x32_10(lanes & MASK, np.full(N, (row0 + r) | (stream_id << 28), np.uint64),
                              np.full(N, c, np.uint64), lanes >> np.uint64(32), seed, it)
            for j in range(4):
                if c * 4 + j < K:
                    out[r, c * 4 + j] = w[j]
    return out


def uniform_from_raw(raw):
    return (raw >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def normal_from_raw(raw):
    """Box-Muller on consecutive pairs (k even -> cos branch, k odd -> sin branch), float64 math."""
    rows, K, N = raw.shape
    out = np.zeros((rows, K, N))
    assert K % 2 == 0, "pass raw words with K padded to an even count (pairs share one Box-Muller draw)"
    u = ((raw >> np.uint32(8)).astype(np.float64) + 0.5) / 16777216.0
    for k in range(0, K - 1, 2):
        rad = np.sqrt(-2.0 * np.log(u[:, k]))
        out[:, k] = rad * np.cos(2 * np.pi * u[:, k + 1])
        out[:, k + 1] = rad * np.sin(2 * np.pi * u[:, k + 1])
    return out
