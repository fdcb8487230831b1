"""numpy restatement of the in-kernel noise generator (oracle; test infrastructure only).

Philox4x32-10 (Salmon et al. 2011) keyed by the 64-bit seed, counter = (vec index inside the
sample, sample lo, sample hi, timestep key); the four 32-bit outputs become four N(0,1) values by
two Box-Muller pairs on 24-bit uniforms - exactly csrc/update.cu::philox_normal4.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def _u01(r):
    return ((r >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def _box_muller(a, b):
    u1, u2 = _u01(a), _u01(b)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = (np.float32(2.0) * u2).astype(np.float64) * np.pi
    return (rad * np.cos(ang).astype(np.float32)).astype(np.float32), (rad * np.sin(ang).astype(np.float32)).astype(np.float32)


def philox_normal(batch, per_sample, seed, sample_offset, t):
    """[batch, per_sample] float32 N(0,1): element r of sample b uses counter (r//4, b+offset, t), lane r%4."""
    nvec = (per_sample + 3) // 4
    out = np.empty((batch, nvec * 4), dtype=np.float32)
    vec = np.arange(nvec, dtype=np.uint32)
    for b in range(batch):
        s = int(sample_offset) + b
        r0, r1, r2, r3 = philox4x32_10(vec, np.full(nvec, s & 0xFFFFFFFF, np.uint32), np.full(nvec, s >> 32, np.uint32),
                                       np.full(nvec, t & 0xFFFFFFFF, np.uint32), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        n0, n1 = _box_muller(r0, r1)
        n2, n3 = _box_muller(r2, r3)
        out[b] = np.stack([n0, n1, n2, n3], axis=1).reshape(-1)
    return out[:, :per_sample]
