"""numpy restatement of the noise stream of csrc/mmm_md.cu: Philox4x32-10 keyed by the seed,
counter = (bead, step_lo, step_hi, stream), Box-Muller in FP64 on 32-bit uniforms."""
import numpy as np


def philox4x32_10(c, k):
    """c: (n,4) uint32 counters, k: (2,) uint32 key -> (n,4) uint32."""
    c = c.astype(np.uint64).copy()
    k0, k1 = np.uint64(k[0]), np.uint64(k[1])
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c[:, 0]
        p1 = np.uint64(0xCD9E8D57) * c[:, 2]
        n0 = (p1 >> np.uint64(32)) ^ c[:, 1] ^ k0
        n1 = p1 & m32
        n2 = (p0 >> np.uint64(32)) ^ c[:, 3] ^ k1
        n3 = p0 & m32
        c = np.stack([n0, n1, n2, n3], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c.astype(np.uint32)


def normals3(seed, n, step, stream):
    c = np.zeros((n, 4), dtype=np.uint32)
    c[:, 0] = np.arange(n, dtype=np.uint32)
    c[:, 1] = step & 0xFFFFFFFF
    c[:, 2] = step >> 32
    c[:, 3] = stream
    u = (philox4x32_10(c, (seed & 0xFFFFFFFF, seed >> 32)).astype(np.float64) + 0.5) / 4294967296.0
    r0, r1 = np.sqrt(-2.0 * np.log(u[:, 0])), np.sqrt(-2.0 * np.log(u[:, 2]))
    return np.stack([r0 * np.cos(2 * np.pi * u[:, 1]), r0 * np.sin(2 * np.pi * u[:, 1]), r1 * np.cos(2 * np.pi * u[:, 3])], axis=1)
