"""NumPy oracle of the Philox4x32-10 generator used by the init / sampling kernels."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32).copy() for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0); k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def init_rows_ref(ids, dim, seed, lo, hi):
    """value(id, j) exactly as csrc/fps_core.cu::fps_init_rows_kernel computes it (fp32)."""
    ids = np.asarray(ids, dtype=np.int64)
    stride = (dim + 3) // 4 * 4
    out = np.zeros((len(ids), stride), dtype=np.float32)
    id_lo = (ids & 0xFFFFFFFF).astype(np.uint32)
    id_hi = ((ids >> 32) & 0xFFFFFFFF).astype(np.uint32)
    for q in range(stride // 4):
        r = philox4x32(id_lo, id_hi, np.uint32(q), np.uint32(0), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for j in range(4):
            col = 4 * q + j
            if col < dim:
                u = (r[j] >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
                out[:, col] = np.float32(lo) + np.float32(hi - lo) * u
    return out
