"""Host restatement (numpy, vectorised) of the Philox4x32-10 lambda stream in ``csrc/philox.cuh``.

Used by the tests and by callers that want to know which lambda env ``i`` will see at its ``k``-th reset
without touching the device: ``lambda_stream(seed, env_index, draw_index, ...)``.
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u53(a, b):
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    return (((a >> np.uint64(5)) << np.uint64(26)) | (b >> np.uint64(6))).astype(np.float64) * (1.0 / 9007199254740992.0)


def lambda_stream(seed, env_index, draw_index, re_interval, im_interval, re_lo_override=None):
    """lambda (complex128 array) of global env ``env_index`` at its ``draw_index``-th draw."""
    env_index = np.asarray(env_index, dtype=np.uint64)
    draw_index = np.asarray(draw_index, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    x0, x1, x2, x3 = philox4x32_10(
        (env_index & _MASK).astype(np.uint32), (env_index >> np.uint64(32)).astype(np.uint32),
        draw_index.astype(np.uint32), np.uint32(0), seed & 0xFFFFFFFF, seed >> 32)
    lo = np.float64(re_interval[0]) if re_lo_override is None else np.asarray(re_lo_override, dtype=np.float64)
    re = lo + (np.float64(re_interval[1]) - lo) * u53(x0, x1)
    im = np.float64(im_interval[0]) + (np.float64(im_interval[1]) - np.float64(im_interval[0])) * u53(x2, x3)
    return re + 1j * im
