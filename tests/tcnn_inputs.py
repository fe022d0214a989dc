"""TEST INFRASTRUCTURE: the seeded inputs oracle/tcnn_golden.cu fed to the reference's tiny-cuda-nn
(make_inputs / make_targets there), restated in numpy so the golden fixtures need not carry them."""
import numpy as np


def _lcg_floats(seed, n):
    out = np.empty(n, np.uint32)
    s = seed & 0xFFFFFFFF
    for i in range(n):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        out[i] = s >> 8
    return out.astype(np.float32) * np.float32(1.0 / 16777216.0)


def make_inputs(n, ch, seed=12345):
    f = _lcg_floats(seed, n * 9).reshape(n, 9)
    x = np.zeros((n, ch), np.float32)
    x[:, :3] = f[:, :3] - np.float32(0.5)
    for v in range(2):
        a = f[:, 3 + 3 * v:6 + 3 * v] * np.float32(2) - np.float32(1)
        l = np.sqrt((a[:, 0] * a[:, 0] + a[:, 1] * a[:, 1]).astype(np.float32) + a[:, 2] * a[:, 2]).astype(np.float32)
        bad = l < np.float32(1e-6)
        a = np.where(bad[:, None], np.array([1, 0, 0], np.float32), a)
        l = np.where(bad, np.float32(1), l)
        x[:, 3 + 3 * v:6 + 3 * v] = a / l[:, None]
    return x


def make_targets(n, seed=777):
    return (_lcg_floats(seed, n * 3) * np.float32(2)).reshape(n, 3)
