"""numpy float32 restatement of ATen's CPU inner-dim sum order -- a checker for the checker:
it pins the summation order that emojivoice_b200/csrc/aten_sum.h implements on the device."""
import numpy as np

F = np.float32


def _ceil_log2(x):
    if x <= 2:
        return 1
    return int(x - 1).bit_length()


def _row_sum(v):  # v: (size, W) float32 -> (W,)
    size, W = v.shape
    n4 = size // 4
    lp = max(4, _ceil_log2(n4) // 4)
    step, mask = 1 << lp, (1 << lp) - 1
    acc = np.zeros((4, 4, W), dtype=F)
    i = 0
    while i + step <= n4:
        for _ in range(step):
            for k in range(4):
                acc[0, k] = acc[0, k] + v[4 * i + k]
            i += 1
        for L in range(1, 4):
            acc[L] = acc[L] + acc[L - 1]
            acc[L - 1] = 0
            if i & (mask << (L * lp)):
                break
    while i < n4:
        for k in range(4):
            acc[0, k] = acc[0, k] + v[4 * i + k]
        i += 1
    for L in range(1, 4):
        acc[0] = acc[0] + acc[L]
    for r in range(4 * n4, size):
        acc[0, 0] = acc[0, 0] + v[r]
    for k in range(1, 4):
        acc[0, 0] = acc[0, 0] + acc[0, k]
    return acc[0, 0]


def sum_f32(x):
    x = np.asarray(x, dtype=F)
    n = x.shape[0]
    if n < 8:
        return float(_row_sum(x.reshape(n, 1))[0])
    nv = n // 8
    lanes = _row_sum(x[: nv * 8].reshape(nv, 8))
    s = F(0)
    for k in range(nv * 8, n):
        s = F(s + x[k])
    for l in range(8):
        s = F(s + lanes[l])
    return float(s)
