"""shared inputs of the intra rough-search tests: reference lines + original blocks"""
import numpy as np


def make_case(rng, log2n, bd, kind):
    """(line[4N+1], org[N, N]) int16; kinds cover flat / extreme / alternating lines (clipping of the edge filter,
    DC rounding) and smooth content (where the [1 2 1] smoothing and the angular interpolation matter)"""
    n, mx = 1 << log2n, (1 << bd) - 1
    if kind == "random":
        line = rng.integers(0, mx + 1, 4 * n + 1)
    elif kind == "max":
        line = np.full(4 * n + 1, mx)
    elif kind == "zero":
        line = np.zeros(4 * n + 1)
    elif kind == "alternate":
        line = (np.arange(4 * n + 1) % 2) * mx
    elif kind == "step":
        line = np.where(np.arange(4 * n + 1) < 2 * n, 0, mx)
    else:       # smooth
        line = np.clip(np.cumsum(rng.integers(-3, 4, 4 * n + 1)) + mx // 2, 0, mx)
    org = rng.integers(0, mx + 1, (n, n)) if kind != "smooth" else np.clip(line[2 * n + 1:3 * n + 1][None, :] + rng.integers(-4, 5, (n, n)), 0, mx)
    return line.astype(np.int16), np.ascontiguousarray(org.astype(np.int16))


KINDS = ("random", "smooth", "max", "zero", "alternate", "step", "random", "smooth")


def oracle_rough(orc, line, org, log2n, bd, above=1, left=1, want_preds=False):
    from oracle import ptr
    n = 1 << log2n
    sad = np.zeros(35, np.uint32)
    preds = np.zeros((35, n, n), np.int16)
    orc.orc_intra_rough(ptr(line), ptr(org), n, log2n, above, left, bd, ptr(sad), ptr(preds))
    return (sad, preds) if want_preds else sad
