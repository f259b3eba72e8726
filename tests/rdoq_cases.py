"""Seeded RDOQ test cases shared by the oracle pin (test_oracle_vs_ref.py), the golden vectors and the GPU
parity tests: CABAC bit-estimate tables shaped like the reference's (TComTrQuant.h:59-72, entries are
-log2(p) * 2^15 of a binary context, TEncSbac::estBit -> ContextModel::getEntropyBits) and transform
coefficient blocks whose levels span 0 .. tens, so that every branch of xRateDistOptQuant is taken
(level decrement, zeroed coefficient groups, moved last position, sign-data hiding)."""
from __future__ import annotations

import numpy as np

import oracle


def make_est(rng) -> "oracle.EstBits":
    est = oracle.EstBits()

    def pair():
        p = float(rng.uniform(0.02, 0.98))
        return int(round(-np.log2(p) * 32768)), int(round(-np.log2(1.0 - p) * 32768))

    for name, n in (("sig_cg", 2), ("sig", 42), ("greater_one", 24), ("level_abs", 6), ("block_cbp", 15),
                    ("block_root_cbp", 4)):
        arr = getattr(est, name)
        for i in range(n):
            arr[i][0], arr[i][1] = pair()
    for i in range(32):
        # last-position prefix: a truncated unary code, cost grows with the group index
        est.last_x[i] = int(rng.integers(8000, 60000)) + 9000 * min(i, 9)
        est.last_y[i] = int(rng.integers(8000, 60000)) + 9000 * min(i, 9)
    est.scan_zigzag[0], est.scan_zigzag[1] = pair()
    est.scan_non_zigzag[0], est.scan_non_zigzag[1] = pair()
    return est


def quant_step(log2: int, per: int, bd: int) -> float:
    """coefficient magnitude that quantises to level 1 (qbits - 14 bits, flat scale ~ 2^14)"""
    return float(2.0 ** (per + 15 - bd - log2))


def make_coef(rng, log2: int, per: int, bd: int, kind: int) -> np.ndarray:
    n = 1 << log2
    step = quant_step(log2, per, bd)
    yy, xx = np.mgrid[0:n, 0:n]
    if kind == 0:       # typical residual: energy decays with frequency, most high frequencies near zero
        amp = 6.0 * step * np.exp(-(xx + yy) / (0.35 * n + 1.0))
        c = rng.laplace(0.0, 1.0, (n, n)) * amp
    elif kind == 1:     # dense, levels around 0..3 everywhere (group zero-out and last-position decisions)
        c = rng.laplace(0.0, 0.8 * step, (n, n))
    elif kind == 2:     # large levels (Rice parameter growth, escape codes)
        c = rng.laplace(0.0, 40.0 * step, (n, n))
        c[rng.random((n, n)) < 0.5] = 0
    elif kind == 3:     # everything rounds to zero
        c = rng.uniform(-0.45 * step, 0.45 * step, (n, n))
    elif kind == 4:     # isolated coefficients near the rounding threshold
        c = np.zeros((n, n))
        k = max(1, n * n // 24)
        idx = rng.integers(0, n * n, k)
        c.flat[idx] = rng.choice([-1.0, 1.0], k) * rng.uniform(0.5, 2.6, k) * step
    else:               # extremes: the Int64 product saturates at MAX_INT - (1 << (qbits-1))
        c = rng.integers(-32768, 32768, (n, n)).astype(np.float64)
        c[rng.random((n, n)) < 0.3] = 0
    return np.clip(np.rint(c), -32768, 32767).astype(np.int32).reshape(-1)


def scan_idx_for(is_intra: int, is_luma: int, n: int, intra_dir: int) -> int:
    """TComDataCU::getCoefScanIdx (TComDataCU.cpp:4014-4066), ours: 0 diag, 1 hor, 2 ver"""
    if not is_intra:
        return 0
    multi = (n in (4, 8)) if is_luma else (n == 4)
    if not multi:
        return 0
    return 1 if abs(intra_dir - 26) < 5 else (2 if abs(intra_dir - 10) < 5 else 0)


def cbf_ctx_for(is_intra: int, is_luma: int, tr_idx: int) -> int:
    """TComTrQuant.cpp:2103-2115 + TComDataCU::getCtxQtCbf (SIMPLE_LUMA_CBF_CTX_DERIVATION)"""
    if (not is_intra) and is_luma and tr_idx == 0:
        return -1
    return (1 if tr_idx == 0 else 0) if is_luma else 5 + tr_idx


def lambda_for(qp: int) -> float:
    """an encoder-like Lagrangian (TEncSlice.cpp: 0.57 * 2^((QP-12)/3) scaled by a QP factor)"""
    return 0.57 * 2.0 ** ((qp - 12) / 3.0) * 0.4624
