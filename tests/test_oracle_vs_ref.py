"""Pin the C restatement (oracle/hm_oracle*.c) against the reference's OWN compiled functions
(oracle/_ref/libhmref.so = /root/reference TLibCommon sources + forwarding shim).

CPU only.  This is the "oracle pinned" gate of the task: the reference ships no golden vectors
(SURVEY.md section 4), so the pin is the reference itself run here, plus tests/golden/.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import ptr

PU_SHAPES = [(64, 64), (64, 32), (32, 64), (64, 16), (64, 48), (16, 64), (48, 64),
             (32, 32), (32, 16), (16, 32), (32, 8), (32, 24), (8, 32), (24, 32),
             (16, 16), (16, 8), (8, 16), (16, 4), (16, 12), (4, 16), (12, 16),
             (8, 8), (8, 4), (4, 8)]


def rand_pels(rng, shape, bd, signed_bipred=False):
    if signed_bipred:
        return rng.integers(-((1 << bd) - 1), 2 * ((1 << bd) - 1) + 1, size=shape, dtype=np.int64).astype(np.int16)
    return rng.integers(0, 1 << bd, size=shape, dtype=np.int64).astype(np.int16)


@pytest.mark.parametrize("bd", [8, 10])
def test_distortion(orc, hmref, bd):
    rng = np.random.default_rng(1234 + bd)
    hmref.ref_init(bd)
    bi = bd - 8
    for (w, h) in PU_SHAPES:
        for trial in range(3):
            org = rand_pels(rng, (64, 64), bd, signed_bipred=(trial == 2))
            cur = rand_pels(rng, (80, 96), bd)
            if trial == 1:  # extremes
                org[:] = (1 << bd) - 1
                cur[:] = 0
            for ss in (0, 1):
                if ss == 1 and h <= 8:
                    continue
                a = orc.orc_sad(ptr(org), 64, ptr(cur, 96 * 3 + 5), 96, w, h, ss, bi)
                b = hmref.ref_sad_me(ptr(org), 64, ptr(cur, 96 * 3 + 5), 96, w, h, ss)
                assert a == b, (w, h, ss, trial)
            for had in (0, 1):
                a = (orc.orc_hads if had else orc.orc_sad)(ptr(org), 64, ptr(cur, 96 + 1), 96, w, h, *([bi] if had else [0, bi]))
                b = hmref.ref_dist_frac(ptr(org), 64, ptr(cur, 96 + 1), 96, w, h, had)
                assert a == b, (w, h, had, trial)
            for df in (1, 8, 22):
                a = orc.orc_get_dist_part(ptr(cur, 7), 96, ptr(org), 64, w, h, df, bi)
                b = hmref.ref_get_dist_part(ptr(cur, 7), 96, ptr(org), 64, w, h, df)
                assert a == b, (w, h, df, trial)
            a = orc.orc_calc_had(ptr(org), 64, ptr(cur), 96, w, h, bi)
            b = hmref.ref_calc_had(ptr(org), 64, ptr(cur), 96, w, h)
            assert a == b
    # chroma-sized SSE blocks (2x2 .. 32x32) through getDistPart
    for n in (2, 4, 8, 16, 32):
        org = rand_pels(rng, (64, 64), bd)
        cur = rand_pels(rng, (80, 96), bd)
        assert orc.orc_get_dist_part(ptr(cur), 96, ptr(org), 64, n, n, 1, bi) == \
            hmref.ref_get_dist_part(ptr(cur), 96, ptr(org), 64, n, n, 1)


def test_mv_cost(orc, hmref):
    hmref.ref_init(8)
    rng = np.random.default_rng(7)
    for v in list(range(-300, 301)) + [-(1 << 14), (1 << 14) - 1, 4096, -4097]:
        assert orc.orc_mv_component_bits(v) == hmref.ref_component_bits(v)
    for lam in (0.5, 4.7, 57.908390, 141.3, 1000.25):
        for _ in range(200):
            x, y = (int(t) for t in rng.integers(-200, 200, 2))
            px, py = (int(t) for t in rng.integers(-300, 300, 2))
            for scale in (0, 1, 2):
                lc = C.c_uint32()
                b = hmref.ref_mv_cost(lam, x, y, scale, px, py, C.byref(lc))
                assert orc.orc_lambda_motion_sad(lam) == lc.value
                assert orc.orc_mv_cost(lc.value, x, y, scale, px, py) == b


@pytest.mark.parametrize("bd", [8, 10])
def test_interpolation_primitives(orc, hmref, bd):
    rng = np.random.default_rng(99 + bd)
    hmref.ref_init(bd)
    S = 96
    for trial in range(4):
        pel = rand_pels(rng, (96, S), bd)
        if trial == 1:
            pel[:] = (1 << bd) - 1
        if trial == 2:
            pel[::2] = 0
            pel[1::2] = (1 << bd) - 1
        mid = rng.integers(-12272, 12209, size=(96, S)).astype(np.int16)  # 14-bit intermediates
        for (w, h) in [(64, 64), (65, 72), (17, 9), (4, 4), (8, 4), (2, 2), (33, 40)]:
            off = 8 * S + 8
            for frac in range(4):
                for last in (0, 1):
                    a = np.zeros((80, 80), np.int16); b = np.zeros((80, 80), np.int16)
                    orc.orc_filter_hor_luma(ptr(pel, off), S, ptr(a), 80, w, h, frac, last, bd)
                    hmref.ref_filter_hor_luma(ptr(pel, off), S, ptr(b), 80, w, h, frac, last)
                    assert np.array_equal(a, b), ("hl", w, h, frac, last)
                for first in (0, 1):
                    for last in (0, 1):
                        src = pel if first else mid
                        a = np.zeros((80, 80), np.int16); b = np.zeros((80, 80), np.int16)
                        orc.orc_filter_ver_luma(ptr(src, off), S, ptr(a), 80, w, h, frac, first, last, bd)
                        hmref.ref_filter_ver_luma(ptr(src, off), S, ptr(b), 80, w, h, frac, first, last)
                        assert np.array_equal(a, b), ("vl", w, h, frac, first, last)
            for frac in range(8):
                for last in (0, 1):
                    a = np.zeros((80, 80), np.int16); b = np.zeros((80, 80), np.int16)
                    orc.orc_filter_hor_chroma(ptr(pel, off), S, ptr(a), 80, w, h, frac, last, bd)
                    hmref.ref_filter_hor_chroma(ptr(pel, off), S, ptr(b), 80, w, h, frac, last)
                    assert np.array_equal(a, b), ("hc", w, h, frac, last)
                for first in (0, 1):
                    for last in (0, 1):
                        src = pel if first else mid
                        a = np.zeros((80, 80), np.int16); b = np.zeros((80, 80), np.int16)
                        orc.orc_filter_ver_chroma(ptr(src, off), S, ptr(a), 80, w, h, frac, first, last, bd)
                        hmref.ref_filter_ver_chroma(ptr(src, off), S, ptr(b), 80, w, h, frac, first, last)
                        assert np.array_equal(a, b), ("vc", w, h, frac, first, last)


def test_dct_tables_and_scans(orc, hmref):
    hmref.ref_init(8)
    for n in (4, 8, 16, 32):
        a = np.zeros(n * n, np.int16); b = np.zeros(n * n, np.int16)
        orc.orc_dct_matrix(n, a); hmref.ref_dct_matrix(n, b)
        assert np.array_equal(a, b), n
    for log2 in (2, 3, 4, 5):
        for scan in (0, 1, 2):
            a = np.zeros(1 << (2 * log2), np.uint32); b = np.zeros_like(a)
            orc.orc_scan(scan, log2, a); hmref.ref_scan(scan, log2, b)
            assert np.array_equal(a, b), (log2, scan)


@pytest.mark.parametrize("bd", [8, 10])
def test_transforms(orc, hmref, bd):
    rng = np.random.default_rng(5 + bd)
    hmref.ref_init(bd)
    bi = bd - 8
    amp = (1 << bd) - 1
    for n in (4, 8, 16, 32):
        for trial in range(6):
            if trial == 0:
                blk = np.full(n * n, amp, np.int16)
            elif trial == 1:
                blk = np.full(n * n, -amp, np.int16)
            elif trial == 2:
                blk = ((np.indices((n, n)).sum(0) % 2) * 2 * amp - amp).astype(np.int16).reshape(-1)
            elif trial == 3:   # full int16 range exercises the (short) wrap of the forward passes
                blk = rng.integers(-32768, 32768, n * n).astype(np.int16)
            else:
                blk = rng.integers(-amp, amp + 1, n * n).astype(np.int16)
            for shift in (1, 2, 5, 7, 9, 11, 12):
                a = np.zeros(n * n, np.int16); b = np.zeros(n * n, np.int16)
                orc.orc_partial_butterfly(n, blk, a, shift, n); hmref.ref_partial_butterfly(n, blk.copy(), b, shift, n)
                assert np.array_equal(a, b), ("fwd", n, shift, trial)
                orc.orc_partial_butterfly_inverse(n, blk, a, shift, n); hmref.ref_partial_butterfly_inverse(n, blk.copy(), b, shift, n)
                assert np.array_equal(a, b), ("inv", n, shift, trial)
            for dst in ((0, 1) if n == 4 else (0,)):
                a = np.zeros(n * n, np.int16); b = np.zeros(n * n, np.int16)
                orc.orc_xTrMxN(blk, a, n, n, dst, bi); hmref.ref_xTrMxN(blk.copy(), b, n, n, dst)
                assert np.array_equal(a, b), ("xTr", n, dst, trial)
                orc.orc_xITrMxN(blk, a, n, n, dst, bi); hmref.ref_xITrMxN(blk.copy(), b, n, n, dst)
                assert np.array_equal(a, b), ("xITr", n, dst, trial)
                # strided wrappers
                resi = rng.integers(-amp, amp + 1, (40, 64)).astype(np.int16)
                ca = np.zeros(n * n, np.int32); cb = np.zeros(n * n, np.int32)
                orc.orc_xT(dst, ptr(resi, 64 * 3 + 2), 64, ca, n, n, bi); hmref.ref_xT(dst, ptr(resi, 64 * 3 + 2), 64, cb, n, n)
                assert np.array_equal(ca, cb)
                co = rng.integers(-70000, 70000, n * n).astype(np.int32)
                ra = np.zeros((40, 64), np.int16); rb = np.zeros((40, 64), np.int16)
                orc.orc_xIT(dst, co, ptr(ra, 65), 64, n, n, bi); hmref.ref_xIT(dst, co.copy(), ptr(rb, 65), 64, n, n)
                assert np.array_equal(ra, rb)
    # 4x4 DST primitives and transform skip
    for trial in range(20):
        blk = rng.integers(-32768, 32768, 16).astype(np.int16)
        for shift in (1, 3, 7, 8, 12):
            a = np.zeros(16, np.int16); b = np.zeros(16, np.int16)
            orc.orc_fast_forward_dst(blk, a, shift); hmref.ref_fast_forward_dst(blk.copy(), b, shift)
            assert np.array_equal(a, b)
            orc.orc_fast_inverse_dst(blk, a, shift); hmref.ref_fast_inverse_dst(blk.copy(), b, shift)
            assert np.array_equal(a, b)
        resi = rng.integers(-amp, amp + 1, (8, 16)).astype(np.int16)
        ca = np.zeros(16, np.int32); cb = np.zeros(16, np.int32)
        orc.orc_transform_skip(ptr(resi, 17), 16, ca, 4, 4, bd); hmref.ref_transform_skip(ptr(resi, 17), 16, cb, 4, 4)
        assert np.array_equal(ca, cb)
        co = rng.integers(-40000, 40000, 16).astype(np.int32)
        ra = np.zeros((8, 16), np.int16); rb = np.zeros((8, 16), np.int16)
        orc.orc_itransform_skip(co, ptr(ra, 1), 16, 4, 4, bd); hmref.ref_itransform_skip(co.copy(), ptr(rb, 1), 16, 4, 4)
        assert np.array_equal(ra, rb)


@pytest.mark.parametrize("bd", [8, 10])
def test_quant_dequant(orc, hmref, bd):
    rng = np.random.default_rng(77 + bd)
    hmref.ref_init(bd)
    qp_bd_offset = 6 * (bd - 8)
    for n in (4, 8, 16, 32):
        log2 = int(np.log2(n))
        for qp in (0, 1, 17, 22, 27, 32, 37, 45, 51):
            for is_luma in ((1, 0) if n < 32 else (1,)):   # 4:2:0: no 32x32 chroma TU exists
                pa, ra_ = C.c_int(), C.c_int(); pb, rb_ = C.c_int(), C.c_int()
                orc.orc_set_qp(qp, is_luma, qp_bd_offset, 0, C.byref(pa), C.byref(ra_))
                hmref.ref_set_qp(qp, is_luma, qp_bd_offset, 0, C.byref(pb), C.byref(rb_))
                assert (pa.value, ra_.value) == (pb.value, rb_.value)
                for trial in range(3):
                    scale = [40, 2000, 32767][trial]
                    coef = rng.integers(-scale, scale + 1, n * n).astype(np.int32)
                    if trial == 2:
                        coef[rng.integers(0, n * n, n)] = 0
                    for (islice, icu, ldir, sh, arl) in [(1, 1, 26, 1, 1), (0, 0, 1, 1, 0), (0, 1, 10, 1, 1), (1, 1, 0, 0, 0)]:
                        base_qp = qp if trial != 1 else max(0, qp - 3)
                        scan_idx = 0
                        if icu and is_luma and n in (4, 8):
                            scan_idx = 1 if abs(ldir - 26) < 5 else (2 if abs(ldir - 10) < 5 else 0)
                        if icu and (not is_luma) and n == 4:
                            scan_idx = 1 if abs(ldir - 26) < 5 else (2 if abs(ldir - 10) < 5 else 0)
                        scan = np.zeros(n * n, np.uint32)
                        orc.orc_scan(scan_idx, log2, scan)
                        bper, brem = C.c_int(), C.c_int()
                        orc.orc_set_qp(base_qp, is_luma, qp_bd_offset, 0, C.byref(bper), C.byref(brem))
                        qpar = oracle.QuantParam(pa.value, ra_.value, bper.value, islice, sh, arl, bd)
                        qa = np.zeros(n * n, np.int32); qb = np.zeros(n * n, np.int32)
                        aa = np.zeros(n * n, np.int32); ab = np.zeros(n * n, np.int32)
                        sa, sb = C.c_uint32(0), C.c_uint32(0)
                        orc.orc_quant(coef, qa, ptr(aa), n, n, C.byref(qpar), scan, C.byref(sa))
                        hmref.ref_quant(coef.copy(), qb, ab, n, n, qp, base_qp, qp_bd_offset, is_luma, islice, icu, ldir, sh, arl, C.byref(sb))
                        assert sa.value == sb.value, (n, qp, trial)
                        assert np.array_equal(qa, qb), (n, qp, is_luma, trial, islice, icu, ldir)
                        if arl:
                            assert np.array_equal(aa, ab)
                    lev = rng.integers(-40000, 40000, n * n).astype(np.int32)
                    da = np.zeros(n * n, np.int32); db = np.zeros(n * n, np.int32)
                    orc.orc_dequant(lev, da, n, n, pa.value, ra_.value, bd)
                    hmref.ref_dequant(lev.copy(), db, n, n, qp, qp_bd_offset, is_luma)
                    assert np.array_equal(da, db), (n, qp, is_luma)


def test_extend_border(orc, hmref):
    hmref.ref_init(8)
    rng = np.random.default_rng(3)
    w, h, mx, my = 40, 24, 80, 80
    stride = w + 2 * mx
    a = np.zeros((h + 2 * my, stride), np.int16)
    a[my:my + h, mx:mx + w] = rng.integers(0, 256, (h, w))
    b = a.copy()
    orc.orc_extend_border(ptr(a, my * stride + mx), stride, w, h, mx, my)
    hmref.ref_extend_border(ptr(b, my * stride + mx), stride, w, h, mx, my)
    assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------------------
# motion search and motion compensation: the restatement against the reference's own TEncSearch /
# TComPrediction members (xSetSearchRange, xTZSearch, xPatternSearch, xPatternSearchFracDIF,
# xPredInterLumaBlk / xPredInterChromaBlk, TComYuv::addAvg)
def _census():
    out = []
    for depth in range(4):
        s = 64 >> depth
        for cy in range(0, 64, s):
            for cx in range(0, 64, s):
                parts = [(0, 0, s, s), (0, 0, s, s // 2), (0, s // 2, s, s // 2), (0, 0, s // 2, s), (s // 2, 0, s // 2, s)]
                if s >= 16:
                    q = s // 4
                    parts += [(0, 0, s, q), (0, q, s, s - q), (0, 0, s, s - q), (0, s - q, s, q),
                              (0, 0, q, s), (q, 0, s - q, s), (0, 0, s - q, s), (s - q, 0, q, s)]
                for (x, y, w, h) in parts:
                    out.append((cx + x, cy + y, w, h, cx, cy))
    return out


def _padded(rng, w, h, bd, kind):
    import synth
    if kind == "noise":
        return synth.random_pic(rng, w, h, bd)
    seq = synth.make_sequence(w, h, 2, seed=int(rng.integers(1, 1 << 30)), bit_depth=bd)
    return synth.to_hostpic(seq[0], w, h), synth.to_hostpic(seq[1], w, h)


def test_census_matches_abi_order(orc):
    c = np.zeros((593, 6), np.int16)
    orc.orc_census(c.ctypes.data_as(C.c_void_p))
    assert [tuple(int(v) for v in r) for r in c] == _census()


@pytest.mark.parametrize("bd", [8, 10])
def test_motion_search_vs_reference(orc, hmref, bd):
    W, H = 416, 240
    hmref.ref_init(bd)
    rng = np.random.default_rng(400 + bd)
    census = _census()
    lam = 57.908390
    lc = orc.orc_lambda_motion_sad(lam)
    checked = 0
    for content in ("synthetic", "noise"):
        if content == "noise":
            cur, ref = _padded(rng, W, H, bd, "noise"), _padded(rng, W, H, bd, "noise")
        else:
            ref, cur = _padded(rng, W, H, bd, "synthetic")
        for fen in (1, 0):
            for had in (1, 0):
                hmref.ref_me_setup(W, H, 64, fen, had)
                for trial in range(70 if content == "synthetic" else 30):
                    ctu = int(rng.integers(0, 7 * 4))
                    x0, y0 = (ctu % 7) * 64, (ctu // 7) * 64
                    px, py, w, h, cux, cuy = census[int(rng.integers(0, 593))]
                    x, y = x0 + px, y0 + py
                    if x + w > W or y + h > H:
                        continue
                    predx, predy = (int(v) for v in rng.integers(-60, 61, 2))
                    if trial % 5 == 4:
                        predx, predy = (int(v) for v in rng.integers(-520, 521, 2))
                    if trial % 17 == 16:
                        predx, predy = -4000, 4000
                    g = oracle.CuGeom(W, H, x0 + cux, y0 + cuy, 64)
                    lx, ty, rx, by = C.c_int(), C.c_int(), C.c_int(), C.c_int()
                    orc.orc_set_search_range(C.byref(g), predx, predy, 64, C.byref(lx), C.byref(ty), C.byref(rx), C.byref(by))
                    r4 = (C.c_int * 4)()
                    hmref.ref_set_search_range(x0 + cux, y0 + cuy, predx, predy, 64, r4)
                    assert (lx.value, ty.value, rx.value, by.value) == tuple(r4)
                    o = ptr(cur.buf_y, cur.origin(0) + y * cur.stride + x)
                    r = ptr(ref.buf_y, ref.origin(0) + y * ref.stride + x)
                    # integer stage: TZ on the full window; exhaustive on a +-4 window (bi-pred refinement size)
                    e = oracle.MeResult()
                    orc.orc_tz_search(C.byref(g), o, cur.stride, r, ref.stride, w, h, lx.value, ty.value, rx.value, by.value,
                                      64, fen, bd - 8, lc, predx, predy, predx, predy, C.byref(e))
                    o3 = (C.c_int * 3)()
                    hmref.ref_int_search(1, o, cur.stride, r, ref.stride, w, h, x0 + cux, y0 + cuy, lx.value, ty.value, rx.value,
                                         by.value, lam, predx, predy, predx, predy, o3)
                    assert (e.mvx, e.mvy, e.sad) == (o3[0], o3[1], o3[2] & 0xffffffff), ("tz", content, fen, x, y, w, h, predx, predy)
                    l4 = (C.c_int * 4)()
                    hmref.ref_set_search_range(x0 + cux, y0 + cuy, predx, predy, 4, l4)
                    orc.orc_pattern_search(o, cur.stride, r, ref.stride, w, h, l4[0], l4[1], l4[2], l4[3], fen, bd - 8, lc, predx, predy,
                                           C.byref(e))
                    hmref.ref_int_search(0, o, cur.stride, r, ref.stride, w, h, x0 + cux, y0 + cuy, l4[0], l4[1], l4[2], l4[3],
                                         lam, predx, predy, predx, predy, o3)
                    assert (e.mvx, e.mvy, e.sad) == (o3[0], o3[1], o3[2] & 0xffffffff), ("full", content, fen, x, y, w, h)
                    # fractional stage at the integer MV the search found
                    f = oracle.FracResult()
                    orc.orc_frac_search(o, cur.stride, r, ref.stride, w, h, e.mvx, e.mvy, had, bd - 8, bd, lc, predx, predy, C.byref(f))
                    o5 = (C.c_int * 5)()
                    hmref.ref_frac_search(o, cur.stride, r, ref.stride, w, h, e.mvx, e.mvy, lam, predx, predy, o5)
                    assert (f.halfx, f.halfy, f.qtrx, f.qtry, f.cost) == (o5[0], o5[1], o5[2], o5[3], o5[4] & 0xffffffff), \
                        ("frac", content, had, x, y, w, h, e.mvx, e.mvy)
                    checked += 1
    assert checked > 300


@pytest.mark.parametrize("bd", [8, 10])
def test_motion_compensation_vs_reference(orc, hmref, bd):
    W, H = 416, 240
    hmref.ref_init(bd)
    hmref.ref_me_setup(W, H, 64, 1, 1)
    rng = np.random.default_rng(500 + bd)
    ref = _padded(rng, W, H, bd, "noise")
    assert ref.stride == W + 160 and ref.buf_y.shape[0] == H + 160
    hmref.ref_mc_set_ref(ptr(ref.buf_y), ptr(ref.buf_u), ptr(ref.buf_v), W, H)
    census = _census()
    n = 0
    for trial in range(300):
        ctu = int(rng.integers(0, 7 * 3))
        x0, y0 = (ctu % 7) * 64, (ctu // 7) * 64
        px, py, w, h, cux, cuy = census[int(rng.integers(0, 593))]
        x, y = x0 + px, y0 + py
        if x + w > W or y + h > H:
            continue
        mv = [int(v) for v in rng.integers(-300, 300, 2)]
        if trial % 7 == 0:
            mv[0] &= ~3
        if trial % 11 == 0:
            mv[1] &= ~3
        if trial % 13 == 0:
            mv = [-4000, 4000]
        g = oracle.CuGeom(W, H, x0 + cux, y0 + cuy, 64)
        a, b = C.c_int(mv[0]), C.c_int(mv[1])
        orc.orc_clip_mv(C.byref(g), C.byref(a), C.byref(b))
        mvx, mvy = a.value, b.value
        outs = {}
        for bi in (0, 1):
            ey = np.zeros((h, w), np.int16); eu = np.zeros((h // 2, w // 2), np.int16); ev = np.zeros_like(eu)
            orc.orc_pred_inter_luma_blk(ptr(ref.buf_y, ref.origin(0) + y * ref.stride + x), ref.stride, mvx, mvy, w, h, ptr(ey), w, bi, bd)
            co = ref.origin(1) + (y // 2) * ref.cstride + x // 2
            orc.orc_pred_inter_chroma_blk(ptr(ref.buf_u, co), ref.cstride, mvx, mvy, w, h, ptr(eu), w // 2, bi, bd)
            orc.orc_pred_inter_chroma_blk(ptr(ref.buf_v, co), ref.cstride, mvx, mvy, w, h, ptr(ev), w // 2, bi, bd)
            ry = np.zeros((h, w), np.int16); ru = np.zeros((h // 2, w // 2), np.int16); rv = np.zeros_like(ru)
            hmref.ref_mc_pu(x, y, w, h, mvx, mvy, bi, ptr(ry), ptr(ru), ptr(rv))
            assert np.array_equal(ey, ry) and np.array_equal(eu, ru) and np.array_equal(ev, rv), (bi, x, y, w, h, mvx, mvy)
            outs[bi] = (ey, eu, ev)
        # bi-prediction average of the 14-bit prediction with a shifted copy of itself
        a14 = outs[1]
        b14 = tuple(np.ascontiguousarray(np.roll(p, 1, axis=1)) for p in a14)
        oy = np.zeros((h, w), np.int16); ou = np.zeros((h // 2, w // 2), np.int16); ov = np.zeros_like(ou)
        hmref.ref_add_avg(ptr(a14[0]), ptr(a14[1]), ptr(a14[2]), ptr(b14[0]), ptr(b14[1]), ptr(b14[2]), w, h, ptr(oy), ptr(ou), ptr(ov))
        for pa, pb, po, (ww, hh) in ((a14[0], b14[0], oy, (w, h)), (a14[1], b14[1], ou, (w // 2, h // 2)), (a14[2], b14[2], ov, (w // 2, h // 2))):
            e = np.zeros((hh, ww), np.int16)
            orc.orc_add_avg(ptr(pa), ww, ptr(pb), ww, ptr(e), ww, ww, hh, bd)
            assert np.array_equal(e, po)
        n += 1
    assert n > 150


@pytest.mark.parametrize("bd", [8, 10])
def test_rdoq(orc, hmref, bd):
    """xRateDistOptQuant (TComTrQuant.cpp:1719-2305): levels, uiAbsSum and ARL coefficients of the restatement
    equal the reference's for every TU size, scan, cbf context, with and without sign-data hiding."""
    import rdoq_cases as rc
    rng = np.random.default_rng(4242 + bd)
    hmref.ref_init(bd)
    qp_bd_offset = 6 * (bd - 8)
    changed = zeroed = 0
    for n in (4, 8, 16, 32):
        log2 = int(np.log2(n))
        for qp in (4, 22, 27, 32, 37, 51):
            est = rc.make_est(rng)
            for is_luma in ((1, 0) if n < 32 else (1,)):
                per, rem = C.c_int(), C.c_int()
                orc.orc_set_qp(qp, is_luma, qp_bd_offset, 0, C.byref(per), C.byref(rem))
                for kind in range(6):
                    for (icu, ldir, tr_idx, sh, arl) in [(0, 1, 0, 1, 1), (0, 1, 1, 1, 0), (1, 26, 0, 1, 1), (1, 10, 1, 1, 0),
                                                         (1, 1, 0, 0, 0), (1, 34, 2, 1, 0)]:
                        coef = rc.make_coef(rng, log2, per.value, bd, kind)
                        lam = rc.lambda_for(qp) * float(rng.uniform(0.5, 2.0))
                        scan_idx = rc.scan_idx_for(icu, is_luma, n, ldir)
                        scan = np.zeros(n * n, np.uint32)
                        orc.orc_scan(scan_idx, log2, scan)
                        par = oracle.RdoqParam(log2, is_luma, scan_idx, per.value, rem.value, bd,
                                               rc.cbf_ctx_for(icu, is_luma, tr_idx), sh, arl, lam)
                        qa = np.zeros(n * n, np.int32); qb = np.zeros(n * n, np.int32)
                        aa = np.full(n * n, -7, np.int32); ab = np.full(n * n, -7, np.int32)
                        sa, sb = C.c_uint32(0), C.c_uint32(0)
                        orc.orc_rdoq(coef, qa, ptr(aa), C.byref(par), C.byref(est), scan, C.byref(sa))
                        hmref.ref_rdoq(coef.copy(), qb, ab, n, qp, qp_bd_offset, is_luma, icu, ldir, tr_idx, sh, arl, lam,
                                       C.byref(est), C.byref(sb))
                        key = (n, qp, is_luma, kind, icu, ldir, tr_idx, sh, arl)
                        assert sa.value == sb.value, key
                        assert np.array_equal(qa, qb), key
                        assert np.array_equal(aa, ab), key
                        # how much RDOQ moved away from plain rounding (the test must exercise the decisions)
                        plain = (np.abs(coef).astype(np.int64) * [26214, 23302, 20560, 18396, 16384, 14564][rem.value]
                                 + (1 << (14 + per.value + 15 - bd - log2 - 1))) >> (14 + per.value + 15 - bd - log2)
                        changed += int(np.count_nonzero(np.abs(qa) != np.minimum(plain, 1 << 30)))
                        zeroed += int(np.count_nonzero((qa == 0) & (plain > 0)))
    assert changed > 5000 and zeroed > 3000, (changed, zeroed)


def test_me_frame_ctu_against_reference_driver(orc, hmref):
    """the census-wide search of whole CTUs (593 PUs x 2 references: xSetSearchRange + xTZSearch + xPatternSearchFracDIF):
    restatement (orc_me_frame_ctu) == the reference's own TEncSearch looped by ref_me_frame_ctu, including a CTU at the
    right / bottom picture border (clipped windows, PUs outside the picture)"""
    import synth
    if not hasattr(hmref, "ref_me_frame_ctu"):
        pytest.skip("libhmref.so predates ref_me_frame_ctu")
    W, H = 416, 240
    hmref.ref_init(8)
    hmref.ref_me_setup(W, H, 64, 1, 1)
    seq = synth.make_sequence(W, H, 3)
    cur = synth.to_hostpic(seq[2], W, H)
    refs = [synth.to_hostpic(seq[1], W, H), synth.to_hostpic(seq[0], W, H)]
    census = np.zeros((593, 6), np.int16)
    orc.orc_census(census.ctypes.data_as(C.c_void_p))
    rp = (C.c_void_p * 2)(*[ptr(r.buf_y, r.origin(0)).value for r in refs])
    lam = 44.7
    lc = orc.orc_lambda_motion_sad(lam)
    rng = np.random.default_rng(12)
    ctus_x = (W + 63) // 64
    for ctu in (8, ctus_x - 1, ctus_x * 3 + 2):          # interior, right border, bottom (partial) row
        x0, y0 = (ctu % ctus_x) * 64, (ctu // ctus_x) * 64
        pred = rng.integers(-30, 31, 4).astype(np.int32)
        oi = np.zeros((2 * 593, 4), np.int32); of = np.zeros((2 * 593, 5), np.int32)
        hmref.ref_me_frame_ctu(ptr(cur.buf_y, cur.origin(0)), rp, 2, cur.stride, W, H, x0, y0, ptr(pred), lam, 64, ptr(census), ptr(oi), ptr(of))
        ires = (oracle.MeResult * (2 * 593))(); fres = (oracle.FracResult * (2 * 593))()
        orc.orc_me_frame_ctu(ptr(cur.buf_y, cur.origin(0)), rp, 2, cur.stride, W, H, x0, y0, ptr(pred), lc, 64, 1, 1, 1, 8, ires, fres)
        valid = 0
        for i in range(2 * 593):
            if oi[i, 3] == 0:
                assert ires[i].n_sads == 0
                continue
            valid += 1
            assert (ires[i].mvx, ires[i].mvy, ires[i].sad) == tuple(int(v) for v in oi[i, :3]), (ctu, i)
            assert (fres[i].halfx, fres[i].halfy, fres[i].qtrx, fres[i].qtry, fres[i].cost) == tuple(int(v) for v in of[i]), (ctu, i)
        assert valid > 300


@pytest.mark.parametrize("bd", [8, 10])
def test_intra_rough_search(orc, hmref, bd):
    """the restated predIntraLumaAng (planar, DC + edge filter, 33 angular modes with the projected side
    reference, the pure horizontal / vertical edge filter and its clip) and calcHAD == the reference's own compiled
    TComPrediction / TComPattern::getPredictorPtr / TComRdCost for every PU size, every mode, sample by sample"""
    import intra_cases
    if not hasattr(hmref, "ref_intra_rough"):
        pytest.skip("libhmref.so predates ref_intra_rough")
    hmref.ref_init(bd)
    rng = np.random.default_rng(4242 + bd)
    for log2n in (2, 3, 4, 5, 6):
        n = 1 << log2n
        for kind in intra_cases.KINDS:
            line, org = intra_cases.make_case(rng, log2n, bd, kind)
            filt = np.zeros_like(line)
            orc.orc_intra_filter_line(ptr(line), n, ptr(filt))
            for above, left in ((1, 1), (1, 0), (0, 1), (0, 0)):
                osad, opred = intra_cases.oracle_rough(orc, line, org, log2n, bd, above, left, want_preds=True)
                rsad, rpred = np.zeros(35, np.uint32), np.zeros((35, n, n), np.int16)
                hmref.ref_intra_rough(ptr(line), ptr(filt), ptr(org), n, log2n, above, left, ptr(rsad), ptr(rpred))
                bad = [m for m in range(35) if not np.array_equal(opred[m], rpred[m])]
                assert not bad, (log2n, kind, above, left, bad)
                assert np.array_equal(osad, rsad)
