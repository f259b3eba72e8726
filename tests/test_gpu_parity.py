"""GPU parity: libthevc_cuda.so (through the C ABI) against the CPU oracle on the same seeded inputs.

Bit-exact bar everywhere (integer / index work).  Every test needs a B200 and goes through
thevc_b200.TLibCuda -> ctypes -> extern "C"; the oracle (oracle/libhm_oracle.so, itself pinned against
the compiled reference by tests/test_oracle_vs_ref.py and tests/golden/) is only the checker.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import ptr as optr
import synth
from thevc_b200 import TLibCuda, capi
from thevc_b200.capi import DistJob, FracJob, MeJob, PU, QuantCfg, TU

pytestmark = pytest.mark.gpu

PU_SHAPES = [(64, 64), (64, 32), (32, 64), (64, 16), (64, 48), (16, 64), (48, 64),
             (32, 32), (32, 16), (16, 32), (32, 8), (32, 24), (8, 32), (24, 32),
             (16, 16), (16, 8), (8, 16), (16, 4), (16, 12), (4, 16), (12, 16),
             (8, 8), (8, 4), (4, 8)]
W, H = 416, 240


@pytest.fixture(scope="module")
def orc():
    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="module")
def ctx8():
    t = TLibCuda(W, H, 8, num_slots=8)
    yield t
    t.close()


@pytest.fixture(scope="module")
def ctx10():
    t = TLibCuda(W, H, 10, num_slots=6)
    yield t
    t.close()


def _pic(rng, bd, kind="random"):
    if kind == "random":
        return synth.random_pic(rng, W, H, bd)
    raise ValueError(kind)


# ----------------------------------------------------------------------------------- pictures
def test_upload_download_extend(ctx8, orc):
    rng = np.random.default_rng(1)
    p = synth.random_pic(rng, W, H, 8, extend=False)
    ctx8.upload(0, p, with_margin=False)
    ctx8.extend_border(0)
    got = ctx8.download(0, with_margin=True)
    p.extend_border()
    q = synth.random_pic(np.random.default_rng(1), W, H, 8, extend=False)
    orc.orc_extend_border(optr(q.buf_y, q.origin(0)), q.stride, W, H, q.mx, q.my)
    assert np.array_equal(q.buf_y, p.buf_y)
    assert np.array_equal(got.buf_y, p.buf_y)
    assert np.array_equal(got.buf_u, p.buf_u)
    assert np.array_equal(got.buf_v, p.buf_v)


@pytest.mark.parametrize("bd", [8, 10])
def test_region_ops(ctx8, ctx10, orc, bd):
    t = ctx8 if bd == 8 else ctx10
    rng = np.random.default_rng(2 + bd)
    a, b = _pic(rng, bd), _pic(rng, bd)
    t.upload(0, a); t.upload(1, b); t.upload(2, a)
    for plane, (x, y, w, h) in [(0, (64, 32, 64, 64)), (1, (8, 4, 32, 32)), (2, (100, 60, 8, 4)), (0, (0, 0, W, H))]:
        ha, hb = a.plane(plane), b.plane(plane)
        st = a.stride if plane == 0 else a.cstride
        off = a.origin(plane) + y * st + x
        # subtract
        t.subtract(3, 0, 1, plane, x, y, w, h)
        exp = np.zeros_like(ha)
        orc.orc_subtract(optr(ha, off), st, optr(hb, off), st, optr(exp, off), st, w, h)
        got = t.download(3).plane(plane)
        ys, xs = np.divmod(off, st)
        assert np.array_equal(got[ys:ys + h, xs:xs + w], exp[ys:ys + h, xs:xs + w])
        # addClip of (a, a-b) = clip(2a-b)
        t.add_clip(4, 0, 3, plane, x, y, w, h)
        exp2 = np.zeros_like(ha)
        orc.orc_add_clip(optr(ha, off), st, optr(exp, off), st, optr(exp2, off), st, w, h, bd)
        got2 = t.download(4).plane(plane)
        assert np.array_equal(got2[ys:ys + h, xs:xs + w], exp2[ys:ys + h, xs:xs + w])
        # removeHighFreq: dst = 2*dst - a   (dst = slot 2 holds a copy of a)
        t.upload(2, a)
        t.remove_high_freq(2, 1, plane, x, y, w, h)
        exp3 = ha.copy()
        orc.orc_remove_high_freq(optr(exp3, off), st, optr(hb, off), st, w, h)
        got3 = t.download(2).plane(plane)
        assert np.array_equal(got3[ys:ys + h, xs:xs + w], exp3[ys:ys + h, xs:xs + w])


# ----------------------------------------------------------------------------------- distortion
@pytest.mark.parametrize("bd", [8, 10])
def test_dist_block_dropin(ctx8, ctx10, orc, bd):
    t = ctx8 if bd == 8 else ctx10
    bi = bd - 8
    rng = np.random.default_rng(10 + bd)
    maxv = (1 << bd) - 1
    for (w, h) in PU_SHAPES:
        for trial in range(3):
            org = rng.integers(0, maxv + 1, (64, 64)).astype(np.int16)
            cur = rng.integers(0, maxv + 1, (80, 96)).astype(np.int16)
            if trial == 1:
                org[:] = maxv; cur[:] = 0
            if trial == 2:   # bi-pred ME target 2*org - pred, unclipped (TComYuv.cpp:603)
                org = rng.integers(-maxv, 2 * maxv + 1, (64, 64)).astype(np.int16)
            co = 96 * 3 + 5
            for ss in (0, 1):
                if ss and h <= 8:
                    continue
                assert t.dist_block(capi.DIST_SAD, org, 0, 64, cur, co, 96, w, h, ss) == \
                    orc.orc_sad(optr(org), 64, optr(cur, co), 96, w, h, ss, bi), (w, h, ss, trial)
            assert t.dist_block(capi.DIST_SSE, org, 0, 64, cur, co, 96, w, h) == orc.orc_sse(optr(org), 64, optr(cur, co), 96, w, h, bi)
            assert t.dist_block(capi.DIST_HADS, org, 0, 64, cur, co, 96, w, h) == orc.orc_hads(optr(org), 64, optr(cur, co), 96, w, h, bi)
    for n in (2, 4, 8, 16, 32):     # chroma-sized blocks
        org = rng.integers(0, maxv + 1, (64, 64)).astype(np.int16)
        cur = rng.integers(0, maxv + 1, (80, 96)).astype(np.int16)
        assert t.dist_block(capi.DIST_SSE, org, 0, 64, cur, 0, 96, n, n) == orc.orc_sse(optr(org), 64, optr(cur), 96, n, n, bi)
        assert t.dist_block(capi.DIST_HADS, org, 0, 64, cur, 0, 96, n, n) == orc.orc_hads(optr(org), 64, optr(cur), 96, n, n, bi)


@pytest.mark.parametrize("bd", [8, 10])
def test_dist_batch_on_pictures(ctx8, ctx10, orc, bd):
    t = ctx8 if bd == 8 else ctx10
    bi = bd - 8
    rng = np.random.default_rng(20 + bd)
    a, b = _pic(rng, bd), _pic(rng, bd)
    t.upload(0, a); t.upload(1, b)
    jobs, exp = [], []
    for k in range(600):
        w, h = PU_SHAPES[k % len(PU_SHAPES)]
        kind = k % 3
        plane = 0 if k % 5 else int(rng.integers(1, 3))
        pw, ph = (W, H) if plane == 0 else (W // 2, H // 2)
        if plane:
            w, h = max(2, w // 2), max(2, h // 2)
        ox, oy = int(rng.integers(0, pw - w + 1)), int(rng.integers(0, ph - h + 1))
        m = a.mx if plane == 0 else a.cmx
        cx, cy = int(rng.integers(-m, pw + m - w + 1)), int(rng.integers(-m, ph + m - h + 1))
        ss = 1 if (kind == 0 and h > 8 and k % 2) else 0
        jobs.append(DistJob(kind, 0, plane, ox, oy, 1, plane, cx, cy, w, h, ss))
        st = a.stride if plane == 0 else a.cstride
        oo = a.origin(plane) + oy * st + ox
        cc = b.origin(plane) + cy * st + cx
        pa, pb = a.plane(plane), b.plane(plane)
        if kind == 0:
            exp.append(orc.orc_sad(optr(pa, oo), st, optr(pb, cc), st, w, h, ss, bi))
        elif kind == 1:
            exp.append(orc.orc_sse(optr(pa, oo), st, optr(pb, cc), st, w, h, bi))
        else:
            exp.append(orc.orc_hads(optr(pa, oo), st, optr(pb, cc), st, w, h, bi))
    got = t.dist_batch(jobs)
    assert np.array_equal(got, np.array(exp, np.uint32))
    assert len(t.dist_batch([])) == 0


# ----------------------------------------------------------------------------------- interpolation
@pytest.mark.parametrize("bd", [8, 10])
def test_filter_dropins(ctx8, ctx10, orc, bd):
    t = ctx8 if bd == 8 else ctx10
    rng = np.random.default_rng(30 + bd)
    S = 96
    pel = rng.integers(0, 1 << bd, (96, S)).astype(np.int16)
    mid = rng.integers(-12272, 12209, (96, S)).astype(np.int16)
    off = 8 * S + 8
    for (w, h) in [(64, 64), (65, 72), (17, 9), (4, 4), (2, 2)]:
        for frac in range(4):
            for last in (0, 1):
                a = np.zeros((80, 80), np.int16); b = np.zeros((80, 80), np.int16)
                orc.orc_filter_hor_luma(optr(pel, off), S, optr(a), 80, w, h, frac, last, bd)
                t.filter_hor_luma(pel, off, S, b, 0, 80, w, h, frac, last)
                assert np.array_equal(a, b), ("hl", w, h, frac, last)
            for first in (0, 1):
                for last in (0, 1):
                    src = pel if first else mid
                    a = np.zeros((80, 80), np.int16); b = np.zeros((80, 80), np.int16)
                    orc.orc_filter_ver_luma(optr(src, off), S, optr(a), 80, w, h, frac, first, last, bd)
                    t.filter_ver_luma(src, off, S, b, 0, 80, w, h, frac, first, last)
                    assert np.array_equal(a, b), ("vl", w, h, frac, first, last)
        for frac in range(8):
            for last in (0, 1):
                a = np.zeros((80, 80), np.int16); b = np.zeros((80, 80), np.int16)
                orc.orc_filter_hor_chroma(optr(pel, off), S, optr(a), 80, w, h, frac, last, bd)
                t.filter_hor_chroma(pel, off, S, b, 0, 80, w, h, frac, last)
                assert np.array_equal(a, b), ("hc", w, h, frac, last)
            for first in (0, 1):
                for last in (0, 1):
                    src = pel if first else mid
                    a = np.zeros((80, 80), np.int16); b = np.zeros((80, 80), np.int16)
                    orc.orc_filter_ver_chroma(optr(src, off), S, optr(a), 80, w, h, frac, first, last, bd)
                    t.filter_ver_chroma(src, off, S, b, 0, 80, w, h, frac, first, last)
                    assert np.array_equal(a, b), ("vc", w, h, frac, first, last)


def _clip_mv(orc, x, y, mvx, mvy):
    g = oracle.CuGeom(W, H, (x // 64) * 64, (y // 64) * 64, 64)
    a, b = C.c_int(mvx), C.c_int(mvy)
    orc.orc_clip_mv(C.byref(g), C.byref(a), C.byref(b))
    return a.value, b.value


def _oracle_mc(orc, refs, pu, bd):
    """expected Y,U,V blocks of one PU (TComPrediction::motionCompensation)"""
    out = []
    for plane in range(3):
        sh = 1 if plane else 0
        w, h, x, y = pu.w >> sh, pu.h >> sh, pu.x >> sh, pu.y >> sh
        fn = orc.orc_pred_inter_chroma_blk if plane else orc.orc_pred_inter_luma_blk
        lists = [(s, mx, my) for (s, mx, my) in ((pu.ref_slot0, pu.mvx0, pu.mvy0), (pu.ref_slot1, pu.mvx1, pu.mvy1)) if s >= 0]
        bi = len(lists) == 2
        tmp = []
        for (s, mx, my) in lists:
            r = refs[s]
            st = r.stride if plane == 0 else r.cstride
            d = np.zeros((h, w), np.int16)
            # the chroma restatement takes the LUMA PU size (it halves internally, TComPrediction.cpp:606-607)
            fn(optr(r.plane(plane), r.origin(plane) + y * st + x), st, mx, my, pu.w if plane else w, pu.h if plane else h,
               optr(d), w, 1 if bi else 0, bd)
            tmp.append(d)
        if bi:
            d = np.zeros((h, w), np.int16)
            orc.orc_add_avg(optr(tmp[0]), w, optr(tmp[1]), w, optr(d), w, w, h, bd)
            out.append(d)
        else:
            out.append(tmp[0])
    return out


@pytest.mark.parametrize("bd", [8, 10])
def test_mc_batch(ctx8, ctx10, orc, bd):
    t = ctx8 if bd == 8 else ctx10
    rng = np.random.default_rng(40 + bd)
    refs = {0: _pic(rng, bd), 1: _pic(rng, bd)}
    t.upload(0, refs[0]); t.upload(1, refs[1])
    # one PU per CTU and round (non-overlapping inside a batch); shapes, list usage and MV classes cycle
    k = 0
    for rnd in range(8):
        pus = []
        for cy in range(0, H - 63, 64):
            for cx in range(0, W - 63, 64):
                w, h = PU_SHAPES[k % len(PU_SHAPES)]
                k += 1
                px, py = cx + int(rng.integers(0, (64 - w) // 4 + 1)) * 4, cy + int(rng.integers(0, (64 - h) // 4 + 1)) * 4
                mode = k % 3
                mv = [int(v) for v in rng.integers(-300, 300, 4)]
                if k % 7 == 0:
                    mv = [4 * (mv[0] // 4), 4 * (mv[1] // 4), mv[2], 4 * (mv[3] // 4)]   # integer / 1-D cases
                if k % 11 == 0:
                    mv = [-4000, -4000, 4000, 4000]                                       # clipped far outside
                m0 = _clip_mv(orc, px, py, mv[0], mv[1])
                m1 = _clip_mv(orc, px, py, mv[2], mv[3])
                if mode == 0:
                    pus.append(PU(px, py, w, h, 0, m0[0], m0[1], -1, 0, 0))
                elif mode == 1:
                    pus.append(PU(px, py, w, h, -1, 0, 0, 1, m1[0], m1[1]))
                else:
                    pus.append(PU(px, py, w, h, 0, m0[0], m0[1], 1, m1[0], m1[1]))
        t.mc_batch(2, pus)
        got = t.download(2)
        for pu in pus:
            exp = _oracle_mc(orc, refs, pu, bd)
            for plane in range(3):
                sh = 1 if plane else 0
                g = (got.y, got.u, got.v)[plane][pu.y >> sh:(pu.y + pu.h) >> sh, pu.x >> sh:(pu.x + pu.w) >> sh]
                assert np.array_equal(g, exp[plane]), (plane, pu.x, pu.y, pu.w, pu.h, pu.mvx0, pu.mvy0, pu.mvx1, pu.mvy1)


@pytest.mark.parametrize("bd", [8, 10])
def test_mc_block_dropin(ctx8, ctx10, orc, bd):
    """tvc_mc_block = xPredInterUni (luma + both chroma planes) into caller buffers, final and 14-bit forms"""
    t = ctx8 if bd == 8 else ctx10
    rng = np.random.default_rng(45 + bd)
    ref = _pic(rng, bd)
    t.upload(1, ref)
    for k in range(60):
        w, h = PU_SHAPES[k % len(PU_SHAPES)]
        x, y = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
        mv = [int(v) for v in rng.integers(-300, 300, 2)]
        if k % 7 == 0:
            mv[0] &= ~3
        if k % 5 == 0:
            mv[1] &= ~3
        mvx, mvy = _clip_mv(orc, x, y, mv[0], mv[1])
        for bi in (0, 1):
            gy, gu, gv = t.mc_block(1, x, y, w, h, mvx, mvy, bool(bi))
            ey = np.zeros((h, w), np.int16); eu = np.zeros((h // 2, w // 2), np.int16); ev = np.zeros_like(eu)
            orc.orc_pred_inter_luma_blk(optr(ref.buf_y, ref.origin(0) + y * ref.stride + x), ref.stride, mvx, mvy, w, h, optr(ey), w, bi, bd)
            co = ref.origin(1) + (y // 2) * ref.cstride + x // 2
            orc.orc_pred_inter_chroma_blk(optr(ref.buf_u, co), ref.cstride, mvx, mvy, w, h, optr(eu), w // 2, bi, bd)
            orc.orc_pred_inter_chroma_blk(optr(ref.buf_v, co), ref.cstride, mvx, mvy, w, h, optr(ev), w // 2, bi, bd)
            assert np.array_equal(gy, ey) and np.array_equal(gu, eu) and np.array_equal(gv, ev), (k, bi, x, y, w, h, mvx, mvy)


# ----------------------------------------------------------------------------------- integer ME
def _ctu_pus():
    """all 593 PU rectangles of one CTU, relative to the CTU origin"""
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
                    out.append((cx + x, cy + y, w, h))
    return out


def test_pu_census():
    assert len(_ctu_pus()) == 593
    assert {(w, h) for (_, _, w, h) in _ctu_pus()} == set(PU_SHAPES)


def test_me_sad_tables(ctx8, orc):
    t = ctx8
    seq = synth.make_sequence(W, H, 3)
    cur = synth.to_hostpic(seq[2], W, H)
    refs = [synth.to_hostpic(seq[1], W, H), synth.random_pic(np.random.default_rng(5), W, H, 8)]
    t.upload(0, cur); t.upload(1, refs[0]); t.upload(2, refs[1])
    nctu = t.ctus_x * t.ctus_y
    rng = np.random.default_rng(50)
    centers = np.zeros((2, nctu, 2), np.int32)
    centers[1] = rng.integers(-40, 41, (nctu, 2))
    t.me_prepass(0, [1, 2], centers)
    pus = _ctu_pus()
    ctus_y_full = H // 64          # the last CTU row of 416x240 is partial (48 rows): PUs must lie inside the picture
    for trial in range(120):
        ri = trial % 2
        ctu = int(rng.integers(0, nctu))
        cx0, cy0 = (ctu % t.ctus_x) * 64, (ctu // t.ctus_x) * 64
        px, py, w, h = pus[int(rng.integers(0, len(pus)))]
        if cx0 + px + w > W or cy0 + py + h > H:
            continue
        x, y = cx0 + px, cy0 + py
        # candidates: corners of the table window, centre, random
        ccx, ccy = int(centers[ri, ctu, 0]), int(centers[ri, ctu, 1])
        # the library clamps centres so the window stays inside the padded plane; mirror the clamp
        lo_x, hi_x = -cur.mx - cx0 + 64, W + cur.mx - cx0 - 64 - 64
        lo_y, hi_y = -cur.my - cy0 + 64, H + cur.my - cy0 - 64 - 64
        ccx = min(max(ccx, lo_x), max(hi_x, lo_x)); ccy = min(max(ccy, lo_y), max(hi_y, lo_y))
        cand = [(ccx - 64, ccy - 64), (ccx + 64, ccy + 64), (ccx + 64, ccy - 64), (ccx - 64, ccy + 64), (ccx, ccy)]
        cand += [(ccx + int(a), ccy + int(b)) for a, b in rng.integers(-64, 65, (40, 2))]
        for fen in (0, 1):
            got = t.me_table_lookup(ri, x, y, w, h, fen, np.array(cand, np.int16))
            ss = 1 if (fen and h > 8) else 0
            r = refs[ri]
            for (mvx, mvy), g in zip(cand, got):
                e = orc.orc_sad(optr(cur.buf_y, cur.origin(0) + y * cur.stride + x), cur.stride,
                                optr(r.buf_y, r.origin(0) + (y + mvy) * r.stride + x + mvx), r.stride, w, h, ss, 0)
                assert g == e, (ri, ctu, x, y, w, h, fen, mvx, mvy)
    assert ctus_y_full == 3


def _me_jobs(orc, rng, cur, n, mode, srange, ref_index=0, ref_slot=1, lam=57.9):
    pus = _ctu_pus()
    jobs, meta = [], []
    lc = orc.orc_lambda_motion_sad(lam)
    nctu_x, nctu_y = (W + 63) // 64, (H + 63) // 64
    while len(jobs) < n:
        ctu = int(rng.integers(0, nctu_x * nctu_y))
        cx0, cy0 = (ctu % nctu_x) * 64, (ctu // nctu_x) * 64
        px, py, w, h = pus[int(rng.integers(0, len(pus)))]
        x, y = cx0 + px, cy0 + py
        if x + w > W or y + h > H:
            continue
        # CU origin for clipMv = the CU that owns the PU; use the enclosing CU of the PU's depth
        predx, predy = (int(v) for v in rng.integers(-60, 61, 2))
        if len(jobs) % 5 == 4:      # far predictor: the zero vector lies outside the search window
            predx, predy = (int(v) for v in rng.integers(-520, 521, 2))
        g = oracle.CuGeom(W, H, cx0, cy0, 64)     # m_uiCUPelX/Y of the CTU-level TComDataCU
        lx, ty, rx, by = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        orc.orc_set_search_range(C.byref(g), predx, predy, srange, C.byref(lx), C.byref(ty), C.byref(rx), C.byref(by))
        sx, sy = C.c_int(predx), C.c_int(predy)
        orc.orc_clip_mv(C.byref(g), C.byref(sx), C.byref(sy))
        fen = int(rng.integers(0, 2))
        jobs.append(MeJob(ref_index, ref_slot, x, y, w, h, mode, fen, srange, lx.value, ty.value, rx.value, by.value,
                          predx, predy, sx.value >> 2, sy.value >> 2, lc))
        meta.append((g, predx, predy))
    return jobs, meta, lc


def _oracle_me(orc, cur, ref, job, meta, bi=0):
    g, predx, predy = meta
    res = oracle.MeResult()
    o = optr(cur.buf_y, cur.origin(0) + job.y * cur.stride + job.x)
    r = optr(ref.buf_y, ref.origin(0) + job.y * ref.stride + job.x)
    if job.mode == capi.ME_FULL:
        orc.orc_pattern_search(o, cur.stride, r, ref.stride, job.w, job.h, job.lx, job.ty, job.rx, job.by, job.fen, bi,
                               job.lambda_cost, predx, predy, C.byref(res))
    else:
        orc.orc_tz_search(C.byref(g), o, cur.stride, r, ref.stride, job.w, job.h, job.lx, job.ty, job.rx, job.by,
                          job.search_range, job.fen, bi, job.lambda_cost, predx, predy, predx, predy, C.byref(res))
    return res


@pytest.mark.parametrize("use_tables", [False, True])
def test_me_search_tz(ctx8, orc, use_tables):
    t = ctx8
    seq = synth.make_sequence(W, H, 3)
    cur = synth.to_hostpic(seq[2], W, H)
    ref = synth.to_hostpic(seq[1], W, H)
    t.upload(0, cur); t.upload(1, ref)
    if use_tables:
        t.me_prepass(0, [1], None)
    rng = np.random.default_rng(60)
    jobs, meta, _ = _me_jobs(orc, rng, cur, 300, capi.ME_TZ, 64)
    got = t.me_search_batch(0, jobs, use_tables)
    for j, m, g in zip(jobs, meta, got):
        e = _oracle_me(orc, cur, ref, j, m)
        assert (g.mvx, g.mvy, g.sad, g.n_sads) == (e.mvx, e.mvy, e.sad, e.n_sads), (j.x, j.y, j.w, j.h, j.fen, m[1], m[2])


def test_me_search_tz_noise(ctx8, orc):
    """i.i.d. noise: worst-case TZ paths (raster scan, many refinement rounds)"""
    t = ctx8
    rng = np.random.default_rng(61)
    cur, ref = _pic(rng, 8), _pic(rng, 8)
    t.upload(0, cur); t.upload(1, ref)
    t.me_prepass(0, [1], None)
    jobs, meta, _ = _me_jobs(orc, rng, cur, 120, capi.ME_TZ, 64)
    for use_tables in (False, True):
        got = t.me_search_batch(0, jobs, use_tables)
        for j, m, g in zip(jobs, meta, got):
            e = _oracle_me(orc, cur, ref, j, m)
            assert (g.mvx, g.mvy, g.sad, g.n_sads) == (e.mvx, e.mvy, e.sad, e.n_sads), (use_tables, j.x, j.y, j.w, j.h)


def test_me_search_full_bipred_window(ctx8, orc):
    """xPatternSearch over the +-4 bi-pred refinement window (TEncSearch.cpp:4227; bipredSearchRange 4)"""
    t = ctx8
    seq = synth.make_sequence(W, H, 2)
    cur = synth.to_hostpic(seq[1], W, H)
    ref = synth.to_hostpic(seq[0], W, H)
    t.upload(0, cur); t.upload(1, ref)
    t.me_prepass(0, [1], None)
    rng = np.random.default_rng(62)
    jobs, meta, _ = _me_jobs(orc, rng, cur, 100, capi.ME_FULL, 4)
    for use_tables in (False, True):
        got = t.me_search_batch(0, jobs, use_tables)
        for j, m, g in zip(jobs, meta, got):
            e = _oracle_me(orc, cur, ref, j, m)
            assert (g.mvx, g.mvy, g.sad, g.n_sads) == (e.mvx, e.mvy, e.sad, e.n_sads)


@pytest.mark.parametrize("bd", [8, 10])
def test_me_bipred_refinement(ctx8, ctx10, orc, bd):
    """tvc_me_bipred = xMotionEstimation's bBi branch: the pattern is the host's block 2 * org - pred(other list) (values beyond the
    pel range), xPatternSearch over the +-4 window around the list's current vector, then xPatternSearchFracDIF -- one device call"""
    t = ctx8 if bd == 8 else ctx10
    rng = np.random.default_rng(90 + bd)
    ref = _pic(rng, bd)
    t.upload(1, ref)
    maxv = (1 << bd) - 1
    lc = orc.orc_lambda_motion_sad(47.0)
    n = 0
    for (w, h) in PU_SHAPES[::2] + [(64, 64), (8, 4)]:
        x, y = int(rng.integers(0, (W - w) // 4 + 1)) * 4, int(rng.integers(0, (H - h) // 4 + 1)) * 4
        cx, cy = int(rng.integers(-20, 21)), int(rng.integers(-20, 21))
        # 2 * org - pred: anywhere in [-maxv, 2 * maxv]; make it resemble the reference block near (cx, cy) so that the search has a minimum
        tgt = np.zeros((h, w), np.int16)
        for r in range(h):
            o = ref.origin(0) + (y + cy + 1 + r) * ref.stride + x + cx - 1
            tgt[r] = ref.buf_y.reshape(-1)[o:o + w]
        tgt = np.clip(2 * tgt.astype(np.int32) - rng.integers(0, maxv + 1, (h, w)), -maxv, 2 * maxv).astype(np.int16)
        predx, predy = int(rng.integers(-90, 91)), int(rng.integers(-90, 91))
        job = MeJob(-1, 1, x, y, w, h, capi.ME_FULL, 1, 64, cx - 4, cy - 4, cx + 4, cy + 4, predx, predy, 0, 0, lc)
        gi, gf = t.me_bipred(3, tgt, job, hadamard=True)
        e = oracle.MeResult()
        r0 = optr(ref.buf_y, ref.origin(0) + y * ref.stride + x)
        orc.orc_pattern_search(optr(tgt), w, r0, ref.stride, w, h, cx - 4, cy - 4, cx + 4, cy + 4, 1, bd - 8, lc, predx, predy, C.byref(e))
        assert (gi.mvx, gi.mvy, gi.sad, gi.n_sads) == (e.mvx, e.mvy, e.sad, e.n_sads), (w, h)
        f = oracle.FracResult()
        orc.orc_frac_search(optr(tgt), w, r0, ref.stride, w, h, e.mvx, e.mvy, 1, bd - 8, bd, lc, predx, predy, C.byref(f))
        assert (gf.halfx, gf.halfy, gf.qtrx, gf.qtry, gf.cost_half, gf.cost) == (f.halfx, f.halfy, f.qtrx, f.qtry, f.cost_half, f.cost), (w, h)
        n += 1
    assert n >= 12


def test_me_search_10bit_direct(ctx10, orc):
    t = ctx10
    rng = np.random.default_rng(63)
    cur, ref = _pic(rng, 10), _pic(rng, 10)
    t.upload(0, cur); t.upload(1, ref)
    jobs, meta, _ = _me_jobs(orc, rng, cur, 60, capi.ME_TZ, 64)
    got = t.me_search_batch(0, jobs, False)
    for j, m, g in zip(jobs, meta, got):
        e = _oracle_me(orc, cur, ref, j, m, bi=2)
        assert (g.mvx, g.mvy, g.sad, g.n_sads) == (e.mvx, e.mvy, e.sad, e.n_sads)


@pytest.fixture(params=[1, 0], ids=["group-search", "sad-tables"])
def ctx8_form(request, ctx8):
    """both forms of the fast integer stage: the group search (default; SADs on demand from the staged window) and the round-1
    form with full SAD tables in HBM (TVC_ME_FUSED=0)"""
    ctx8.L.tvc_me_set_fused(ctx8.h, request.param)
    assert ctx8.L.tvc_me_uses_tables(ctx8.h) == 1 - request.param
    yield ctx8
    ctx8.L.tvc_me_set_fused(ctx8.h, -1)


def test_me_frame_prepass(ctx8_form, orc):
    """frame pre-pass = for every census PU x CTU x reference: xSetSearchRange + xTZSearch + xPatternSearchFracDIF"""
    t = ctx8_form
    seq = synth.make_sequence(W, H, 3)
    cur = synth.to_hostpic(seq[2], W, H)
    refs = [synth.to_hostpic(seq[1], W, H), synth.to_hostpic(seq[0], W, H)]
    t.upload(0, cur); t.upload(1, refs[0]); t.upload(2, refs[1])
    nctu = t.ctus_x * t.ctus_y
    rng = np.random.default_rng(64)
    pred = rng.integers(-48, 49, (2, nctu, 2)).astype(np.int32)
    pred[1, 0] = (-4000, 4000)          # clipped by clipMv
    lc = orc.orc_lambda_motion_sad(57.9)
    census = t.me_census()
    assert [tuple(int(v) for v in c[:4]) for c in census] == _ctu_pus()
    ires, fres = t.me_frame(0, [1, 2], pred, lc)
    ires0, _ = t.me_frame(0, [1, 2], pred, lc, use_tables=False, do_frac=False)
    assert np.array_equal(ires, ires0)
    checked = 0
    for ri in range(2):
        ref = refs[ri]
        for ctu in range(nctu):
            x0, y0 = (ctu % t.ctus_x) * 64, (ctu // t.ctus_x) * 64
            ks = rng.choice(593, 24, replace=False) if ctu % 3 else range(593)
            if ctu % 3 == 0 and ctu > 6:
                ks = rng.choice(593, 60, replace=False)
            for k in ks:
                px, py, w, h, cux, cuy = (int(v) for v in census[k])
                x, y = x0 + px, y0 + py
                gi, gf = ires[ri, ctu, k], fres[ri, ctu, k]
                if x + w > W or y + h > H:
                    assert gi["n_sads"] == 0
                    continue
                g = oracle.CuGeom(W, H, x0 + cux, y0 + cuy, 64)
                predx, predy = int(pred[ri, ctu, 0]), int(pred[ri, ctu, 1])
                lx, ty, rx, by = C.c_int(), C.c_int(), C.c_int(), C.c_int()
                orc.orc_set_search_range(C.byref(g), predx, predy, 64, C.byref(lx), C.byref(ty), C.byref(rx), C.byref(by))
                o = optr(cur.buf_y, cur.origin(0) + y * cur.stride + x)
                r = optr(ref.buf_y, ref.origin(0) + y * ref.stride + x)
                e = oracle.MeResult()
                orc.orc_tz_search(C.byref(g), o, cur.stride, r, ref.stride, w, h, lx.value, ty.value, rx.value, by.value,
                                  64, 1, 0, lc, predx, predy, predx, predy, C.byref(e))
                assert (gi["mvx"], gi["mvy"], gi["sad"], gi["n_sads"]) == (e.mvx, e.mvy, e.sad, e.n_sads), (ri, ctu, k)
                f = oracle.FracResult()
                orc.orc_frac_search(o, cur.stride, r, ref.stride, w, h, e.mvx, e.mvy, 1, 0, 8, lc, predx, predy, C.byref(f))
                assert (gf["halfx"], gf["halfy"], gf["qtrx"], gf["qtry"], gf["cost_half"], gf["cost"]) == \
                    (f.halfx, f.halfy, f.qtrx, f.qtry, f.cost_half, f.cost), (ri, ctu, k)
                checked += 1
    assert checked > 2000


@pytest.mark.parametrize("kind", ["far-motion", "noise", "static-with-far-predictor"])
def test_me_group_search_hard_content(ctx8, orc, kind):
    """the group search where the PUs of a CTU disagree: content that moved far from the predictor (first sweep lands at distance
    > 5: raster stage + several refinement rounds), pure noise (every PU ends somewhere else: hundreds of distinct candidates per
    round) and a predictor far from a static scene (the zero vector wins: sweeps centred outside the staged window, candidates
    from the reference plane in global memory).  Every PU of every CTU against the per-PU kernel (SADs straight from the
    pictures, itself oracle-checked), a sample against the oracle."""
    t = ctx8
    rng = np.random.default_rng({"far-motion": 71, "noise": 72, "static-with-far-predictor": 73}[kind])
    nctu = t.ctus_x * t.ctus_y
    base = synth.random_pic(rng, W + 128, H + 128, 8, extend=False).y
    k3 = np.ones(3) / 3.0
    sm = np.apply_along_axis(lambda r: np.convolve(r, k3, mode="same"), 1, base.astype(np.float64))
    sm = np.apply_along_axis(lambda c: np.convolve(c, k3, mode="same"), 0, sm)
    cur, ref = synth.random_pic(rng, W, H, 8, extend=False), synth.random_pic(rng, W, H, 8, extend=False)
    if kind == "far-motion":
        cur.y[:] = np.rint(sm[64:64 + H, 64:64 + W]); ref.y[:] = np.rint(sm[64 - 29:64 - 29 + H, 64 + 37:64 + 37 + W])
        pred = rng.integers(-8, 9, (1, nctu, 2)).astype(np.int32)
    elif kind == "noise":
        pred = rng.integers(-60, 61, (1, nctu, 2)).astype(np.int32)
    else:
        cur.y[:] = np.rint(sm[64:64 + H, 64:64 + W]); ref.y[:] = cur.y + rng.integers(-2, 3, cur.y.shape)
        pred = np.zeros((1, nctu, 2), np.int32)
        pred[0, :, 0] = rng.choice([-600, 520, 300], nctu); pred[0, :, 1] = rng.choice([-380, 410, 280], nctu)
    cur.extend_border(); ref.extend_border()
    t.upload(0, cur); t.upload(1, ref)
    lc = orc.orc_lambda_motion_sad(33.0)
    t.L.tvc_me_set_fused(t.h, 1)
    try:
        ires, _ = t.me_frame(0, [1], pred, lc, do_frac=False)
        stats = t.me_frame_stats()
    finally:
        t.L.tvc_me_set_fused(t.h, -1)
    ires0, _ = t.me_frame(0, [1], pred, lc, use_tables=False, do_frac=False)
    bad = np.argwhere(ires != ires0)
    assert bad.size == 0, (kind, bad[:5], ires[tuple(bad[0])], ires0[tuple(bad[0])])
    print(kind, stats, "mean candidates per search", float(ires["n_sads"][ires["n_sads"] > 0].mean()))
    census = t.me_census()
    checked = 0
    for ctu in (int(v) for v in rng.choice(nctu, 6, replace=False)):
        x0, y0 = (ctu % t.ctus_x) * 64, (ctu // t.ctus_x) * 64
        for k in (int(v) for v in rng.choice(593, 12, replace=False)):
            px, py, w, h, cux, cuy = (int(v) for v in census[k])
            x, y = x0 + px, y0 + py
            if x + w > W or y + h > H:
                continue
            g = oracle.CuGeom(W, H, x0 + cux, y0 + cuy, 64)
            predx, predy = int(pred[0, ctu, 0]), int(pred[0, ctu, 1])
            lx, ty, rx, by = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            orc.orc_set_search_range(C.byref(g), predx, predy, 64, C.byref(lx), C.byref(ty), C.byref(rx), C.byref(by))
            e = oracle.MeResult()
            orc.orc_tz_search(C.byref(g), optr(cur.buf_y, cur.origin(0) + y * cur.stride + x), cur.stride,
                              optr(ref.buf_y, ref.origin(0) + y * ref.stride + x), ref.stride, w, h, lx.value, ty.value, rx.value, by.value,
                              64, 1, 0, lc, predx, predy, predx, predy, C.byref(e))
            gi = ires[0, ctu, k]
            assert (gi["mvx"], gi["mvy"], gi["sad"], gi["n_sads"]) == (e.mvx, e.mvy, e.sad, e.n_sads), (kind, ctu, k)
            checked += 1
    assert checked > 40


def test_me_frame_packed(ctx8, orc):
    """tvc_me_frame_packed: the 16-byte records hold exactly what tvc_me_frame returns in 40 bytes (minus cost_half / n_sads)"""
    t = ctx8
    seq = synth.make_sequence(W, H, 2)
    t.upload(0, synth.to_hostpic(seq[1], W, H)); t.upload(1, synth.to_hostpic(seq[0], W, H))
    nctu = t.ctus_x * t.ctus_y
    pred = np.random.default_rng(66).integers(-30, 31, (1, nctu, 2)).astype(np.int32)
    lc = orc.orc_lambda_motion_sad(50.0)
    ires, fres = t.me_frame(0, [1], pred, lc)
    from thevc_b200.capi import MeFrameCfg, ptr
    out = np.zeros((1, nctu, 593), capi.ME_PACKED_DTYPE)
    refs = (C.c_int * 1)(1)
    cfg = MeFrameCfg(64, 1, 1, 1, 1, lc)
    rc_ = t.L.tvc_me_frame_packed(t.h, 0, 1, refs, ptr(pred), C.byref(cfg), ptr(out))
    assert rc_ == 0, t.L.tvc_last_error(t.h)
    ok = ires["n_sads"] > 0
    assert ok.sum() > 5000 and (~ok).sum() > 0
    for name, src in (("mvx", ires), ("mvy", ires), ("sad", ires), ("halfx", fres), ("halfy", fres), ("qtrx", fres), ("qtry", fres), ("cost", fres)):
        assert np.array_equal(out[name][ok].astype(np.int64), src[name][ok].astype(np.int64)), name
    assert np.all(out["sad"][~ok] == 0xFFFFFFFF) and np.all(out["cost"][~ok] == 0xFFFFFFFF)


def test_me_ctu_group(ctx8_form, orc):
    """tvc_me_ctu: one (CTU, reference) census group with an explicit predictor.  The frame pre-pass run with that
    predictor for the CTU gives the same 593 results (itself oracle-checked above); predictors different from the one the
    SAD tables were centred on are served too (candidates outside the table window are evaluated from the pictures),
    checked against the oracle; and the table-less form agrees."""
    t = ctx8_form
    seq = synth.make_sequence(W, H, 3)
    cur = synth.to_hostpic(seq[2], W, H)
    refs = [synth.to_hostpic(seq[1], W, H), synth.to_hostpic(seq[0], W, H)]
    t.upload(0, cur); t.upload(1, refs[0]); t.upload(2, refs[1])
    nctu = t.ctus_x * t.ctus_y
    rng = np.random.default_rng(65)
    pred = rng.integers(-40, 41, (2, nctu, 2)).astype(np.int32)
    lc = orc.orc_lambda_motion_sad(41.3)
    ires, fres = t.me_frame(0, [1, 2], pred, lc)          # leaves the SAD tables of this picture behind
    census = t.me_census()
    for (ri, ctu) in [(0, 0), (1, 9), (0, nctu - 1), (1, t.ctus_x * (t.ctus_y - 1))]:
        gi, gf = t.me_ctu(0, ri, 1 + ri, ctu, pred[ri, ctu], lc)
        assert np.array_equal(gi, ires[ri, ctu]) and np.array_equal(gf, fres[ri, ctu]), (ri, ctu)
        gi2, gf2 = t.me_ctu(0, -1, 1 + ri, ctu, pred[ri, ctu], lc, use_tables=False)
        assert np.array_equal(gi2, gi) and np.array_equal(gf2, gf)
    # a predictor 30 pels away from the table centre: mixed table / direct evaluation
    ri, ctu = 1, 8
    p2 = (int(pred[ri, ctu, 0]) + 120, int(pred[ri, ctu, 1]) - 96)
    gi, gf = t.me_ctu(0, ri, 1 + ri, ctu, p2, lc)
    x0, y0 = (ctu % t.ctus_x) * 64, (ctu // t.ctus_x) * 64
    ref = refs[ri]
    for k in rng.choice(593, 80, replace=False):
        px, py, w, h, cux, cuy = (int(v) for v in census[k])
        x, y = x0 + px, y0 + py
        if x + w > W or y + h > H:
            assert gi[k]["n_sads"] == 0
            continue
        g = oracle.CuGeom(W, H, x0 + cux, y0 + cuy, 64)
        lx, ty, rx, by = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        orc.orc_set_search_range(C.byref(g), p2[0], p2[1], 64, C.byref(lx), C.byref(ty), C.byref(rx), C.byref(by))
        o = optr(cur.buf_y, cur.origin(0) + y * cur.stride + x)
        r = optr(ref.buf_y, ref.origin(0) + y * ref.stride + x)
        e = oracle.MeResult()
        orc.orc_tz_search(C.byref(g), o, cur.stride, r, ref.stride, w, h, lx.value, ty.value, rx.value, by.value,
                          64, 1, 0, lc, p2[0], p2[1], p2[0], p2[1], C.byref(e))
        assert (gi[k]["mvx"], gi[k]["mvy"], gi[k]["sad"], gi[k]["n_sads"]) == (e.mvx, e.mvy, e.sad, e.n_sads), k
        f = oracle.FracResult()
        orc.orc_frac_search(o, cur.stride, r, ref.stride, w, h, e.mvx, e.mvy, 1, 0, 8, lc, p2[0], p2[1], C.byref(f))
        assert (gf[k]["halfx"], gf[k]["halfy"], gf[k]["qtrx"], gf[k]["qtry"], gf[k]["cost"]) == (f.halfx, f.halfy, f.qtrx, f.qtry, f.cost), k


def test_me_ctu_async_tickets(ctx8, orc):
    """tvc_me_ctu_async / tvc_me_ctu_fetch: eight groups in flight on the side stream while synchronous calls run on the main one;
    every fetched result equals the synchronous call's, a ticket that is re-used drops its old result, fetching an idle ticket fails"""
    from thevc_b200.tlibcuda import TvcError
    t = ctx8
    seq = synth.make_sequence(W, H, 3)
    t.upload(0, synth.to_hostpic(seq[2], W, H)); t.upload(1, synth.to_hostpic(seq[1], W, H)); t.upload(2, synth.to_hostpic(seq[0], W, H))
    nctu = t.ctus_x * t.ctus_y
    rng = np.random.default_rng(67)
    lc = orc.orc_lambda_motion_sad(44.0)
    asks = [(int(rng.integers(0, 2)), int(rng.integers(0, nctu)), (int(rng.integers(-40, 41)), int(rng.integers(-40, 41)))) for _ in range(8)]
    for tk, (ri, ctu, pred) in enumerate(asks):
        t.me_ctu_async(tk, 0, ri, 1 + ri, ctu, pred, lc)
    sync = [t.me_ctu(0, ri, 1 + ri, ctu, pred, lc) for (ri, ctu, pred) in asks]        # main stream, while the tickets are in flight
    for tk in (3, 0, 7, 1, 2, 6, 5, 4):
        gi, gf = t.me_ctu_fetch(tk)
        assert np.array_equal(gi, sync[tk][0]) and np.array_equal(gf, sync[tk][1]), tk
    with pytest.raises(TvcError):
        t.me_ctu_fetch(2)                                    # nothing in flight any more
    t.me_ctu_async(5, 0, 0, 1, 0, (4, -8), lc)
    t.me_ctu_async(5, 0, 1, 2, 3, (-12, 16), lc)             # re-used before the fetch: the first result is dropped
    gi, gf = t.me_ctu_fetch(5)
    ei, ef = t.me_ctu(0, 1, 2, 3, (-12, 16), lc)
    assert np.array_equal(gi, ei) and np.array_equal(gf, ef)


# ----------------------------------------------------------------------------------- fractional ME
@pytest.mark.parametrize("bd", [8, 10])
def test_me_frac(ctx8, ctx10, orc, bd):
    t = ctx8 if bd == 8 else ctx10
    bi = bd - 8
    if bd == 8:
        seq = synth.make_sequence(W, H, 2)
        cur, ref = synth.to_hostpic(seq[1], W, H), synth.to_hostpic(seq[0], W, H)
    else:
        rng0 = np.random.default_rng(7)
        cur, ref = _pic(rng0, bd), _pic(rng0, bd)
    t.upload(0, cur); t.upload(1, ref)
    rng = np.random.default_rng(70 + bd)
    pus = _ctu_pus()
    lc = orc.orc_lambda_motion_sad(57.9)
    jobs = []
    while len(jobs) < 200:
        ctu = int(rng.integers(0, t.ctus_x * t.ctus_y))
        cx0, cy0 = (ctu % t.ctus_x) * 64, (ctu // t.ctus_x) * 64
        px, py, w, h = pus[int(rng.integers(0, len(pus)))]
        x, y = cx0 + px, cy0 + py
        if x + w > W or y + h > H:
            continue
        imvx, imvy = (int(v) for v in rng.integers(-20, 21, 2))
        if len(jobs) % 9 == 0:   # window edge: integer MV at the clipMv limit
            imvx, imvy = -(64 + 8) - cx0 + 1, -(64 + 8) - cy0 + 1
        predx, predy = (int(v) for v in rng.integers(-90, 91, 2))
        jobs.append(FracJob(1, x, y, w, h, imvx, imvy, predx, predy, lc, int(len(jobs) % 4 != 3)))
    got = t.me_frac_batch(0, jobs)
    for j, g in zip(jobs, got):
        e = oracle.FracResult()
        orc.orc_frac_search(optr(cur.buf_y, cur.origin(0) + j.y * cur.stride + j.x), cur.stride,
                            optr(ref.buf_y, ref.origin(0) + j.y * ref.stride + j.x), ref.stride, j.w, j.h,
                            j.imvx, j.imvy, j.hadamard, bi, bd, j.lambda_cost, j.predx, j.predy, C.byref(e))
        assert (g.halfx, g.halfy, g.qtrx, g.qtry, g.cost_half, g.cost) == (e.halfx, e.halfy, e.qtrx, e.qtry, e.cost_half, e.cost), \
            (j.x, j.y, j.w, j.h, j.imvx, j.imvy, j.hadamard)


def test_me_frac_bipred_target_range(ctx8, orc):
    """fractional search on an 8-bit bi-prediction target (2 * org - pred: -255 .. 510), extremes included: the two-candidates-
    per-Hadamard path keeps every coefficient inside int16 (64 * 510), and content beyond that range takes the single path"""
    t = ctx8
    rng = np.random.default_rng(91)
    from thevc_b200.tlibcuda import HostPic
    ref = _pic(rng, 8)
    for case in range(2):
        cur = HostPic(W, H)
        lo, hi = (-255, 510) if case == 0 else (-2000, 2000)
        cur.y[:] = rng.integers(lo, hi + 1, cur.y.shape).astype(np.int16)
        cur.y[0:64, 0:64] = hi                    # saturated blocks against whatever the reference holds
        cur.y[64:128, 0:64] = lo
        cur.extend_border()
        t.upload(0, cur); t.upload(1, ref)
        pus = _ctu_pus()
        lc = orc.orc_lambda_motion_sad(33.0)
        jobs = []
        for i in range(120):
            ctu = int(rng.integers(0, t.ctus_x * (t.ctus_y - 1))) if i >= 26 else (0 if i < 13 else t.ctus_x)
            cx0, cy0 = (ctu % t.ctus_x) * 64, (ctu // t.ctus_x) * 64
            px, py, w, h = pus[i % 13] if i < 26 else pus[int(rng.integers(0, len(pus)))]
            imvx, imvy = (int(v) for v in rng.integers(-12, 13, 2))
            jobs.append(FracJob(1, cx0 + px, cy0 + py, w, h, imvx, imvy, int(rng.integers(-40, 41)), int(rng.integers(-40, 41)), lc, 1))
        got = t.me_frac_batch(0, jobs)
        for j, g in zip(jobs, got):
            e = oracle.FracResult()
            orc.orc_frac_search(optr(cur.buf_y, cur.origin(0) + j.y * cur.stride + j.x), cur.stride,
                                optr(ref.buf_y, ref.origin(0) + j.y * ref.stride + j.x), ref.stride, j.w, j.h,
                                j.imvx, j.imvy, 1, 0, 8, j.lambda_cost, j.predx, j.predy, C.byref(e))
            assert (g.halfx, g.halfy, g.qtrx, g.qtry, g.cost_half, g.cost) == (e.halfx, e.halfy, e.qtrx, e.qtry, e.cost_half, e.cost), \
                (case, j.x, j.y, j.w, j.h)


# ----------------------------------------------------------------------------------- transform / quant
def _tu_list(rng, bd, n_per_size=40):
    """non-overlapping TUs over the three planes, grouped by ascending size"""
    tus = []
    off = 0
    used = [np.zeros((H, W), bool), np.zeros((H // 2, W // 2), bool), np.zeros((H // 2, W // 2), bool)]
    for log2 in (2, 3, 4, 5):
        n = 1 << log2
        cnt = 0
        while cnt < n_per_size:
            plane = int(rng.integers(0, 3)) if log2 < 5 else 0
            pw, ph = (W, H) if plane == 0 else (W // 2, H // 2)
            x, y = int(rng.integers(0, pw // n)) * n, int(rng.integers(0, ph // n)) * n
            if used[plane][y:y + n, x:x + n].any():
                continue
            used[plane][y:y + n, x:x + n] = True
            flags = 0
            if log2 == 2:
                r = cnt % 5
                flags = capi.TU_DST if (r == 1 and plane == 0) else (capi.TU_SKIP if r == 2 else 0)
            qp = int(rng.integers(0, 52))
            qp_bd = 6 * (bd - 8)
            per, rem = (qp + qp_bd) // 6, (qp + qp_bd) % 6
            base_per = max(0, per - int(rng.integers(0, 2)))
            scan_idx = int(rng.integers(0, 3)) if log2 <= 3 else 0
            tus.append(TU(plane, x, y, log2, flags, scan_idx, per, rem, base_per, off))
            off += n * n
            cnt += 1
    return tus, off


@pytest.mark.parametrize("bd", [8, 10])
def test_tq_batches(ctx8, ctx10, orc, bd):
    t = ctx8 if bd == 8 else ctx10
    bi = bd - 8
    rng = np.random.default_rng(80 + bd)
    amp = (1 << bd) - 1
    resi = synth.HostPic(W, H)
    resi.y[:] = rng.integers(-amp, amp + 1, (H, W)) // 4
    resi.u[:] = rng.integers(-amp, amp + 1, (H // 2, W // 2)) // 8
    resi.v[:] = rng.integers(-amp, amp + 1, (H // 2, W // 2))
    pred = _pic(rng, bd)
    t.upload(0, resi); t.upload(1, pred)
    tus, elems = _tu_list(rng, bd)

    def resi_ptr(pic, tu):
        st = pic.stride if tu.plane == 0 else pic.cstride
        return optr(pic.plane(tu.plane), pic.origin(tu.plane) + tu.y * st + tu.x), st

    # forward transform only
    coef = t.fwd_transform_batch(0, tus, elems)
    exp = np.zeros(elems, np.int32)
    for tu in tus:
        n = 1 << tu.log2_size
        p, st = resi_ptr(resi, tu)
        c = np.zeros(n * n, np.int32)
        if tu.flags & capi.TU_SKIP:
            orc.orc_transform_skip(p, st, c, n, n, bd)
        else:
            orc.orc_xT(1 if tu.flags & capi.TU_DST else 0, p, st, c, n, n, bi)
        exp[tu.coef_offset:tu.coef_offset + n * n] = c
    assert np.array_equal(coef, exp)

    # transform + quant (+ sign hiding, ARL)
    for (islice, sh, arl) in [(1, 1, 1), (0, 1, 0), (0, 0, 1)]:
        qc = QuantCfg(islice, sh, arl)
        lev, arlc, abs_sum = t.fwd_tq_batch(0, tus, qc, elems, want_arl=bool(arl))
        elev = np.zeros(elems, np.int32); earl = np.zeros(elems, np.int32); eabs = np.zeros(len(tus), np.uint32)
        for i, tu in enumerate(tus):
            n = 1 << tu.log2_size
            c = exp[tu.coef_offset:tu.coef_offset + n * n].copy()
            scan = np.zeros(n * n, np.uint32)
            orc.orc_scan(tu.scan_idx, tu.log2_size, scan)
            qp = oracle.QuantParam(tu.qp_per, tu.qp_rem, tu.base_per, islice, sh, arl, bd)
            q = np.zeros(n * n, np.int32); a = np.zeros(n * n, np.int32); s = C.c_uint32(0)
            orc.orc_quant(c, q, optr(a), n, n, C.byref(qp), scan, C.byref(s))
            elev[tu.coef_offset:tu.coef_offset + n * n] = q
            earl[tu.coef_offset:tu.coef_offset + n * n] = a
            eabs[i] = s.value
        assert np.array_equal(abs_sum, eabs), (islice, sh, arl)
        assert np.array_equal(lev, elev), (islice, sh, arl)
        if arl:
            assert np.array_equal(arlc, earl)

    # dequant + inverse transform + reconstruction
    levels = np.zeros(elems, np.int32)
    for tu in tus:
        n = 1 << tu.log2_size
        levels[tu.coef_offset:tu.coef_offset + n * n] = rng.integers(-300, 301, n * n) * (rng.random(n * n) < 0.3)
    levels[tus[0].coef_offset] = 40000          # clip16 of the level
    t.inv_tq_batch(2, 1, 3, tus, levels)
    got_resi, got_rec = t.download(2), t.download(3)
    for tu in tus:
        n = 1 << tu.log2_size
        dq = np.zeros(n * n, np.int32)
        orc.orc_dequant(levels[tu.coef_offset:tu.coef_offset + n * n].copy(), dq, n, n, tu.qp_per, tu.qp_rem, bd)
        r = np.zeros((n, n), np.int16)
        if tu.flags & capi.TU_SKIP:
            orc.orc_itransform_skip(dq, optr(r), n, n, n, bd)
        else:
            orc.orc_xIT(1 if tu.flags & capi.TU_DST else 0, dq, optr(r), n, n, n, bi)
        g = got_resi.plane(tu.plane)
        st = got_resi.stride if tu.plane == 0 else got_resi.cstride
        oy, ox = np.divmod(got_resi.origin(tu.plane), st)
        assert np.array_equal(g[oy + tu.y:oy + tu.y + n, ox + tu.x:ox + tu.x + n], r), (tu.plane, tu.x, tu.y, n, tu.flags)
        pp = pred.plane(tu.plane)[oy + tu.y:oy + tu.y + n, ox + tu.x:ox + tu.x + n]
        rec = np.clip(pp.astype(np.int32) + r, 0, amp).astype(np.int16)
        gr = got_rec.plane(tu.plane)[oy + tu.y:oy + tu.y + n, ox + tu.x:ox + tu.x + n]
        assert np.array_equal(gr, rec)


@pytest.mark.parametrize("bd", [8, 10])
def test_single_tu_dropins(ctx8, ctx10, orc, bd):
    t = ctx8 if bd == 8 else ctx10
    bi = bd - 8
    rng = np.random.default_rng(90 + bd)
    amp = (1 << bd) - 1
    for n in (4, 8, 16, 32):
        for dst in ((0, 1) if n == 4 else (0,)):
            for trial in range(3):
                resi = rng.integers(-amp, amp + 1, (40, 64)).astype(np.int16)
                if trial == 1:
                    resi[:] = amp
                c = t.xT(dst, resi, 64 * 3 + 2, 64, n)
                e = np.zeros(n * n, np.int32)
                orc.orc_xT(dst, optr(resi, 64 * 3 + 2), 64, e, n, n, bi)
                assert np.array_equal(c, e), ("xT", n, dst, trial)
                co = rng.integers(-70000 if trial == 2 else -3000, 70000 if trial == 2 else 3000, n * n).astype(np.int32)
                ra = np.zeros((40, 64), np.int16); rb = np.zeros((40, 64), np.int16)
                t.xIT(dst, co, ra, 65, 64, n)
                orc.orc_xIT(dst, co, optr(rb, 65), 64, n, n, bi)
                assert np.array_equal(ra, rb), ("xIT", n, dst, trial)
        for (per, rem) in [(0, 0), (3, 4), (6, 1), (8, 5)]:
            q = rng.integers(-40000, 40000, n * n).astype(np.int32)
            e = np.zeros(n * n, np.int32)
            orc.orc_dequant(q, e, n, n, per, rem, bd)
            assert np.array_equal(t.xDeQuant(q, n, per, rem), e), ("dq", n, per, rem)


# ----------------------------------------------------------------------------------- errors
def test_error_behaviour(ctx8):
    from thevc_b200 import TvcError
    with pytest.raises(TvcError):
        ctx8.dist_batch([DistJob(0, 0, 0, 0, 0, 99, 0, 0, 0, 8, 8, 0)])          # bad slot
    with pytest.raises(TvcError):
        ctx8.mc_batch(2, [PU(0, 0, 64, 64, 0, 4 * 4000, 0, -1, 0, 0)])            # unclipped MV outside the margin
    with pytest.raises(TvcError):
        ctx8.xT(0, np.zeros((8, 8), np.int16), 0, 8, 5)                           # unsupported size
    assert ctx8.launch_count() > 0


@pytest.mark.parametrize("bd", [8, 10])
def test_pred_cost_batch(ctx8, ctx10, orc, bd):
    """merge / AMVP candidate evaluation = luma motion compensation of a candidate + HAD or SAD against the original
    block (xGetInterPredictionError, TEncSearch.cpp:3059-3081; xGetTemplateCost :4057-4118).  Candidates of one PU overlap, every
    PU shape, uni / bi, fractional / integer / far-clipped MVs, 8- and 10-bit."""
    t = ctx8 if bd == 8 else ctx10
    rng = np.random.default_rng(140 + bd)
    cur = _pic(rng, bd)
    refs = {1: _pic(rng, bd), 2: _pic(rng, bd)}
    t.upload(0, cur); t.upload(1, refs[1]); t.upload(2, refs[2])
    pus = []
    for k in range(5 * len(PU_SHAPES)):
        w, h = PU_SHAPES[k % len(PU_SHAPES)]
        cx, cy = int(rng.integers(0, W // 64)) * 64, int(rng.integers(0, H // 64)) * 64          # a PU never crosses its CTU
        px, py = cx + int(rng.integers(0, (64 - w) // 4 + 1)) * 4, cy + int(rng.integers(0, (64 - h) // 4 + 1)) * 4
        for cand in range(5):                                 # five candidates of the same PU, as xMergeEstimation walks them
            mv = [int(v) for v in rng.integers(-200, 200, 4)]
            if cand == 3:
                mv = [4 * (mv[0] // 4), 4 * (mv[1] // 4), mv[2], 4 * (mv[3] // 4)]
            if cand == 4 and k % 6 == 0:
                mv = [-4000, 4000, 4000, -4000]
            m0, m1 = _clip_mv(orc, px, py, mv[0], mv[1]), _clip_mv(orc, px, py, mv[2], mv[3])
            mode = (k + cand) % 3
            pus.append(PU(px, py, w, h, 1, m0[0], m0[1], -1, 0, 0) if mode == 0 else
                       PU(px, py, w, h, -1, 0, 0, 2, m1[0], m1[1]) if mode == 1 else
                       PU(px, py, w, h, 1, m0[0], m0[1], 2, m1[0], m1[1]))
    for kind, fn in ((capi.DIST_HADS, orc.orc_hads), (capi.DIST_SAD, orc.orc_sad_generic)):
        got = t.pred_cost_batch(0, kind, pus)
        for i, pu in enumerate(pus):
            pred = np.ascontiguousarray(_oracle_mc(orc, refs, pu, bd)[0])
            exp = fn(optr(cur.buf_y, cur.origin(0) + pu.y * cur.stride + pu.x), cur.stride, optr(pred), pu.w, pu.w, pu.h, bd - 8)
            assert int(got[i]) == int(exp), (kind, i, pu.x, pu.y, pu.w, pu.h, pu.ref_slot0, pu.ref_slot1)
    with pytest.raises(Exception):
        t.pred_cost_batch(0, capi.DIST_SSE, pus[:1])


@pytest.mark.parametrize("bd", [8, 10])
def test_ctu_cost_grids_serve_every_pu(ctx8, ctx10, orc, bd):
    """the look-up form of SURVEY 8f-3: from the three prefix-sum grids of one (CTU, reference, MV) the SAD and the SATD of EVERY
    PU of the 593-PU census with that motion, equal to predicting that PU alone and running xGetSAD / xGetHADs on it (interior CTU,
    partial right / bottom CTUs, MVs clipped at all four picture edges, integer / half / quarter positions)"""
    t = ctx8 if bd == 8 else ctx10
    rng = np.random.default_rng(170 + bd)
    cur, ref = _pic(rng, bd), _pic(rng, bd)
    t.upload(0, cur); t.upload(1, ref)
    census = t.me_census()
    cases = [(64, 64, 13, -7), (128, 64, 0, 0), (192, 128, 8, 5), (384, 0, 2, 0), (384, 192, 0, 3), (0, 192, -9, 30),
             (0, 0, -4000, -4000), (384, 192, 4000, 4000), (0, 192, -4000, 4000), (384, 0, 4000, -4000), (256, 128, 7, 7)]
    jobs = np.zeros(len(cases), capi.GRID_JOB_DTYPE)
    for i, (x0, y0, mvx, mvy) in enumerate(cases):
        jobs[i] = (1, x0, y0, mvx, mvy)
    # a far MV is clipped per CU in the reference; the CTU-level clip is what a 64x64 PU would use -- the grid must hold for it and,
    # with the clamped window reads, for any MV value at all
    for i, (x0, y0, mvx, mvy) in enumerate(cases):
        if abs(mvx) > 1000:
            jobs[i]["mvx"], jobs[i]["mvy"] = _clip_mv(orc, x0, y0, mvx, mvy)
    sad4, had4, had8 = t.ctu_cost_grids(0, jobs)

    def S(I, x, y, w, h):
        return int(I[y + h, x + w]) - int(I[y, x + w]) - int(I[y + h, x]) + int(I[y, x])
    checked = 0
    for i, j in enumerate(jobs):
        x0, y0, mvx, mvy = int(j["x0"]), int(j["y0"]), int(j["mvx"]), int(j["mvy"])
        for (px, py, w, h) in census[:, :4]:
            px, py, w, h = int(px), int(py), int(w), int(h)
            if x0 + px + w > W or y0 + py + h > H:
                continue
            pred = np.zeros((h, w), np.int16)
            orc.orc_pred_inter_luma_blk(optr(ref.buf_y, ref.origin(0) + (y0 + py) * ref.stride + x0 + px), ref.stride, mvx, mvy, w, h,
                                        optr(pred), w, 0, bd)
            o = optr(cur.buf_y, cur.origin(0) + (y0 + py) * cur.stride + x0 + px)
            assert S(sad4[i], px // 4, py // 4, w // 4, h // 4) >> (bd - 8) == orc.orc_sad_generic(o, cur.stride, optr(pred), w, w, h, bd - 8)
            got = S(had8[i], px // 8, py // 8, w // 8, h // 8) if (w % 8 == 0 and h % 8 == 0) else S(had4[i], px // 4, py // 4, w // 4, h // 4)
            if w % 8 == 0 and h % 8 == 0:
                assert px % 8 == 0 and py % 8 == 0
            assert got >> (bd - 8) == orc.orc_hads(o, cur.stride, optr(pred), w, w, h, bd - 8), (i, px, py, w, h)
            checked += 1
    assert checked > 4000
