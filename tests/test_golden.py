"""Golden vectors generated from the reference's own compiled code (tests/golden/make_golden.py ->
hm72_golden.npz).  CPU part: the oracle restatement reproduces them (this is what pins the oracle on
a box without /root/reference).  GPU part (-m gpu): the CUDA path reproduces them through the C ABI."""
import ctypes as C
import os
import sys
import zlib

import numpy as np
import pytest

import oracle
from oracle import ptr

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "hm72_golden.npz"))


def _me_pics(bd):
    cur, ref = mg.me_inputs(bd)
    crcs = [mg.crc(cur.buf_y), mg.crc(ref.buf_y), mg.crc(ref.buf_u), mg.crc(ref.buf_v)]
    if crcs != [int(v) for v in G["bd%d_me_crc" % bd]]:
        pytest.skip("seeded input regeneration differs on this numpy build; motion goldens not comparable")
    return cur, ref


@pytest.mark.parametrize("bd", [8, 10])
def test_oracle_reproduces_reference_goldens(orc, bd):
    tag = "bd%d_" % bd
    bi = bd - 8
    org, cur, org_bi = G[tag + "dist_org"], G[tag + "dist_cur"], G[tag + "dist_org_bi"]
    k = 0
    for (w, h) in mg.PU_SHAPES:
        for o in (org, org_bi):
            co = 96 * 3 + 5
            e = G[tag + "dist_out"][k]; k += 1
            assert (int(e[0]), int(e[1])) == (w, h)
            assert orc.orc_sad(ptr(o), 64, ptr(cur, co), 96, w, h, 0, bi) == e[2]
            if h > 8:
                assert orc.orc_sad(ptr(o), 64, ptr(cur, co), 96, w, h, 1, bi) == e[3]
            assert orc.orc_sse(ptr(o), 64, ptr(cur, co), 96, w, h, bi) == e[4]
            assert orc.orc_hads(ptr(o), 64, ptr(cur, co), 96, w, h, bi) == e[5]
    pel, mid, outs = G[tag + "if_pel"], G[tag + "if_mid"], G[tag + "if_out"]
    off, w, h, k = 8 * 48 + 8, 17, 9, 0
    for frac in range(4):
        for last in (0, 1):
            d = np.zeros((h, w), np.int16)
            orc.orc_filter_hor_luma(ptr(pel, off), 48, ptr(d), w, w, h, frac, last, bd)
            assert np.array_equal(d, outs[k]); k += 1
        for first in (0, 1):
            for last in (0, 1):
                d = np.zeros((h, w), np.int16)
                orc.orc_filter_ver_luma(ptr(pel if first else mid, off), 48, ptr(d), w, w, h, frac, first, last, bd)
                assert np.array_equal(d, outs[k]); k += 1
    for frac in range(8):
        for last in (0, 1):
            d = np.zeros((h, w), np.int16)
            orc.orc_filter_hor_chroma(ptr(pel, off), 48, ptr(d), w, w, h, frac, last, bd)
            assert np.array_equal(d, outs[k]); k += 1
        for first in (0, 1):
            for last in (0, 1):
                d = np.zeros((h, w), np.int16)
                orc.orc_filter_ver_chroma(ptr(pel if first else mid, off), 48, ptr(d), w, w, h, frac, first, last, bd)
                assert np.array_equal(d, outs[k]); k += 1
    for n in (4, 8, 16, 32):
        resi = G[tag + "tq_resi%d" % n]
        log2 = int(np.log2(n))
        for dst in ((0, 1) if n == 4 else (0,)):
            c = np.zeros(n * n, np.int32)
            orc.orc_xT(dst, ptr(resi), n, c, n, n, bi)
            assert np.array_equal(c, G[tag + "tq_coef%d_%d" % (n, dst)])
            r = np.zeros((n, n), np.int16)
            orc.orc_xIT(dst, c, ptr(r), n, n, n, bi)
            assert np.array_equal(r, G[tag + "tq_back%d_%d" % (n, dst)])
        c = G[tag + "tq_coef%d_0" % n]
        scan = np.zeros(n * n, np.uint32)
        orc.orc_scan(0, log2, scan)
        for qp, islice in ((22, 1), (32, 0), (37, 0)):
            per, rem = C.c_int(), C.c_int()
            orc.orc_set_qp(qp, 1, 6 * bi, 0, C.byref(per), C.byref(rem))
            qpar = oracle.QuantParam(per.value, rem.value, per.value, islice, 1, 1, bd)
            q = np.zeros(n * n, np.int32); a = np.zeros(n * n, np.int32); s = C.c_uint32(0)
            orc.orc_quant(np.ascontiguousarray(c), q, ptr(a), n, n, C.byref(qpar), scan, C.byref(s))
            assert np.array_equal(q, G[tag + "tq_lev%d_qp%d" % (n, qp)])
            assert np.array_equal(a, G[tag + "tq_arl%d_qp%d" % (n, qp)])
            assert s.value == int(G[tag + "tq_abs%d_qp%d" % (n, qp)][0])
            d = np.zeros(n * n, np.int32)
            orc.orc_dequant(q, d, n, n, per.value, rem.value, bd)
            assert np.array_equal(d, G[tag + "tq_deq%d_qp%d" % (n, qp)])
    # motion search + compensation
    cur_p, ref_p = _me_pics(bd)
    lc = orc.orc_lambda_motion_sad(mg.LAMBDA)
    for job, exp, mc in zip(G[tag + "me_jobs"], G[tag + "me_out"], G[tag + "mc_crc"]):
        x, y, w, h, cux, cuy, predx, predy, lx, ty, rx, by = (int(v) for v in job)
        g = oracle.CuGeom(mg.W, mg.H, cux, cuy, 64)
        o = ptr(cur_p.buf_y, cur_p.origin(0) + y * cur_p.stride + x)
        r = ptr(ref_p.buf_y, ref_p.origin(0) + y * ref_p.stride + x)
        e = oracle.MeResult()
        orc.orc_tz_search(C.byref(g), o, cur_p.stride, r, ref_p.stride, w, h, lx, ty, rx, by, 64, 1, bi, lc, predx, predy, predx, predy, C.byref(e))
        f = oracle.FracResult()
        orc.orc_frac_search(o, cur_p.stride, r, ref_p.stride, w, h, e.mvx, e.mvy, 1, bi, bd, lc, predx, predy, C.byref(f))
        assert (e.mvx, e.mvy, e.sad, f.halfx, f.halfy, f.qtrx, f.qtry, f.cost) == tuple(int(v) for v in exp)
        mvx, mvy = (e.mvx << 2) + (f.halfx << 1) + f.qtrx, (e.mvy << 2) + (f.halfy << 1) + f.qtry
        py_ = np.zeros((h, w), np.int16); pu_ = np.zeros((h // 2, w // 2), np.int16); pv_ = np.zeros_like(pu_)
        orc.orc_pred_inter_luma_blk(ptr(ref_p.buf_y, ref_p.origin(0) + y * ref_p.stride + x), ref_p.stride, mvx, mvy, w, h, ptr(py_), w, 0, bd)
        co = ref_p.origin(1) + (y // 2) * ref_p.cstride + x // 2
        orc.orc_pred_inter_chroma_blk(ptr(ref_p.buf_u, co), ref_p.cstride, mvx, mvy, w, h, ptr(pu_), w // 2, 0, bd)
        orc.orc_pred_inter_chroma_blk(ptr(ref_p.buf_v, co), ref_p.cstride, mvx, mvy, w, h, ptr(pv_), w // 2, 0, bd)
        assert (mg.crc(py_), mg.crc(pu_), mg.crc(pv_)) == tuple(int(v) for v in mc)


@pytest.mark.gpu
@pytest.mark.parametrize("bd", [8, 10])
def test_cuda_reproduces_reference_goldens(bd):
    from thevc_b200 import TLibCuda, capi
    from thevc_b200.capi import FracJob, MeJob, PU, QuantCfg, TU
    tag = "bd%d_" % bd
    t = TLibCuda(mg.W, mg.H, bd, num_slots=6)
    org, cur, org_bi = G[tag + "dist_org"], G[tag + "dist_cur"], G[tag + "dist_org_bi"]
    k = 0
    for (w, h) in mg.PU_SHAPES:
        for o in (org, org_bi):
            co = 96 * 3 + 5
            e = G[tag + "dist_out"][k]; k += 1
            o = np.ascontiguousarray(o); c = np.ascontiguousarray(cur)
            assert t.dist_block(capi.DIST_SAD, o, 0, 64, c, co, 96, w, h, 0) == e[2]
            if h > 8:
                assert t.dist_block(capi.DIST_SAD, o, 0, 64, c, co, 96, w, h, 1) == e[3]
            assert t.dist_block(capi.DIST_SSE, o, 0, 64, c, co, 96, w, h) == e[4]
            assert t.dist_block(capi.DIST_HADS, o, 0, 64, c, co, 96, w, h) == e[5]
    pel, mid, outs = np.ascontiguousarray(G[tag + "if_pel"]), np.ascontiguousarray(G[tag + "if_mid"]), G[tag + "if_out"]
    off, w, h, k = 8 * 48 + 8, 17, 9, 0
    for frac in range(4):
        for last in (0, 1):
            d = np.zeros((h, w), np.int16)
            t.filter_hor_luma(pel, off, 48, d, 0, w, w, h, frac, last)
            assert np.array_equal(d, outs[k]); k += 1
        for first in (0, 1):
            for last in (0, 1):
                d = np.zeros((h, w), np.int16)
                t.filter_ver_luma(pel if first else mid, off, 48, d, 0, w, w, h, frac, first, last)
                assert np.array_equal(d, outs[k]); k += 1
    for frac in range(8):
        for last in (0, 1):
            d = np.zeros((h, w), np.int16)
            t.filter_hor_chroma(pel, off, 48, d, 0, w, w, h, frac, last)
            assert np.array_equal(d, outs[k]); k += 1
        for first in (0, 1):
            for last in (0, 1):
                d = np.zeros((h, w), np.int16)
                t.filter_ver_chroma(pel if first else mid, off, 48, d, 0, w, w, h, frac, first, last)
                assert np.array_equal(d, outs[k]); k += 1
    bi = bd - 8
    from thevc_b200.tlibcuda import HostPic
    for n in (4, 8, 16, 32):
        resi = np.ascontiguousarray(G[tag + "tq_resi%d" % n])
        log2 = int(np.log2(n))
        for dst in ((0, 1) if n == 4 else (0,)):
            c = t.xT(dst, resi, 0, n, n)
            assert np.array_equal(c, G[tag + "tq_coef%d_%d" % (n, dst)])
            r = np.zeros((n, n), np.int16)
            t.xIT(dst, c, r, 0, n, n)
            assert np.array_equal(r, G[tag + "tq_back%d_%d" % (n, dst)])
        # fused transform + quant on a picture slot holding the residual at (64, 32)
        pic = HostPic(mg.W, mg.H)
        pic.y[32:32 + n, 64:64 + n] = resi
        t.upload(0, pic)
        for qp, islice in ((22, 1), (32, 0), (37, 0)):
            q = qp + 6 * bi
            tu = TU(0, 64, 32, log2, 0, 0, q // 6, q % 6, q // 6, 0)
            lev, arl, abs_sum = t.fwd_tq_batch(0, [tu], QuantCfg(islice, 1, 1), n * n, want_arl=True)
            assert np.array_equal(lev, G[tag + "tq_lev%d_qp%d" % (n, qp)])
            assert np.array_equal(arl, G[tag + "tq_arl%d_qp%d" % (n, qp)])
            assert abs_sum[0] == int(G[tag + "tq_abs%d_qp%d" % (n, qp)][0])
            assert np.array_equal(t.xDeQuant(lev, n, q // 6, q % 6), G[tag + "tq_deq%d_qp%d" % (n, qp)])
    cur_p, ref_p = _me_pics(bd)
    t.upload(0, cur_p); t.upload(1, ref_p)
    if bd == 8:
        t.me_prepass(0, [1], None)
    lc = int(np.floor(65536.0 * np.sqrt(mg.LAMBDA)))
    jobs = G[tag + "me_jobs"]
    mj = []
    for job in jobs:
        x, y, w, h, cux, cuy, predx, predy, lx, ty, rx, by = (int(v) for v in job)
        # start = clipMv(pred) >> 2 with the CU geometry (TComDataCU.cpp:3505-3517)
        hmax, hmin = (mg.W + 8 - cux - 1) << 2, (-64 - 8 - cux + 1) * 4
        vmax, vmin = (mg.H + 8 - cuy - 1) << 2, (-64 - 8 - cuy + 1) * 4
        sx, sy = min(hmax, max(hmin, predx)) >> 2, min(vmax, max(vmin, predy)) >> 2
        mj.append(MeJob(0, 1, x, y, w, h, capi.ME_TZ, 1, 64, lx, ty, rx, by, predx, predy, sx, sy, lc))
    res = t.me_search_batch(0, mj, use_tables=(bd == 8))
    fj = [FracJob(1, j.x, j.y, j.w, j.h, r.mvx, r.mvy, j.predx, j.predy, lc, 1) for j, r in zip(mj, res)]
    fres = t.me_frac_batch(0, fj)
    pus = []
    for j, r, f, exp in zip(mj, res, fres, G[tag + "me_out"]):
        assert (r.mvx, r.mvy, r.sad, f.halfx, f.halfy, f.qtrx, f.qtry, f.cost) == tuple(int(v) for v in exp)
        pus.append(PU(j.x, j.y, j.w, j.h, 1, (r.mvx << 2) + (f.halfx << 1) + f.qtrx, (r.mvy << 2) + (f.halfy << 1) + f.qtry, -1, 0, 0))
    for pu, mc in zip(pus, G[tag + "mc_crc"]):     # PUs overlap: one launch each
        t.mc_batch(2, [pu])
        p = t.download(2, with_margin=False)
        got = (mg.crc(p.y[pu.y:pu.y + pu.h, pu.x:pu.x + pu.w]), mg.crc(p.u[pu.y // 2:(pu.y + pu.h) // 2, pu.x // 2:(pu.x + pu.w) // 2]),
               mg.crc(p.v[pu.y // 2:(pu.y + pu.h) // 2, pu.x // 2:(pu.x + pu.w) // 2]))
        assert got == tuple(int(v) for v in mc)
    t.close()


# ----------------------------------------------------------------------------------- RDOQ goldens
def _rdoq_cases():
    import make_rdoq_golden as mr
    g = np.load(os.path.join(HERE, "golden", "rdoq_golden.npz"))
    ests = []
    for row in g["est"]:
        e = oracle.EstBits()
        C.memmove(C.byref(e), np.ascontiguousarray(row, np.int32).ctypes.data, C.sizeof(e))
        ests.append(e)
    cases = [dict(zip(mr.COLS, (int(v) for v in p))) for p in g["params"]]
    return g, ests, cases


def test_oracle_reproduces_rdoq_goldens(orc):
    """xRateDistOptQuant outputs of the compiled reference (rdoq_golden.npz) == the C restatement"""
    g, ests, cases = _rdoq_cases()
    for c, lam in zip(cases, g["lambdas"]):
        nn = 1 << (2 * c["log2"])
        o = c["offset"]
        scan = np.zeros(nn, np.uint32)
        orc.orc_scan(c["scan_idx"], c["log2"], scan)
        par = oracle.RdoqParam(c["log2"], c["is_luma"], c["scan_idx"], c["per"], c["rem"], c["bd"], c["cbf_ctx"], c["sign_hide"],
                               c["use_arl"], float(lam))
        q = np.zeros(nn, np.int32); a = np.zeros(nn, np.int32); s = C.c_uint32(0)
        orc.orc_rdoq(np.ascontiguousarray(g["coef"][o:o + nn]), q, ptr(a), C.byref(par), C.byref(ests[c["est"]]), scan, C.byref(s))
        assert np.array_equal(q, g["levels"][o:o + nn]) and s.value == c["abs_sum"], c
        assert np.array_equal(a, g["arl"][o:o + nn]), c


@pytest.mark.gpu
@pytest.mark.parametrize("bd", [8, 10])
def test_cuda_reproduces_rdoq_goldens(bd):
    from thevc_b200 import TLibCuda
    from thevc_b200.capi import EstBits, QuantCfg, RdoqTU
    g, ests, cases = _rdoq_cases()
    abi_est = []
    for e in ests:
        x = EstBits(); C.memmove(C.byref(x), C.byref(e), C.sizeof(x)); abi_est.append(x)
    t = TLibCuda(mg.W, mg.H, bd, num_slots=1)
    try:
        for sh in (0, 1):
            for arl in (0, 1):
                sel = [(c, float(l)) for c, l in zip(cases, g["lambdas"]) if c["bd"] == bd and c["sign_hide"] == sh and c["use_arl"] == arl]
                tus = [RdoqTU(c["log2"], c["is_luma"], c["scan_idx"], c["per"], c["rem"], c["cbf_ctx"], c["est"], c["offset"], l) for c, l in sel]
                lev, ga, sums = t.rdoq_batch(tus, abi_est, QuantCfg(0, sh, arl), g["coef"])
                for i, (c, _) in enumerate(sel):
                    nn, o = 1 << (2 * c["log2"]), c["offset"]
                    assert np.array_equal(lev[o:o + nn], g["levels"][o:o + nn]) and int(sums[i]) == c["abs_sum"], c
                    if arl:
                        assert np.array_equal(ga[o:o + nn], g["arl"][o:o + nn]), c
    finally:
        t.close()


@pytest.mark.parametrize("bd", [8, 10])
def test_oracle_reproduces_intra_rough_goldens(orc, bd):
    """tests/golden/intra_rough.npz: reference samples, original blocks and the 35 uiSad values dumped by the reference
    encoder's own estIntraPredQT (make_intra_golden.py): pins the reference-sample smoothing + the whole mode loop"""
    g = np.load(os.path.join(HERE, "golden", "intra_rough.npz"))
    key = "bd%d" % bd
    log2s, lines, orgs, sads = g[key + "_log2"], g[key + "_lines"], g[key + "_orgs"], g[key + "_sads"]
    assert set(int(v) for v in log2s) == {2, 3, 4, 5, 6}
    lo = oo = 0
    for i, log2n in enumerate(log2s):
        n = 1 << int(log2n)
        line = np.ascontiguousarray(lines[lo:lo + 4 * n + 1])
        org = np.ascontiguousarray(orgs[oo:oo + n * n])
        lo += 4 * n + 1; oo += n * n
        sad = np.zeros(35, np.uint32)
        orc.orc_intra_rough(ptr(line), ptr(org), n, int(log2n), 1, 1, bd, ptr(sad), None)
        assert np.array_equal(sad, sads[i]), (i, int(log2n))
    assert lo == len(lines) and oo == len(orgs)
