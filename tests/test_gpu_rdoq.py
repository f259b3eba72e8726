"""GPU parity of RDOQ: k_rdoq through the C ABI (tvc_rdoq_batch, tvc_xRateDistOptQuant) against the CPU oracle's
orc_rdoq (pinned against the compiled reference's xRateDistOptQuant by tests/test_oracle_vs_ref.py::test_rdoq and
tests/golden/rdoq_*.npz).  Bit-exact levels, uiAbsSum and ARL coefficients."""
import ctypes as C

import numpy as np
import pytest

import oracle
import rdoq_cases as rc
from thevc_b200 import TLibCuda
from thevc_b200.capi import EstBits, QuantCfg, RdoqTU

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    oracle.build()
    return oracle.lib()


def _to_abi_est(e: "oracle.EstBits") -> EstBits:
    out = EstBits()
    C.memmove(C.byref(out), C.byref(e), C.sizeof(out))
    return out


def _build_batch(rng, orc, bd, n_tus, sizes=(2, 3, 4, 5), kinds=range(6)):
    qp_bd_offset = 6 * (bd - 8)
    n_est = 5
    ests = [rc.make_est(rng) for _ in range(n_est)]
    tus, coefs, meta = [], [], []
    off = 0
    for i in range(n_tus):
        log2 = int(rng.choice(sizes))
        n = 1 << log2
        is_luma = 1 if log2 == 5 else int(rng.integers(0, 2))
        qp = int(rng.choice([4, 17, 22, 27, 32, 37, 44, 51]))
        per, rem = C.c_int(), C.c_int()
        orc.orc_set_qp(qp, is_luma, qp_bd_offset, 0, C.byref(per), C.byref(rem))
        icu = int(rng.integers(0, 2))
        ldir = int(rng.choice([0, 1, 10, 26, 34, 8, 30]))
        tr_idx = int(rng.integers(0, 3))
        scan_idx = rc.scan_idx_for(icu, is_luma, n, ldir)
        lam = rc.lambda_for(qp) * float(rng.uniform(0.4, 2.5))
        kind = int(rng.choice(list(kinds)))
        coef = rc.make_coef(rng, log2, per.value, bd, kind)
        ei = int(rng.integers(0, n_est))
        tus.append(RdoqTU(log2, is_luma, scan_idx, per.value, rem.value, rc.cbf_ctx_for(icu, is_luma, tr_idx), ei, off, lam))
        coefs.append(coef)
        meta.append((log2, is_luma, scan_idx, per.value, rem.value, rc.cbf_ctx_for(icu, is_luma, tr_idx), ei, off, lam, kind))
        off += n * n
    return ests, tus, np.concatenate(coefs), meta


def _oracle_batch(orc, bd, ests, meta, coef, sign_hide, use_arl):
    lev = np.zeros(coef.size, np.int32)
    arl = np.zeros(coef.size, np.int32)
    sums = np.zeros(len(meta), np.uint32)
    for i, (log2, is_luma, scan_idx, per, rem, cbf, ei, off, lam, kind) in enumerate(meta):
        nn = 1 << (2 * log2)
        scan = np.zeros(nn, np.uint32)
        orc.orc_scan(scan_idx, log2, scan)
        par = oracle.RdoqParam(log2, is_luma, scan_idx, per, rem, bd, cbf, sign_hide, use_arl, lam)
        q = np.zeros(nn, np.int32); a = np.zeros(nn, np.int32); s = C.c_uint32(0)
        orc.orc_rdoq(np.ascontiguousarray(coef[off:off + nn]), q, oracle.ptr(a), C.byref(par), C.byref(ests[ei]), scan, C.byref(s))
        lev[off:off + nn] = q; arl[off:off + nn] = a; sums[i] = s.value
    return lev, arl, sums


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("sign_hide,use_arl", [(1, 1), (0, 0), (1, 0)])
def test_rdoq_batch(orc, bd, sign_hide, use_arl):
    rng = np.random.default_rng(900 + bd + 3 * sign_hide + use_arl)
    t = TLibCuda(416, 240, bd, num_slots=1)
    try:
        ests, tus, coef, meta = _build_batch(rng, orc, bd, 700)
        lev, arl, sums = t.rdoq_batch(tus, [_to_abi_est(e) for e in ests], QuantCfg(0, sign_hide, use_arl), coef)
        elev, earl, esums = _oracle_batch(orc, bd, ests, meta, coef, sign_hide, use_arl)
        bad = [i for i, m in enumerate(meta) if not np.array_equal(lev[m[7]:m[7] + (1 << (2 * m[0]))], elev[m[7]:m[7] + (1 << (2 * m[0]))])]
        assert not bad, ("levels differ for TUs", bad[:10], [meta[i] for i in bad[:3]])
        assert np.array_equal(sums, esums)
        if use_arl:
            assert np.array_equal(arl, earl)
        assert int(np.count_nonzero(elev)) > 10000           # the batch is not degenerate
    finally:
        t.close()


def test_rdoq_edge_cases(orc):
    """all-zero TUs, a single coefficient at every scan position class, saturating inputs, the largest batch shapes"""
    rng = np.random.default_rng(77)
    t = TLibCuda(416, 240, 8, num_slots=1)
    try:
        est = rc.make_est(rng)
        for log2 in (2, 3, 4, 5):
            n = 1 << log2
            per, rem = 5, 2
            step = rc.quant_step(log2, per, 8)
            cases = [np.zeros(n * n, np.int32)]
            for pos in (0, 1, n - 1, n * n - 1, n * (n - 1), (n * n) // 2 + 3):
                c = np.zeros(n * n, np.int32); c[pos] = int(1.2 * step); cases.append(c)
                c = np.zeros(n * n, np.int32); c[pos] = -int(7.7 * step); cases.append(c)
            cases.append(np.full(n * n, 32767, np.int32))
            cases.append(np.full(n * n, -32768, np.int32))
            for scan_idx in ((0, 1, 2) if log2 <= 3 else (0,)):
                for cbf in (-1, 0, 1, 5, 7):
                    for c in cases:
                        scan = np.zeros(n * n, np.uint32)
                        orc.orc_scan(scan_idx, log2, scan)
                        par = oracle.RdoqParam(log2, int(cbf < 5), scan_idx, per, rem, 8, cbf, 1, 1, 37.5)
                        q = np.zeros(n * n, np.int32); a = np.zeros(n * n, np.int32); s = C.c_uint32(0)
                        orc.orc_rdoq(c, q, oracle.ptr(a), C.byref(par), C.byref(est), scan, C.byref(s))
                        gq, ga, gs = t.xRateDistOptQuant(c, n, int(cbf < 5), scan_idx, per, rem, cbf, 1, 1, 37.5, _to_abi_est(est))
                        assert np.array_equal(gq, q) and np.array_equal(ga, a) and gs == s.value, (log2, scan_idx, cbf)
        # argument errors are reported, not executed
        bad = RdoqTU(6, 1, 0, 5, 2, 0, 0, 0, 10.0)
        with pytest.raises(Exception):
            t.rdoq_batch([bad], [_to_abi_est(est)], QuantCfg(0, 1, 0), np.zeros(4096, np.int32))
        bad = RdoqTU(4, 1, 1, 5, 2, 0, 0, 0, 10.0)       # hor scan does not exist for 16x16
        with pytest.raises(Exception):
            t.rdoq_batch([bad], [_to_abi_est(est)], QuantCfg(0, 1, 0), np.zeros(256, np.int32))
    finally:
        t.close()


def test_rdoq_frame_sized_properties(orc):
    """1080p worth of 32x32 luma TUs (2040) + spot checks against the oracle; size-independent properties on all:
    |level| never exceeds the plain-rounding level, uiAbsSum equals the sum of |levels| before sign hiding,
    and with sign hiding every hidden-sign subset has the parity of its first coefficient's sign."""
    rng = np.random.default_rng(5)
    bd = 8
    t = TLibCuda(416, 240, bd, num_slots=1)
    try:
        ests, tus, coef, meta = _build_batch(rng, orc, bd, 2040, sizes=(5,), kinds=(0, 1))
        lev0, _, sums0 = t.rdoq_batch(tus, [_to_abi_est(e) for e in ests], QuantCfg(0, 0, 0), coef)
        lev1, _, sums1 = t.rdoq_batch(tus, [_to_abi_est(e) for e in ests], QuantCfg(0, 1, 0), coef)
        assert np.array_equal(sums0, sums1)               # uiAbsSum is taken before sign hiding
        scan = np.zeros(1024, np.uint32)
        orc.orc_scan(0, 5, scan)
        qs = np.array([26214, 23302, 20560, 18396, 16384, 14564], np.int64)
        for i, m in enumerate(meta):
            off, per, rem = m[7], m[3], m[4]
            c = coef[off:off + 1024].astype(np.int64)
            qbits = 14 + per + 15 - bd - 5
            plain = (np.abs(c) * qs[rem] + (1 << (qbits - 1))) >> qbits
            a0 = np.abs(lev0[off:off + 1024].astype(np.int64))
            assert np.all(a0 <= plain) and int(a0.sum()) == int(sums0[i])
            assert np.all(np.sign(lev0[off:off + 1024]) * np.sign(c) >= 0)
            l1 = lev1[off:off + 1024][scan].reshape(64, 16)
            for sub in range(64):
                nz = np.nonzero(l1[sub])[0]
                if nz.size and nz[-1] - nz[0] >= 4:
                    assert (int(np.abs(l1[sub]).sum()) & 1) == (0 if l1[sub][nz[0]] > 0 else 1), (i, sub)
        idx = rng.choice(len(meta), 60, replace=False)
        elev, _, esums = _oracle_batch(orc, bd, ests, [meta[i] for i in idx], coef, 1, 0)
        for j, i in enumerate(idx):
            off = meta[i][7]
            assert np.array_equal(lev1[off:off + 1024], elev[off:off + 1024]) and sums1[i] == esums[j]
    finally:
        t.close()


@pytest.mark.parametrize("bd", [8, 10])
def test_fwd_rdoq_batch_on_picture(orc, bd):
    """transformNxN with RDOQ on: residual plane -> k_fwd_tq (transform only) -> k_rdoq, coefficients never leave the
    device; against the oracle's xT + xRateDistOptQuant over the same TU list (all three planes, every TU size)."""
    import synth
    from thevc_b200.capi import TU
    from thevc_b200.tlibcuda import HostPic
    rng = np.random.default_rng(31 + bd)
    W, H = 416, 240
    t = TLibCuda(W, H, bd, num_slots=2)
    try:
        pic = HostPic(W, H)
        amp = (1 << bd) // 8
        for pl in range(3):
            p = pic.plane(pl)
            p[:] = np.clip(np.rint(rng.laplace(0, amp, p.shape)), -(1 << bd) + 1, (1 << bd) - 1).astype(np.int16)
        t.upload(0, pic)
        qp = 30
        est = rc.make_est(rng)
        lam_l, lam_c = rc.lambda_for(qp), rc.lambda_for(qp) * 0.8
        tus, rtus, rows = [], [], []
        off = 0
        for log2 in (2, 3, 4, 5):
            n = 1 << log2
            for pl in (0, 1, 2):
                if pl and log2 == 5:
                    continue
                pw, ph = (W, H) if pl == 0 else (W // 2, H // 2)
                for k in range(40):
                    x = int(rng.integers(0, (pw - n) // n + 1)) * n
                    y = int(rng.integers(0, (ph - n) // n + 1)) * n
                    q = qp + 6 * (bd - 8)
                    tus.append(TU(pl, x, y, log2, 0, 0, q // 6, q % 6, q // 6, off))
                    rtus.append(RdoqTU(log2, int(pl == 0), 0, q // 6, q % 6, -1 if pl == 0 else 5, 0, off, lam_l if pl == 0 else lam_c))
                    rows.append((pl, x, y, log2, 0, 0, q // 6, q % 6, q // 6, off))
                    off += n * n
        lev, _, sums = t.fwd_rdoq_batch(0, tus, rtus, [_to_abi_est(est)], QuantCfg(0, 1, 0), off)
        tu_arr = np.array(rows, np.int32)
        elev = np.zeros(off, np.int32); esum = np.zeros(len(rows), np.uint32)
        tri = (C.c_void_p * 3)(*[oracle.ptr(pic.plane(pl), pic.origin(pl)).value for pl in range(3)])
        orc.orc_fwd_rdoq_batch(tri, pic.stride, pic.cstride, len(rows), oracle.ptr(tu_arr), 1, bd, C.byref(est), lam_l, lam_c,
                               oracle.ptr(elev), oracle.ptr(esum))
        assert np.array_equal(lev, elev) and np.array_equal(sums, esum)
        assert np.count_nonzero(elev) > 2000
    finally:
        t.close()


def test_fwd_rdoq_recon_round_trip(orc):
    """tvc_fwd_rdoq_recon_batch == tvc_fwd_rdoq_batch followed by tvc_inv_tq_batch on the returned levels"""
    from thevc_b200.capi import TU
    from thevc_b200.tlibcuda import HostPic
    rng = np.random.default_rng(3)
    W, H, bd = 416, 240, 8
    t = TLibCuda(W, H, bd, num_slots=6)
    try:
        resi, pred = HostPic(W, H), HostPic(W, H)
        for pl in range(3):
            resi.plane(pl)[:] = np.clip(np.rint(rng.laplace(0, 20, resi.plane(pl).shape)), -255, 255).astype(np.int16)
            pred.plane(pl)[:] = rng.integers(0, 256, pred.plane(pl).shape).astype(np.int16)
        t.upload(0, resi); t.upload(1, pred)
        est = rc.make_est(rng)
        tus, rtus = [], []
        off = 0
        for log2 in (2, 3, 4, 5):
            n = 1 << log2
            for pl in (0, 1, 2):
                if pl and log2 == 5:
                    continue
                pw, ph = (W, H) if pl == 0 else (W // 2, H // 2)
                for y in range(0, ph - n + 1, 2 * n):
                    for x in range(0, pw - n + 1, 2 * n):
                        tus.append(TU(pl, x, y, log2, 0, 0, 5, 2, 5, off))
                        rtus.append(RdoqTU(log2, int(pl == 0), 0, 5, 2, -1 if pl == 0 else 5, 0, off, 40.0))
                        off += n * n
        qc = QuantCfg(0, 1, 0)
        lev, _, sums = t.fwd_rdoq_batch(0, tus, rtus, [_to_abi_est(est)], qc, off)
        t.inv_tq_batch(2, 1, 3, tus, lev)
        a_resi, a_rec = t.download(2, with_margin=False), t.download(3, with_margin=False)
        lev2 = np.zeros(off, np.int32); sums2 = np.zeros(len(tus), np.uint32)
        from thevc_b200.tlibcuda import _arr
        from thevc_b200.capi import ptr
        ta, ra = _arr(TU, tus), _arr(RdoqTU, rtus)
        e = _to_abi_est(est)
        rcode = t.L.tvc_fwd_rdoq_recon_batch(t.h, 0, 4, 1, 5, len(tus), C.cast(ta, C.c_void_p), C.cast(ra, C.c_void_p), 1, C.byref(e), C.byref(qc),
                                             ptr(lev2), off, ptr(sums2))
        assert rcode == 0, t.L.tvc_last_error(t.h)
        b_resi, b_rec = t.download(4, with_margin=False), t.download(5, with_margin=False)
        assert np.array_equal(lev, lev2) and np.array_equal(sums, sums2)
        for tu in tus:
            n = 1 << tu.log2_size
            for a, b in ((a_resi, b_resi), (a_rec, b_rec)):
                pa, pb = (a.y, a.u, a.v)[tu.plane], (b.y, b.u, b.v)[tu.plane]
                assert np.array_equal(pa[tu.y:tu.y + n, tu.x:tu.x + n], pb[tu.y:tu.y + n, tu.x:tu.x + n])
        assert np.count_nonzero(lev) > 1000
        # the 16-bit form: same levels as int16, same sums, same reconstruction
        lev16 = np.zeros(off, np.int16); sums3 = np.zeros(len(tus), np.uint32)
        rcode = t.L.tvc_fwd_rdoq_recon_batch16(t.h, 0, 4, 1, 5, len(tus), C.cast(ta, C.c_void_p), C.cast(ra, C.c_void_p), 1, C.byref(e), C.byref(qc),
                                               ptr(lev16), off, ptr(sums3))
        assert rcode == 0, t.L.tvc_last_error(t.h)
        assert np.array_equal(lev16.astype(np.int32), lev) and np.array_equal(sums3, sums)
        c_rec = t.download(5, with_margin=False)
        assert np.array_equal(c_rec.y, b_rec.y) and np.array_equal(c_rec.u, b_rec.u)
    finally:
        t.close()


def test_fwd_rdoq_list_validation_threaded(orc):
    """a picture-sized pair of TU lists (65 536 records: the fused, threaded check of tvc_fwd_rdoq_batch) is accepted when valid and
    rejected -- with the record named -- when one late record is out of range, when the two lists disagree, and when the list is not
    grouped by ascending size; nothing runs on a rejected list (levels stay untouched)"""
    import copy
    from thevc_b200.capi import TU
    from thevc_b200.tlibcuda import TvcError
    W, H = 416, 240
    t = TLibCuda(W, H, 8, num_slots=2)
    try:
        est = rc.make_est(np.random.default_rng(3))
        n = 65536
        tus = [TU(0, (i % 100) * 4, ((i // 100) % 59) * 4, 2, 0, 0, 5, 0, 5, 16 * i) for i in range(n)]
        rtus = [RdoqTU(2, 1, 0, 5, 0, -1, 0, 16 * i, 20.0) for i in range(n)]
        lev, _, _ = t.fwd_rdoq_batch(0, tus, rtus, [_to_abi_est(est)], QuantCfg(0, 1, 0), 16 * n)
        assert lev.shape == (16 * n,)
        for what, mutate, needle in (
                ("late record out of range", lambda a, b: setattr(a[n - 5], "x", W + 400), "TU %d" % (n - 5)),
                ("lists disagree", lambda a, b: setattr(b[n // 2 + 1], "coef_offset", 0), "TU %d" % (n // 2 + 1)),
                ("not ascending", lambda a, b: (setattr(a[10], "log2_size", 3), setattr(b[10], "log2_size", 3)), "ascending")):
            a, b = copy.deepcopy(tus), copy.deepcopy(rtus)
            mutate(a, b)
            with pytest.raises(TvcError) as e:
                t.fwd_rdoq_batch(0, a, b, [_to_abi_est(est)], QuantCfg(0, 1, 0), 16 * n)
            assert needle in str(e.value), (what, str(e.value))
    finally:
        t.close()
