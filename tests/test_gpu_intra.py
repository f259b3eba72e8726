"""GPU parity of the intra 35-mode rough search: tvc_intra_rough / tvc_intra_rough_batch through the C ABI
against the CPU oracle (hm_oracle_intra.c, itself pinned against the reference's compiled predIntraLumaAng + calcHAD by
tests/test_oracle_vs_ref.py and against vectors dumped from the running reference encoder, tests/golden/intra_rough.npz).
Bit-exact: SATDs and every prediction sample."""
import os

import numpy as np
import pytest

import oracle
import intra_cases
from thevc_b200 import TLibCuda, capi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "intra_rough.npz")


@pytest.fixture(scope="module")
def orc():
    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="module", params=[8, 10])
def ctx(request):
    t = TLibCuda(416, 240, request.param, num_slots=2)
    t.bd = request.param
    yield t
    t.close()


@pytest.mark.parametrize("log2n", [2, 3, 4, 5, 6])
def test_intra_rough_single(ctx, orc, log2n):
    rng = np.random.default_rng(100 + log2n + ctx.bd)
    for kind in intra_cases.KINDS:
        line, org = intra_cases.make_case(rng, log2n, ctx.bd, kind)
        sad, preds = ctx.intra_rough(log2n, line, org, want_preds=True)
        osad, opreds = intra_cases.oracle_rough(orc, line, org, log2n, ctx.bd, want_preds=True)
        bad = [m for m in range(35) if not np.array_equal(preds[m], opreds[m])]
        assert not bad, (kind, "prediction differs for modes", bad)
        assert np.array_equal(sad, osad), (kind, sad, osad)


def test_intra_rough_availability_flags(ctx, orc):
    """bAbove / bLeft only steer the DC value and its edge filter (TComPrediction.cpp:127-165,361)"""
    rng = np.random.default_rng(7)
    for log2n in (2, 4):
        line, org = intra_cases.make_case(rng, log2n, ctx.bd, "random")
        for above, left in ((1, 0), (0, 1), (0, 0)):
            sad, preds = ctx.intra_rough(log2n, line, org, above=bool(above), left=bool(left), want_preds=True)
            osad, opreds = intra_cases.oracle_rough(orc, line, org, log2n, ctx.bd, above, left, want_preds=True)
            assert np.array_equal(preds, opreds) and np.array_equal(sad, osad)


def test_intra_rough_batch_mixed(ctx, orc):
    """one call, 400 PUs of mixed sizes; original blocks addressed inside a picture-like plane with its stride"""
    rng = np.random.default_rng(11 + ctx.bd)
    mx = (1 << ctx.bd) - 1
    stride, rows = 200, 136
    plane = rng.integers(0, mx + 1, (rows, stride)).astype(np.int16)
    jobs = np.zeros(400, capi.INTRA_JOB_DTYPE)
    lines = []
    off = 0
    for i in range(len(jobs)):
        log2n = int(rng.integers(2, 7))
        n = 1 << log2n
        x, y = int(rng.integers(0, stride - n + 1)), int(rng.integers(0, rows - n + 1))
        line, _ = intra_cases.make_case(rng, log2n, ctx.bd, intra_cases.KINDS[i % len(intra_cases.KINDS)])
        jobs[i] = (log2n, off, y * stride + x, stride, 1, 1)
        lines.append(line)
        off += len(line)
    lines = np.concatenate(lines)
    sad = ctx.intra_rough_batch(jobs, lines, plane)
    for i, j in enumerate(jobs):
        n = 1 << int(j["log2_size"])
        y, x = divmod(int(j["org_offset"]), stride)
        org = np.ascontiguousarray(plane[y:y + n, x:x + n])
        line = lines[j["line_offset"]:j["line_offset"] + 4 * n + 1]
        assert np.array_equal(sad[i], intra_cases.oracle_rough(orc, line, org, int(j["log2_size"]), ctx.bd)), i


def test_intra_rough_bad_arguments(ctx):
    jobs = np.zeros(1, capi.INTRA_JOB_DTYPE)
    jobs[0] = (7, 0, 0, 64, 1, 1)
    with pytest.raises(RuntimeError):
        ctx.intra_rough_batch(jobs, np.zeros(600, np.int16), np.zeros(64 * 64, np.int16))
    jobs[0] = (3, 0, 0, 8, 1, 1)
    with pytest.raises(RuntimeError):       # line shorter than 4N+1
        ctx.intra_rough_batch(jobs, np.zeros(32, np.int16), np.zeros(64, np.int16))


def test_intra_rough_golden_from_reference_encoder(ctx):
    """vectors dumped by the reference's own estIntraPredQT (m_piYuvExt after initAdiPattern, the original block and
    the uiSad of every mode) while encoding the synthetic 416x240 sequence"""
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/intra_rough.npz absent")
    g = np.load(GOLD)
    key = "bd%d" % ctx.bd
    if key + "_log2" not in g:
        pytest.skip("no vectors for this bit depth")
    log2s, lines, orgs, sads = g[key + "_log2"], g[key + "_lines"], g[key + "_orgs"], g[key + "_sads"]
    lo = oo = 0
    for i, log2n in enumerate(log2s):
        n = 1 << int(log2n)
        line, org = lines[lo:lo + 4 * n + 1], np.ascontiguousarray(orgs[oo:oo + n * n].reshape(n, n))
        lo += 4 * n + 1; oo += n * n
        assert np.array_equal(ctx.intra_rough(int(log2n), line, org), sads[i]), (i, log2n)
