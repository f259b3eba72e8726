"""End-to-end drop-in check (north star): the reference encoder with the TLibCuda hooks
(build/hm/TAppEncoderCuda: xTZSearch, xPatternSearchFracDIF, xT, xIT, xDeQuant, xRateDistOptQuant served by the CUDA
library, SAD tables on) must write the SAME BITSTREAM as the unmodified reference encoder
(oracle/_ref/bin/TAppEncoderStatic), and the reference decoder must accept it with matching picture
hashes.  Both binaries are built from /root/reference by committed recipes (thevc_b200/host/Makefile,
oracle/Makefile) and travel to the GPU box as build outputs; the test skips where they are absent."""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC_CUDA = os.path.join(ROOT, "build", "hm", "TAppEncoderCuda")
ENC_REF = os.path.join(ROOT, "oracle", "_ref", "bin", "TAppEncoderStatic")
DEC_REF = os.path.join(ROOT, "oracle", "_ref", "bin", "TAppDecoderStatic")
DEC_CUDA = os.path.join(ROOT, "build", "hm", "TAppDecoderCuda")
CFG = os.path.join(ROOT, "build", "hm", "cfg")

pytestmark = pytest.mark.gpu


def _need():
    for p in (ENC_CUDA, ENC_REF, DEC_REF, DEC_CUDA):
        if not os.path.exists(p):
            pytest.skip("%s not built (needs /root/reference at build time)" % os.path.relpath(p, ROOT))


def _yuv(path, w, h, n):
    seq = synth.make_sequence(w, h, n)
    with open(path, "wb") as f:
        for y, u, v in seq:
            f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())


def _encode(binary, cfg, yuv, w, h, n, out, env=None, extra=()):
    cmd = [binary, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(n), "-b", out,
           "-o", os.devnull, "--SEIpictureDigest=1"] + list(extra)
    e = dict(os.environ)
    if env:
        e.update(env)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=e)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return r


_POOL = concurrent.futures.ThreadPoolExecutor(max_workers=4)


def _encode_ref_bg(*args, **kw):
    """the unmodified reference encoder runs on a host core while the hooked encoder runs in the test's own thread
    (independent processes, independent outputs); .result() re-raises its assertion"""
    return _POOL.submit(_encode, ENC_REF, *args, **kw)


def _run_bg(cmd, env=None, timeout=1500):
    return _POOL.submit(subprocess.run, cmd, capture_output=True, text=True, timeout=timeout, env=env)


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("cfg,frames,extra", [
    ("encoder_lowdelay_P_main.cfg", 3, ()),                 # C2 at the CPU-runnable size: ME + frac + transforms on the GPU
    ("encoder_intra_main.cfg", 2, ()),                      # C1: transforms / dequant only
    ("encoder_lowdelay_P_main.cfg", 2, ("--RDOQ=0",)),      # non-RDOQ quantiser path of the host around the GPU transforms
    ("encoder_randomaccess_main.cfg", 5, ()),               # C3 (B slices): uni-pred searches on the GPU, bi-pred refinement on the reference path
    ("encoder_intra_he10.cfg", 2, ()),                      # C4: 10-bit internal (bitIncrement 2) transforms / dequant
    ("encoder_lowdelay_main.cfg", 2, ()),                   # low-delay B with 10-bit off: generalised B pictures
])
def test_bitstream_md5_identical_to_reference(tmp_path, cfg, frames, extra):
    _need()
    w, h = 416, 240
    yuv = str(tmp_path / "in.yuv")
    _yuv(yuv, w, h, frames)
    ref_bin, cuda_bin = str(tmp_path / "ref.bin"), str(tmp_path / "cuda.bin")
    ref_job = _encode_ref_bg(cfg, yuv, w, h, frames, ref_bin, extra=extra)
    r = _encode(ENC_CUDA, cfg, yuv, w, h, frames, cuda_bin, env={"TVC_HM": "me,frac,tq,rdoq,mc,tables,bipred,verify"}, extra=extra)
    ref_job.result()
    if "randomaccess" in cfg or cfg == "encoder_lowdelay_main.cfg":
        # B slices: every bi-prediction refinement (xPatternSearch on 2 * org - pred + its xPatternSearchFracDIF) ran on the device
        bl = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda bi-prediction refinement:")]
        assert bl, r.stderr[-800:]
        print(bl[-1])
        f = bl[-1].split()
        assert int(f[3]) > 100 and int(f[6]) == int(f[3]) and int(f[13]) == 0, bl[-1]
    served = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda:")]
    assert served and "kernel launches" in served[-1], r.stderr[-500:]
    print(served[-1])
    if "lowdelay" in cfg:
        assert " 0 xTZSearch" not in served[-1]              # the hooks really ran
    assert " 0 xT," not in served[-1]
    if "--RDOQ=0" not in extra:
        assert " 0 xRateDistOptQuant" not in served[-1]      # RDOQ=1 in every cfg: the quantiser decisions came from k_rdoq
    assert os.path.getsize(ref_bin) > 1000
    assert _md5(cuda_bin) == _md5(ref_bin)
    # the reference decoder accepts the stream and every picture hash matches
    # the three decoder runs are independent processes reading the same stream: started together
    d_job = _run_bg([DEC_REF, "-b", cuda_bin, "-o", str(tmp_path / "dec.yuv")], timeout=600)
    dc_job = _run_bg([DEC_CUDA, "-b", cuda_bin, "-o", str(tmp_path / "dec_cuda.yuv")], env=dict(os.environ, TVC_HM="tq,mc"))
    db_job = _run_bg([DEC_CUDA, "-b", cuda_bin, "-o", str(tmp_path / "dec_batch.yuv")], env=dict(os.environ, TVC_HM="tq,mc,batch,dbk,sao"))
    d = d_job.result()
    assert d.returncode == 0
    assert "ERROR" not in d.stdout and d.stdout.count("(OK)") >= frames
    # C5: the reference decoder with the hooks (xIT, xDeQuant, xPredInterUni on the GPU) reconstructs the same pictures
    dc = dc_job.result()
    assert dc.returncode == 0, dc.stdout[-800:] + dc.stderr[-800:]
    assert "ERROR" not in dc.stdout and dc.stdout.count("(OK)") >= frames
    dserved = [ln for ln in dc.stderr.splitlines() if ln.startswith("TLibCuda:")]
    assert dserved and " 0 xIT" not in dserved[-1], dc.stderr[-500:]
    print("decoder", dserved[-1])
    assert _md5(str(tmp_path / "dec_cuda.yuv")) == _md5(str(tmp_path / "dec.yuv"))
    # C5 as BASELINE.json words it: BATCHED dequant + inverse transform and MC -- inter CUs deferred to one tvc_mc_batch + one
    # tvc_inv_tq_batch per flush (before intra CUs and before the in-loop filters)
    db = db_job.result()                                      # + deblocking and SAO apply on the device
    assert db.returncode == 0, db.stdout[-800:] + db.stderr[-800:]
    assert "ERROR" not in db.stdout and db.stdout.count("(OK)") >= frames
    bl = [ln for ln in db.stderr.splitlines() if ln.startswith("TLibCuda picture batch:")]
    assert bl, db.stderr[-500:]
    print("decoder", bl[-1])
    if "intra" not in cfg:
        assert int(bl[-1].split()[3]) > 100, bl[-1]          # inter CUs really went through the batch
    assert _md5(str(tmp_path / "dec_batch.yuv")) == _md5(str(tmp_path / "dec.yuv"))
    dl = [ln for ln in db.stderr.splitlines() if ln.startswith("TLibCuda deblocking:")]
    assert dl and int(dl[-1].split()[2]) >= frames, db.stderr[-500:]
    print("decoder", dl[-1])
    sl = [ln for ln in db.stderr.splitlines() if ln.startswith("TLibCuda SAO:")]
    if sl:
        print("decoder", sl[-1])


def test_census_lookup_serves_the_cu_loop(tmp_path):
    """The fast configuration: integer + fractional ME of the CU loop served by look-up from census-wide
    (CTU, reference, predictor) batches (tvc_me_ctu), everything else on the host's own code.  Same bitstream as the
    unmodified reference; most searches are look-ups."""
    _need()
    w, h, frames = 416, 240, 4
    yuv = str(tmp_path / "in.yuv")
    _yuv(yuv, w, h, frames)
    ref_bin, cuda_bin = str(tmp_path / "ref.bin"), str(tmp_path / "cuda.bin")
    ref_job = _encode_ref_bg("encoder_lowdelay_P_main.cfg", yuv, w, h, frames, ref_bin)
    # + deblocking on the device: the picture digest SEI inside the bitstream hashes the final reconstruction, so an equal
    # bitstream md5 also proves the device-filtered pictures equal the reference's
    r = _encode(ENC_CUDA, "encoder_lowdelay_P_main.cfg", yuv, w, h, frames, cuda_bin, env={"TVC_HM": "me,frac,tables,dbk,sao"})
    ref_job.result()
    assert _md5(cuda_bin) == _md5(ref_bin)
    dl = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda deblocking:")]
    assert dl and int(dl[-1].split()[2]) == frames, r.stderr[-600:]
    print(dl[-1])
    sl = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda SAO:")]
    assert sl and int(sl[-1].split()[2]) >= frames, r.stderr[-600:]
    print(sl[-1])
    line = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda look-up:")]
    assert line, r.stderr[-600:]
    print(line[-1])
    f = line[-1].split()
    served, total = int(f[2]), int(f[4])
    assert total > 10000 and served > 0.8 * total, line[-1]


@pytest.mark.parametrize("cfg,frames,w,h", [
    ("encoder_intra_main.cfg", 2, 416, 240),          # C1: every PU is intra; 8-bit
    ("encoder_intra_he10.cfg", 1, 416, 240),          # C4 at the CPU-runnable size: 10-bit internal
    ("encoder_lowdelay_P_main.cfg", 2, 208, 120),     # intra PUs tested inside P pictures
])
def test_intra_rough_search_on_device(tmp_path, cfg, frames, w, h):
    """SURVEY 8f-2 inside the real encoder: the 35-mode rough search of estIntraPredQT served by tvc_intra_rough for
    EVERY PU size (intra4: also the 4x4 / 8x8 PUs the fast configuration leaves to the host).  A single differing SATD
    changes a candidate list and with it the bitstream, so an equal md5 pins all of them."""
    _need()
    yuv = str(tmp_path / "in.yuv")
    _yuv(yuv, w, h, frames)
    ref_bin, cuda_bin = str(tmp_path / "ref.bin"), str(tmp_path / "cuda.bin")
    ref_job = _encode_ref_bg(cfg, yuv, w, h, frames, ref_bin)
    r = _encode(ENC_CUDA, cfg, yuv, w, h, frames, cuda_bin, env={"TVC_HM": "intra4"})
    ref_job.result()
    assert _md5(cuda_bin) == _md5(ref_bin)
    il = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda intra rough search:")]
    assert il, r.stderr[-600:]
    print(il[-1])
    assert int(il[-1].split()[4]) > 500 and " 0 smaller PUs" in il[-1], il[-1]


@pytest.mark.parametrize("mode", ["cand", "candgrid"])
@pytest.mark.parametrize("cfg,frames", [
    ("encoder_lowdelay_P_main.cfg", 3),        # P slices: uni-predicted merge candidates, AMVP templates of 4 references
    ("encoder_randomaccess_main.cfg", 5),      # B slices: bi-predicted candidates (addAvg), identical-motion reduction
])
def test_candidate_evaluation_on_device(tmp_path, cfg, frames, mode):
    """SURVEY 8f-3 inside the real encoder: every xMergeEstimation candidate set and every xGetTemplateCost call served by
    tvc_pred_cost_batch; the merge / AMVP decisions, hence the bitstream, are the reference's."""
    _need()
    w, h = 208, 120
    yuv = str(tmp_path / "in.yuv")
    _yuv(yuv, w, h, frames)
    ref_bin, cuda_bin = str(tmp_path / "ref.bin"), str(tmp_path / "cuda.bin")
    ref_job = _encode_ref_bg(cfg, yuv, w, h, frames, ref_bin)
    # cand: one device call per xMergeEstimation / xGetTemplateCost; candgrid: look-up in CTU-wide (CTU, reference, MV) cost grids
    r = _encode(ENC_CUDA, cfg, yuv, w, h, frames, cuda_bin, env={"TVC_HM": mode})
    ref_job.result()
    assert _md5(cuda_bin) == _md5(ref_bin)
    if mode == "candgrid":
        gl = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda candidate look-up:")]
        assert gl, r.stderr[-600:]
        print(gl[-1])
        assert int(gl[-1].split()[3]) > 5 * int(gl[-1].split()[10]), gl[-1]       # mostly look-ups
    cl = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda candidate evaluation:")]
    assert cl, r.stderr[-600:]
    print(cl[-1])
    f = cl[-1].split()
    # candgrid leaves bi-predicted merge sets (most of a B slice's) to the reference's code: only the AMVP templates are certain to be many
    assert (int(f[3]) > 100 or mode == "candgrid") and int(f[9]) > 1000, cl[-1]


@pytest.mark.parametrize("digest", [1, 2, 3])
def test_picture_hash_on_device(tmp_path, digest):
    """SURVEY 8f-4 inside the real encoder and decoder: the decoded-picture-hash SEI (MD5 / CRC / checksum) computed by tvc_pic_hash.
    The SEI payload is part of the bitstream, so an equal md5 proves the encoder's digests; the hooked decoder recomputes them on the
    device and must report (OK) for every picture."""
    _need()
    w, h, frames = 208, 120, 3
    yuv = str(tmp_path / "in.yuv")
    _yuv(yuv, w, h, frames)
    ref_bin, cuda_bin = str(tmp_path / "ref.bin"), str(tmp_path / "cuda.bin")
    ex = ("--SEIpictureDigest=%d" % digest,)
    ref_job = _encode_ref_bg("encoder_lowdelay_P_main.cfg", yuv, w, h, frames, ref_bin, extra=ex)
    r = _encode(ENC_CUDA, "encoder_lowdelay_P_main.cfg", yuv, w, h, frames, cuda_bin, env={"TVC_HM": "hash"}, extra=ex)
    ref_job.result()
    assert _md5(cuda_bin) == _md5(ref_bin)
    assert "TLibCuda picture hash: %d pictures" % frames in r.stderr, r.stderr[-400:]
    d = subprocess.run([DEC_CUDA, "-b", cuda_bin], capture_output=True, text=True, timeout=600, env=dict(os.environ, TVC_HM="hash"))
    assert d.returncode == 0 and "ERROR" not in d.stdout and d.stdout.count("(OK)") == frames, d.stdout[-600:]
    assert "TLibCuda picture hash: %d pictures" % frames in d.stderr, d.stderr[-400:]


def test_psnr_sums_on_device(tmp_path):
    """SURVEY 8f-4, second half: the three UInt64 sums of squared differences of TEncGOP::xCalculateAddPSNR from tvc_pic_ssd
    (TVC_HM=psnr).  The encoder prints the PSNR of every picture: the hooked encoder's lines equal the reference's digit for digit."""
    _need()
    import re
    w, h, frames = 208, 120, 3
    yuv = str(tmp_path / "in.yuv")
    _yuv(yuv, w, h, frames)
    ref_bin, cuda_bin = str(tmp_path / "ref.bin"), str(tmp_path / "cuda.bin")
    ref_job = _encode_ref_bg("encoder_lowdelay_P_main.cfg", yuv, w, h, frames, ref_bin)
    r = _encode(ENC_CUDA, "encoder_lowdelay_P_main.cfg", yuv, w, h, frames, cuda_bin, env={"TVC_HM": "psnr"})
    ref = ref_job.result()
    assert _md5(cuda_bin) == _md5(ref_bin)

    def psnr(out):
        return re.findall(r"\[Y\s+[\d.]+ dB\s+U\s+[\d.]+ dB\s+V\s+[\d.]+ dB\]", out)
    a, b = psnr(r.stdout), psnr(ref.stdout)
    assert len(a) == frames and a == b, (a, b)
    assert "TLibCuda PSNR: the squared-difference sums of %d pictures" % frames in r.stderr, r.stderr[-400:]


def test_frame_prepass_feeds_the_cu_loop(tmp_path):
    """the north star's per-frame batched pre-pass inside the real encoder (TVC_HM=...,frame): from the second inter picture on one
    tvc_me_frame call per picture searches the whole census with the predictor guesses; a (CTU, reference) group whose first real
    AMVP predictor equals its guess takes its 593 integer + fractional results from that call, the others fall back to tvc_me_ctu.
    Same bitstream as the unmodified reference."""
    _need()
    w, h, frames = 416, 240, 5
    yuv = str(tmp_path / "in.yuv")
    _yuv(yuv, w, h, frames)
    ref_bin, cuda_bin = str(tmp_path / "ref.bin"), str(tmp_path / "cuda.bin")
    ref_job = _encode_ref_bg("encoder_lowdelay_P_main.cfg", yuv, w, h, frames, ref_bin)
    r = _encode(ENC_CUDA, "encoder_lowdelay_P_main.cfg", yuv, w, h, frames, cuda_bin, env={"TVC_HM": "me,frac,tables,frame,candgrid"})
    ref_job.result()
    assert _md5(cuda_bin) == _md5(ref_bin)
    fl = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda frame pre-pass:")]
    assert fl, r.stderr[-600:]
    print(fl[-1])
    assert int(fl[-1].split()[3]) > 20, fl[-1]


@pytest.mark.parametrize("case,hm", [
    ("ldp_240_24", "me,frac,tables,frame,candgrid,verify"),      # 24 P pictures: the encoder recycles its picture buffers from POC ~11 on
    ("ra_240_33", "me,frac,tables,candgrid,dbk,sao,verify"),     # random access as the cfg ships it: crosses the CRA at POC 32
    ("ldb_240_17", "me,frac,tables,frame,candgrid"),             # low-delay B: two lists over the same pictures
])
def test_long_sequences_with_recycled_picture_buffers(tmp_path, case, hm):
    """TEncTop::xGetNewPicBuffer reuses the oldest TComPic once GOPSize + maxDecPicBuffering + 2 pictures exist; a device slot that
    still mapped the recycled buffer used to resolve reference pointers to the OLD picture (round-1 ADVICE, high).  Sequences long
    enough to recycle, device slots beyond the first 21 uploads, the frame pre-pass across many pictures; `verify` re-runs every
    look-up as a single search.  md5 against the unmodified reference's single run (tests/golden/hm_md5.json, or the reference
    itself when it is here)."""
    _need()
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_hm_md5 as gold
    cfg, w, h, frames, extra = gold.CASES[case]
    yuv = str(tmp_path / "in.yuv")
    _yuv(yuv, w, h, frames)
    cuda_bin = str(tmp_path / "cuda.bin")
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "hm_md5.json"))).get(case)
    ref_job = None if g else _encode_ref_bg(cfg, yuv, w, h, frames, str(tmp_path / "ref.bin"), extra=extra)
    r = _encode(ENC_CUDA, cfg, yuv, w, h, frames, cuda_bin, env={"TVC_HM": hm}, extra=extra)
    if ref_job:
        ref_job.result()
    want = g["md5"] if g else _md5(str(tmp_path / "ref.bin"))
    assert _md5(cuda_bin) == want, [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda")]
    line = [ln for ln in r.stderr.splitlines() if ln.startswith("TLibCuda look-up:")]
    assert line, r.stderr[-600:]
    print(line[-1])
    f = line[-1].split()
    assert int(f[2]) > 0.8 * int(f[4]), line[-1]
