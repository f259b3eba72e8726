"""Sharding by independent units: frame ranges of all-intra encodes (BASELINE configs[3]) and closed intra periods of
random-access encodes with IDR refresh (configs[2]).  N processes encode disjoint frame ranges, the host
concatenates the Annex-B streams -> byte for byte the single-run stream of the UNMODIFIED reference encoder, and the
reference decoder accepts it with every picture hash (OK).  CPU part: the shards run the reference's own code
(TVC_HM=none; only the POC-offset patch of thevc_b200/host/patch_hm.py is exercised) -- this is the world-size-N host
logic of the path.  GPU part (-m gpu): the shards run with the device hooks on."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import synth
from thevc_b200.host import shard_encode as se

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC_REF = os.path.join(ROOT, "oracle", "_ref", "bin", "TAppEncoderStatic")
DEC_REF = os.path.join(ROOT, "oracle", "_ref", "bin", "TAppDecoderStatic")
CFG = os.path.join(ROOT, "build", "hm", "cfg")


def _need():
    for p in (se.ENC, ENC_REF, DEC_REF):
        if not os.path.exists(p):
            pytest.skip("%s not built (needs /root/reference at build time)" % os.path.relpath(p, ROOT))


def _setup(tmp_path, cfg, w, h, frames, extra=()):
    yuv = str(tmp_path / "in.yuv")
    with open(yuv, "wb") as f:
        for y, u, v in synth.make_sequence(w, h, frames):
            f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
    ref = str(tmp_path / "ref.bin")
    subprocess.run([ENC_REF, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "-b", ref,
                    "-o", os.devnull, "--SEIpictureDigest=1"] + list(extra), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
    return yuv, hashlib.md5(open(ref, "rb").read()).hexdigest()


def test_frame_ranges():
    assert se.frame_ranges(8, 4) == [(0, 2), (2, 2), (4, 2), (6, 2)]
    assert se.frame_ranges(5, 3) == [(0, 1), (1, 2), (3, 2)]
    assert se.frame_ranges(2, 8) == [(0, 1), (1, 1)]                         # more shards than frames: empty ones dropped
    for frames, shards in ((17, 8), (3, 2), (64, 8)):
        r = se.frame_ranges(frames, shards)
        assert r[0][0] == 0 and sum(c for _, c in r) == frames and all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(len(r) - 1))


def test_inter_configurations_are_refused(tmp_path):
    _need()
    with pytest.raises(ValueError):
        se.shard_encode(os.path.join(CFG, "encoder_lowdelay_P_main.cfg"), "none.yuv", 64, 64, 4, 2, str(tmp_path / "o.bin"))
    with pytest.raises(ValueError):         # low delay with periodic IDRs: the pictures of the IDR's GOP are coded BEFORE it and belong to the previous unit
        se.plan_ranges(os.path.join(CFG, "encoder_lowdelay_P_main.cfg"), 64, 2, ("--IntraPeriod=32", "--DecodingRefreshType=2"))
    with pytest.raises(ValueError):         # random access as the cfg ships it: CRA, open GOP -> leading pictures cross the intra period
        se.shard_encode(os.path.join(CFG, "encoder_randomaccess_main.cfg"), "none.yuv", 64, 64, 64, 2, str(tmp_path / "o.bin"))


def test_intra_period_ranges():
    # IntraPeriod 16, GOPSize 8: IDRs at 16, 32; a unit starts at the IDR's first leading picture (P - 7)
    assert se.intra_period_ranges(34, 3, 16, 8) == [(0, 9), (9, 16), (25, 9)]
    assert se.intra_period_ranges(34, 2, 16, 8) == [(0, 9), (9, 25)]
    assert se.intra_period_ranges(34, 8, 16, 8) == [(0, 9), (9, 16), (25, 9)]          # more shards than units
    assert se.intra_period_ranges(32, 4, 16, 8) == [(0, 9), (9, 23)]                   # POC 32 is not coded: no third unit
    assert se.intra_period_ranges(30, 4, 16, 8) == [(0, 9), (9, 21)]                   # trailing partial GOP stays with its unit
    assert se.intra_period_ranges(100, 4, 32, 8) == [(0, 25), (25, 32), (57, 32), (89, 11)]
    cfg = os.path.join(CFG, "encoder_randomaccess_main.cfg")
    if os.path.exists(cfg):
        assert se.cfg_value(cfg, (), "IntraPeriod", "ip") == 32 and se.cfg_value(cfg, (), "GOPSize", "g") == 8
        assert se.cfg_value(cfg, ("--IntraPeriod=16",), "IntraPeriod", "ip") == 16 and se.cfg_value(cfg, ("-ip", "48"), "IntraPeriod", "ip") == 48
        assert se.plan_ranges(cfg, 100, 4, ("--DecodingRefreshType=2",)) == [(0, 25), (25, 32), (57, 32), (89, 11)]


def test_closed_gop_intra_periods_shard_to_the_single_run_stream_cpu(tmp_path):
    """BASELINE configs[2]: random access with IDR refresh; 34 pictures = units [0,9) [9,25) [25,34): the middle shard starts on
    leading pictures (POC 9..15 are coded after IDR 16), the last one ends in a partial GOP"""
    _need()
    cfg, w, h, frames = "encoder_randomaccess_main.cfg", 128, 64, 34
    extra = ["--IntraPeriod=16", "--DecodingRefreshType=2"]
    yuv, ref_md5 = _setup(tmp_path, cfg, w, h, frames, extra)
    for shards in (3, 2):
        out = str(tmp_path / ("out%d.bin" % shards))
        r = se.shard_encode(os.path.join(CFG, cfg), yuv, w, h, frames, shards, out, hm="none", extra=["--SEIpictureDigest=1"] + extra)
        assert r["md5"] == ref_md5, (shards, r)
    assert r["ranges"] == [(0, 9), (9, 25)]
    log = open(str(tmp_path / "shard_001.bin.log")).read()
    pocs = [int(ln.split()[1]) for ln in log.splitlines() if ln.startswith("POC")]
    assert pocs[:8] == [16, 12, 10, 9, 11, 14, 13, 15] and sorted(pocs) == list(range(9, 34)), pocs
    # no decoder leg here: the reference's OWN decoder stops (SIGSEGV) at the first leading picture behind an IDR of the reference
    # encoder's own single-run stream (its POC restarts at the IDR), so the bar for this option is the encoder's bitstream md5


@pytest.mark.parametrize("cfg", ["encoder_intra_main.cfg", "encoder_intra_he10.cfg"])
def test_sharded_stream_equals_single_run_cpu(tmp_path, cfg):
    _need()
    w, h, frames = 208, 120, 5
    yuv, ref_md5 = _setup(tmp_path, cfg, w, h, frames)
    for shards in (2, 3, 8):
        out = str(tmp_path / ("out%d.bin" % shards))
        r = se.shard_encode(os.path.join(CFG, cfg), yuv, w, h, frames, shards, out, hm="none", extra=["--SEIpictureDigest=1"])
        assert r["md5"] == ref_md5, (cfg, shards, r)
    d = subprocess.run([DEC_REF, "-b", out], capture_output=True, text=True, timeout=600)
    assert d.returncode == 0 and "ERROR" not in d.stdout and d.stdout.count("(OK)") == frames


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["encoder_intra_main.cfg", "encoder_intra_he10.cfg"])
def test_sharded_stream_equals_single_run_gpu(tmp_path, cfg):
    """the same with device hooks on in every shard (the intra rough search of every PU size, deblocking, SAO apply)"""
    _need()
    w, h, frames = 208, 120, 3
    yuv, ref_md5 = _setup(tmp_path, cfg, w, h, frames)
    out = str(tmp_path / "out.bin")
    r = se.shard_encode(os.path.join(CFG, cfg), yuv, w, h, frames, 2, out, gpus=[0], hm="intra4,dbk,sao", extra=["--SEIpictureDigest=1"])
    assert r["md5"] == ref_md5, r
    log = open(str(tmp_path / "shard_001.bin.log")).read()
    assert "TLibCuda intra rough search:" in log and "POC    1" in log and "POC    2" in log and "POC    0" not in log


def test_stateful_extra_arguments_are_refused():
    for bad in (["--RateControl=1"], ["-f", "3"], ["-fs", "2"], ["--FrameSkip=4"], ["--TargetBitrate=1000"]):
        with pytest.raises(ValueError):
            se.check_extra(bad)
    se.check_extra(["--SEIpictureDigest=1", "--DecodingRefreshType=2", "--IntraPeriod=16", "--RateControl=0"])
    assert se.shards_per_gpu(1920, 1080, "me,frac,tables") == 16 and se.shards_per_gpu(3840, 2160, "intra16,dbk,sao") == 16
    os.environ["TVC_ME_FUSED"] = "0"          # the round-1 form reserves 34.8 GB of SAD tables per 1080p process
    try:
        assert se.shards_per_gpu(1920, 1080, "me,frac,tables") == 4 and se.shards_per_gpu(416, 240, "me,frac,tables") >= 16
    finally:
        del os.environ["TVC_ME_FUSED"]


@pytest.mark.gpu
def test_intra_period_shards_with_device_hooks_gpu(tmp_path):
    """BASELINE configs[2] at the CPU-runnable size with the ME hooks on in every shard: 33 pictures, IntraPeriod 16, IDR refresh =
    units [0,9) [9,25) [25,33); the concatenation has the md5 of the unmodified reference's single run"""
    _need()
    import json
    cfg, w, h, frames = "encoder_randomaccess_main.cfg", 416, 240, 33
    extra = ["--DecodingRefreshType=2", "--IntraPeriod=16"]
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "hm_md5.json"))).get("ra_240_33_idr_ip16")
    if g:
        yuv = str(tmp_path / "in.yuv")
        with open(yuv, "wb") as f:
            for y, u, v in synth.make_sequence(w, h, frames):
                f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
        ref_md5 = g["md5"]
    else:
        yuv, ref_md5 = _setup(tmp_path, cfg, w, h, frames, extra)
    out = str(tmp_path / "out.bin")
    r = se.shard_encode(os.path.join(CFG, cfg), yuv, w, h, frames, 3, out, hm="me,frac,tables,candgrid,dbk,sao", extra=["--SEIpictureDigest=1"] + extra)
    assert r["md5"] == ref_md5, r
    assert r["ranges"] == [(0, 9), (9, 16), (25, 8)]
    log = open(str(tmp_path / "shard_001.bin.log")).read()
    assert "TLibCuda look-up:" in log and " 0 xTZSearch" not in log
