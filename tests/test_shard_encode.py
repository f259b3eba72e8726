"""Frame sharding of all-intra encodes (BASELINE configs[3]): N processes encode disjoint frame ranges, the host
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


def _setup(tmp_path, cfg, w, h, frames):
    yuv = str(tmp_path / "in.yuv")
    with open(yuv, "wb") as f:
        for y, u, v in synth.make_sequence(w, h, frames):
            f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
    ref = str(tmp_path / "ref.bin")
    subprocess.run([ENC_REF, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "-b", ref,
                    "-o", os.devnull, "--SEIpictureDigest=1"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=900)
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
