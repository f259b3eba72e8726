"""BASELINE.json's configuration matrix at FULL size, on hardware, with the device hooks on (-m gpu):

  C2  cfg/encoder_lowdelay_P_main.cfg   1920x1080, 17 pictures (four references active from POC 4 on; picture buffers recycled)
  C3  cfg/encoder_randomaccess_main.cfg 1920x1080, 33 pictures (crosses the intra period: CRA at POC 32 with its leading pictures)
  configs[2]  the same cfg with --DecodingRefreshType=2, 66 pictures = three closed intra periods sharded over the visible GPUs
  configs[3]  cfg/encoder_intra_he10.cfg 3840x2160 internal 10 bit, 8 pictures = 8 frame shards over the visible GPUs

Every stream must have the md5 of the UNMODIFIED reference encoder's single run.  Those single runs take 5-25 minutes each on one
host core, so their md5s are committed fixtures (tests/golden/hm_md5.json, written by tests/golden/make_hm_md5.py in the container that
has /root/reference); the inputs are the seeded synthetic sequences of tests/synth.py.  The encodes are independent processes: a
session fixture starts all of them together (the group search needs no SAD tables, so they all fit in HBM side by side) and each test
waits for its own.
Every run leaves a JSON record (md5, wall time, fps, hook counters) under gpurun_out/matrix/ -> copied to profiles/.

TVC_SKIP_FULLSIZE=1 skips the whole file (bounded smoke runs)."""
import concurrent.futures
import hashlib
import json
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_hm_md5 as gold  # noqa: E402
from thevc_b200.host import shard_encode as se  # noqa: E402

pytestmark = pytest.mark.gpu

ENC_CUDA = os.path.join(ROOT, "build", "hm", "TAppEncoderCuda")
DEC_REF = os.path.join(ROOT, "oracle", "_ref", "bin", "TAppDecoderStatic")
OUT_DIR = os.path.join(ROOT, "gpurun_out", "matrix")
FAST_HM = "me,frac,tables,frame,candgrid,dbk,sao"


def _golden():
    p = os.path.join(ROOT, "tests", "golden", "hm_md5.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def _hook_lines(text):
    return [ln for ln in text.splitlines() if ln.startswith("TLibCuda")]


def _single(case, workdir, hm):
    """one hooked encoder process over the whole sequence"""
    cfg, w, h, frames, extra = gold.CASES[case]
    os.makedirs(workdir, exist_ok=True)
    yuv, out = os.path.join(workdir, "in.yuv"), os.path.join(workdir, "cuda.bin")
    gold.write_yuv(yuv, w, h, frames)
    t0 = time.perf_counter()
    r = subprocess.run([ENC_CUDA] + gold.encoder_args(cfg, yuv, w, h, frames, out, extra), capture_output=True, text=True,
                       env=dict(os.environ, TVC_HM=hm), timeout=3000)
    wall = time.perf_counter() - t0
    os.remove(yuv)
    rec = {"case": case, "cfg": cfg, "width": w, "height": h, "frames": frames, "extra": extra, "hm": hm, "mode": "single process, 1 GPU",
           "rc": r.returncode, "wall_s": round(wall, 2), "fps": round(frames / wall, 4), "hooks": _hook_lines(r.stderr),
           "picture_seconds": [int(ln.split("[ET")[1].split("]")[0]) for ln in r.stdout.splitlines() if ln.startswith("POC") and "[ET" in ln]}
    if r.returncode == 0:
        rec["md5"] = hashlib.md5(open(out, "rb").read()).hexdigest()
        rec["bytes"] = os.path.getsize(out)
    else:
        rec["error"] = (r.stdout[-600:] + r.stderr[-600:])
    return rec


def _sharded(case, workdir, hm, shards, decode=False):
    cfg, w, h, frames, extra = gold.CASES[case]
    os.makedirs(workdir, exist_ok=True)
    yuv, out = os.path.join(workdir, "in.yuv"), os.path.join(workdir, "out.bin")
    gold.write_yuv(yuv, w, h, frames)
    gpus = se.visible_gpus()
    rec = {"case": case, "cfg": cfg, "width": w, "height": h, "frames": frames, "extra": extra, "hm": hm,
           "mode": "%d shards over %d GPU(s), host concatenation, no collective" % (shards, len(gpus))}
    try:
        r = se.shard_encode(os.path.join(gold.CFG, cfg), yuv, w, h, frames, shards, out, gpus=gpus or None, hm=hm,
                            extra=["--SEIpictureDigest=1"] + list(extra), workdir=workdir)
        rec.update(r)
        rec["rc"] = 0
        logs = [open(os.path.join(workdir, "shard_%03d.bin.log" % i)).read() for i in range(r["shards"])]
        rec["hooks"] = [_hook_lines(t)[-1] if _hook_lines(t) else "" for t in logs]
        if decode:
            d = subprocess.run([DEC_REF, "-b", out], capture_output=True, text=True, timeout=1500)
            rec["decoder_rc"] = d.returncode
            rec["decoder_ok_pictures"] = d.stdout.count("(OK)")
            rec["decoder_errors"] = d.stdout.count("ERROR")
    except Exception as ex:     # recorded; the test asserts on rc
        rec["rc"] = 1
        rec["error"] = repr(ex)[-1200:]
    finally:
        if os.path.exists(yuv):
            os.remove(yuv)
    return rec


@pytest.fixture(scope="session")
def matrix(tmp_path_factory):
    if os.environ.get("TVC_SKIP_FULLSIZE") == "1":
        pytest.skip("TVC_SKIP_FULLSIZE=1")
    for p in (ENC_CUDA, DEC_REF):
        if not os.path.exists(p):
            pytest.skip("%s not built (needs /root/reference at build time)" % os.path.relpath(p, ROOT))
    base = tmp_path_factory.mktemp("matrix")
    os.makedirs(OUT_DIR, exist_ok=True)
    ngpu = max(1, len(se.visible_gpus()))
    pool = concurrent.futures.ThreadPoolExecutor(8)
    t0 = time.perf_counter()
    # TVC_MATRIX_CASES=a,b,...: only these cases run (multi-GPU boxes are charged per GPU: the sharded cases alone there)
    only = [c for c in os.environ.get("TVC_MATRIX_CASES", "").split(",") if c]
    plan = {
        "ldp_1080_17": lambda: _single("ldp_1080_17", str(base / "c2"), FAST_HM),
        "ra_1080_33": lambda: _single("ra_1080_33", str(base / "c3"), FAST_HM),
        "he10_2160_8": lambda: _sharded("he10_2160_8", str(base / "c4"), "intra16,dbk,sao", 8, True),
        "ra_1080_66_idr": lambda: _sharded("ra_1080_66_idr", str(base / "c3s"), FAST_HM, max(3, ngpu)),
        "ra_1080_66_idr_ip16": lambda: _sharded("ra_1080_66_idr_ip16", str(base / "c3t"), FAST_HM, max(5, ngpu)),
    }
    default = ["ldp_1080_17", "ra_1080_33", "he10_2160_8", "ra_1080_66_idr"]
    jobs = {name: pool.submit(plan[name]) for name in (only or default)}
    out = {}

    def get(name):
        if name not in jobs:
            pytest.skip("%s not in TVC_MATRIX_CASES" % name)
        if name not in out:
            rec = jobs[name].result()
            rec["gpus_visible"] = ngpu
            rec["golden_md5"] = _golden().get(name, {}).get("md5")
            rec["golden_reference_wall_s_1core"] = _golden().get(name, {}).get("reference_wall_s_this_container_1core")
            rec["matrix_elapsed_s"] = round(time.perf_counter() - t0, 1)
            json.dump(rec, open(os.path.join(OUT_DIR, "%s_n%d.json" % (name, ngpu)), "w"), indent=1)
            out[name] = rec
        return out[name]
    yield get
    pool.shutdown(wait=False, cancel_futures=True)


def _check(rec):
    assert rec["rc"] == 0, rec.get("error")
    assert rec["golden_md5"], "tests/golden/hm_md5.json has no md5 for %s (run tests/golden/make_hm_md5.py where /root/reference exists)" % rec["case"]
    assert rec["md5"] == rec["golden_md5"], rec
    print(json.dumps({k: rec[k] for k in ("case", "mode", "hm", "wall_s", "fps", "md5") if k in rec}))


def test_c2_lowdelay_p_1080p_17_pictures(matrix):
    rec = matrix("ldp_1080_17")
    _check(rec)
    look = [ln for ln in rec["hooks"] if ln.startswith("TLibCuda look-up:")]
    assert look, rec["hooks"]
    f = look[-1].split()
    assert int(f[2]) > 0.95 * int(f[4]) > 1e6, look[-1]        # the CU loop's searches were served by the device batches


def test_c3_random_access_1080p_33_pictures(matrix):
    rec = matrix("ra_1080_33")
    _check(rec)
    assert any(ln.startswith("TLibCuda deblocking: 33 pictures") for ln in rec["hooks"]), rec["hooks"]


def test_configs2_random_access_1080p_intra_period_shards(matrix):
    rec = matrix("ra_1080_66_idr")
    _check(rec)
    assert rec["ranges"][0] == (0, 25) or rec["ranges"][0] == [0, 25], rec["ranges"]
    assert all("xTZSearch" in ln and " 0 xTZSearch" not in ln for ln in rec["hooks"]), rec["hooks"]


def test_configs3_intra_he10_2160p_frame_shards(matrix):
    rec = matrix("he10_2160_8")
    _check(rec)
    assert rec["shards"] == 8
    assert rec["decoder_rc"] == 0 and rec["decoder_errors"] == 0 and rec["decoder_ok_pictures"] == 8, rec


def test_configs2_random_access_1080p_intra_period_16_shards(matrix):
    """the same sequence with --IntraPeriod=16: five closed intra periods, enough units for five GPUs (runs when asked for with
    TVC_MATRIX_CASES; the cfg's own IntraPeriod 32 gives three units over 66 pictures)"""
    if "ra_1080_66_idr_ip16" not in os.environ.get("TVC_MATRIX_CASES", ""):
        pytest.skip("ra_1080_66_idr_ip16 only with TVC_MATRIX_CASES")
    rec = matrix("ra_1080_66_idr_ip16")
    _check(rec)
    assert rec["shards"] == 5


def test_bench_workload_has_the_encoders_tz_work(matrix):
    """round-1 VERDICT weak #1: bench.py's step must give the TZ search the work the real encoder's searches have on the same kind
    of content.  The hooked 1080p LDP encode above reports its mean candidates per served xTZSearch (the reference's own count);
    the step of bench.py (same synthetic sequence generator, its predictor guesses) must be within a factor of two of it."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from thevc_b200 import TLibCuda
    rec = matrix("ldp_1080_17")
    tz = [ln for ln in rec["hooks"] if ln.startswith("TLibCuda TZ work:")]
    assert tz, rec["hooks"]
    enc_mean = float(tz[-1].split()[3])
    wl = bench.Workload(20261018, pinned=False)
    t = TLibCuda(bench.W, bench.H, 8, num_slots=6)
    try:
        for s_, p in enumerate(wl.pics):
            t.upload(s_, p)
        lc = int(np.floor(65536.0 * np.sqrt(bench.LAMBDA)))
        ires, _ = t.me_frame(0, [1, 2, 3, 4], wl.pred, lc, do_frac=False)
    finally:
        t.close()
    n = ires["n_sads"]
    step_mean = float(n[n > 0].mean())
    print("TZ candidates per search: bench step %.1f, hooked encoder %.1f" % (step_mean, enc_mean))
    assert 0.5 <= step_mean / enc_mean <= 2.0, (step_mean, enc_mean)
