#!/usr/bin/env python
"""Golden bitstream md5s of the UNMODIFIED reference encoder (oracle/_ref/bin/TAppEncoderStatic, built from
/root/reference by oracle/Makefile) for the full-size configurations of BASELINE.json (C2 >= 16 pictures so
that four references are active, C3 >= 33 pictures so that an intra period is crossed, C3 with IDR refresh for the
intra-period shards, C4 3840x2160 internal 10 bit).  A single run of the reference takes 5-50 minutes per case on one host
core, so the md5s are generated HERE (this container has /root/reference) and committed as tests/golden/hm_md5.json; the
-m gpu tests compare the hooked / sharded encoder's stream with them on the GPU box.  Inputs are the seeded synthetic
sequences of tests/synth.py (deterministic: numpy default_rng), written as 8-bit 4:2:0 files.

usage: python tests/golden/make_hm_md5.py [case ...]      (no argument: every case, up to --jobs at a time)
"""
from __future__ import annotations

import argparse
import concurrent.futures
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

OUT = os.path.join(HERE, "hm_md5.json")
ENC_REF = os.path.join(ROOT, "oracle", "_ref", "bin", "TAppEncoderStatic")
CFG = os.path.join(ROOT, "build", "hm", "cfg")

# name -> (cfg, width, height, frames, extra encoder arguments)
CASES = {
    "ldp_1080_17": ("encoder_lowdelay_P_main.cfg", 1920, 1080, 17, []),
    "ra_1080_33": ("encoder_randomaccess_main.cfg", 1920, 1080, 33, []),
    "ra_1080_66_idr": ("encoder_randomaccess_main.cfg", 1920, 1080, 66, ["--DecodingRefreshType=2"]),
    "ra_1080_66_idr_ip16": ("encoder_randomaccess_main.cfg", 1920, 1080, 66, ["--DecodingRefreshType=2", "--IntraPeriod=16"]),
    "he10_2160_8": ("encoder_intra_he10.cfg", 3840, 2160, 8, []),
    "ldp_240_24": ("encoder_lowdelay_P_main.cfg", 416, 240, 24, []),
    "ra_240_33_idr_ip16": ("encoder_randomaccess_main.cfg", 416, 240, 33, ["--DecodingRefreshType=2", "--IntraPeriod=16"]),
    "ra_240_33": ("encoder_randomaccess_main.cfg", 416, 240, 33, []),
    "ldb_240_17": ("encoder_lowdelay_main.cfg", 416, 240, 17, []),
}


def write_yuv(path, w, h, frames, seed=None):
    import numpy as np
    import synth
    seq = synth.make_sequence(w, h, frames) if seed is None else synth.make_sequence(w, h, frames, seed=seed)
    with open(path, "wb") as f:
        for y, u, v in seq:
            f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())


def encoder_args(cfg, yuv, w, h, frames, out, extra):
    return ["-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames), "-b", out,
            "-o", os.devnull, "--SEIpictureDigest=1"] + list(extra)


def run_case(name):
    cfg, w, h, frames, extra = CASES[name]
    with tempfile.TemporaryDirectory() as d:
        yuv, out = os.path.join(d, "in.yuv"), os.path.join(d, "ref.bin")
        write_yuv(yuv, w, h, frames)
        t0 = time.perf_counter()
        r = subprocess.run([ENC_REF] + encoder_args(cfg, yuv, w, h, frames, out, extra), capture_output=True, text=True)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError("%s: reference encoder failed: %s" % (name, r.stdout[-500:] + r.stderr[-500:]))
        data = open(out, "rb").read()
        return name, {"cfg": cfg, "width": w, "height": h, "frames": frames, "extra": extra, "md5": hashlib.md5(data).hexdigest(),
                      "bytes": len(data), "reference_wall_s_this_container_1core": round(wall, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*")
    ap.add_argument("--jobs", type=int, default=6)
    a = ap.parse_args()
    names = a.cases or list(CASES)
    with concurrent.futures.ThreadPoolExecutor(a.jobs) as pool:
        for fut in concurrent.futures.as_completed([pool.submit(run_case, n) for n in names]):
            name, rec = fut.result()
            cur = json.load(open(OUT)) if os.path.exists(OUT) else {}
            cur[name] = rec
            json.dump(cur, open(OUT, "w"), indent=1, sort_keys=True)
            print(name, rec, flush=True)


if __name__ == "__main__":
    main()
