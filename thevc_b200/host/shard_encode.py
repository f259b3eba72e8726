"""Shard an ALL-INTRA encode across N processes / GPUs by frame range and concatenate the bitstreams (BASELINE
configs[3]: "all-intra frames sharded across 8 B200").

With IntraPeriod 1 every picture is coded on its own (I slices, no reference pictures, entropy state reset per slice), so
frame ranges are independent units: shard r encodes input frames [start_r, start_r + count_r) with `-fs start_r -f count_r`
and TVC_POC_OFFSET=start_r (the hooked encoder then numbers its pictures from POC start_r and, for r > 0, writes no
VPS/SPS/PPS; tlibcuda_hm.h).  No collective, no NCCL: the host concatenates the Annex-B streams, and the result is byte for
byte the stream of a single run over all frames (tests/test_shard_encode.py checks the md5 against the unmodified
reference encoder).  Inter configurations are refused: their pictures form one dependency chain per sequence.

    python -m thevc_b200.host.shard_encode --cfg build/hm/cfg/encoder_intra_main.cfg -i in.yuv -wdt 1920 -hgt 1080 \
        --frames 16 --shards 8 -o out.bin [--gpus 0,1,...] [--hm intra16] [-- extra encoder arguments]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ENC = os.path.join(ROOT, "build", "hm", "TAppEncoderCuda")


def frame_ranges(frames: int, shards: int):
    """contiguous, as equal as possible: [(start, count)] with empty shards dropped"""
    out = []
    for r in range(shards):
        a, b = frames * r // shards, frames * (r + 1) // shards
        if b > a:
            out.append((a, b - a))
    return out


def cfg_is_all_intra(path: str) -> bool:
    m = re.search(r"^\s*IntraPeriod\s*:\s*(-?\d+)", open(path, encoding="latin-1").read(), re.M)
    return bool(m) and int(m.group(1)) == 1


def shard_encode(cfg, yuv, w, h, frames, shards, out, gpus=None, hm="none", extra=(), encoder=ENC, workdir=None):
    if not cfg_is_all_intra(cfg) or any(a.startswith("--IntraPeriod") or a == "-ip" for a in extra):
        raise ValueError("frame sharding needs an all-intra configuration (IntraPeriod 1): inter pictures depend on their references")
    if not os.path.exists(encoder):
        raise RuntimeError("%s is not built (make -C thevc_b200/host hm)" % encoder)
    workdir = workdir or os.path.dirname(os.path.abspath(out))
    ranges = frame_ranges(frames, shards)
    procs = []
    t0 = time.perf_counter()
    for r, (start, count) in enumerate(ranges):
        part = os.path.join(workdir, "shard_%03d.bin" % r)
        env = dict(os.environ, TVC_POC_OFFSET=str(start), TVC_HM=hm)
        if gpus:
            env["CUDA_VISIBLE_DEVICES"] = str(gpus[r % len(gpus)])
        cmd = [encoder, "-c", cfg, "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-fs", str(start), "-f", str(count), "-b", part, "-o", os.devnull] + list(extra)
        log = open(part + ".log", "w")
        procs.append((part, log, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT, env=env)))
    for part, log, p in procs:
        rc = p.wait()
        log.close()
        if rc != 0:
            raise RuntimeError("shard %s failed (%d): %s" % (part, rc, open(part + ".log").read()[-800:]))
    wall = time.perf_counter() - t0
    md5 = hashlib.md5()
    with open(out, "wb") as fo:                      # the whole exchange step of this path: concatenation on the host
        for part, _, _ in procs:
            data = open(part, "rb").read()
            fo.write(data)
            md5.update(data)
    return {"frames": frames, "shards": len(ranges), "ranges": ranges, "wall_s": wall, "fps": frames / wall, "md5": md5.hexdigest(),
            "bytes": os.path.getsize(out)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", required=True)
    ap.add_argument("-i", "--input", required=True)
    ap.add_argument("-wdt", type=int, required=True)
    ap.add_argument("-hgt", type=int, required=True)
    ap.add_argument("--frames", type=int, required=True)
    ap.add_argument("--shards", type=int, required=True)
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--gpus", default="", help="comma list of device indices, one per shard (round robin)")
    ap.add_argument("--hm", default="none", help="TVC_HM hook list of the shards (none = the reference's own code)")
    ap.add_argument("extra", nargs="*")
    a = ap.parse_args()
    gpus = [int(g) for g in a.gpus.split(",") if g != ""] or None
    print(json.dumps(shard_encode(a.cfg, a.input, a.wdt, a.hgt, a.frames, a.shards, a.out, gpus, a.hm, a.extra)))


if __name__ == "__main__":
    sys.exit(main())
