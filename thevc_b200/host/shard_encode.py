"""Shard an encode across N processes / GPUs by independent units and concatenate the bitstreams.

* ALL-INTRA (BASELINE configs[3]: "all-intra frames sharded across 8 B200"): with IntraPeriod 1 every picture is coded on its own
  (I slices, no reference pictures, entropy state reset per slice), so frame ranges are the units: shard r encodes input frames
  [start_r, start_r + count_r) with `-fs start_r -f count_r` and TVC_POC_OFFSET=start_r (the hooked encoder then numbers its
  pictures from POC start_r and, for r > 0, writes no VPS/SPS/PPS; tlibcuda_hm.h).
* CLOSED-GOP RANDOM ACCESS (BASELINE configs[2]: "independent intra periods sharded across 2/4/8 B200"): with
  `--DecodingRefreshType=2` the picture at every multiple P of IntraPeriod is an IDR (TEncGOP::getNalUnitType, TLE/TEncGOP.cpp:1728-1742).
  It is coded FIRST in its GOP (POC P-G+1 .. P for GOPSize G) and marks every earlier picture unused
  (decodingRefreshMarking, :288), so the G-1 leading pictures coded after it reference only it and each other: the unit is the
  frame range [P-G+1, P+IntraPeriod-G+1).  A shard that starts at input frame P-G+1 with TVC_POC_OFFSET = P-G+1 does not take
  the first-picture exception (`iPOCLast == 0`, :206,1435; TEncTop.cpp:388), collects a full GOP and codes exactly the single
  run's pictures in the single run's order; the encoder state that crosses pictures (m_iLastIDR :218, the SAO depth rates
  TEncSampleAdaptiveOffset.cpp:1503,1788, the CABAC table choice TEncSbac.cpp:175) is set by the IDR before it is read.  With the
  default `DecodingRefreshType 1` (CRA, open GOP) the leading pictures reference the previous intra period: refused.

No collective, no NCCL: the host concatenates the Annex-B streams, and the result is byte for byte the stream of a single run
over all frames (tests/test_shard_encode.py checks the md5 against the unmodified reference encoder).  Low-delay
configurations (IntraPeriod -1) are refused: their pictures form one dependency chain per sequence.

    python -m thevc_b200.host.shard_encode --cfg build/hm/cfg/encoder_intra_main.cfg -i in.yuv -wdt 1920 -hgt 1080 \
        --frames 16 --shards 8 -o out.bin [--gpus 0,1,...] [--hm intra16] [-- extra encoder arguments]
    python -m thevc_b200.host.shard_encode --cfg build/hm/cfg/encoder_randomaccess_main.cfg ... --frames 128 --shards 4 \
        --hm me,frac,tables,candgrid -o out.bin -- --DecodingRefreshType=2
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


def cfg_value(path: str, extra, key: str, short: str | None = None, default: int | None = None):
    """integer value of a cfg key, command-line arguments (`--Key=v`, `--Key v`, `-short v`) overriding the file as in TAppEncCfg"""
    val = default
    m = re.search(r"^\s*%s\s*:\s*(-?\d+)" % re.escape(key), open(path, encoding="latin-1").read(), re.M)
    if m:
        val = int(m.group(1))
    extra = list(extra)
    for i, a in enumerate(extra):
        if a.startswith("--%s=" % key):
            val = int(a.split("=", 1)[1])
        elif (a == "--" + key or (short and a == "-" + short)) and i + 1 < len(extra):
            val = int(extra[i + 1])
    return val


def cfg_is_all_intra(path: str, extra=()) -> bool:
    return cfg_value(path, extra, "IntraPeriod", "ip") == 1


def intra_period_ranges(frames: int, shards: int, intra_period: int, gop: int):
    """closed-GOP units grouped into contiguous shards: unit k >= 1 starts at frame k * IntraPeriod - (GOPSize - 1), the first
    leading picture of the IDR at k * IntraPeriod; an IDR that lies beyond the last frame opens no unit"""
    starts = [0] + [p - (gop - 1) for p in range(intra_period, frames, intra_period)]
    bounds = starts + [frames]
    units = len(starts)
    out = []
    for r in range(shards):
        a, b = units * r // shards, units * (r + 1) // shards
        if b > a:
            out.append((bounds[a], bounds[b] - bounds[a]))
    return out


def cfg_idr_leads_its_gop(path: str, gop: int) -> bool:
    """the GOP entry coded first (Frame1) is the GOP's last picture in output order, the slot an IDR at a multiple of GOPSize falls in
    (hierarchical-B random access; in the low-delay structures Frame1 is POC 1 and pictures before the IDR are coded before it)"""
    m = re.search(r"^\s*Frame1\s*:\s*\S+\s+(\d+)", open(path, encoding="latin-1").read(), re.M)
    return bool(m) and int(m.group(1)) == gop


def plan_ranges(cfg, frames, shards, extra=()):
    ip = cfg_value(cfg, extra, "IntraPeriod", "ip")
    if ip == 1:
        return frame_ranges(frames, shards)
    gop = cfg_value(cfg, extra, "GOPSize", "g", 1)
    if ip is not None and ip > 1 and cfg_value(cfg, extra, "DecodingRefreshType", "dr", 0) == 2 and gop > 1 and ip > gop and ip % gop == 0 \
            and cfg_idr_leads_its_gop(cfg, gop):     # TAppEncCfg.cpp:513-516 wants IntraPeriod > GOPSize for periodic IDRs
        return intra_period_ranges(frames, shards, ip, gop)
    raise ValueError("sharding needs independent units: an all-intra configuration (IntraPeriod 1) or closed intra periods "
                     "(IntraPeriod a multiple of GOPSize with --DecodingRefreshType=2); other inter pictures depend on the previous unit")


STATEFUL_OPTIONS = ("RateControl", "RateCtrl", "TargetBitrate", "NumLCUInUnit", "FrameSkip", "FramesToBeEncoded")


def check_extra(extra):
    """options that carry state across pictures (rate control) or move the frame window (-f / -fs: the shards set their own)
    break the byte-identical concatenation and are refused"""
    for a in extra:
        key = a.lstrip("-").split("=", 1)[0]
        if a in ("-f", "-fs") or key in STATEFUL_OPTIONS:
            if key.startswith("Rate") and "=" in a and a.split("=", 1)[1] in ("0", "false"):
                continue
            raise ValueError("shard_encode: '%s' is not allowed in the extra encoder arguments (state across pictures / frame window)" % a)


def visible_gpus():
    """device indices of this box (nvidia-smi -L); [] without a driver"""
    try:
        r = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30)
        return [i for i, ln in enumerate(r.stdout.splitlines()) if ln.startswith("GPU ")] if r.returncode == 0 else []
    except (OSError, subprocess.TimeoutExpired):
        return []


def shards_per_gpu(w, h, hm, hbm_gb=150):
    """how many shard processes one GPU holds.  The default integer search (group search: windows staged in shared memory) needs no
    table memory; only the round-1 form (TVC_ME_FUSED=0 with the `tables` hook) reserves the SAD tables of four references per
    process (17.04 MB per (CTU, reference): 34.8 GB at 1080p)"""
    if "tables" not in hm.split(",") or os.environ.get("TVC_ME_FUSED", "1") != "0":
        return 16
    nctu = ((w + 63) // 64) * ((h + 63) // 64)
    return max(1, int(hbm_gb * 1e9 // (nctu * 4 * 17.04e6 + 2e9)))


def shard_encode(cfg, yuv, w, h, frames, shards, out, gpus=None, hm="none", extra=(), encoder=ENC, workdir=None, max_per_gpu=None):
    check_extra(extra)
    ranges = plan_ranges(cfg, frames, shards, extra)
    if not os.path.exists(encoder):
        raise RuntimeError("%s is not built (make -C thevc_b200/host hm)" % encoder)
    workdir = workdir or os.path.dirname(os.path.abspath(out))
    device_hooks = hm not in ("", "none")
    if device_hooks and not gpus:
        gpus = visible_gpus()
        if not gpus:
            raise RuntimeError("shard_encode: device hooks (%s) need a GPU and none is visible" % hm)
    cap = max_per_gpu or (shards_per_gpu(w, h, hm) if device_hooks else len(ranges))
    load = {g: 0 for g in (gpus or [None])}
    pending = list(enumerate(ranges))
    running, parts = [], [None] * len(ranges)
    t0 = time.perf_counter()
    shard_wall = [0.0] * len(ranges)

    def kill_all():
        for _, _, _, _, p, _ in running:
            if p.poll() is None:
                p.kill()
        for _, _, _, log, p, _ in running:
            p.wait()
            log.close()

    try:
        while pending or running:
            # start every shard a GPU has room for (least loaded first)
            while pending:
                g = min(load, key=lambda k: load[k])
                if load[g] >= cap:
                    break
                r, (start, count) = pending.pop(0)
                part = os.path.join(workdir, "shard_%03d.bin" % r)
                env = dict(os.environ, TVC_POC_OFFSET=str(start), TVC_HM=hm)
                if g is not None:
                    env["CUDA_VISIBLE_DEVICES"] = str(g)
                cmd = [encoder, "-c", cfg, "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-fs", str(start), "-f", str(count), "-b", part,
                       "-o", os.devnull] + list(extra)
                log = open(part + ".log", "w")
                running.append((r, g, part, log, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT, env=env), time.perf_counter()))
                load[g] += 1
                parts[r] = part
            done = [x for x in running if x[4].poll() is not None]
            if not done:
                time.sleep(0.05)
                continue
            for x in done:
                r, g, part, log, p, ts = x
                running.remove(x)
                log.close()
                load[g] -= 1
                shard_wall[r] = time.perf_counter() - ts
                if p.returncode != 0:
                    raise RuntimeError("shard %s failed (%d): %s" % (part, p.returncode, open(part + ".log").read()[-800:]))
    except BaseException:
        kill_all()          # one failed shard ends the job: the others are stopped, their logs closed
        raise
    wall = time.perf_counter() - t0
    md5 = hashlib.md5()
    with open(out, "wb") as fo:                      # the whole exchange step of this path: concatenation on the host
        for part in parts:
            data = open(part, "rb").read()
            fo.write(data)
            md5.update(data)
    return {"frames": frames, "shards": len(ranges), "ranges": ranges, "wall_s": wall, "fps": frames / wall, "md5": md5.hexdigest(),
            "bytes": os.path.getsize(out), "gpus": gpus, "shards_per_gpu_limit": cap, "shard_wall_s": [round(x, 2) for x in shard_wall], "hm": hm}


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
    ap.add_argument("--max-per-gpu", type=int, default=0, help="shard processes per GPU at a time (default: derived from the SAD-table size)")
    ap.add_argument("extra", nargs="*")
    a = ap.parse_args()
    gpus = [int(g) for g in a.gpus.split(",") if g != ""] or None
    print(json.dumps(shard_encode(a.cfg, a.input, a.wdt, a.hgt, a.frames, a.shards, a.out, gpus, a.hm, a.extra, max_per_gpu=a.max_per_gpu or None)))


if __name__ == "__main__":
    sys.exit(main())
