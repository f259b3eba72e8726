"""Generate tests/golden/intra_rough.npz from the REFERENCE ITSELF: the reference encoder (build/hm/TAppEncoderCuda with
TVC_HM=intradump: no CUDA, the reference's own estIntraPredQT loop runs predIntraLumaAng + calcHAD) writes, for a sample
of the PUs of every size, the unfiltered reference samples initAdiPattern left in m_piYuvExt (as one line, see
include/thevc_cuda.h), the original block and the uiSad of each of the 35 modes (TEncSearch.cpp:2530-2537).  The
smoothing of the reference samples, the choice of the smoothed buffer per mode, all 35 predictors and the Hadamard cost
are therefore pinned by what the running reference computed.  Run in the build container:

    python tests/golden/make_intra_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
import synth  # noqa: E402

ENC = os.path.join(ROOT, "build", "hm", "TAppEncoderCuda")
CFG = os.path.join(ROOT, "build", "hm", "cfg")


def read_dump(path):
    raw = open(path, "rb").read()
    recs, o = [], 0
    while o < len(raw):
        log2n, bd = (int(v) for v in np.frombuffer(raw, np.int32, 2, o)); o += 8
        n = 1 << log2n
        line = np.frombuffer(raw, np.int16, 4 * n + 1, o); o += (4 * n + 1) * 2
        org = np.frombuffer(raw, np.int16, n * n, o); o += n * n * 2
        sad = np.frombuffer(raw, np.uint32, 35, o); o += 140
        recs.append((log2n, bd, line, org, sad))
    return recs


def run(cfg, w, h, frames, extra, d, seed, per_size, stride):
    yuv = os.path.join(d, "in.yuv")
    seq = synth.make_sequence(w, h, frames, seed=seed)
    with open(yuv, "wb") as f:
        for y, u, v in seq:
            f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
    dump = os.path.join(d, "dump"); os.makedirs(dump, exist_ok=True)
    for f in os.listdir(dump):
        os.remove(os.path.join(dump, f))
    env = dict(os.environ, TVC_HM="intradump", TVC_INTRA_DUMP=dump, TVC_INTRA_DUMP_PER_SIZE=str(per_size), TVC_INTRA_DUMP_STRIDE=str(stride))
    subprocess.run([ENC, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames),
                    "-b", os.path.join(d, "o.bin")] + list(extra), check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return read_dump(os.path.join(dump, "intra_rough.bin"))


def main():
    out = {}
    with tempfile.TemporaryDirectory() as d:
        # picture borders (substituted neighbours), interior PUs, intra PUs inside P pictures; 8-bit and 10-bit internal
        cases = [("encoder_intra_main.cfg", 416, 240, 1, ("--QP=32",), 1, 30, 37),
                 ("encoder_lowdelay_P_main.cfg", 208, 120, 2, ("--QP=30",), 2, 10, 53),
                 ("encoder_intra_he10.cfg", 208, 120, 1, ("--QP=30",), 3, 30, 17)]
        by_bd = {}
        for cfg, w, h, frames, extra, seed, per_size, stride in cases:
            recs = run(cfg, w, h, frames, extra, d, seed, per_size, stride)
            sizes = {}
            for log2n, bd, line, org, sad in recs:
                by_bd.setdefault(bd, []).append((log2n, line, org, sad))
                sizes[1 << log2n] = sizes.get(1 << log2n, 0) + 1
            print(cfg, "%dx%d" % (w, h), "PUs per size:", dict(sorted(sizes.items())))
        for bd, recs in by_bd.items():
            k = "bd%d" % bd
            out[k + "_log2"] = np.array([r[0] for r in recs], np.int32)
            out[k + "_lines"] = np.concatenate([r[1] for r in recs])
            out[k + "_orgs"] = np.concatenate([r[2] for r in recs])
            out[k + "_sads"] = np.stack([r[3] for r in recs])
    path = os.path.join(HERE, "intra_rough.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
