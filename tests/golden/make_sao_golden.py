"""Generate tests/golden/sao_golden.npz from the REFERENCE ITSELF: the reference encoder (build/hm/TAppEncoderCuda with
TVC_HM=saodump: no CUDA, the reference's own TComSampleAdaptiveOffset code filters) writes, per colour component, the
plane before and after SAOProcess and the per-CTU records (type, edge-offset table, band offsets) its processSaoUnitAll
resolved.  Run in the build container:

    python tests/golden/make_sao_golden.py
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
    w, h, bd, comp, ctu, ctus_x, n, poc = (int(v) for v in np.frombuffer(raw, np.int32, 8))
    o = 32
    units = np.frombuffer(raw, np.int16, n * 38, o).reshape(n, 38); o += n * 76
    before = np.frombuffer(raw, np.int16, w * h, o).reshape(h, w); o += w * h * 2
    after = np.frombuffer(raw, np.int16, w * h, o).reshape(h, w); o += w * h * 2
    assert o == len(raw)
    return dict(w=w, h=h, bd=bd, comp=comp, ctu=ctu, ctus_x=ctus_x, poc=poc, units=units, before=before, after=after)


def run(cfg, w, h, frames, extra, d, seed):
    yuv = os.path.join(d, "in.yuv")
    seq = synth.make_sequence(w, h, frames, seed=seed)
    with open(yuv, "wb") as f:
        for y, u, v in seq:
            f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
    dump = os.path.join(d, "dump"); os.makedirs(dump, exist_ok=True)
    for f in os.listdir(dump):
        os.remove(os.path.join(dump, f))
    env = dict(os.environ, TVC_HM="saodump", TVC_SAO_DUMP=dump, TVC_SAO_DUMP_PLANES=str(3 * frames))
    subprocess.run([ENC, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames),
                    "-b", os.path.join(d, "o.bin")] + list(extra), check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return [read_dump(os.path.join(dump, f)) for f in sorted(os.listdir(dump))]


def main():
    out = {}
    n = 0
    types = set()
    with tempfile.TemporaryDirectory() as d:
        cases = [("encoder_lowdelay_P_main.cfg", 416, 240, 2, ("--QP=34",), 1),
                 ("encoder_intra_main.cfg", 208, 120, 2, ("--QP=40",), 2),
                 ("encoder_intra_main.cfg", 200, 120, 1, ("--QP=30",), 4),          # width not a multiple of 64/16: partial CTUs
                 ("encoder_intra_he10.cfg", 208, 120, 1, ("--QP=36",), 3)]          # 10-bit internal
        for cfg, w, h, frames, extra, seed in cases:
            for rec in run(cfg, w, h, frames, extra, d, seed):
                tag = "c%d_" % n
                out[tag + "hdr"] = np.array([rec["w"], rec["h"], rec["bd"], rec["comp"], rec["ctu"], rec["ctus_x"], rec["poc"]], np.int32)
                out[tag + "units"] = rec["units"]
                out[tag + "before"] = rec["before"]
                out[tag + "delta"] = (rec["after"].astype(np.int32) - rec["before"]).astype(np.int16)
                used = sorted(set(int(t) for t in rec["units"][:, 0]))
                types.update(used)
                print(cfg, "%dx%d" % (rec["w"], rec["h"]), "bd", rec["bd"], "comp", rec["comp"], "POC", rec["poc"], "types", used,
                      "samples changed", int(np.count_nonzero(rec["after"] != rec["before"])))
                n += 1
    out["count"] = np.array([n], np.int32)
    path = os.path.join(HERE, "sao_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", n, "planes; SAO types seen:", sorted(types))


if __name__ == "__main__":
    main()
