"""Generate tests/golden/deblock_golden.npz from the REFERENCE ITSELF: the reference encoder (build/hm/TAppEncoderCuda
with TVC_HM=dbkdump: no CUDA, every filter decision and sample is the reference's own TComLoopFilter code) writes, per
picture, the reconstruction before and after loopFilterPic and the edge-unit records (bs, qp, flags) its xEdgeFilterLuma
acted on.  Run in the build container (needs build/hm, i.e. /root/reference at build time):

    python tests/golden/make_deblock_golden.py
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
    w, h, bd, beta, tc, poc = np.frombuffer(raw, np.int32, 6)
    o = 24
    nv, nh = ((w + 7) >> 3) * ((h + 3) >> 2), ((w + 3) >> 2) * ((h + 7) >> 3)
    ver = np.frombuffer(raw, np.uint8, nv * 4, o).reshape(nv, 4); o += nv * 4
    hor = np.frombuffer(raw, np.uint8, nh * 4, o).reshape(nh, 4); o += nh * 4
    planes = []
    for k in range(6):
        pw, ph = (w, h) if k % 3 == 0 else (w // 2, h // 2)
        planes.append(np.frombuffer(raw, np.int16, pw * ph, o).reshape(ph, pw)); o += pw * ph * 2
    assert o == len(raw)
    return dict(w=int(w), h=int(h), bd=int(bd), beta=int(beta), tc=int(tc), poc=int(poc), ver=ver, hor=hor, before=planes[:3], after=planes[3:])


def run(cfg, w, h, frames, extra, d):
    yuv = os.path.join(d, "in.yuv")
    seq = synth.make_sequence(w, h, frames)
    with open(yuv, "wb") as f:
        for y, u, v in seq:
            f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
    dump = os.path.join(d, "dump"); os.makedirs(dump, exist_ok=True)
    for f in os.listdir(dump):
        os.remove(os.path.join(dump, f))
    env = dict(os.environ, TVC_HM="dbkdump", TVC_DBK_DUMP=dump, TVC_DBK_DUMP_PICS=str(frames))
    subprocess.run([ENC, "-c", os.path.join(CFG, cfg), "-i", yuv, "-wdt", str(w), "-hgt", str(h), "-fr", "30", "-f", str(frames),
                    "-b", os.path.join(d, "o.bin")] + list(extra), check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return [read_dump(os.path.join(dump, f)) for f in sorted(os.listdir(dump))]


def main():
    out = {}
    n = 0
    with tempfile.TemporaryDirectory() as d:
        cases = [("encoder_lowdelay_P_main.cfg", 416, 240, 3, ("--QP=30",)),                        # I + 2 P pictures, 8-bit
                 ("encoder_lowdelay_P_main.cfg", 208, 120, 2, ("--QP=38", "--DeblockingFilterControlPresent=1", "--LoopFilterOffsetInPPS=1", "--LoopFilterBetaOffset_div2=2", "--LoopFilterTcOffset_div2=-1")),
                 ("encoder_intra_he10.cfg", 208, 120, 1, ("--QP=34",))]                             # 10-bit internal
        for cfg, w, h, frames, extra in cases:
            for rec in run(cfg, w, h, frames, extra, d):
                tag = "c%d_" % n
                out[tag + "hdr"] = np.array([rec["w"], rec["h"], rec["bd"], rec["beta"], rec["tc"], rec["poc"]], np.int32)
                out[tag + "ver"], out[tag + "hor"] = rec["ver"], rec["hor"]
                for k, nm in enumerate("yuv"):
                    out[tag + "before_" + nm] = rec["before"][k]
                    out[tag + "delta_" + nm] = (rec["after"][k].astype(np.int32) - rec["before"][k]).astype(np.int16)
                changed = sum(int(np.count_nonzero(rec["after"][k] != rec["before"][k])) for k in range(3))
                print(cfg, "%dx%d" % (rec["w"], rec["h"]), "bd", rec["bd"], "POC", rec["poc"], "units", int((rec["ver"][:, 0] > 0).sum() + (rec["hor"][:, 0] > 0).sum()),
                      "bs2", int((rec["ver"][:, 0] > 1).sum() + (rec["hor"][:, 0] > 1).sum()), "samples changed", changed, "offsets", rec["beta"], rec["tc"])
                n += 1
    out["count"] = np.array([n], np.int32)
    path = os.path.join(HERE, "deblock_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", n, "pictures")


if __name__ == "__main__":
    main()
