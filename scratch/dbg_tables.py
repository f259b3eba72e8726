import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, ctypes as C
import synth, oracle
from oracle import ptr as optr
from thevc_b200 import TLibCuda, capi, TvcError
orc = oracle.lib()
W, H = 416, 240
def run(tag, nrefs, centers):
    t = TLibCuda(W, H, 8, num_slots=6)
    seq = synth.make_sequence(W, H, 3)
    cur = synth.to_hostpic(seq[2], W, H)
    refs = [synth.to_hostpic(seq[1], W, H), synth.to_hostpic(seq[0], W, H)]
    t.upload(0, cur); t.upload(1, refs[0]); t.upload(2, refs[1])
    try:
        t.me_prepass(0, [1, 2][:nrefs], centers)
        t.sync()
    except TvcError as e:
        print(tag, "FAILED", e); return
    # check 4x4 blocks everywhere for a handful of candidates
    bad = 0; tot = 0
    rng = np.random.default_rng(0)
    nctu = t.ctus_x * t.ctus_y
    for ri in range(nrefs):
        for ctu in range(nctu):
            cx0, cy0 = (ctu % t.ctus_x) * 64, (ctu // t.ctus_x) * 64
            cc = centers[ri, ctu] if centers is not None else (0, 0)
            lo_x, hi_x = -80 - cx0 + 64, W + 80 - cx0 - 128
            lo_y, hi_y = -80 - cy0 + 64, H + 80 - cy0 - 128
            ccx = min(max(int(cc[0]), lo_x), max(hi_x, lo_x)); ccy = min(max(int(cc[1]), lo_y), max(hi_y, lo_y))
            cand = [(ccx + int(a), ccy + int(b)) for a, b in rng.integers(-64, 65, (6, 2))] + [(ccx+64, ccy+64), (ccx-64, ccy-64), (ccx+64, ccy-3), (ccx+5, ccy+64)]
            for by in range(16):
                for bx in range(16):
                    x, y = cx0 + bx * 4, cy0 + by * 4
                    if x + 4 > W or y + 4 > H: continue
                    got = t.me_table_lookup(ri, x, y, 4, 4, 0, np.array(cand, np.int16))
                    for (mvx, mvy), g in zip(cand, got):
                        r = refs[ri]
                        e = orc.orc_sad(optr(cur.buf_y, cur.origin(0) + y * cur.stride + x), cur.stride,
                                        optr(r.buf_y, r.origin(0) + (y + mvy) * r.stride + x + mvx), r.stride, 4, 4, 0, 0)
                        tot += 1
                        if g != e:
                            bad += 1
                            if bad < 12: print(tag, "mismatch ref", ri, "ctu", ctu, "bx,by", bx, by, "d", mvx - ccx, mvy - ccy, "got", g, "exp", e)
    print(tag, "checked", tot, "bad", bad)
    t.close()
nctu = 7 * 4
run("A: 2 refs zero centres", 2, None)
c = np.zeros((1, nctu, 2), np.int32); c[0, :, 0] = 5; c[0, :, 1] = 3
run("B: 1 ref unaligned centres", 1, c)
c = np.zeros((1, nctu, 2), np.int32); c[0, :, 0] = 16; c[0, :, 1] = 3
run("C: 1 ref 16-aligned x centre", 1, c)
