"""wall-clock breakdown of bench.py's e2e step (host-pointer ABI), each call followed by a device synchronise"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from bench import *
from thevc_b200 import TLibCuda
from thevc_b200.capi import MeFrameCfg, QuantCfg, ptr
wl = bench.Workload(20261018, pinned=True)
stream = torch.cuda.Stream()
t = TLibCuda(W, H, BD, num_slots=NUM_SLOTS, device=0, stream=stream.cuda_stream)
L, h = t.L, t.h
lc = int(np.floor(65536.0 * np.sqrt(LAMBDA)))
for s, p in enumerate(wl.pics): t.upload(s, p)
refs = (C.c_int * NUM_REFS)(*range(1, NUM_REFS + 1))
mcfg = MeFrameCfg(SEARCH_RANGE, 1, 1, 1, 1, lc); qc = QuantCfg(0, 1, 0)
n_pu, n_tu = len(wl.pus), len(wl.tus)
def pinned(shape, dtype):
    tt = torch.zeros(shape, dtype=dtype).pin_memory(); return tt, tt.numpy()
_k1, ires_h = pinned((NUM_REFS * wl.nctu * 593, 4), torch.int32)
_k2, fres_h = pinned((NUM_REFS * wl.nctu * 593, 6), torch.int32)
_k3, levels_h = pinned((wl.coef_elems,), torch.int32)
_k4, abs_h = pinned((n_tu,), torch.int32)
recon_h = type(wl.pics[0])(W, H, alloc=wl.alloc)
planes_wh = [(0, W, H), (1, W // 2, H // 2), (2, W // 2, H // 2)]
acc = {}
def tm(name, f):
    t0 = time.perf_counter(); f(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0]); a[0] += (t1 - t0) * 1e3; a[1] += (t2 - t1) * 1e3
def step():
    tm("upload0", lambda: t.upload(0, wl.pics[0]))
    tm("upload1", lambda: t.upload(1, wl.pics[1]))
    tm("me_frame", lambda: L.tvc_me_frame(h, SLOT_CUR, NUM_REFS, refs, ptr(wl.pred), C.byref(mcfg), ptr(ires_h), ptr(fres_h)))
    tm("mc", lambda: L.tvc_mc_batch(h, SLOT_PRED, n_pu, ptr(wl.pus)))
    tm("subtract", lambda: [L.tvc_pic_subtract(h, SLOT_RESI, SLOT_CUR, SLOT_PRED, pl, 0, 0, pw, ph) for pl, pw, ph in planes_wh])
    tm("fwd_rdoq_recon", lambda: L.tvc_fwd_rdoq_recon_batch(h, SLOT_RESI, SLOT_RESI2, SLOT_PRED, SLOT_RECON, n_tu, ptr(wl.tus), ptr(wl.rtus), 1, ptr(wl.est_bytes),
                                      C.byref(qc), ptr(levels_h), wl.coef_elems, ptr(abs_h)))
    tm("download", lambda: t.download(SLOT_RECON, into=recon_h))
with torch.cuda.stream(stream):
    step(); step(); acc.clear()
    for _ in range(5): step()
for k, v in acc.items(): print("%-16s call %.3f ms + sync %.3f ms" % (k, v[0] / 5, v[1] / 5))
print("total %.3f" % (sum(v[0] + v[1] for v in acc.values()) / 5))
print("bytes: ires %d fres %d levels %d pic %d pus %d tus %d rtus %d" % (ires_h.nbytes, fres_h.nbytes, levels_h.nbytes, wl.pic_bytes(), wl.pus.nbytes, wl.tus.nbytes, wl.rtus.nbytes))
