"""frame pre-pass at 1080p: serial vs pipelined form (TVC_ME_PIPE = chunks per reference, 0 = serial); prints ms per picture and a checksum"""
import sys, os, json, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
import bench
from thevc_b200 import TLibCuda
from thevc_b200.capi import MeFrameCfg, ptr
wl = bench.Workload(20261018, pinned=False)
stream = torch.cuda.Stream()
t = TLibCuda(bench.W, bench.H, 8, num_slots=9, stream=stream.cuda_stream)
for s, p in enumerate(wl.pics): t.upload(s, p)
lc = int(np.floor(65536.0 * np.sqrt(bench.LAMBDA)))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
refs = (C.c_int * 4)(1, 2, 3, 4)
mcfg = MeFrameCfg(64, 1, 1, 1, 1, lc)
pi, pf = C.c_void_p(), C.c_void_p()
def step():
    rc = t.L.tvc_me_frame_dev(t.h, 0, 4, refs, ptr(wl.pred), C.byref(mcfg), C.byref(pi), C.byref(pf))
    assert rc == 0, t.L.tvc_last_error(t.h)
with torch.cuda.stream(stream):
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): step()
    e1.record(stream)
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
ires, fres = t.me_frame(0, [1, 2, 3, 4], wl.pred, lc)
print("PIPE=%s ms/picture %.3f checksum %d %d %d %d" % (os.environ.get("TVC_ME_PIPE", "default"), ms, int(ires["sad"].sum()), int(ires["n_sads"].sum()),
      int(fres["cost"].sum()), int(fres["qtrx"].astype(np.int64).sum() * 7 + fres["qtry"].astype(np.int64).sum() * 3 + fres["halfx"].astype(np.int64).sum())))
