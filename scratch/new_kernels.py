"""launches the round-1 'next row' kernels at 1080p scale (for ncu captures and a quick CUDA-event timing):
k_ctu_cost_grids (one MV per CTU and reference), k_pred_cost (the picture's PU list, HAD), k_intra_rough (every intra PU), hashes"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch
import bench
from thevc_b200 import TLibCuda, capi
from thevc_b200.capi import ptr
wl = bench.Workload(20261018, pinned=False)
stream = torch.cuda.Stream()
t = TLibCuda(bench.W, bench.H, 8, num_slots=9, stream=stream.cuda_stream)
for s, p in enumerate(wl.pics): t.upload(s, p)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
jobs = np.zeros(wl.nctu * 4, capi.GRID_JOB_DTYPE)
k = 0
for r in range(4):
    for c in range(wl.nctu):
        jobs[k] = (r + 1, (c % wl.ctus_x) * 64, (c // wl.ctus_x) * 64, int(wl.pred[r, c, 0]), int(wl.pred[r, c, 1])); k += 1
d_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1).copy()).cuda()
d_out = torch.zeros(len(jobs) * capi.GRID_WORDS, dtype=torch.int32, device="cuda")
pus = wl.pus
d_pus = torch.from_numpy(pus.view(np.uint8).reshape(-1).copy()).cuda()
d_dist = torch.zeros(len(pus), dtype=torch.int32, device="cuda")
def ev_time(f):
    with torch.cuda.stream(stream):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps): f()
        e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def ck(rc):
    assert rc == 0, t.L.tvc_last_error(t.h)
ms = ev_time(lambda: ck(t.L.tvc_ctu_cost_grids_dev(t.h, 0, len(jobs), C.c_void_p(d_jobs.data_ptr()), C.c_void_p(d_out.data_ptr()))))
print("k_ctu_cost_grids: %d (CTU, reference, MV) grids in %.3f ms (%.2f us each)" % (len(jobs), ms, ms * 1e3 / len(jobs)))
ms = ev_time(lambda: ck(t.L.tvc_pred_cost_batch_dev(t.h, 0, capi.DIST_HADS, len(pus), C.c_void_p(d_pus.data_ptr()), C.c_void_p(d_dist.data_ptr()))))
print("k_pred_cost: %d candidates (HAD) in %.3f ms" % (len(pus), ms))
print("intra_rough:", bench.intra_rough_leg(t, wl, 0, cpu_sample=8))
for m in (3, 2): t.pic_hash(0, m)
t.pic_ssd(0, 1)
