"""micro-benchmark of the ME frame pre-pass phases at 1080p (prints CUDA-event phase times)"""
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np
import bench
from thevc_b200 import TLibCuda
wl = bench.Workload(20261018, pinned=False)
t = TLibCuda(bench.W, bench.H, 8, num_slots=9)
for s, p in enumerate(wl.pics): t.upload(s, p)
lc = int(np.floor(65536.0 * np.sqrt(bench.LAMBDA)))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
frac = int(os.environ.get("DO_FRAC", "1"))
t.me_frame(0, [1, 2, 3, 4], wl.pred, lc, do_frac=bool(frac))
t.prof_enable(True); t.prof_read(True)
for _ in range(reps):
    ires, fres = t.me_frame(0, [1, 2, 3, 4], wl.pred, lc, do_frac=bool(frac))
ph = t.prof_read(True)
print(json.dumps({k: round(v[0] / reps, 3) for k, v in ph.items() if v[1]}), "checksum", int(ires["sad"].sum()), int(ires["n_sads"].sum()), int(fres["cost"].sum()))
