import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import bench
from thevc_b200 import TLibCuda
wl = bench.Workload(20261018, pinned=False)
t = TLibCuda(bench.W, bench.H, 8, num_slots=6)
for s, p in enumerate(wl.pics):
    t.upload(s, p)
lc = int(np.floor(65536.0 * np.sqrt(bench.LAMBDA)))
ires, fres = t.me_frame(0, [1, 2, 3, 4], wl.pred, lc, do_frac=False)
v = ires["n_sads"] > 0
out = {}
dist = []
top1 = []
for r in range(4):
    for c in range(ires.shape[1]):
        m = v[r, c]
        if not m.any():
            continue
        mv = ires["mvx"][r, c][m].astype(np.int64) * 1000 + ires["mvy"][r, c][m]
        u, cnt = np.unique(mv, return_counts=True)
        dist.append(len(u)); top1.append(cnt.max() / m.sum())
dist = np.array(dist); top1 = np.array(top1)
ns = ires["n_sads"][v]
print(json.dumps({"groups": int(len(dist)), "distinct_final_mv_per_group": {"mean": float(dist.mean()), "median": float(np.median(dist)), "p90": float(np.percentile(dist, 90)), "max": int(dist.max())},
                  "share_of_most_common_mv": {"mean": float(top1.mean()), "p10": float(np.percentile(top1, 10))},
                  "n_sads_per_job": {"mean": float(ns.mean()), "median": float(np.median(ns)), "p90": float(np.percentile(ns, 90))},
                  "jobs_with_raster": float((ns > 700).mean())}))
t.close()
