#!/bin/bash
# round 2, call Z2: k_me_group launched longest-group-first by the previous call's CTA durations (TVC_GROUP_ORDER=4, default) against the interleaved order (2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
cat > /tmp/order_me.py <<'P'
import hashlib, os, sys, numpy as np
sys.path.insert(0, '.')
import bench
from thevc_b200 import TLibCuda
wl = bench.Workload(20261018, pinned=False)
t = TLibCuda(bench.W, bench.H, 8, num_slots=6)
for s_, p in enumerate(wl.pics): t.upload(s_, p)
lc = int(np.floor(65536.0 * np.sqrt(bench.LAMBDA)))
for _ in range(3): t.me_frame(0, [1, 2, 3, 4], wl.pred, lc, do_frac=False)
t.prof_enable(True); t.prof_read(reset=True)
for _ in range(8): ires, _f = t.me_frame(0, [1, 2, 3, 4], wl.pred, lc, do_frac=False)
ms = t.prof_read(reset=True)
t.prof_enable(False)
ires, fres = t.me_frame(0, [1, 2, 3, 4], wl.pred, lc)
print("order", os.environ.get("TVC_GROUP_ORDER", "default"), "me_search ms", ms["me_search"][0] / 8, "hash", hashlib.md5(ires.tobytes()).hexdigest(), hashlib.md5(fres.tobytes()).hexdigest())
t.close()
P
timeout 300 python /tmp/order_me.py 2>&1 | tail -1 | tee $O/r02z2_order.log
TVC_GROUP_ORDER=2 timeout 300 python /tmp/order_me.py 2>&1 | tail -1 | tee -a $O/r02z2_order.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "me_" > $O/r02z2_parity.log 2>&1; echo "parity rc=$?"; tail -2 $O/r02z2_parity.log
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 0 --cpu-enc-frames 0 > $O/r02z2_bench.json 2> $O/r02z2_bench.err; echo "bench rc=$?"; tail -3 $O/r02z2_bench.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/r02z2_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['detail']['phase_ms_per_step'])
P
