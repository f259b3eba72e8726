#!/bin/bash
# round 2, call Z: CTA -> group order of k_me_group (TVC_GROUP_ORDER 0..3): kernel time per order, result hash; then the short bench
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
for order in (0, 1, 2, 3, 0, 3):
    os.environ["TVC_GROUP_ORDER"] = str(order)
    for _ in range(2): t.me_frame(0, [1, 2, 3, 4], wl.pred, lc, do_frac=False)
    t.prof_enable(True); t.prof_read(reset=True)
    for _ in range(6): ires, _f = t.me_frame(0, [1, 2, 3, 4], wl.pred, lc, do_frac=False)
    ms = t.prof_read(reset=True)
    t.prof_enable(False)
    print("order", order, "me_search ms", ms["me_search"][0] / 6, "hash", hashlib.md5(ires.tobytes()).hexdigest())
t.close()
P
timeout 300 python /tmp/order_me.py 2>&1 | tail -7 | tee $O/r02z_order.log
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 0 --cpu-enc-frames 0 > $O/r02z_bench.json 2> $O/r02z_bench.err; echo "bench rc=$?"; tail -3 $O/r02z_bench.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/r02z_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['detail']['phase_ms_per_step'])
P
