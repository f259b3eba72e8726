#!/bin/bash
# round 2, call X: k_me_group variant 14 as the default + per-PU fractional kernels over lists of the unserved jobs: hash, parity, bench, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
cat > /tmp/hash_me.py <<'P'
import hashlib, sys, numpy as np
sys.path.insert(0, '.')
import bench
from thevc_b200 import TLibCuda
wl = bench.Workload(20261018, pinned=False)
t = TLibCuda(bench.W, bench.H, 8, num_slots=6)
for s_, p in enumerate(wl.pics): t.upload(s_, p)
lc = int(np.floor(65536.0 * np.sqrt(bench.LAMBDA)))
ires, fres = t.me_frame(0, [1, 2, 3, 4], wl.pred, lc)
print("hash", hashlib.md5(ires.tobytes()).hexdigest(), hashlib.md5(fres.tobytes()).hexdigest())
t.close()
P
timeout 300 python /tmp/hash_me.py 2>&1 | tail -1 | tee $O/r02x_hash.log
TVC_FRAC_LISTS=0 timeout 300 python /tmp/hash_me.py 2>&1 | tail -1 | sed 's/^/lists=0 /' | tee -a $O/r02x_hash.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q -m gpu > $O/r02x_parity.log 2>&1; echo "parity rc=$?"; tail -3 $O/r02x_parity.log
timeout 900 python bench.py --steps 10 --warmup 3 --hm-frames 0 --no-cpu --cpu-enc-frames 0 > $O/r02x_bench.json 2> $O/r02x_bench.err; echo "bench rc=$?"; tail -3 $O/r02x_bench.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['detail']['phase_ms_per_step'])
P
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --hm-frames 0 --cpu-enc-frames 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02x_launches.csv $CMD > $O/r02x_ncu_l.log 2>&1; echo "ncu list rc=$?"
