#!/bin/bash
# round 2, call U: code-size variants of k_me_group (TVC_GRP_VAR 0..3: SAD loop / CU cost routine behind calls), timing + result hash
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
for v in 0 1 2 3 6; do
  cp build/variants/libthevc_cuda_v$v.so thevc_b200/lib/libthevc_cuda.so
  timeout 300 python /tmp/hash_me.py 2>&1 | tail -1 | sed "s/^/v$v /" | tee -a $O/r02u_hash.log
  timeout 600 python bench.py --steps 8 --warmup 3 --hm-frames 0 --no-cpu --cpu-enc-frames 0 > $O/r02u_bench_v$v.json 2> $O/r02u_bench_v$v.err; echo "v$v bench rc=$?"
done
TVC_GROUP_MINB=3 timeout 600 python bench.py --steps 8 --warmup 3 --hm-frames 0 --no-cpu --cpu-enc-frames 0 > $O/r02u_bench_v6_minb3.json 2> $O/r02u_bench_v6_minb3.err; echo "v6 minb3 bench rc=$?"
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02u_bench_v*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['detail']['phase_ms_per_step']['me_search'], d['detail']['phase_ms_per_step']['me_frac'])
    except Exception as e:
        print(f, 'ERR', e)
P
