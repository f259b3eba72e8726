#!/bin/bash
# round 2, call W: k_me_group with per-warp statistics sums (variants 11, 14) and record prefetch (14p): timing, result hash, ncu of the fastest
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
for v in $VARIANTS; do
  cp build/variants/libthevc_cuda_v$v.so thevc_b200/lib/libthevc_cuda.so
  timeout 300 python /tmp/hash_me.py 2>&1 | tail -1 | sed "s/^/v$v /" | tee -a $O/r02w_hash.log
  timeout 600 python bench.py --steps 8 --warmup 3 --hm-frames 0 --no-cpu --cpu-enc-frames 0 > $O/r02w_bench_v$v.json 2> $O/r02w_bench_v$v.err; echo "v$v bench rc=$?"
done
BEST=$(python - <<'P'
import json, glob
best=None
for f in sorted(glob.glob('gpurun_out/r02w_bench_v*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ms = d['detail']['phase_ms_per_step']['me_search']
        import sys
        print(f, d['ms_per_step'], ms, d['detail']['phase_ms_per_step']['me_frac'], file=sys.stderr)
        v = f.split('_v')[-1].split('.')[0]
        if best is None or ms < best[0]: best = (ms, v)
    except Exception as e:
        pass
print(best[1] if best else '')
P
)
echo "best variant: $BEST"
if [ -n "$BEST" ]; then
  cp build/variants/libthevc_cuda_v$BEST.so thevc_b200/lib/libthevc_cuda.so
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_me_group -c 1 -o $O/r02w_grp_v$BEST python bench.py --steps 1 --warmup 1 --no-cpu --hm-frames 0 --cpu-enc-frames 0 > $O/r02w_ncu.log 2>&1; echo "ncu rc=$?"
fi
