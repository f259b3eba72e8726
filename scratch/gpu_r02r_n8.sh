#!/bin/bash
# round 2, second 8-GPU call: the bench line at N = 8 with the ranks pinned to their share of the host cores (end-to-end scaling)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
nproc > $O/r02r_nproc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02r_bench_n8.json 2> $O/r02r_bench_n8.err; echo "bench8 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --hm-frames 0 > $O/r02r_bench_n1_same_box.json 2> $O/r02r_bench_n1.err; echo "bench1 rc=$?"
python - <<'P'
import json
for f in ("gpurun_out/r02r_bench_n8.json", "gpurun_out/r02r_bench_n1_same_box.json"):
    b = json.loads(open(f).read().strip().splitlines()[-1]); print(f, b["n_gpus"], b["value"], b["ms_per_step"], b["e2e"])
P
