#!/bin/bash
# round 2, 2-GPU call: configs[2] (RA 1080p, IDR refresh, cfg's own IntraPeriod 32: three closed intra periods as TWO shards over two GPUs), bench at N = 2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O/matrix
nvidia-smi -L > $O/r02q_gpus.txt; nproc >> $O/r02q_gpus.txt
python - <<'P' > $O/r02q_shard_n2.log 2>&1
import json, os, sys, tempfile, time
sys.path.insert(0, "tests"); sys.path.insert(0, "tests/golden")
import make_hm_md5 as gold
from thevc_b200.host import shard_encode as se
case = "ra_1080_66_idr"
cfg, w, h, frames, extra = gold.CASES[case]
with tempfile.TemporaryDirectory() as d:
    yuv = os.path.join(d, "in.yuv"); gold.write_yuv(yuv, w, h, frames)
    r = se.shard_encode(os.path.join(gold.CFG, cfg), yuv, w, h, frames, 2, os.path.join(d, "out.bin"), gpus=[0, 1],
                        hm="me,frac,tables,frame,candgrid,dbk,sao", extra=["--SEIpictureDigest=1"] + list(extra), workdir=d)
g = json.load(open("tests/golden/hm_md5.json"))[case]
r.update({"case": case, "cfg": cfg, "width": w, "height": h, "extra": extra, "golden_md5": g["md5"], "md5_equal": r["md5"] == g["md5"],
          "golden_reference_wall_s_1core": g["reference_wall_s_this_container_1core"], "mode": "2 shards over 2 GPUs, host concatenation, no collective"})
json.dump(r, open("gpurun_out/matrix/ra_1080_66_idr_n2.json", "w"), indent=1)
print(json.dumps(r))
assert r["md5_equal"]
P
echo "shard rc=$?"; tail -c 700 $O/r02q_shard_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02q_bench_n2.json 2> $O/r02q_bench_n2.err; echo "bench2 rc=$?"
tail -c 400 $O/r02q_bench_n2.json
