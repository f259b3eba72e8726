#!/bin/bash
# 1080p: N shards on N GPUs (frame ranges, POC offset) vs the unmodified reference's single run; prints fps and md5 equality
#   all-intra (default):                 scratch/shard1080.sh 8 8
#   closed intra periods, random access: CFG=encoder_randomaccess_main.cfg HM=me,frac,tables,candgrid,dbk,sao EXTRA=--DecodingRefreshType=2 scratch/shard1080.sh 4 128
cd "$(dirname "$0")/.."
N=${1:-2}; F=${2:-4}; CFG=${CFG:-encoder_intra_main.cfg}; HM=${HM:-intra16,dbk,sao}; EXTRA=${EXTRA:-}
D=gpurun_out/shard1080; mkdir -p $D
python - <<PY
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, synth
seq = synth.make_sequence(1920,1080,$F)
with open('$D/in.yuv','wb') as f:
    for y,u,v in seq:
        f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
PY
( t0=$(date +%s.%N); oracle/_ref/bin/TAppEncoderStatic -c build/hm/cfg/$CFG -i $D/in.yuv -wdt 1920 -hgt 1080 -fr 30 -f $F -b $D/ref.bin -o /dev/null --SEIpictureDigest=1 $EXTRA > $D/ref.log 2>&1; python3 -c "import sys; print(float(sys.argv[1]) - float(sys.argv[2]))" $(date +%s.%N) $t0 > $D/ref.time ) &
python -m thevc_b200.host.shard_encode --cfg build/hm/cfg/$CFG -i $D/in.yuv -wdt 1920 -hgt 1080 --frames $F --shards $N -o $D/out.bin \
    --gpus $(seq -s, 0 $((N-1))) --hm $HM -- --SEIpictureDigest=1 $EXTRA > $D/shard.json
wait
echo "reference single run: $(cat $D/ref.time) s for $F frames; md5 $(md5sum < $D/ref.bin)" | tee $D/summary.txt
cat $D/shard.json | tee -a $D/summary.txt
grep -h "TLibCuda intra\|TLibCuda deblocking" $D/shard_00*.bin.log | head -4
rm -f $D/in.yuv $D/*.bin
