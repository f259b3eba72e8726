#!/bin/bash
# 1080p LDP encode: unmodified reference vs the hooked encoder (ME by census look-up); prints wall times and md5s
set -e
cd "$(dirname "$0")/.."
N=${1:-3}
D=gpurun_out/hm1080; mkdir -p $D
python - <<PY
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, synth
seq = synth.make_sequence(1920,1080,$N)
with open('$D/in.yuv','wb') as f:
    for y,u,v in seq:
        f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
PY
ARGS="-c build/hm/cfg/${CFG:-encoder_lowdelay_P_main.cfg} -i $D/in.yuv -wdt 1920 -hgt 1080 -fr 30 -f $N --SEIpictureDigest=1"
( time oracle/_ref/bin/TAppEncoderStatic $ARGS -b $D/ref.bin > $D/ref.log ) 2> $D/ref.time &
( time env TVC_HM=${TVC_HM:-me,frac,tables} build/hm/TAppEncoderCuda $ARGS -b $D/cuda.bin > $D/cuda.log 2> $D/cuda.err ) 2> $D/cuda.time
wait
grep POC $D/ref.log | sed 's/\[MD5.*//' ; grep real $D/ref.time
grep POC $D/cuda.log | sed 's/\[MD5.*//'; grep real $D/cuda.time; grep TLibCuda $D/cuda.err
md5sum $D/ref.bin $D/cuda.bin
rm -f $D/in.yuv
