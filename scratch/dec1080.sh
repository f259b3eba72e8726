#!/bin/bash
# 1080p LDP decode: unmodified reference decoder vs the hooked decoder with the picture batch; wall times and md5s
cd "$(dirname "$0")/.."
N=${1:-5}
D=gpurun_out/dec1080; mkdir -p $D
python - <<PY
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, synth
seq = synth.make_sequence(1920,1080,$N)
with open('$D/in.yuv','wb') as f:
    for y,u,v in seq:
        f.write(y.astype(np.uint8).tobytes()); f.write(u.astype(np.uint8).tobytes()); f.write(v.astype(np.uint8).tobytes())
PY
oracle/_ref/bin/TAppEncoderStatic -c build/hm/cfg/encoder_lowdelay_P_main.cfg -i $D/in.yuv -wdt 1920 -hgt 1080 -fr 30 -f $N --SEIpictureDigest=1 --QP=${QP:-32} -b $D/str.bin > $D/enc.log
ls -la $D/str.bin
( time oracle/_ref/bin/TAppDecoderStatic -b $D/str.bin -o $D/ref.yuv > $D/ref.log ) 2> $D/ref.time
( time env TVC_HM=tq,mc,batch build/hm/TAppDecoderCuda -b $D/str.bin -o $D/cuda.yuv > $D/cuda.log 2> $D/cuda.err ) 2> $D/cuda.time
( time env TVC_HM=tq,mc build/hm/TAppDecoderCuda -b $D/str.bin -o $D/cuda2.yuv > $D/cuda2.log 2> $D/cuda2.err ) 2> $D/cuda2.time
grep real $D/ref.time $D/cuda.time $D/cuda2.time; grep -c "(OK)" $D/ref.log $D/cuda.log $D/cuda2.log; grep TLibCuda $D/cuda.err $D/cuda2.err
md5sum $D/ref.yuv $D/cuda.yuv $D/cuda2.yuv
rm -f $D/in.yuv $D/*.yuv
