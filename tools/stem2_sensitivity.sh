#!/bin/bash
# What the conv2 kernel's time is made of: the switches build (python asr-ttl-mtl_b200/build.py --trace) with parts of the
# kernel taken out - B200MEL_C2_FLAGS: 1 no GELU / stores, 2 no MMAs, 4 no input-frame copies.  The results are garbage; only
# the times mean something.     bash tools/stem2_sensitivity.sh > gpurun_out/stem2_sensitivity.txt
cd "$(dirname "$0")/.."
export B200MEL_LIB=$PWD/asr-ttl-mtl_b200/lib/libb200mel_switches.so
for f in 0 1 2 4 3 5 6 7; do
  echo "flags=$f: $(B200MEL_C2_FLAGS=$f REPS=10 python tools/stem2_bench.py 2>&1 | grep 'conv2 + GELU')"
done
