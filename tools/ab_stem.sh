# A/B of two library builds for the stem kernels on one box: lib/libb200mel_old.so (build the other source into it) vs the product.   bash tools/ab_stem.sh
for round in 1 2 3; do
for name in old base; do
  lib=$PWD/asr-ttl-mtl_b200/lib/libb200mel_$name.so; [ "$name" = base ] && lib=$PWD/asr-ttl-mtl_b200/lib/libb200mel.so
  echo "$name: $(B200MEL_LIB=$lib REPS=30 python tools/stem_bench.py 2>/dev/null | grep 'stem kernel') | $(B200MEL_LIB=$lib REPS=30 python tools/stem2_bench.py 2>/dev/null | grep 'conv1 + GELU')"
done; done
