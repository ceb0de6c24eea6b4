# A/B timing of library builds on one box: bash tools/ab_libs.sh name1 name2 ...   (lib/libb200mel_<name>.so; "base" = the product)
for round in 1 2; do
for name in "$@"; do
  lib=asr-ttl-mtl_b200/lib/libb200mel_$name.so; [ "$name" = base ] && lib=asr-ttl-mtl_b200/lib/libb200mel.so
  echo -n "$name: "; BLOCKS=5 B200MEL_LIB=$lib python tools/tc_trace.py 2>/dev/null | tail -1
done; done
