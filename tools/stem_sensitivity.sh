# what the stem kernel costs without its GELU (1), its stores (2), its loads (4): switches build, timing only
for f in 0 1 2 4 3 6 7; do echo -n "flags=$f: "; B200MEL_LIB=asr-ttl-mtl_b200/lib/libb200mel_switches.so B200MEL_STEM_FLAGS=$f python tools/stem_bench.py 2>/dev/null | grep "stem kernel"; done
