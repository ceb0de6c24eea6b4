for f in 0 16 32 48 128 144 176; do B200MEL_LIB=asr-ttl-mtl_b200/lib/libb200mel_switches.so B200MEL_TC_FLAGS=$f python tools/tc_trace.py 2>/dev/null | tail -1; done
