"""Bring-up aid (GPU box): B200MEL_TC_TRACE=1 python tools/tc_trace.py  -> CTA 0's hand-over timeline (third call)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asr_ttl_mtl_b200 as b
x = (0.1 * torch.randn(256, 480000, device="cuda"))
for _ in range(3):
    print("---- call", file=sys.stderr, flush=True)
    y = b.log_mel_spectrogram_batch(x, n_mels=int(os.environ.get("N_MELS", "80")), variant="tcgen05")
    torch.cuda.synchronize()
