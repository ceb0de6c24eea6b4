"""GPU box: int16 PCM waveforms resident in HBM through the batched front-end."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asr_ttl_mtl_b200 as b
g = torch.Generator("cuda").manual_seed(5)
bufs = [(3277.0 * torch.randn(256, 480000, device="cuda", generator=g)).clamp_(-32768, 32767).to(torch.int16) for _ in range(2)]
for out_dtype in (torch.float32, torch.float16):
    for _ in range(3): b.log_mel_spectrogram_batch(bufs[0], out_dtype=out_dtype)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): b.log_mel_spectrogram_batch(bufs[i & 1], out_dtype=out_dtype)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    algo = 256 * (480000 * 2 + 80 * 3000 * (4 if out_dtype == torch.float32 else 2))
    print(f"int16 in, {out_dtype} out: {ms:.3f} ms per 256 clips = {256*30/3600/ms*1e3:.0f} audio-h/s, {algo/ms/1e6:.0f} GB/s algorithmic = {algo/ms/1e6/6544.7*100:.1f} % of roofline")
