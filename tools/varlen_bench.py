"""GPU box: BASELINE config 4 - variable-length clips (1-30 s) zero-padded to 30 s like dataset.py:85 - through the
batched front-end, (i) API-faithful (padded rows in) and (ii) with the `lengths` fast path.  Audio-hours count REAL seconds."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asr_ttl_mtl_b200 as b

B = int(os.environ.get("B", "256"))
rng = np.random.default_rng(4321)
lens = rng.integers(16000, 480001, size=B)
g = torch.Generator("cuda").manual_seed(4321)
bufs = []
for _ in range(2):
    x = (0.1 * torch.randn(B, 480000, device="cuda", generator=g)).clamp_(-1, 1)
    for i, n in enumerate(lens):
        x[i, n:] = 0.0
    bufs.append(x)
lens_t = torch.from_numpy(lens.astype(np.int32)).cuda()
real_hours = float(lens.sum()) / 16000 / 3600

def timed(fn, reps=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for variant in ("tcgen05", "fft"):
    ms_pad = timed(lambda i: b.log_mel_spectrogram_batch(bufs[i & 1], variant=variant))
    ms_len = timed(lambda i: b.log_mel_spectrogram_batch(bufs[i & 1], lengths=lens_t, variant=variant))
    a = b.log_mel_spectrogram_batch(bufs[0], variant=variant); c = b.log_mel_spectrogram_batch(bufs[0], lengths=lens_t, variant=variant)
    print(f"{variant}: padded rows {ms_pad:.3f} ms ({real_hours/ms_pad*1e3:.0f} real audio-h/s, {B*30/3600/ms_pad*1e3:.0f} padded), "
          f"lengths path {ms_len:.3f} ms ({real_hours/ms_len*1e3:.0f} real audio-h/s); equal={torch.equal(a, c)}")
