"""GPU box: a few small calls of every kernel (every entry point once; a quick sanity run after a kernel change).

    python tools/all_paths_probe.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import asr_ttl_mtl_b200 as b200  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(3)
x = 0.1 * torch.randn(5, 48000 + 37 * 160, generator=g, device=dev)
lens = torch.tensor([0, 201, 20000, 48000, 60000], dtype=torch.int32, device=dev)
for variant in ("tcgen05", "fft"):
    for n_mels in (80, 128):
        a = b200.log_mel_spectrogram_batch(x, n_mels=n_mels, variant=variant)
        b = b200.log_mel_spectrogram_batch(x, n_mels=n_mels, variant=variant, lengths=lens)
        c = b200.log_mel_spectrogram(x[0], n_mels, padding=480000)
        torch.cuda.synchronize()
pcm = (x * 32768).round().clamp(-32768, 32767).to(torch.int16)
p = b200.log_mel_spectrogram_batch(pcm, lengths=lens)
h = b200.log_mel_spectrogram_batch(x, out_dtype=torch.float16)
w = torch.randn(384, 80, 3, device=dev) * 0.05
bias = torch.randn(384, device=dev) * 0.05
s = b200.log_mel_encoder_stem(x, w, bias, lengths=lens)
win = b200.mel_windows(c, [0, 100, 3000], [3000, 50, 100])
host = b200.log_mel_spectrogram(np.zeros(16000, dtype=np.float32))
torch.cuda.synchronize()
print("all paths probe done", float(a.sum()), float(p.sum()), float(s.sum()), float(win.float().sum()), float(host.sum()))
