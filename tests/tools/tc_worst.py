"""GPU box: on the full BASELINE batch (256 x 30 s, 0.1 randn), where do the two STFT variants differ most, and how far
is each from the float64 oracle and from the fp32 reference port there?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import asr_ttl_mtl_b200 as b
from oracle import logmel_oracle as orc

g = torch.Generator("cuda").manual_seed(1234)
x = (0.1 * torch.randn(256, 480000, device="cuda", generator=g)).clamp_(-1, 1)
for n_mels in (80, 128):
    tc = b.log_mel_spectrogram_batch(x, n_mels=n_mels, variant="tcgen05")
    ff = b.log_mel_spectrogram_batch(x, n_mels=n_mels, variant="fft")
    diff = (tc - ff).abs()
    per_clip = diff.flatten(1).max(dim=1).values
    worst = torch.topk(per_clip, 3).indices.tolist()
    print(f"n_mels={n_mels}: max|tc-fft| = {diff.max().item():.3e}; worst clips {worst}")
    for c in worst:
        xc = x[c].cpu().numpy()
        f64 = orc.logmel_f64(xc, n_mels)
        ref = orc.logmel_f32_port(xc, n_mels).numpy()
        e_tc = np.abs(tc[c].cpu().numpy() - f64).max(); e_ff = np.abs(ff[c].cpu().numpy() - f64).max(); e_ref = np.abs(ref - f64).max()
        r_tc = np.abs(tc[c].cpu().numpy() - ref).max(); r_ff = np.abs(ff[c].cpu().numpy() - ref).max()
        print(f"  clip {c}: |tc-f64| {e_tc:.3e}  |fft-f64| {e_ff:.3e}  |ref-f64| {e_ref:.3e}   |tc-ref| {r_tc:.3e}  |fft-ref| {r_ff:.3e}")
