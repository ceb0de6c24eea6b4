"""Bring-up aid (GPU box): run the tcgen05 variant at several batch sizes, each in its own process so that a
device fault in one does not hide the others; prints max |tc - fft| and the time per call."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, torch
sys.path.insert(0, %r)
import asr_ttl_mtl_b200 as b
B, L, M = %d, %d, %d
g = torch.Generator('cuda').manual_seed(1)
x = 0.1 * torch.randn(B, L, device='cuda', generator=g)
y = b.log_mel_spectrogram_batch(x, n_mels=M, variant='tcgen05')
torch.cuda.synchronize()
r = b.log_mel_spectrogram_batch(x, n_mels=M, variant='fft')
torch.cuda.synchronize()
err = (y - r).abs().max().item()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = b.log_mel_spectrogram_batch(x, n_mels=M, variant='tcgen05')
e1.record(); torch.cuda.synchronize()
print(f'B={B} L={L} M={M}: max|tc-fft|={err:.3e}  {e0.elapsed_time(e1)/10:.4f} ms/call', flush=True)
"""

def main():
    cases = [(1, 480000, 80), (8, 480000, 80), (24, 480000, 80), (256, 480000, 80), (256, 480000, 128), (64, 16000 * 7 + 123, 80)]
    for B, L, M in cases:
        p = subprocess.run([sys.executable, "-c", CHILD % (ROOT, B, L, M)], capture_output=True, text=True, timeout=300)
        tail = (p.stdout + p.stderr).strip().splitlines()[-4:]
        print(f"[rc={p.returncode}] " + " | ".join(tail), flush=True)

if __name__ == "__main__":
    main()
