import sys, os, torch
sys.path.insert(0, "/root/repo")
import asr_ttl_mtl_b200 as b
x = (0.1 * torch.randn(256, 480000, device="cuda"))
y = b.log_mel_spectrogram_batch(x, n_mels=80, variant="tcgen05")
torch.cuda.synchronize()
