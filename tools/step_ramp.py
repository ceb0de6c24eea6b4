"""Per-step device time of the first steps after a synchronize (events between the steps): what the first steps of a short
timed region cost compared with the steady state.      python tools/step_ramp.py     # env: STEPS (40), IDLE_MS (0)"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import asr_ttl_mtl_b200 as b200  # noqa: E402

dev = "cuda:0"
steps = int(os.environ.get("STEPS", 40))
g = torch.Generator(device=dev).manual_seed(1234)
inputs = [(0.1 * torch.randn(256, 480000, generator=g, device=dev)).clamp_(-1, 1) for _ in range(2)]
outs = [torch.empty(256, 80, 3000, device=dev) for _ in range(2)]
for i in range(5):
    b200.log_mel_spectrogram_batch(inputs[i & 1], out=outs[i & 1])
for idle_ms in (0.0, 5.0, 50.0):
    torch.cuda.synchronize()
    time.sleep(idle_ms / 1e3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        b200.log_mel_spectrogram_batch(inputs[i & 1], out=outs[i & 1])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    print(f"idle {idle_ms:5.1f} ms before: first 8 steps " + " ".join(f"{m:.3f}" for m in ms[:8]) +
          f" | mean of 9..{steps} {sum(ms[8:]) / len(ms[8:]):.4f} | total/steps {sum(ms) / steps:.4f}")
