"""Time the encoder-stem kernel (conv1 + GELU, TF32 tensor cores) against torch's cudnn conv + gelu, and the fused
front-end + stem against the two-step path.  CUDA events, L2-sized working set (the output alone is > 1 GB).

    python tools/stem_bench.py            # env: CLIPS (256), N_STATE (384), REPS (10)
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import asr_ttl_mtl_b200 as b200  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    clips = int(os.environ.get("CLIPS", 256))
    n_state = int(os.environ.get("N_STATE", 384))
    reps = int(os.environ.get("REPS", 10))
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(1)
    wave = torch.randn(clips, 480000, generator=g, device=dev) * 0.1
    w = (torch.rand(n_state, 80, 3, generator=g, device=dev) * 2 - 1) / 240 ** 0.5
    bias = (torch.rand(n_state, generator=g, device=dev) * 2 - 1) / 240 ** 0.5
    mel = b200.log_mel_spectrogram_batch(wave)
    out = torch.empty(clips, n_state, 3000, device=dev)
    t_front = timed(lambda: b200.log_mel_spectrogram_batch(wave, out=mel), reps)
    t_stem = timed(lambda: b200.encoder_stem(mel, w, bias, out=out), reps)
    t_fused = timed(lambda: b200.log_mel_encoder_stem(wave, w, bias), reps)
    torch.backends.cudnn.allow_tf32 = True
    t_torch_tf32 = timed(lambda: F.gelu(F.conv1d(mel, w, bias, padding=1)), reps)
    torch.backends.cudnn.allow_tf32 = False
    t_torch_fp32 = timed(lambda: F.gelu(F.conv1d(mel, w, bias, padding=1)), reps)
    half = timed(lambda: F.gelu(F.conv1d(mel.half(), w.half(), bias.half(), padding=1)), reps)
    bytes_stem = mel.numel() * 4 + out.numel() * 4
    flops = 2.0 * clips * 3000 * n_state * 240
    print(f"clips {clips} n_state {n_state}")
    print(f"front-end                 {t_front:8.4f} ms")
    print(f"stem kernel               {t_stem:8.4f} ms  {bytes_stem / t_stem / 1e6:8.1f} GB/s  {flops / t_stem / 1e9:8.1f} TFLOP/s")
    print(f"front-end + stem (fused)  {t_fused:8.4f} ms  (two-step {t_front + t_stem:.4f})")
    print(f"torch conv1d+gelu tf32    {t_torch_tf32:8.4f} ms")
    print(f"torch conv1d+gelu fp32    {t_torch_fp32:8.4f} ms")
    print(f"torch conv1d+gelu fp16 (incl. casts) {half:8.4f} ms")
    err = (out[:4].double() - F.gelu(F.conv1d(mel[:4].double(), w.double(), bias.double(), padding=1))).abs().max().item()
    # what the memory system gives a stream with the stem's read : write mix (1 : 4.8): a plain fill of the output and a
    # copy of a mel-sized block next to it, on the same buffers
    t_fill = timed(lambda: out.fill_(1.0), reps)
    t_mix = timed(lambda: (out.fill_(1.0), mel.clone()), reps)
    print(f"fill of the stem's output {t_fill:8.4f} ms  {out.numel() * 4 / t_fill / 1e6:8.1f} GB/s written")
    print(f"fill + clone of the mel   {t_mix:8.4f} ms  {(out.numel() * 4 + 2 * mel.numel() * 4) / t_mix / 1e6:8.1f} GB/s")
    print(f"|stem - f64| on 4 clips {err:.2e}")


if __name__ == "__main__":
    main()
