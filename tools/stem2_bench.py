"""Time the whole encoder stem (conv1 + GELU -> half, frames major; conv2 stride 2 + GELU + positional embedding as a
kind::f16 GEMM) against torch's cudnn pair of convolutions, per kernel through the library's own CUDA events.

    python tools/stem2_bench.py            # env: CLIPS (256), N_STATE (384), REPS (10)
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import asr_ttl_mtl_b200 as b200  # noqa: E402
from asr_ttl_mtl_b200 import _native  # noqa: E402
from stem_bench import timed  # noqa: E402


def main():
    clips = int(os.environ.get("CLIPS", 256))
    n_state = int(os.environ.get("N_STATE", 384))
    reps = int(os.environ.get("REPS", 10))
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(1)
    wave = torch.randn(clips, 480000, generator=g, device=dev) * 0.1
    w1 = (torch.rand(n_state, 80, 3, generator=g, device=dev) * 2 - 1) / 240 ** 0.5
    b1 = (torch.rand(n_state, generator=g, device=dev) * 2 - 1) / 240 ** 0.5
    w2 = (torch.rand(n_state, n_state, 3, generator=g, device=dev) * 2 - 1) / (3 * n_state) ** 0.5
    b2 = (torch.rand(n_state, generator=g, device=dev) * 2 - 1) / (3 * n_state) ** 0.5
    pos = torch.rand(1500, n_state, generator=g, device=dev)
    packed = b200.pack_conv2_weight(w2)
    mel = b200.log_mel_spectrogram_batch(wave)
    lib = _native.load()
    h1 = torch.empty(clips, 3000, n_state, dtype=torch.float16, device=dev)
    out = torch.empty(clips, 1500, n_state, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def conv1():
        _native.check(lib.b200mel_stem_conv1_gelu_fm16_device(mel.data_ptr(), None, 0, clips, 80, 3000, w1.data_ptr(), b1.data_ptr(),
                                                              n_state, h1.data_ptr(), s))

    def conv2():
        _native.check(lib.b200mel_stem_conv2_gelu_device(h1.data_ptr(), clips, 3000, packed.data_ptr(), b2.data_ptr(), pos.data_ptr(),
                                                         n_state, out.data_ptr(), 0, s))

    def conv2_plain():
        _native.check(lib.b200mel_stem_conv2_gelu_device(h1.data_ptr(), clips, 3000, packed.data_ptr(), b2.data_ptr(), None,
                                                         n_state, out.data_ptr(), 0, s))

    t1 = timed(conv1, reps)
    t2 = timed(conv2, reps)
    t2_plain = timed(conv2_plain, reps)
    out16 = torch.empty(clips, 1500, n_state, dtype=torch.float16, device=dev)
    t2_half = timed(lambda: _native.check(lib.b200mel_stem_conv2_gelu_device(
        h1.data_ptr(), clips, 3000, packed.data_ptr(), b2.data_ptr(), pos.data_ptr(), n_state, out16.data_ptr(), _native.FLAG_OUT_F16, s)), reps)
    del out16
    t_stem = timed(lambda: b200.encoder_stem2(mel, w1, b1, packed, b2, pos), reps)
    t_all = timed(lambda: b200.log_mel_encoder_stem2(wave, w1, b1, packed, b2, pos), reps)
    flops2 = 2.0 * clips * 1500 * n_state * n_state * 3
    bytes1 = mel.numel() * 4 + h1.numel() * 2
    bytes2 = h1.numel() * 2 + out.numel() * 4
    print(f"clips {clips} n_state {n_state}")
    print(f"conv1 + GELU -> half [B, 3000, {n_state}]   {t1:8.4f} ms  {bytes1 / t1 / 1e6:8.1f} GB/s")
    print(f"conv2 + GELU + pos -> [B, 1500, {n_state}]  {t2:8.4f} ms  {flops2 / t2 / 1e9:8.1f} TFLOP/s (f16 MMA)  {bytes2 / t2 / 1e6:8.1f} GB/s")
    print(f"conv2 + GELU + pos -> half                {t2_half:8.4f} ms")
    print(f"conv2 + GELU without pos                  {t2_plain:8.4f} ms")
    print(f"encoder_stem2 (mel in)              {t_stem:8.4f} ms")
    print(f"log_mel_encoder_stem2 (waveform in) {t_all:8.4f} ms  = {clips * 30 / 3600 / (t_all * 1e-3):8.1f} audio-hours/s")

    def torch_pair(x, a1, c1, a2, c2, p):
        y = F.gelu(F.conv1d(x, a1, c1, padding=1))
        return F.gelu(F.conv1d(y, a2, c2, stride=2, padding=1)).permute(0, 2, 1) + p

    torch.backends.cudnn.allow_tf32 = True
    t_tf32 = timed(lambda: torch_pair(mel, w1, b1, w2, b2, pos), reps)
    torch.backends.cudnn.allow_tf32 = False
    t_fp32 = timed(lambda: torch_pair(mel, w1, b1, w2, b2, pos), reps)
    mh, w1h, b1h, w2h, b2h, ph = (t.half() for t in (mel, w1, b1, w2, b2, pos))
    t_fp16 = timed(lambda: torch_pair(mh, w1h, b1h, w2h, b2h, ph), reps)
    print(f"torch stem (model.py:193-197) tf32  {t_tf32:8.4f} ms")
    print(f"torch stem fp32                     {t_fp32:8.4f} ms")
    print(f"torch stem fp16 (inputs already half) {t_fp16:8.4f} ms")
    got = b200.encoder_stem2(mel[:2], w1, b1, packed, b2, pos)
    want = torch_pair(mel[:2].double(), w1.double(), b1.double(), w2.double(), b2.double(), pos.double())
    print(f"|stem2 - f64| on 2 clips {(got.double() - want).abs().max().item():.2e}")


if __name__ == "__main__":
    main()
