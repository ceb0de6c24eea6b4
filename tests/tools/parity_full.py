"""GPU box: parity of BOTH STFT kernels at the full BASELINE batch sizes, every clip, every value.

    python tests/tools/parity_full.py [--clips 256] [--emul 2]

For configs 2 / 3 (256 x 30 s of 0.1 randn, 80 / 128 mel) prints the worst |gpu - ref|, |gpu - f64| and |ref - f64| over
all values, how many values are more than 1e-4 from the reference, and - for the clips where the tcgen05 kernel is
furthest from float64 - how far the kernel is from the CPU emulator with an exact accumulator and with the truncating
model of the hardware accumulator (tests/emul/emul_tc.cpp).  Then the amplitude ladder (x 1e-6 ... x 32768) and a
config-4 sample (variable-length clips).  `ref` is the fp32 port of the reference's operators, per utterance.
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import asr_ttl_mtl_b200 as b  # noqa: E402
from oracle import logmel_oracle as orc  # noqa: E402
from oracle import signals  # noqa: E402


def emulator():
    lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emul", "libemul_tc.so"))
    fp = ctypes.POINTER(ctypes.c_float)
    lib.emul_tc_logmel.argtypes = [fp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, fp, fp, ctypes.c_int]

    def run(x, n_mels, model):
        lib.emul_tc_accumulate_model(model)
        f = np.ascontiguousarray(orc.reference_filters(n_mels))
        out = np.zeros((n_mels, len(x) // 160), np.float32)
        assert lib.emul_tc_logmel(x.ctypes.data_as(fp), len(x), len(x), 0, n_mels, f.ctypes.data_as(fp),
                                  out.ctypes.data_as(fp), 1) == 0
        return out

    return run


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=256)
    ap.add_argument("--emul", type=int, default=2, help="clips to run through the CPU emulator per n_mels")
    args = ap.parse_args()
    emul = emulator()
    g = torch.Generator("cuda").manual_seed(1234)
    x = (0.1 * torch.randn(args.clips, 480000, device="cuda", generator=g)).clamp_(-1, 1)
    xc = x.cpu().numpy()
    for n_mels in (80, 128):
        t0 = time.time()
        got = {v: b.log_mel_spectrogram_batch(x, n_mels=n_mels, variant=v).cpu().numpy() for v in ("tcgen05", "fft")}
        ref = orc.logmel_f32_port_per_utterance(torch.from_numpy(xc), n_mels).numpy()
        f64 = np.stack([orc.logmel_f64(c, n_mels) for c in xc])
        print(f"== config {'2' if n_mels == 80 else '3'}: {args.clips} x 30 s, n_mels={n_mels}, {ref.size / 1e6:.1f} M values "
              f"({time.time() - t0:.0f} s)")
        e_ref = np.abs(ref - f64)
        print(f"   |ref - f64|      max {e_ref.max():.3e}   values > 5e-5: {(e_ref > 5e-5).sum()}")
        # the reference's OWN operator sequence (audio.py:146-156) run on the GPU (torch.stft -> cuFFT, matmul -> cuBLAS,
        # fp32 without TF32) against its CPU run: what two fp32 evaluations of the same formulas differ by at this size
        filt = torch.from_numpy(orc.reference_filters(n_mels)).cuda()
        win = torch.hann_window(orc.N_FFT, device="cuda")
        saved = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        on_gpu = []
        for row in x:
            stft = torch.stft(row, orc.N_FFT, orc.HOP_LENGTH, window=win, return_complex=True)
            ls = torch.clamp(filt @ (stft[..., :-1].abs() ** 2), min=1e-10).log10()
            on_gpu.append(((torch.maximum(ls, ls.max() - 8.0) + 4.0) / 4.0).cpu().numpy())
        torch.backends.cuda.matmul.allow_tf32 = saved
        on_gpu = np.stack(on_gpu)
        d_self = np.abs(on_gpu - ref)
        print(f"   reference operators on cuda vs on cpu: |ref_cuda - ref_cpu| max {d_self.max():.3e} (> 1e-4: {(d_self > 1e-4).sum()} values)   "
              f"|ref_cuda - f64| max {np.abs(on_gpu - f64).max():.3e}")
        for v in ("tcgen05", "fft"):
            d_ref, d_f64 = np.abs(got[v] - ref), np.abs(got[v] - f64)
            closer = (d_f64 <= e_ref + 2.5e-5)
            print(f"   {v:8s} |gpu - ref| max {d_ref.max():.3e} (> 1e-4: {(d_ref > 1e-4).sum()} values; of those with the "
                  f"reference itself > 5e-5 from f64: {((d_ref > 1e-4) & (e_ref > 5e-5)).sum()})   |gpu - f64| max {d_f64.max():.3e} "
                  f"(> 1e-4: {(d_f64 > 1e-4).sum()}, > 5e-5: {(d_f64 > 5e-5).sum()})   gpu at least as close to f64 as ref + 2.5e-5: "
                  f"{closer.mean() * 100:.6f} %")
        per_clip = np.abs(got["tcgen05"] - f64).reshape(args.clips, -1).max(axis=1)
        for c in np.argsort(-per_clip)[:args.emul]:
            e0, e1 = emul(xc[c], n_mels, 0), emul(xc[c], n_mels, 1)
            print(f"   clip {c}: |tc - f64| {per_clip[c]:.3e}  |tc - emul(exact acc)| {np.abs(got['tcgen05'][c] - e0).max():.3e}  "
                  f"|tc - emul(RZ model)| {np.abs(got['tcgen05'][c] - e1).max():.3e}  bit-equal to the RZ model: "
                  f"{(got['tcgen05'][c] == e1).mean() * 100:.2f} %   |emul(exact) - f64| {np.abs(e0 - f64[c]).max():.3e}  "
                  f"|emul(RZ) - f64| {np.abs(e1 - f64[c]).max():.3e}  |ref - f64| {e_ref[c].max():.3e}")

    print("== every signal family at full length: 8 clips of 30 s each (80 clips), both mel counts, every value")
    kinds = list(signals.KINDS)
    clips = np.stack([signals.make_signal(k, 480000, 7000 + 10 * i + j) for i, k in enumerate(kinds) for j in range(8)])
    xk = torch.from_numpy(clips).cuda()
    for n_mels in (80, 128):
        ref = orc.logmel_f32_port_per_utterance(torch.from_numpy(clips), n_mels).numpy()
        f64 = np.stack([orc.logmel_f64(c, n_mels) for c in clips])
        for v in ("tcgen05", "fft"):
            y = b.log_mel_spectrogram_batch(xk, n_mels=n_mels, variant=v).cpu().numpy()
            per_kind = [f"{k} {np.abs(y[8 * i:8 * i + 8] - ref[8 * i:8 * i + 8]).max():.1e}" for i, k in enumerate(kinds)]
            print(f"   n_mels={n_mels} {v:8s} |gpu - ref| max {np.abs(y - ref).max():.3e}  |gpu - f64| max {np.abs(y - f64).max():.3e}  "
                  f"|ref - f64| max {np.abs(ref - f64).max():.3e}   per family: " + ", ".join(per_kind))

    print("== amplitude ladder: 4 clips of 5 s of randn x scale, 80 mel")
    for scale in (1e-6, 1e-4, 1e-2, 1.0, 100.0, 3276.8, 32768.0, 1e6):
        clips = np.stack([(scale * np.random.default_rng(50 + i).standard_normal(80000)).astype(np.float32) for i in range(4)])
        ref = orc.logmel_f32_port_per_utterance(torch.from_numpy(clips), 80).numpy()
        f64 = np.stack([orc.logmel_f64(c, 80) for c in clips])
        row = f"   x{scale:<8g} |ref - f64| {np.abs(ref - f64).max():.2e}"
        for v in ("tcgen05", "fft"):
            y = b.log_mel_spectrogram_batch(torch.from_numpy(clips).cuda(), n_mels=80, variant=v).cpu().numpy()
            row += f"   {v}: |gpu - ref| {np.abs(y - ref).max():.2e} |gpu - f64| {np.abs(y - f64).max():.2e} finite {np.isfinite(y).all()}"
        print(row)

    print("== loud / quiet mix inside one clip: 1 s at x1, 4 s at x1e-3 (quiet part within the 80 dB window), 80 mel")
    rng = np.random.default_rng(7)
    clip = np.concatenate([rng.standard_normal(16000), 1e-3 * rng.standard_normal(64000)]).astype(np.float32) * 0.1
    ref, f64 = orc.logmel_f32_port(clip, 80).numpy(), orc.logmel_f64(clip, 80)
    for v in ("tcgen05", "fft"):
        y = b.log_mel_spectrogram(torch.from_numpy(clip).cuda(), 80).cpu().numpy() if v == "tcgen05" else \
            b.log_mel_spectrogram_batch(torch.from_numpy(clip[None]).cuda(), n_mels=80, variant=v)[0].cpu().numpy()
        print(f"   {v}: |gpu - ref| {np.abs(y - ref).max():.2e}  |gpu - f64| {np.abs(y - f64).max():.2e}   |ref - f64| {np.abs(ref - f64).max():.2e}")

    print("== config 4 sample: 64 clips of 1-30 s, zero-padded to 30 s (lengths fast path), 80 mel")
    lens = signals.variable_lengths(64)
    clips = np.zeros((64, 480000), np.float32)
    for i, n in enumerate(lens):
        clips[i, :n] = signals.make_signal("gauss", int(n), 900 + i)
    ref = orc.logmel_f32_port_per_utterance(torch.from_numpy(clips), 80).numpy()
    for v in ("tcgen05", "fft"):
        y = b.log_mel_spectrogram_batch(torch.from_numpy(clips).cuda(), n_mels=80, variant=v, lengths=torch.from_numpy(lens)).cpu().numpy()
        y2 = b.log_mel_spectrogram_batch(torch.from_numpy(clips).cuda(), n_mels=80, variant=v).cpu().numpy()
        print(f"   {v}: |gpu(lengths) - ref| {np.abs(y - ref).max():.3e}   |gpu(padded rows) - ref| {np.abs(y2 - ref).max():.3e}   lengths == padded: {np.array_equal(y, y2)}")


if __name__ == "__main__":
    main()
