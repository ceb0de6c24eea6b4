"""Bring-up aid (GPU box): the hand-over timeline of CTA 0 of the tcgen05 kernel, and the kernel time with the bring-up
switches of the trace build (python asr-ttl-mtl_b200/build.py --trace):

    B200MEL_LIB=asr-ttl-mtl_b200/lib/libb200mel_trace.so B200MEL_TC_TRACE=20 python tools/tc_trace.py
    B200MEL_LIB=... B200MEL_TC_FLAGS=1 python tools/tc_trace.py      # 1: fixed scale step, 2: count without the fence
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asr_ttl_mtl_b200 as b
n_mels = int(os.environ.get("N_MELS", "80"))
B = int(os.environ.get("CLIPS", "256"))
x = [(0.1 * torch.randn(B, 480000, device="cuda")) for _ in range(1 if os.environ.get("SAME") else 2)]
x = x * 2
if os.environ.get("ZEROPAD"):   # config 4: 1-30 s clips zero-padded to 30 s
    import numpy as np
    lens = torch.from_numpy(np.random.default_rng(4321).integers(16000, 480001, size=B)).cuda()
    for v in x:
        v.masked_fill_(torch.arange(480000, device="cuda")[None, :] >= lens[:, None], 0.0)
if os.environ.get("PCM"):
    x = [(v * 32768).round().clamp(-32768, 32767).to(torch.int16) for v in x]
for i in range(3):
    print("---- call", file=sys.stderr, flush=True)
    y = b.log_mel_spectrogram_batch(x[i & 1], n_mels=n_mels, variant="tcgen05")
    torch.cuda.synchronize()
if not os.environ.get("B200MEL_TC_TRACE"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    from asr_ttl_mtl_b200 import _native
    for i in range(20):
        b.log_mel_spectrogram_batch(x[i & 1], n_mels=n_mels, variant="tcgen05", out=y)
    torch.cuda.synchronize()
    # BLOCKS blocks of 200 launches (run-to-run noise on this pool is +-2 %: compare medians, not single blocks)
    blocks = []
    for _ in range(int(os.environ.get("BLOCKS", "1"))):
        _native.profile_enable(True); _native.profile_collect()
        for i in range(200):
            b.log_mel_spectrogram_batch(x[i & 1], n_mels=n_mels, variant="tcgen05", out=y)
        torch.cuda.synchronize()
        ms, n = _native.profile_collect()["tcgen05_pass"]
        blocks.append(ms / n)
    blocks.sort()
    extra = f" (min {blocks[0]:.4f}, median {blocks[len(blocks) // 2]:.4f}, max {blocks[-1]:.4f} over {len(blocks)} blocks)" if len(blocks) > 1 else ""
    print(f"flags={os.environ.get('B200MEL_TC_FLAGS', '0')} n_mels={n_mels} pcm={bool(os.environ.get('PCM'))}: {blocks[len(blocks) // 2]:.4f} ms per {B} clips (kernel events, {n} launches){extra}")
