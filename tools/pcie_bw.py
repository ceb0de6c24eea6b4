"""GPU box: pinned-memory PCIe bandwidth, each direction alone and both at once (bounds the host-buffer `e2e` number)."""
import time, torch
n_in, n_out = 491_520_000, 245_760_000
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
t = timed(lambda: d_in.copy_(h_in, non_blocking=True)); print(f"H2D alone  {n_in/t/1e9:.1f} GB/s ({t*1e3:.2f} ms)")
t = timed(lambda: h_out.copy_(d_out, non_blocking=True)); print(f"D2H alone  {n_out/t/1e9:.1f} GB/s ({t*1e3:.2f} ms)")
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
t = timed(both); print(f"both at once: {t*1e3:.2f} ms for 491.5 MB in + 245.8 MB out -> {256*30/3600/t:.0f} audio-hours/s ceiling")
