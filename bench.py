#!/usr/bin/env python
"""Benchmark of the log-mel front-end hot path (BASELINE.json metric: audio-hours/sec of log-mel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n-mels 80|128] [--impl reference]

Own arm (default).  One *step* = one pass of the hot path over one batch of 256 synthetic 30 s clips through the
public API `log_mel_spectrogram_batch`, i.e. through the C ABI.
  N = 1   BASELINE config 2 (256 x 30 s, 80 mel; `--n-mels 128` = config 3), two rotating 491.5 MB input buffers.
          `value`      whole-job throughput, inputs resident in HBM (CUDA events on the launching stream);
          `roofline`   the dominant kernel's own duration (events the library records around each launch) against the
                       measured HBM peak, with the tensor-core (FLOP) bound of the shipped variant beside it and the
                       binding one named; `traffic` = DRAM bytes per launch from the committed ncu capture
                       (profiles/traffic.json, written by tools/ncu_traffic.py), null if there is none for this workload;
          `sustained`  >= 2 s of back-to-back steps with the clock samples taken while they ran;
          `configs`    config 3 (128 mel) and config 4 (variable-length 1-30 s clips of a 1,737-clip epoch, padded rows
                       and the `lengths` fast path, batches of 16 as the trainer takes them and one 256-clip batch);
          `e2e`        the same call with HOST buffers (pinned): float32 in -> float32 mel back on the host; beside it the
                       consumer's shape (speech_disorder/trainer.py:393 wants the mel ON the GPU): int16 PCM / float32
                       in, mel stays on the device, one scalar per step read back; and the host-copy ceiling measured
                       with every rank copying at once;
          `cpu_baseline`  the reference's own `log_mel_spectrogram` (vendored unmodified into baseline/_ref by
                       baseline/vendor_reference.py; the oracle port if that copy is absent) per clip as
                       speech_disorder/dataset.py:82-89 does, on the host cores: all threads, 1 thread, and - informative -
                       the same function with device="cuda" (cuFFT + cuBLAS eager).
  N > 1   BASELINE config 5: 65,536 clips sharded data-parallel (`shard_range`), rank r generates its shard on the device in
          256-clip chunks (seed 1234 + global chunk index), one chunk per step, the `[256, 80, 3000]` CUDA batch handed
          over as the trainer's `batch['mels']`; no collective on the data path, per-GPU work per step fixed ("weak").

Reference arm (`--impl reference`).  Times the reference's CPU implementation of the path (baseline/_ref, else the
oracle port) per clip with all host threads, 256-clip steps built once outside the timed region; rank 0 alone runs it.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CLIP_SECONDS = 30.0
N_SAMPLES = 480000
N_FRAMES = 3000
DEFAULT_BATCH = 256
TOTAL_CLIPS_CONFIG5 = 65536
TILE_FRAMES = 128

THROTTLE_BITS = {
    0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
    0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
    0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
}


def bytes_per_clip(n_mels: int, in_bytes: int = 4, out_bytes: int = 4) -> int:
    """Algorithmic HBM bytes per 30 s clip: waveform read once + log-mel written once."""
    return N_SAMPLES * in_bytes + n_mels * N_FRAMES * out_bytes


def tensor_flops_per_clip() -> float:
    """Tensor-core work of the shipped (tcgen05) variant per 30 s clip: 24 tiles of 128 frames x 4 units x 20 MMAs of
    128 x 104 x 16 (6 K-steps x 3 products + 2 leftover steps), 2 flop per multiply-add."""
    tiles = -(-N_FRAMES // TILE_FRAMES)
    return tiles * 4 * 20 * (128 * 104 * 16) * 2.0


def load_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tflops": float(p["bf16_tflops"]), "tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json: hbm_gbs, bf16_tflops - the f16 MMA rate)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tflops": 2250.0, "tflops_sustained": 2250.0, "source": "fallback (B200_PROFILING.md: 6.65 TB/s, 2.25 PFLOP/s dense f16)"}


def load_traffic(kernel: str, n_mels: int, batch: int, out_dtype: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this exact workload, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            entry = json.load(f).get(f"{kernel}|{n_mels}|{batch}|{out_dtype}")
        return (float(entry["dram_bytes_per_launch"]), entry["source"]) if entry else (None, None)
    except Exception:
        return None, None


# ---- distributed helpers (also exercised by the gloo CPU test) --------------------------------
def dist_ready() -> bool:
    return torch.distributed.is_available() and torch.distributed.is_initialized()


def max_over_ranks(value: float, device="cuda") -> float:
    if not dist_ready():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cuda") -> float:
    if not dist_ready():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM)
    return float(t.item())


def barrier() -> None:
    if dist_ready():
        torch.distributed.barrier()


def shard_chunks(total_clips: int, rank: int, world: int, chunk: int):
    """Config 5: the chunk-sized pieces of rank `rank`'s contiguous shard, as (global chunk index, first clip, clips)."""
    from asr_ttl_mtl_b200 import shard_range

    begin, end = shard_range(total_clips, rank, world)
    out = []
    for first in range(begin, end, chunk):
        out.append((first // chunk, first, min(chunk, end - first)))
    return out


class ClockSampler:
    """Polls NVML for SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._handle = None
        self._nvml = None
        self.error = None

    def prepare(self):
        """NVML start-up (tens of milliseconds): done BEFORE the barrier + synchronize in front of a timed region, so that the
        GPU does not sit idle between the synchronize and the first timed launch."""
        if self._handle is not None or self.error:
            return self
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception as e:
            self.error = repr(e)
        return self

    def __enter__(self):
        self.prepare()
        if self._handle is None:
            return self
        pynvml, handle = self._nvml, self._handle

        def poll():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                    self.reasons |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle))
                except Exception as e:  # keep the bench alive; report the gap
                    self.error = repr(e)
                    return
                time.sleep(0.002)

        self._thread = threading.Thread(target=poll, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(2.0)

    def summary(self) -> dict:
        reasons = [name for bit, name in THROTTLE_BITS.items() if self.reasons & bit and name != "gpu_idle"]
        out = {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": reasons,
            "samples": len(self.samples),
        }
        if self.error:
            out["error"] = self.error
        return out


# ---- the CPU arm: the reference's own function (vendored), or the oracle port of its operators ----------------
def reference_function():
    """(callable(np.float32 [L], n_mels) -> tensor, pad_or_trim, kind): the vendored, unmodified reference
    `log_mel_spectrogram` / `pad_or_trim` (whisper/audio.py:110-157, :65-88) if baseline/_ref exists, else the oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import vendor_reference

        ref = vendor_reference.load()
    except Exception:
        ref = None
    if ref is not None:
        return (lambda x, n_mels, device=None: ref.log_mel_spectrogram(x, n_mels, device=device)), ref.pad_or_trim, "reference"
    from oracle import logmel_oracle

    return (lambda x, n_mels, device=None: logmel_oracle.logmel_f32_port(x, n_mels)), logmel_oracle.pad_or_trim_oracle, "port"


def cpu_clips(count: int, seed: int = 0) -> list:
    """`count` distinct 30 s clips of 0.1 randn (BASELINE config 1 / 2 distribution), built OUTSIDE any timed region."""
    rng = np.random.default_rng(seed)
    return [(0.1 * rng.standard_normal(N_SAMPLES)).astype(np.float32) for _ in range(count)]


def cpu_loop(fn, pad_or_trim, clips: list, n_mels: int, repeat_until_s: float = 0.0) -> tuple[int, float]:
    """Per-clip loop exactly like speech_disorder/dataset.py:82-89 (pad_or_trim -> log_mel_spectrogram); returns (clips, seconds)."""
    done = 0
    t0 = time.perf_counter()
    while True:
        for x in clips:
            fn(pad_or_trim(x, N_SAMPLES), n_mels)
            done += 1
        if time.perf_counter() - t0 >= repeat_until_s:
            break
    return done, time.perf_counter() - t0


def cpu_baseline(n_mels: int, seconds: float, with_cuda: bool) -> dict:
    fn, pad_or_trim, kind = reference_function()
    clips = cpu_clips(32)
    threads = torch.get_num_threads()
    cpu_loop(fn, pad_or_trim, clips[:4], n_mels)                                  # warm the FFT plans / filter cache
    done, elapsed = cpu_loop(fn, pad_or_trim, clips, n_mels, repeat_until_s=seconds)
    out = {
        "value": done * CLIP_SECONDS / 3600.0 / elapsed,
        "unit": "audio-hours/s",
        "cores": threads,
        "host_cpus": os.cpu_count(),
        "kind": kind,
        "sample": f"{done} clips of 30 s (32 distinct), per-clip loop (dataset.py:82-89), {elapsed:.1f} s of CPU work, n_mels={n_mels}, "
                  f"torch {torch.__version__}",
        "ms_per_clip": 1e3 * elapsed / done,
    }
    # BASELINE config 1: ONE clip, 0.1 randn(480000) from the CPU generator with seed 0, median of 20 calls after 2 warm-ups
    g0 = torch.Generator().manual_seed(0)
    one = (0.1 * torch.randn(N_SAMPLES, generator=g0)).numpy()
    for _ in range(2):
        fn(one, n_mels)
    laps = []
    for _ in range(20):
        t0 = time.perf_counter()
        fn(one, n_mels)
        laps.append(time.perf_counter() - t0)
    med = statistics.median(laps)
    out["config1"] = {"workload": "one synthetic 30 s clip, reference CPU path (BASELINE config 1)", "ms_median_of_20": 1e3 * med,
                      "value": CLIP_SECONDS / 3600.0 / med, "unit": "audio-hours/s", "cores": threads}
    # BASELINE.md §3 run A: one thread
    torch.set_num_threads(1)
    try:
        cpu_loop(fn, pad_or_trim, clips[:2], n_mels)
        d1, e1 = cpu_loop(fn, pad_or_trim, clips, n_mels, repeat_until_s=min(4.0, seconds))
        out["one_thread"] = {"value": d1 * CLIP_SECONDS / 3600.0 / e1, "ms_per_clip": 1e3 * e1 / d1, "cores": 1, "sample": f"{d1} clips"}
    finally:
        torch.set_num_threads(threads)
    # BASELINE.md §3 run D (informative): the reference function itself with device="cuda" (cuFFT + cuBLAS, eager)
    if with_cuda and kind == "reference" and torch.cuda.is_available():
        try:
            dev_clips = [torch.from_numpy(c).cuda() for c in clips]
            for x in dev_clips[:4]:
                fn(x, n_mels)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                for x in dev_clips:
                    fn(x, n_mels)
            e1.record()
            torch.cuda.synchronize()
            n = 4 * len(dev_clips)
            ms = e0.elapsed_time(e1)
            out["reference_on_cuda"] = {"value": n * CLIP_SECONDS / 3600.0 / (ms / 1e3), "ms_per_clip": ms / n,
                                        "sample": f"{n} clips already on the GPU, one eager call per clip (audio.py:143-156 with a CUDA tensor)"}
        except Exception as e:  # informative leg only
            out["reference_on_cuda"] = {"error": repr(e)}
    return out


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)            # torchrun exports OMP_NUM_THREADS=1: the arm uses every host thread at every N
    fn, pad_or_trim, kind = reference_function()
    clips_per_step = args.batch
    clips = cpu_clips(clips_per_step)          # one step = the whole 256-clip batch, built once outside the timed region
    for _ in range(min(args.warmup, 2)):
        cpu_loop(fn, pad_or_trim, clips[:16], args.n_mels)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        n, _ = cpu_loop(fn, pad_or_trim, clips, args.n_mels)
        done += n
    elapsed = time.perf_counter() - t0
    value = done * CLIP_SECONDS / 3600.0 / elapsed
    line = {
        "impl": "reference",
        "metric": f"audio-hours/sec log-mel ({args.n_mels} mel, 30 s clips)",
        "value": value,
        "unit": "audio-hours/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / max(args.steps, 1),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args, clips_per_step, "cpu", int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {
            "value": value, "unit": "audio-hours/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(),
            "kind": kind,
            "sample": f"{clips_per_step} distinct clips of 30 s per step, per-clip loop as dataset.py:82-89 "
                      f"({'baseline/_ref: the unmodified whisper/audio.py' if kind == 'reference' else 'oracle port of its torch operators'}), "
                      f"torch {torch.__version__}",
        },
        "e2e": {"value": value, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch: int, where: str, world: int) -> dict:
    if world > 1:
        workload = (f"{TOTAL_CLIPS_CONFIG5} synthetic 30 s 16 kHz clips sharded data-parallel over {world} GPUs (BASELINE config 5), "
                    f"n_mels={args.n_mels}, fp32, one {batch}-clip chunk of the rank's shard per step")
    else:
        workload = (f"batch of {batch} synthetic 30 s 16 kHz clips, n_mels={args.n_mels}, fp32 "
                    f"(BASELINE config {'2' if args.n_mels == 80 else '3'})")
    return {
        "workload": workload,
        "clips_per_step_per_gpu": batch,
        "n_samples": N_SAMPLES,
        "n_mels": args.n_mels,
        "parallelism": f"dp{world} (utterance shards, no collective)",
        "l2": "inputs larger than L2: 491.5 MB of waveform per step, a different input buffer every step" if where == "gpu"
              else "n/a (cpu)",
        "variant": args.variant,
        "arithmetic": "fp32 folds + fp32 accumulation; the DFT products as fp16 hi/lo operands, three products per value "
                      "(the compensation of 3xTF32 at the f16 MMA rate), per-32-frame power-of-two pre-scale",
    }


def timed_steps(step, steps: int, local_rank: int):
    """K steps between two events on the current stream, with clock sampling; returns (ms, clock summary)."""
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank).prepare()
    barrier()
    torch.cuda.synchronize()
    with sampler as clocks:
        start.record()
        for i in range(steps):
            step(i)
        stop.record()
        torch.cuda.synchronize()
    barrier()
    return start.elapsed_time(stop), clocks.summary()


def host_copy_ceiling(device, mbytes: int = 256, reps: int = 6) -> dict:
    """What the host link gives THIS rank while every rank copies at once: pinned H2D alone, and H2D + D2H together."""
    n = mbytes * (1 << 20)
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n // 2, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=device), torch.empty(n // 2, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)
    out = {}
    for name, both in (("h2d_alone_gbs", False), ("h2d_with_d2h_gbs", True)):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            if both:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        out[name] = reps * n / dt / 1e9
        if both:
            out["d2h_with_h2d_gbs"] = reps * (n // 2) / dt / 1e9
    return out


def run_own_arm(args) -> None:
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)

    import __graft_entry__ as entry

    entry.build()
    import asr_ttl_mtl_b200 as b200
    from asr_ttl_mtl_b200 import _native

    B, n_mels = args.batch, args.n_mels
    warmup = max(args.warmup, 3)
    out_dtype = torch.float16 if args.out_dtype == "f16" else torch.float32

    def make_chunk(seed: int, clips: int = B) -> torch.Tensor:
        gen = torch.Generator(device=device).manual_seed(seed)
        return (0.1 * torch.randn(clips, N_SAMPLES, generator=gen, device=device, dtype=torch.float32)).clamp_(-1.0, 1.0)

    if world == 1:
        # config 2 / 3: two rotating 491.5 MB buffers (> the 126 MB L2)
        inputs = [make_chunk(1234), make_chunk(1235)]
        shard_note = None
    else:
        # config 5: this rank's shard of the 65,536 clips, chunk by chunk (as many chunks as the run touches; >= 2)
        chunks = shard_chunks(TOTAL_CLIPS_CONFIG5, rank, world, B)
        touched = chunks[:max(2, min(len(chunks), args.steps + warmup, args.max_resident_chunks))]
        inputs = [make_chunk(1234 + gidx, n) for gidx, _first, n in touched]
        shard_note = {"clips_total": TOTAL_CLIPS_CONFIG5, "clips_per_rank": sum(n for _g, _f, n in chunks),
                      "chunks_per_rank": len(chunks), "chunks_resident": len(inputs), "first_clip_of_rank0": chunks[0][1] if rank == 0 else None}
    outs = [torch.empty(B, n_mels, N_FRAMES, device=device, dtype=out_dtype) for _ in range(2)]

    def step(i: int) -> torch.Tensor:
        # the `[B, n_mels, 3000]` CUDA batch the trainer takes as batch['mels'] (trainer.py:393); two rotating output buffers
        return b200.log_mel_spectrogram_batch(inputs[i % len(inputs)], n_mels=n_mels, out=outs[i & 1], variant=args.variant, out_dtype=out_dtype)

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()

    # ---- timed region: K steps, device time on the launching stream, max over ranks ----
    launches_before = b200.gpu_launches()
    local_ms, clocks = timed_steps(step, args.steps, local_rank)
    launches = b200.gpu_launches() - launches_before
    total_ms = max_over_ranks(local_ms)
    total_clips = sum_over_ranks(float(B * args.steps))
    value = total_clips * CLIP_SECONDS / 3600.0 / (total_ms / 1e3)

    # ---- same K steps with the library's per-launch events: the dominant kernel's own duration ----
    peaks = load_peaks()

    def kernel_profile(fn, steps: int, mels: int, batch: int, out_bytes: int, step_ms: float) -> dict:
        """The dominant kernel's duration.  `step_ms`: per-step time of the timed region (events around K back-to-back steps on
        the launching stream).  A second pass of the same K steps with the library's events around EVERY launch gives each
        kernel's share of a step - and a bracketed duration that includes the launch latency a kernel sees when an event sits
        in front of it (a few microseconds: it can exceed the whole back-to-back step).  The kernel's duration is therefore
        taken as min(bracketed, step_ms x share)."""
        _native.profile_enable(True)
        _native.profile_collect()
        for i in range(steps):
            fn(i)
        torch.cuda.synchronize()
        prof = _native.profile_collect()
        _native.profile_enable(False)
        kind = "tcgen05_pass" if prof["tcgen05_pass"][1] else "fft_pass"
        ms, n = prof[kind]
        norm_ms, _ = prof["normalise"]
        algo = batch * bytes_per_clip(mels, out_bytes=out_bytes)
        share = ms / (ms + norm_ms) if ms + norm_ms > 0 else None
        bracketed = ms / max(n, 1)
        in_step = step_ms * share if share else bracketed
        kernel_ms = min(bracketed, in_step) if bracketed > 0 else in_step
        achieved = algo / (kernel_ms / 1e3) / 1e9 if kernel_ms > 0 else None
        rec = {
            "bound": "hbm", "kernel": f"logmel_{kind}", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"] if achieved else None,
            "algorithmic_bytes_per_launch": algo, "kernel_ms_per_launch": kernel_ms, "launches": n,
            "kernel_ms_bracketed": bracketed, "kernel_ms_in_step": in_step,
            "kernel_ms_method": "min(events around each launch, per-step time of the timed region x the kernel's share of the per-launch event times)",
            "kernel_share_of_step": share,
            "peak_source": peaks["source"],
        }
        if kind == "tcgen05_pass" and kernel_ms > 0:
            flops = batch * tensor_flops_per_clip()
            tf = flops / (kernel_ms / 1e3) / 1e12
            rec["tensor"] = {"bound": "tensor", "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": tf / peaks["tflops"],
                             "flops_per_launch": flops,
                             "note": "f16 MMAs actually issued (3 products per value); peak = measured cuBLAS bf16 rate"}
            # which bound allows fewer clips per second: algorithmic bytes at the HBM peak, or the issued MMAs at the tensor peak
            t_hbm, t_tensor = algo / (peaks["hbm_gbs"] * 1e9), flops / (peaks["tflops"] * 1e12)
            rec["binding"] = "tensor" if t_tensor > t_hbm else "hbm"
            rec["frac_of_binding"] = max(t_hbm, t_tensor) / (kernel_ms / 1e3)
        return rec

    roofline = kernel_profile(step, args.steps, n_mels, B, 2 if args.out_dtype == "f16" else 4, local_ms / args.steps)
    traffic, traffic_source = (args.traffic_bytes, "--traffic-bytes") if args.traffic_bytes else load_traffic(
        roofline["kernel"].replace("logmel_", ""), n_mels, B, args.out_dtype)
    roofline["traffic"] = traffic
    roofline["traffic_source"] = traffic_source
    roofline["step_achieved"] = B * bytes_per_clip(n_mels, out_bytes=2 if args.out_dtype == "f16" else 4) / (local_ms / args.steps / 1e3) / 1e9
    roofline["step_frac"] = roofline["step_achieved"] / peaks["hbm_gbs"]

    extra = {}
    if world == 1 and not args.quick:
        # ---- the other single-GPU configs of BASELINE.json ----
        configs = {}
        other = 128 if n_mels == 80 else 80
        out_o = torch.empty(B, other, N_FRAMES, device=device)

        def step_other(i: int):
            return b200.log_mel_spectrogram_batch(inputs[i & 1], n_mels=other, out=out_o, variant=args.variant)

        for i in range(3):
            step_other(i)
        ms_o, clocks_o = timed_steps(step_other, args.steps, local_rank)
        r_o = kernel_profile(step_other, args.steps, other, B, 4, ms_o / args.steps)
        configs[f"config{'3' if other == 128 else '2'}"] = {
            "workload": f"batch of {B} synthetic 30 s clips, n_mels={other}", "value": B * args.steps * CLIP_SECONDS / 3600.0 / (ms_o / 1e3),
            "unit": "audio-hours/s", "ms_per_step": ms_o / args.steps, "roofline_frac": r_o["frac"], "kernel_ms_per_launch": r_o["kernel_ms_per_launch"],
            "tensor_frac": (r_o.get("tensor") or {}).get("frac"), "clocks": clocks_o}
        del out_o

        # config 4: 1-30 s clips (U{16000..480000} samples, default_rng(4321)) of a 1,737-clip epoch, zero-padded to 30 s
        lens = np.random.default_rng(4321).integers(16000, N_SAMPLES + 1, size=1737).astype(np.int64)
        first = torch.from_numpy(lens[:B].astype(np.int32)).to(device)
        var = inputs[0].clone()
        var.masked_fill_(torch.arange(N_SAMPLES, device=device)[None, :] >= first[:, None], 0.0)   # what pad_or_trim hands over
        real_s = float(lens[:B].sum()) / 16000.0
        out_v = torch.empty(B, n_mels, N_FRAMES, device=device)
        c4 = {"workload": "variable-length 1-30 s clips zero-padded to 30 s (BASELINE config 4), first 256 clips of the 1,737-clip epoch",
              "mean_clip_seconds": real_s / B}
        for name, kw, rows in (("padded_rows_256", {}, B), ("lengths_256", {"lengths": first}, B),
                               ("lengths_batches_of_16", {"lengths": first}, 16)):
            def step_v(i: int, kw=kw, rows=rows):
                for b0 in range(0, B, rows):
                    k2 = {k: v[b0:b0 + rows] for k, v in kw.items()}
                    b200.log_mel_spectrogram_batch(var[b0:b0 + rows], n_mels=n_mels, out=out_v[b0:b0 + rows], variant=args.variant, **k2)
            for i in range(3):
                step_v(i)
            n_v = max(3, args.steps // 2)
            ms_v, _ = timed_steps(step_v, n_v, local_rank)
            c4[name] = {"ms_per_256_clips": ms_v / n_v, "real_audio_hours_per_s": real_s * n_v / 3600.0 / (ms_v / 1e3),
                        "padded_audio_hours_per_s": B * n_v * CLIP_SECONDS / 3600.0 / (ms_v / 1e3)}
        # the trainer's batches of 16 are launch-bound (three launches of ~35 us of work each): the same sixteen calls captured
        # once in a CUDA graph and replayed
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph):
            for b0 in range(0, B, 16):
                b200.log_mel_spectrogram_batch(var[b0:b0 + 16], n_mels=n_mels, out=out_v[b0:b0 + 16], variant=args.variant, lengths=first[b0:b0 + 16])
        for _ in range(3):
            graph.replay()
        n_g = max(3, args.steps // 2)
        ms_g, _ = timed_steps(lambda i: graph.replay(), n_g, local_rank)
        c4["lengths_batches_of_16_cuda_graph"] = {"ms_per_256_clips": ms_g / n_g, "real_audio_hours_per_s": real_s * n_g / 3600.0 / (ms_g / 1e3),
                                                  "padded_audio_hours_per_s": B * n_g * CLIP_SECONDS / 3600.0 / (ms_g / 1e3)}
        del graph
        configs["config4"] = c4
        del var, out_v

        # config 1, this arm: ONE 30 s clip through the drop-in call the unchanged dataset path makes (dataset.py:82-89):
        # NumPy in, CPU tensor out - H2D, kernel and D2H of a single utterance, wall clock per call
        g0 = torch.Generator().manual_seed(0)
        one = (0.1 * torch.randn(N_SAMPLES, generator=g0)).numpy()
        for _ in range(3):
            b200.log_mel_spectrogram(one, n_mels)
        laps = []
        for _ in range(30):
            t0 = time.perf_counter()
            b200.log_mel_spectrogram(one, n_mels)
            laps.append(time.perf_counter() - t0)
        one_dev = torch.from_numpy(one).to(device)
        for _ in range(3):
            b200.log_mel_spectrogram(one_dev, n_mels)
        torch.cuda.synchronize()
        dev_laps = []
        for _ in range(30):
            t0 = time.perf_counter()
            b200.log_mel_spectrogram(one_dev, n_mels)
            torch.cuda.synchronize()
            dev_laps.append(time.perf_counter() - t0)
        configs["config1"] = {
            "workload": "one synthetic 30 s clip through log_mel_spectrogram (BASELINE config 1, this arm)",
            "host_in_host_out_ms_median_of_30": 1e3 * statistics.median(laps),
            "device_in_device_out_ms_median_of_30": 1e3 * statistics.median(dev_laps),
            "value": CLIP_SECONDS / 3600.0 / statistics.median(laps), "unit": "audio-hours/s"}
        if n_mels == 80:
            # the step behind the path: the encoder stem of model.py:193-197 on the same 256 clips, n_state 384
            # (Whisper tiny's) - conv1 + GELU, conv2 (stride 2) + GELU, permute, positional embedding; torch's cudnn pair beside it
            import torch.nn.functional as F

            n_state = 384
            gs = torch.Generator(device=device).manual_seed(77)
            w1 = (torch.rand(n_state, 80, 3, generator=gs, device=device) * 2 - 1) / 240 ** 0.5
            b1 = (torch.rand(n_state, generator=gs, device=device) * 2 - 1) / 240 ** 0.5
            w2 = (torch.rand(n_state, n_state, 3, generator=gs, device=device) * 2 - 1) / (3 * n_state) ** 0.5
            b2 = (torch.rand(n_state, generator=gs, device=device) * 2 - 1) / (3 * n_state) ** 0.5
            pos = torch.rand(N_FRAMES // 2, n_state, generator=gs, device=device)
            packed = b200.pack_conv2_weight(w2)
            mel_s = b200.log_mel_spectrogram_batch(inputs[0], n_mels=80)

            def torch_stem(i: int):
                y = F.gelu(F.conv1d(mel_s, w1, b1, padding=1))
                return F.gelu(F.conv1d(y, w2, b2, stride=2, padding=1)).permute(0, 2, 1) + pos

            stem_steps = max(3, min(args.steps, 20))
            rows_s = {}
            for name, fn in (("encoder_stem2_ms", lambda i: b200.encoder_stem2(mel_s, w1, b1, packed, b2, pos)),
                             ("log_mel_encoder_stem2_ms", lambda i: b200.log_mel_encoder_stem2(inputs[i & 1], w1, b1, packed, b2, pos)),
                             ("torch_cudnn_tf32_stem_ms", torch_stem)):
                for i in range(3):
                    fn(i)
                ms_s, _ = timed_steps(fn, stem_steps, local_rank)
                rows_s[name] = ms_s / stem_steps
            configs["stem"] = {
                "workload": f"encoder stem (model.py:193-197) behind the front-end, {B} clips, n_state {n_state}: mel [B, 80, 3000] -> float32 [B, 1500, {n_state}]",
                **rows_s, "conv2_gflop": 2.0 * B * (N_FRAMES // 2) * n_state * n_state * 3 / 1e9,
                "waveform_to_stem_audio_hours_per_s": B * CLIP_SECONDS / 3600.0 / (rows_s["log_mel_encoder_stem2_ms"] / 1e3),
                "arithmetic": "conv1 TF32 operands, conv2 IEEE-half operands (TF32's significand), float32 accumulation in tensor memory"}
            del mel_s, packed, pos
            torch.cuda.empty_cache()
        extra["configs"] = configs

        # ---- sustained: >= 2 s of back-to-back steps, clocks sampled while they run (last: it leaves the GPU at its power cap) ----
        est = local_ms / args.steps
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / est) + 1)
        sus_ms, sus_clocks = timed_steps(step, n_sus, local_rank)
        extra["sustained"] = {"seconds": sus_ms / 1e3, "steps": n_sus, "ms_per_step": sus_ms / n_sus,
                              "value": B * n_sus * CLIP_SECONDS / 3600.0 / (sus_ms / 1e3), "unit": "audio-hours/s", "clocks": sus_clocks}

    # ---- end to end: host (pinned) buffers through the same public API, copies inside the timed region ----
    e2e_batch = args.e2e_batch
    host_in = torch.empty(e2e_batch, N_SAMPLES, dtype=torch.float32).pin_memory()
    host_in.copy_(inputs[0][:e2e_batch])
    host_pcm = (host_in * 32768.0).round().clamp_(-32768, 32767).to(torch.int16).pin_memory()
    host_out = torch.empty(e2e_batch, n_mels, N_FRAMES, dtype=out_dtype).pin_memory()
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    out_bytes = 2 if args.out_dtype == "f16" else 4

    def round_trip(src) -> float:
        for _ in range(2):
            b200.log_mel_spectrogram_batch(src, n_mels=n_mels, out=host_out, variant=args.variant, out_dtype=out_dtype)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            b200.log_mel_spectrogram_batch(src, n_mels=n_mels, out=host_out, variant=args.variant, out_dtype=out_dtype)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    def consumer(src) -> float:
        """What the trainer needs (trainer.py:393): the batch arrives in pinned host memory, the mel stays on the GPU; one
        scalar per step (the batch mean) comes back.  Copies on a side stream, two staging buffers, so step i + 1's copy
        overlaps step i's kernel."""
        copy_stream = torch.cuda.Stream(device)
        stage = [torch.empty_like(src, device=device) for _ in range(2)]
        mels = outs if e2e_batch == B else [torch.empty(e2e_batch, n_mels, N_FRAMES, device=device, dtype=out_dtype) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        metric = torch.zeros(e2e_steps + 2, dtype=torch.float32).pin_memory()
        main = torch.cuda.current_stream(device)

        def run(n: int):
            for i in range(n):
                s = i & 1
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(free[s])
                    stage[s].copy_(src, non_blocking=True)
                    ready[s].record(copy_stream)
                main.wait_event(ready[s])
                mel = b200.log_mel_spectrogram_batch(stage[s], n_mels=n_mels, out=mels[s], variant=args.variant, out_dtype=out_dtype)
                free[s].record(main)
                metric[i].copy_(mel.float().mean() if i == n - 1 else mel[0, 0, :8].float().mean(), non_blocking=True)
        for ev in free:
            ev.record(main)
        run(2)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(e2e_steps)
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0)

    e2e_clips = sum_over_ranks(float(e2e_batch * e2e_steps))

    def rate(seconds: float) -> float:
        return e2e_clips * CLIP_SECONDS / 3600.0 / seconds

    e2e = {
        "value": rate(round_trip(host_in)),
        "unit": "audio-hours/s",
        "h2d_bytes_per_step": e2e_batch * N_SAMPLES * 4,
        "d2h_bytes_per_step": e2e_batch * n_mels * N_FRAMES * out_bytes,
        "clips_per_step_per_gpu": e2e_batch,
        "steps": e2e_steps,
        "api": "log_mel_spectrogram_batch(pinned CPU float32 tensor) -> b200mel_logmel_host -> pinned CPU mel (chunked H2D / compute / D2H on three streams)",
        # fed with int16 PCM (what load_audio decodes before it scales by 1/32768, audio.py:62): half the bytes in, bit-equal output
        "pcm16_input": {"value": rate(round_trip(host_pcm)), "unit": "audio-hours/s", "h2d_bytes_per_step": e2e_batch * N_SAMPLES * 2,
                        "d2h_bytes_per_step": e2e_batch * n_mels * N_FRAMES * out_bytes},
        # the consumer's shape: mel stays on the device, one scalar per step comes back
        "consumer_pcm16": {"value": rate(consumer(host_pcm)), "unit": "audio-hours/s", "h2d_bytes_per_step": e2e_batch * N_SAMPLES * 2,
                           "d2h_bytes_per_step": 4, "api": "pinned int16 batch .to(device, non_blocking) on a side stream -> log_mel_spectrogram_batch -> CUDA mel"},
        "consumer_f32": {"value": rate(consumer(host_in)), "unit": "audio-hours/s", "h2d_bytes_per_step": e2e_batch * N_SAMPLES * 4,
                         "d2h_bytes_per_step": 4},
    }
    ceiling = host_copy_ceiling(device)
    ceiling["ranks_copying_at_once"] = world
    ceiling["round_trip_f32_ceiling_audio_hours_per_s"] = world * e2e_batch * CLIP_SECONDS / 3600.0 / max(
        e2e["h2d_bytes_per_step"] / (ceiling["h2d_with_d2h_gbs"] * 1e9), e2e["d2h_bytes_per_step"] / (ceiling["d2h_with_h2d_gbs"] * 1e9))
    ceiling["consumer_pcm16_ceiling_audio_hours_per_s"] = world * e2e_batch * CLIP_SECONDS / 3600.0 / (e2e_batch * N_SAMPLES * 2 / (ceiling["h2d_alone_gbs"] * 1e9))
    e2e["host_copy_ceiling"] = ceiling

    if rank == 0:
        cpu = cpu_baseline(n_mels, args.cpu_seconds, with_cuda=True) if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": f"audio-hours/sec log-mel ({n_mels} mel, 30 s clips)",
            "value": value,
            "unit": "audio-hours/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": warmup,
            "ms_per_step": total_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, B, "gpu", world),
            "clips_per_s": total_clips / (total_ms / 1e3),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if shard_note:
            line["config"]["shard"] = shard_note
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--n-mels", type=int, default=80, choices=[80, 128])
    ap.add_argument("--batch", type=int, default=DEFAULT_BATCH, help="clips per step per GPU")
    ap.add_argument("--variant", default="auto", choices=["auto", "fft", "tcgen05"])
    ap.add_argument("--out-dtype", default="f32", choices=["f32", "f16"],
                    help="f16: the float32 result rounded to half (SURVEY 8 f3, what transcribe feeds the fp16 model)")
    ap.add_argument("--e2e-batch", type=int, default=DEFAULT_BATCH)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--max-resident-chunks", type=int, default=48, help="config 5: chunks of the rank's shard kept in HBM (24 GB at 48)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the sustained record and the extra configs")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel (overrides profiles/traffic.json)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
