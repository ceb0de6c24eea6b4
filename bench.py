#!/usr/bin/env python
"""Benchmark of the log-mel front-end hot path (BASELINE.json metric: audio-hours/sec of log-mel).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n-mels 80|128] [--impl reference]

Own arm (default).  One *step* = one pass of the hot path over one batch of 256 synthetic
30 s clips (BASELINE config 2; `--n-mels 128` gives config 3), through the public API
`log_mel_spectrogram_batch`, i.e. through the C ABI.  `value` is whole-job throughput with the
inputs resident in HBM (device-timed with CUDA events on the launching stream, max over
ranks); `e2e` is the same call with HOST buffers (pinned), H2D and D2H inside the timed
region; `roofline` is the dominant kernel's own duration (events the library records around
each launch) against the measured HBM peak; `cpu_baseline` times the oracle port (the
reference's torch operators) on the host cores over a bounded sample.  Multi-GPU is utterance-
level data parallelism: every rank runs its own shard, no collective on the data path (weak
scaling: 256 clips per rank per step).

Reference arm (`--impl reference`).  Times the reference's CPU implementation of the path —
the oracle port `oracle/logmel_oracle.logmel_f32_port`, which calls the same fp32 PyTorch
operators as whisper/audio.py:146-156 (the reference is Python and cannot travel to the GPU
box) — per clip as `speech_disorder/dataset.py:82-89` does, with all host threads.
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
L2_BYTES = 126e6

THROTTLE_BITS = {
    0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
    0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
    0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
}


def bytes_per_clip(n_mels: int, in_bytes: int = 4, out_bytes: int = 4) -> int:
    """Algorithmic HBM bytes per 30 s clip: waveform read once + log-mel written once."""
    return N_SAMPLES * in_bytes + n_mels * N_FRAMES * out_bytes


def load_peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


class ClockSampler:
    """Polls NVML for SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self.error = None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))

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
        except Exception as e:
            self.error = repr(e)
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


# ---- the CPU arm: oracle port == the reference's torch operators --------------------------------
def cpu_clip_loop(n_mels: int, clips: int, seconds_cap: float, seed: int = 0):
    """Per-clip loop exactly like speech_disorder/dataset.py:82-89 (pad_or_trim -> log_mel_spectrogram)."""
    from oracle import logmel_oracle, signals

    rng_clips = [signals.make_signal("gauss", N_SAMPLES, seed + i) for i in range(min(clips, 4))]
    logmel_oracle.logmel_f32_port(rng_clips[0], n_mels)  # warm MKL plans / filter cache
    done = 0
    t0 = time.perf_counter()
    while done < clips:
        x = logmel_oracle.pad_or_trim_oracle(rng_clips[done % len(rng_clips)], N_SAMPLES)
        logmel_oracle.logmel_f32_port(x, n_mels)
        done += 1
        if time.perf_counter() - t0 > seconds_cap:
            break
    return done, time.perf_counter() - t0


def cpu_baseline(n_mels: int, seconds_cap: float = 12.0) -> dict:
    done, elapsed = cpu_clip_loop(n_mels, clips=100000, seconds_cap=seconds_cap)
    return {
        "value": done * CLIP_SECONDS / 3600.0 / elapsed,
        "unit": "audio-hours/s",
        "cores": torch.get_num_threads(),
        "host_cpus": os.cpu_count(),
        "kind": "port",
        "sample": f"{done} clips of 30 s, per-clip loop (dataset.py:82-89), {elapsed:.1f} s of CPU work, n_mels={n_mels}",
        "ms_per_clip": 1e3 * elapsed / done,
    }


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    clips_per_step = args.ref_clips_per_step
    for _ in range(args.warmup):
        cpu_clip_loop(args.n_mels, clips_per_step, 1e9)
    done = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n, _ = cpu_clip_loop(args.n_mels, clips_per_step, 1e9)
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
        "config": workload_config(args, clips_per_step, "cpu"),
        "cpu_baseline": {
            "value": value, "unit": "audio-hours/s", "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(),
            "kind": "port",
            "sample": f"{clips_per_step} clips of 30 s per step (bounded sample of the {args.batch}-clip batch), "
                      f"per-clip loop as dataset.py:82-89, torch {torch.__version__}",
        },
        "e2e": {"value": value, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch: int, where: str) -> dict:
    return {
        "workload": f"batch of {args.batch} synthetic 30 s 16 kHz clips, n_mels={args.n_mels}, fp32 "
                    f"(BASELINE config {'2' if args.n_mels == 80 else '3'})",
        "clips_per_step_per_gpu": batch,
        "n_samples": N_SAMPLES,
        "n_mels": args.n_mels,
        "parallelism": f"dp{args.gpus} (utterance shards, no collective)",
        "l2": "inputs larger than L2: 491.5 MB of waveform per step, 2 rotating input buffers" if where == "gpu"
              else "n/a (cpu)",
        "variant": args.variant,
    }


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
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    inputs = [
        (0.1 * torch.randn(B, N_SAMPLES, generator=gen, device=device, dtype=torch.float32)).clamp_(-1.0, 1.0)
        for _ in range(2)
    ]
    out_dtype = torch.float16 if args.out_dtype == "f16" else torch.float32
    out = torch.empty(B, n_mels, N_FRAMES, device=device, dtype=out_dtype)

    def step(i: int) -> None:
        b200.log_mel_spectrogram_batch(inputs[i & 1], n_mels=n_mels, out=out, variant=args.variant, out_dtype=out_dtype)

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()

    # ---- timed region: K steps, device time on the launching stream, max over ranks ----
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches_before = b200.gpu_launches()
    barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        start.record()
        for i in range(args.steps):
            step(i)
        stop.record()
        torch.cuda.synchronize()
    barrier()
    launches = b200.gpu_launches() - launches_before
    local_ms = start.elapsed_time(stop)
    total_ms = max_over_ranks(local_ms)
    total_clips = sum_over_ranks(float(B * args.steps))
    value = total_clips * CLIP_SECONDS / 3600.0 / (total_ms / 1e3)

    # ---- same K steps with the library's per-launch events: the dominant kernel's own duration ----
    _native.profile_enable(True)
    _native.profile_collect()
    for i in range(args.steps):
        step(i)
    torch.cuda.synchronize()
    prof = _native.profile_collect()
    _native.profile_enable(False)
    peak_gbs, peak_src = load_peaks()
    fused_kind = "tcgen05_pass" if prof["tcgen05_pass"][1] else "fft_pass"
    fused_ms, fused_launches = prof[fused_kind]
    norm_ms, norm_launches = prof["normalise"]
    algo_bytes_step = B * bytes_per_clip(n_mels, out_bytes=2 if args.out_dtype == "f16" else 4)
    achieved = algo_bytes_step * args.steps / (fused_ms / 1e3) / 1e9 if fused_ms > 0 else None
    # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this exact workload
    traffic, traffic_source = args.traffic_bytes, "--traffic-bytes" if args.traffic_bytes else None
    if traffic is None and fused_kind == "tcgen05_pass" and n_mels == 80 and B == DEFAULT_BATCH and args.out_dtype == "f32":
        traffic = 492.030720e6 + 217.395456e6
        traffic_source = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum (profiles/r01_tc_final_ncu_full_summary.txt)"
    roofline = {
        "bound": "hbm",
        "kernel": f"logmel_{fused_kind}",
        "achieved": achieved,
        "peak": peak_gbs,
        "unit": "GB/s",
        "frac": achieved / peak_gbs if achieved else None,
        "traffic": traffic,
        "traffic_source": traffic_source,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes_step * args.steps / max(fused_launches, 1),
        "kernel_ms_per_launch": fused_ms / max(fused_launches, 1),
        "launches": fused_launches,
        "kernel_share_of_step": fused_ms / (fused_ms + norm_ms) if fused_ms + norm_ms > 0 else None,
        "normalise_ms_per_step": norm_ms / max(args.steps, 1),
        "step_achieved": algo_bytes_step / (local_ms / args.steps / 1e3) / 1e9,
        "step_frac": algo_bytes_step / (local_ms / args.steps / 1e3) / 1e9 / peak_gbs,
    }

    # ---- end to end: host (pinned) buffers through the same public API, copies inside the timed region ----
    e2e_batch = args.e2e_batch
    host_in = torch.empty(e2e_batch, N_SAMPLES, dtype=torch.float32).pin_memory()
    host_in.copy_(inputs[0][:e2e_batch])
    host_out = torch.empty(e2e_batch, n_mels, N_FRAMES, dtype=out_dtype).pin_memory()
    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    for _ in range(2):
        b200.log_mel_spectrogram_batch(host_in, n_mels=n_mels, out=host_out, variant=args.variant, out_dtype=out_dtype)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        b200.log_mel_spectrogram_batch(host_in, n_mels=n_mels, out=host_out, variant=args.variant, out_dtype=out_dtype)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_clips = sum_over_ranks(float(e2e_batch * e2e_steps))
    e2e = {
        "value": e2e_clips * CLIP_SECONDS / 3600.0 / e2e_s,
        "unit": "audio-hours/s",
        "h2d_bytes_per_step": e2e_batch * N_SAMPLES * 4,
        "d2h_bytes_per_step": e2e_batch * n_mels * N_FRAMES * (2 if args.out_dtype == "f16" else 4),
        "clips_per_step_per_gpu": e2e_batch,
        "steps": e2e_steps,
        "api": "log_mel_spectrogram_batch(pinned CPU tensor) -> b200mel_logmel_host",
        "bound": "PCIe 5.0 x16: 491.5 MB in + 245.8 MB out per step, both directions at once take 9.5 ms on this pool "
                 "(tools/pcie_bw.py) = 225 audio-hours/s",
    }
    # The same call fed with int16 PCM (what load_audio decodes before it scales by 1/32768, audio.py:62):
    # half the host-to-device bytes, bit-equal output.  Reported beside the fp32 number, not instead of it.
    host_pcm = (host_in * 32768.0).round().clamp_(-32768, 32767).to(torch.int16).pin_memory()
    for _ in range(2):
        b200.log_mel_spectrogram_batch(host_pcm, n_mels=n_mels, out=host_out, variant=args.variant, out_dtype=out_dtype)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        b200.log_mel_spectrogram_batch(host_pcm, n_mels=n_mels, out=host_out, variant=args.variant, out_dtype=out_dtype)
    torch.cuda.synchronize()
    pcm_s = max_over_ranks(time.perf_counter() - t0)
    e2e["pcm16_input"] = {"value": e2e_clips * CLIP_SECONDS / 3600.0 / pcm_s, "unit": "audio-hours/s",
                          "h2d_bytes_per_step": e2e_batch * N_SAMPLES * 2}

    if rank == 0:
        cpu = cpu_baseline(n_mels, args.cpu_seconds) if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": f"audio-hours/sec log-mel ({n_mels} mel, 30 s clips)",
            "value": value,
            "unit": "audio-hours/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": workload_config(args, B, "gpu"),
            "clips_per_s": total_clips / (total_ms / 1e3),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": clocks.summary(),
        }
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
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-clips-per-step", type=int, default=16)
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
