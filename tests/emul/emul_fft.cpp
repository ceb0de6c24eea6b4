// CPU choreography emulator of the FFT-variant kernel (TEST INFRASTRUCTURE, not a fallback).
//
// It runs the very phase functions the sm_100a kernel runs (csrc/logmel_core.cuh),
// with the kernel's tile geometry, shared-memory layout and barrier placement: each
// __syncthreads() of logmel_fft.cu is a loop boundary here.  The CPU test suite uses
// it to check the index maps, butterflies, reflect padding and the max/normalise
// logic against the oracle in a container that has no GPU.  The product never loads it.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../asr-ttl-mtl_b200/csrc/tables.h"

using namespace b200mel;

extern "C" int emul_fft_logmel(const float* audio, int64_t n_samples, int64_t valid, int64_t right_pad,
                               int n_mels, const float* filters, float* out, int do_normalise) {
    DeviceTables tab;
    if (build_tables(n_mels, filters, &tab) != kTablesOk) return 5;
    const int64_t total = n_samples + (right_pad > 0 ? right_pad : 0);
    if (total <= kHalfWin) return 3;
    const int n_frames = static_cast<int>(total / kHop);
    const int tiles = (n_frames + kTileFrames - 1) / kTileFrames;
    if (valid > n_samples) valid = n_samples;

    std::vector<float> s_audio(kAudioTile);
    std::vector<float2> s_work(kGroups * kGroupStride);
    float* s_P = reinterpret_cast<float*>(s_work.data());  // aliases the FFT scratch, as in the kernel
    std::vector<float> s_S(tab.n_rows * kSStride, 0.f);  // last row is the all-zero row
    std::vector<float2> regs(kThreads * kRadix);
    uint32_t clip_key = 0;

    for (int tile = 0; tile < tiles; ++tile) {
        const int t0 = tile * kTileFrames;
        const int64_t s0 = static_cast<int64_t>(t0) * kHop - kHalfWin;
        for (int i = 0; i < kAudioTile; ++i) {
            const int64_t s = s0 + i;
            float v = 0.f;
            if (s < total + kHalfWin) {
                const int64_t idx = reflect_source_index(s, total);
                if (idx >= 0 && idx < valid) v = audio[idx];
            }
            s_audio[i] = v;
        }
        for (int tid = 0; tid < kThreads; ++tid) {  // phase 1
            float win_half[kRadix];
            for (int n1 = 0; n1 < kRadix; ++n1) win_half[n1] = tab.win_half[kRadix * n1 + tid % kRadix];
            phase_fft_first<float>(tid, s_audio.data(), win_half, tab.twiddle, s_work.data());
        }
        auto R = [&](int tid) -> float2(&)[kRadix] { return *reinterpret_cast<float2(*)[kRadix]>(&regs[tid * kRadix]); };
        for (int tid = 0; tid < kThreads; ++tid) phase_fft_second_load(tid, s_work.data(), R(tid));
        for (int tid = 0; tid < kThreads; ++tid) phase_fft_second_store(tid, R(tid), s_work.data());
        for (int tid = 0; tid < kThreads; ++tid) phase_power_load(tid, s_work.data(), R(tid));
        for (int tid = 0; tid < kThreads; ++tid) phase_power_store(tid, R(tid), s_P);
        for (int tid = 0; tid < kThreads; ++tid) phase_mel_sweep(tid, s_P, tab.sweep, s_S.data());
        const int frames_valid = n_frames - t0 < kTileFrames ? n_frames - t0 : kTileFrames;
        for (int tid = 0; tid < kThreads; ++tid) {
            const uint32_t k = phase_finish(tid, n_mels, s_S.data(), tab.row_off, frames_valid, out + t0, n_frames);
            if (k > clip_key) clip_key = k;
        }
    }
    if (do_normalise) {
        const float g = max_key_decode(clip_key);
        for (int64_t i = 0; i < static_cast<int64_t>(n_mels) * n_frames; ++i) out[i] = normalise(out[i], g);
    }
    return 0;
}

// 20-point DFT on its own, for a direct check against numpy.fft
extern "C" void emul_dft20(const float* in_ri, float* out_ri) {
    float2 x[kRadix], X[kRadix];
    for (int i = 0; i < kRadix; ++i) x[i] = make_float2(in_ri[2 * i], in_ri[2 * i + 1]);
    dft20(x, X);
    for (int i = 0; i < kRadix; ++i) { out_ri[2 * i] = X[i].x; out_ri[2 * i + 1] = X[i].y; }
}

extern "C" uint32_t emul_key_encode(float v) { return max_key_encode(v); }
extern "C" float emul_key_decode(uint32_t k) { return max_key_decode(k); }
