// Internal launcher interface between the C ABI (b200mel_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "logmel_core.cuh"
#include "tables.h"
#include "tc_tables.h"

namespace b200mel {

struct LogmelArgs {
    const void* audio;       // device, [batch, n_samples] rows, pitch stride_b elements
    int64_t stride_b;
    int64_t n_samples;       // samples per row present in memory
    int64_t total;           // n_samples + right zero pad: the length torch.stft sees
    const int32_t* lengths;  // device int32 [batch] or nullptr
    int64_t batch;           // utterances in this launch
    int n_frames;            // T = total / 160
    int n_mels;
    float* out;              // device, [batch, n_mels, T] (IEEE half when out_f16, tcgen05 variant only)
    int out_f16;
    uint32_t* max_keys;      // device, [batch] (or [1] with global_max), order-preserving keys
    uint32_t* done_counters; // device, [batch]: epilogue warps that have finished a tile of the utterance (fused normalise:
                             // whoever brings the count to 8 x tiles normalises the utterance)
    uint32_t* tile_counter;  // device, [1]: the persistent kernel's tile queue head
    uint32_t* min_keys;      // device, [batch]: ~key of the utterance's smallest log10 value (tcgen05 variant: decides
                             // whether the dynamic-range clamp touches the utterance at all)
    uint32_t* tile_keys;     // device, [tiles of 128 frames][2] or nullptr: max key and ~min key of every tile (tcgen05)
    int global_max;
    int fused_norm;          // the normalisation happens inside the front-end kernel (no pass 2)
    int n_rows;              // rows of the mel partial-sum tile (DeviceTables::n_rows)
    const DeviceTables* tables;  // device
};

// FFT variant, one persistent launch: log10 mel + per-utterance max keys (+ in-place normalise when
// a.fused_norm).  The counters in `a` must be zero when the kernel starts.
cudaError_t launch_fft_fused(const LogmelArgs& a, int dtype, cudaStream_t stream);
// tcgen05 variant (the folded DFT as GEMMs on the tensor cores), one persistent launch: (log10 mel + 4) / 4 and the
// extremes of every utterance and tile, for launch_tc_finish (a.fused_norm is not used).  The counters and keys in `a`
// must be zero when the kernel starts.  tables: device copy of the constant matrices.
cudaError_t launch_tc_pass1(const LogmelArgs& a, const TcTables* tables, int dtype, cudaStream_t stream);
// Finish kernel of the tcgen05 variant, launched right behind launch_tc_pass1 on the same stream: applies the clamp at
// max - 8 tile by tile from the extremes pass 1 left (float32 or half output; one max per utterance or per call).
cudaError_t launch_tc_finish(const LogmelArgs& a, cudaStream_t stream);
// Code (0 = none) and CTA of a hand-over inside the tcgen05 kernel that timed out (a protocol bug: the kernel then ran to
// its end with garbage in that launch's output instead of hanging); synchronises the device.
unsigned tc_kernel_fault(unsigned* cta);

// Pass 2 (shared by all variants): out = (max(out, g - 8) + 4) / 4.
cudaError_t launch_normalise(float* out, const uint32_t* max_keys, int64_t batch, int64_t elems_per_clip,
                             int global_max, cudaStream_t stream);

// Encoder stem (stem_conv.cu): out = gelu(conv1d(x, weight, bias, kernel 3, padding 1)), x = mel clamped at max - 8 on
// load when max_keys is given (then mel is the front-end's output before launch_tc_finish).  n_mels 80, n_state % 128 == 0.
// out_fm16 != nullptr: the result goes there instead, as half [batch, n_frames rounded up to even, n_state] - frames major,
// the operand layout of launch_stem_conv2_gelu (`out` is not used).
cudaError_t launch_stem_conv1_gelu(const float* mel, const uint32_t* max_keys, const uint32_t* tile_keys, int global_max, int64_t batch,
                                   int n_frames, const float* weight, const float* bias, int n_state, float* out, void* out_fm16, cudaStream_t stream);

// Encoder stem, second layer (stem_conv2.cu): out[b, t, n] = gelu(conv1d(h, weight, bias, kernel 3, stride 2, padding 1))[b, n, t]
// (+ pos[t, n]), t < frames_padded / 2.  h_fm16: half [batch, frames_padded (even), n_state] as launch_stem_conv1_gelu leaves
// it; weight_f16: half [3, n_state, n_state] (tap, out, in); n_state % 128 == 0; out: float32, or IEEE half with out_f16.
cudaError_t launch_stem_conv2_gelu(const void* h_fm16, int64_t batch, int frames_padded, const void* weight_f16, const float* bias, const float* pos,
                                   int n_state, void* out, int out_f16, cudaStream_t stream);

// Window cut behind the front-end (mel_windows.cu): out[w, m, j] = mel[m, seeks[w] + j] for j < sizes[w] (nullptr: the whole
// window), zeros behind - transcribe.py:282-286 for n_windows windows at once, float32 or half.
cudaError_t launch_mel_windows(const float* mel, int n_mels, int64_t n_frames, const int32_t* seeks, const int32_t* sizes, int n_windows,
                               int window_frames, void* out, int out_f16, cudaStream_t stream);

uint64_t launches_so_far();
void count_launch(unsigned n = 1);

// Optional event bracketing of a launch (b200mel_profile_enable); kind indexes B200MEL_PROFILE_KINDS.
struct ProfileScope {
    ProfileScope(int kind, cudaStream_t stream);
    ~ProfileScope();
    cudaEvent_t start_ = nullptr, stop_ = nullptr;
    cudaStream_t stream_ = nullptr;
    int kind_ = 0;
};

}  // namespace b200mel
