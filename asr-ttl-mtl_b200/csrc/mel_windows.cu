// The window cut of the decoding loop behind the front-end (reference whisper/transcribe.py:282-286, :150):
//
//     mel_segment = pad_or_trim(mel[:, seek : seek + segment_size], N_FRAMES).to(device).to(dtype)
//
// for MANY windows of one long utterance in one launch: out[w, m, j] = mel[m, seek[w] + j] for j < size[w], else 0 -
// straight into zero-padded [n_windows, n_mels, window_frames] windows, float32 or IEEE half (the rounding of
// .to(torch.float16)).  Pure data movement, HBM bound: every thread moves four consecutive frames (the source is read
// with scalar, warp-coalesced loads because `seek` is arbitrary; the destination with one 16- or 8-byte store).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace b200mel {

namespace {

constexpr int kWinThreads = 256;

template <typename OutT> struct Out4;
template <> struct Out4<float> {
    static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
    static __device__ __forceinline__ void store1(float* p, float a) { *p = a; }
};
template <> struct Out4<__half> {
    static __device__ __forceinline__ void store(__half* p, float a, float b, float c, float d) {
        union { __half2 h[2]; uint2 u; } v;
        v.h[0] = __floats2half2_rn(a, b);
        v.h[1] = __floats2half2_rn(c, d);
        *reinterpret_cast<uint2*>(p) = v.u;
    }
    static __device__ __forceinline__ void store1(__half* p, float a) { *p = __float2half_rn(a); }
};

template <typename OutT>
__global__ void __launch_bounds__(kWinThreads) mel_windows_kernel(const float* __restrict__ mel, int64_t n_frames, const int32_t* __restrict__ seeks,
                                                                   const int32_t* __restrict__ sizes, int window_frames, OutT* __restrict__ out, int vector_ok) {
    const int w = blockIdx.z, m = blockIdx.y;
    const int64_t seek = seeks[w];
    int64_t size = sizes != nullptr ? sizes[w] : window_frames;
    // what mel[:, seek : seek + size] keeps (Python slice semantics for 0 <= seek) and what pad_or_trim then trims
    if (size > window_frames) size = window_frames;
    if (seek + size > n_frames) size = n_frames - seek;
    if (seek < 0 || size < 0) size = 0;
    const float* src = mel + static_cast<int64_t>(m) * n_frames + seek;
    OutT* dst = out + (static_cast<int64_t>(w) * gridDim.y + m) * window_frames;
    const int j0 = 4 * (blockIdx.x * kWinThreads + threadIdx.x);
    if (j0 >= window_frames) return;
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (j0 + i < size) ? __ldg(src + j0 + i) : 0.0f;
    if (vector_ok && j0 + 4 <= window_frames) {
        Out4<OutT>::store(dst + j0, v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (j0 + i < window_frames) Out4<OutT>::store1(dst + j0 + i, v[i]);
    }
}

}  // namespace

cudaError_t launch_mel_windows(const float* mel, int n_mels, int64_t n_frames, const int32_t* seeks, const int32_t* sizes, int n_windows,
                               int window_frames, void* out, int out_f16, cudaStream_t stream) {
    if (n_windows <= 0 || n_mels <= 0 || window_frames <= 0) return cudaSuccess;
    const dim3 grid((window_frames + 4 * kWinThreads - 1) / (4 * kWinThreads), n_mels, n_windows);
    const size_t elem = out_f16 ? 2 : 4;
    const int vector_ok = (window_frames % 4 == 0 && reinterpret_cast<uintptr_t>(out) % (4 * elem) == 0) ? 1 : 0;
    ProfileScope profile(3, stream);
    if (out_f16)
        mel_windows_kernel<__half><<<grid, kWinThreads, 0, stream>>>(mel, n_frames, seeks, sizes, window_frames, static_cast<__half*>(out), vector_ok);
    else
        mel_windows_kernel<float><<<grid, kWinThreads, 0, stream>>>(mel, n_frames, seeks, sizes, window_frames, static_cast<float*>(out), vector_ok);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200mel
