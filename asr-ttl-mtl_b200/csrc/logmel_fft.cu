// FFT variant of the fused log-mel front-end for sm_100a.
//
// One CTA = 320 threads = 16 frame pairs = 32 frames of one utterance.  A pass fuses
// reflect-padded framing, the Hann window, a 400-point FFT per frame pair (two real
// frames ride one complex FFT), the power spectrum, the banded mel projection,
// log10 with the 1e-10 clamp and the per-utterance max (warp REDUX + one atomicMax
// per CTA).  Reference: whisper/audio.py:145-155.  The (max-8, (x+4)/4) step of
// audio.py:155-156 needs the finished max and runs as the second kernel below, over
// data that is still L2-resident (the host launches in L2-sized chunks).
#include <cuda_runtime.h>

#include "kernels.h"

namespace b200mel {

namespace {

template <typename InT> __device__ __forceinline__ float load_sample(const InT* p);
template <> __device__ __forceinline__ float load_sample<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_sample<int16_t>(const int16_t* p) {
    return static_cast<float>(__ldg(p)) * (1.0f / 32768.0f);  // audio.py:62
}

constexpr int kWarps = kThreads / 32;

// Shared memory carve-up (bytes): audio tile | per-pair scratch | twiddles | out tile | mel taps | bands
__host__ __device__ constexpr size_t fft_smem_bytes(int n_mels) {
    return sizeof(float) * kAudioTile + sizeof(float2) * kGroups * kGroupStride + sizeof(float2) * kNFFT +
           sizeof(float) * n_mels * kOutStride + sizeof(float) * kMaxMelWeights + sizeof(int) * kMaxMels;
}

template <typename InT>
__global__ void __launch_bounds__(kThreads, 2) logmel_fft_pass1_kernel(const LogmelArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_audio = reinterpret_cast<float*>(smem_raw);
    float2* s_work = reinterpret_cast<float2*>(s_audio + kAudioTile);
    float2* s_tw = s_work + kGroups * kGroupStride;
    float* s_out = reinterpret_cast<float*>(s_tw + kNFFT);
    float* s_melw = s_out + a.n_mels * kOutStride;
    int* s_band = reinterpret_cast<int*>(s_melw + kMaxMelWeights);
    __shared__ uint32_t s_key[kWarps];

    const int tid = threadIdx.x;
    const int tiles_per_clip = (a.n_frames + kTileFrames - 1) / kTileFrames;
    const int64_t clip = blockIdx.x / tiles_per_clip;
    const int tile = blockIdx.x - static_cast<int>(clip) * tiles_per_clip;
    const int t0 = tile * kTileFrames;
    const DeviceTables* __restrict__ tab = a.tables;

    // constant operands -> shared memory / registers
    for (int i = tid; i < kNFFT; i += kThreads) s_tw[i] = tab->twiddle[i];
    for (int i = tid; i < kMaxMelWeights; i += kThreads) s_melw[i] = tab->mel_weights[i];
    for (int i = tid; i < a.n_mels; i += kThreads) s_band[i] = tab->mel_band[i];
    float win_half[kRadix];
    {
        const int j = tid % kRadix;
#pragma unroll
        for (int n1 = 0; n1 < kRadix; ++n1) win_half[n1] = tab->win_half[kRadix * n1 + j];
    }

    // stage the reflect-padded, zero-extended audio tile (coalesced loads)
    {
        const InT* __restrict__ row = static_cast<const InT*>(a.audio) + clip * a.stride_b;
        int64_t valid = a.n_samples;
        if (a.lengths != nullptr) {
            const int64_t len = a.lengths[clip];
            valid = len < 0 ? 0 : (len < valid ? len : valid);
        }
        const int64_t s0 = static_cast<int64_t>(t0) * kHop - kHalfWin;
        for (int i = tid; i < kAudioTile; i += kThreads) {
            const int64_t s = s0 + i;
            float v = 0.f;
            if (s < a.total + kHalfWin) {
                const int64_t idx = reflect_source_index(s, a.total);
                if (idx >= 0 && idx < valid) v = load_sample<InT>(row + idx);
            }
            s_audio[i] = v;
        }
    }
    __syncthreads();

    float2 r[kRadix];
    phase_fft_first(tid, s_audio, win_half, s_tw, s_work);
    __syncthreads();
    phase_fft_second_load(tid, s_work, r);
    __syncthreads();
    phase_fft_second_store(tid, r, s_work);
    __syncthreads();
    phase_power_load(tid, s_work, r);
    __syncthreads();
    phase_power_store(tid, r, s_work);
    __syncthreads();
    const int frames_valid = min(kTileFrames, a.n_frames - t0);
    uint32_t key = phase_mel_log(tid, a.n_mels, s_work, s_band, s_melw, s_out, frames_valid);

    // per-utterance max: warp REDUX, then one atomicMax per CTA
    key = __reduce_max_sync(0xffffffffu, key);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) s_key[warp] = key;
    __syncthreads();
    if (tid == 0) {
        uint32_t k = s_key[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) k = max(k, s_key[w]);
        atomicMax(a.max_keys + (a.global_max ? 0 : clip), k);
    }

    // coalesced store of the [n_mels, 32] tile: one 128-byte row segment per warp instruction
    const int t = t0 + lane;
    if (t < a.n_frames) {
        float* __restrict__ dst = a.out + (clip * a.n_mels) * static_cast<int64_t>(a.n_frames) + t;
        for (int m = warp; m < a.n_mels; m += kWarps)
            dst[static_cast<int64_t>(m) * a.n_frames] = s_out[m * kOutStride + lane];
    }
}

// Pass 2: dynamic-range clamp + affine map, in place (audio.py:155-156).
constexpr int kNormThreads = 256;
constexpr int kNormElemsPerBlock = kNormThreads * 4 * 4;  // 4 float4 per thread

__global__ void __launch_bounds__(kNormThreads) logmel_normalise_kernel(float* __restrict__ out,
                                                                         const uint32_t* __restrict__ max_keys,
                                                                         int64_t elems_per_clip, int blocks_per_clip,
                                                                         int global_max, int vec_ok) {
    const int64_t clip = blockIdx.x / blocks_per_clip;
    const int blk = blockIdx.x - static_cast<int>(clip) * blocks_per_clip;
    const float g = max_key_decode(max_keys[global_max ? 0 : clip]);
    float* base = out + clip * elems_per_clip;
    const int64_t begin = static_cast<int64_t>(blk) * kNormElemsPerBlock;
    const int64_t end = min(begin + kNormElemsPerBlock, elems_per_clip);
    if (vec_ok) {
        float4* v = reinterpret_cast<float4*>(base);
        const int64_t vb = begin / 4, ve = end / 4;  // elems_per_clip % 4 == 0 here
        for (int64_t i = vb + threadIdx.x; i < ve; i += kNormThreads) {
            float4 x = v[i];
            x.x = normalise(x.x, g); x.y = normalise(x.y, g);
            x.z = normalise(x.z, g); x.w = normalise(x.w, g);
            v[i] = x;
        }
    } else {
        for (int64_t i = begin + threadIdx.x; i < end; i += kNormThreads) base[i] = normalise(base[i], g);
    }
}

}  // namespace

cudaError_t launch_fft_pass1(const LogmelArgs& a, int dtype, cudaStream_t stream) {
    const int tiles_per_clip = (a.n_frames + kTileFrames - 1) / kTileFrames;
    const int64_t blocks = a.batch * tiles_per_clip;
    if (blocks <= 0) return cudaSuccess;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const size_t smem = fft_smem_bytes(a.n_mels);
    cudaError_t err;
    ProfileScope profile(0, stream);
    if (dtype == 0) {
        static bool attr_done_f32 = false;  // benign race: idempotent
        if (!attr_done_f32) {
            err = cudaFuncSetAttribute(logmel_fft_pass1_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(fft_smem_bytes(kMaxMels)));
            if (err != cudaSuccess) return err;
            attr_done_f32 = true;
        }
        logmel_fft_pass1_kernel<float><<<static_cast<unsigned>(blocks), kThreads, smem, stream>>>(a);
    } else {
        static bool attr_done_s16 = false;
        if (!attr_done_s16) {
            err = cudaFuncSetAttribute(logmel_fft_pass1_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(fft_smem_bytes(kMaxMels)));
            if (err != cudaSuccess) return err;
            attr_done_s16 = true;
        }
        logmel_fft_pass1_kernel<int16_t><<<static_cast<unsigned>(blocks), kThreads, smem, stream>>>(a);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_normalise(float* out, const uint32_t* max_keys, int64_t batch, int64_t elems_per_clip,
                             int global_max, cudaStream_t stream) {
    if (batch <= 0 || elems_per_clip <= 0) return cudaSuccess;
    const int blocks_per_clip = static_cast<int>((elems_per_clip + kNormElemsPerBlock - 1) / kNormElemsPerBlock);
    const int64_t blocks = batch * blocks_per_clip;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const int vec_ok = (elems_per_clip % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    ProfileScope profile(1, stream);
    logmel_normalise_kernel<<<static_cast<unsigned>(blocks), kNormThreads, 0, stream>>>(
        out, max_keys, elems_per_clip, blocks_per_clip, global_max, vec_ok);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200mel
