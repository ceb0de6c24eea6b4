// FFT variant of the fused log-mel front-end for sm_100a (reference: whisper/audio.py:145-156).
//
// Persistent kernel: 2 CTAs of 320 threads per SM pull 32-frame tiles of the batch from a
// global queue (clip-major order).  Per tile a CTA
//   0. has the tile's 5360 waveform samples in shared memory — brought in by ONE bulk-TMA copy
//      (cp.async.bulk -> mbarrier) that was issued during the previous tile's math; tiles that
//      touch a clip edge (reflect padding, zero tail, `lengths`) or are not 16-byte aligned are
//      staged by a generic path instead;
//   1-3. runs 16 complex 400-point FFTs (two real frames each, 20 x 20 Cooley-Tukey with 4 x 5
//      prime-factor butterflies in registers) and splits them into 32 power spectra;
//   4. projects them onto the mel bank with a warp-uniform sweep (warp = 20 bins, lane = frame);
//   5. takes log10(max(.,1e-10)), stores the [n_mels, 32] tile with 128-byte row segments and
//      folds the tile maximum into the utterance's max key (warp REDUX + one atomicMax per warp).
// The last warp to finish an utterance (per-clip completion counter) flags it, and its CTA
// applies max(x, g-8), (x+4)/4 in place while the clip's 0.96 MB is still L2-resident, so the
// whole front-end is one launch whose DRAM traffic is the algorithmic read + write.
#include <cuda_runtime.h>

#include "kernels.h"

namespace b200mel {

namespace {

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk tensor-memory-accelerator copy (SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename InT> __device__ __forceinline__ InT load_raw(const InT* p) { return __ldg(p); }

// Where a tile's samples come from.
struct TileSource {
    int64_t clip;
    int t0;            // first frame of the tile
    int64_t s0;        // index of the tile's first sample in the zero-extended waveform (may be < 0)
    int64_t valid;     // samples of this row that are real
    bool all_zero;     // every sample of the tile is a zero of the tail: log-mel is the clamp floor
    bool bulk;         // whole tile is real, in range and 16-byte aligned: one bulk copy
};

template <typename InT>
__device__ __forceinline__ TileSource tile_source(const LogmelArgs& a, int tile, int tiles_per_clip) {
    TileSource s;
    s.clip = tile / tiles_per_clip;
    s.t0 = (tile - static_cast<int>(s.clip) * tiles_per_clip) * kTileFrames;
    s.s0 = static_cast<int64_t>(s.t0) * kHop - kHalfWin;
    s.valid = a.n_samples;
    if (a.lengths != nullptr) {
        const int64_t len = a.lengths[s.clip];
        s.valid = len < 0 ? 0 : (len < s.valid ? len : s.valid);
    }
    const int64_t s_end = s.s0 + kAudioTile;  // one past the last staged sample
    // smallest source index any staged sample maps to (right-edge reflection folds back)
    int64_t lowest = s.s0;
    if (s_end > a.total) {
        const int64_t folded = 2 * (a.total - 1) - (s_end - 1);
        lowest = folded < lowest ? folded : lowest;
    }
    s.all_zero = s.s0 >= 0 && lowest >= s.valid;
    const InT* first = static_cast<const InT*>(a.audio) + s.clip * a.stride_b + s.s0;
    s.bulk = s.s0 >= 0 && s_end <= s.valid && (reinterpret_cast<uintptr_t>(first) & 15u) == 0;
    return s;
}

// Generic staging: reflect padding at both clip ends, zeros beyond `valid`.
template <typename InT>
__device__ __forceinline__ void stage_generic(const LogmelArgs& a, const TileSource& s, InT* s_audio, int tid) {
    const InT* __restrict__ row = static_cast<const InT*>(a.audio) + s.clip * a.stride_b;
    for (int i = tid; i < kAudioTile; i += kThreads) {
        const int64_t pos = s.s0 + i;
        InT v = InT(0);
        if (pos < a.total + kHalfWin) {
            const int64_t idx = reflect_source_index(pos, a.total);
            if (idx >= 0 && idx < s.valid) v = load_raw(row + idx);
        }
        s_audio[i] = v;
    }
}

// In-place dynamic-range clamp + affine map of one finished utterance (audio.py:155-156).
__device__ __forceinline__ void normalise_clip(float* __restrict__ base, int64_t elems, float g, int tid) {
    if ((elems & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & 15u) == 0) {
        float4* v = reinterpret_cast<float4*>(base);
        const int64_t n4 = elems >> 2;
        int64_t i = tid;
        for (; i + 3 * kThreads < n4; i += 4 * kThreads) {
            float4 x0 = __ldcg(v + i), x1 = __ldcg(v + i + kThreads);
            float4 x2 = __ldcg(v + i + 2 * kThreads), x3 = __ldcg(v + i + 3 * kThreads);
            x0.x = normalise(x0.x, g); x0.y = normalise(x0.y, g); x0.z = normalise(x0.z, g); x0.w = normalise(x0.w, g);
            x1.x = normalise(x1.x, g); x1.y = normalise(x1.y, g); x1.z = normalise(x1.z, g); x1.w = normalise(x1.w, g);
            x2.x = normalise(x2.x, g); x2.y = normalise(x2.y, g); x2.z = normalise(x2.z, g); x2.w = normalise(x2.w, g);
            x3.x = normalise(x3.x, g); x3.y = normalise(x3.y, g); x3.z = normalise(x3.z, g); x3.w = normalise(x3.w, g);
            v[i] = x0; v[i + kThreads] = x1; v[i + 2 * kThreads] = x2; v[i + 3 * kThreads] = x3;
        }
        for (; i < n4; i += kThreads) {
            float4 x = __ldcg(v + i);
            x.x = normalise(x.x, g); x.y = normalise(x.y, g); x.z = normalise(x.z, g); x.w = normalise(x.w, g);
            v[i] = x;
        }
    } else {
        for (int64_t i = tid; i < elems; i += kThreads) base[i] = normalise(__ldcg(base + i), g);
    }
}

// Shared memory carve-up (bytes), all offsets 16-byte aligned.
struct SmemLayout {
    int audio, work, s_tile, twiddle, sweep, rows, total;
};
__host__ __device__ constexpr int align16(int v) { return (v + 15) & ~15; }
__host__ __device__ constexpr SmemLayout smem_layout(int sample_bytes, int n_rows) {
    SmemLayout l{};
    l.audio = 0;
    l.work = align16(kAudioTile * sample_bytes);
    l.s_tile = l.work + static_cast<int>(sizeof(float2)) * kGroups * kGroupStride;
    l.twiddle = l.s_tile + align16(static_cast<int>(sizeof(float)) * n_rows * kSStride);
    l.sweep = l.twiddle + static_cast<int>(sizeof(float2)) * kNFFT;
    l.rows = l.sweep + static_cast<int>(sizeof(MelSweepEntry)) * kUsedBins;
    l.total = l.rows + 2 * static_cast<int>(sizeof(int)) * kMaxMels;
    return l;
}
static_assert(sizeof(float) * kUsedBins * kPStride <= sizeof(float2) * kGroups * kGroupStride,
              "the power tile aliases the FFT scratch");

template <typename InT>
__global__ void __launch_bounds__(kThreads, 2) logmel_fft_fused_kernel(const LogmelArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const SmemLayout L = smem_layout(static_cast<int>(sizeof(InT)), a.n_rows);
    InT* s_audio = reinterpret_cast<InT*>(smem_raw + L.audio);
    float2* s_work = reinterpret_cast<float2*>(smem_raw + L.work);
    float* s_P = reinterpret_cast<float*>(smem_raw + L.work);  // aliases s_work (phase 3b onwards)
    float* s_S = reinterpret_cast<float*>(smem_raw + L.s_tile);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw + L.twiddle);
    MelSweepEntry* s_sweep = reinterpret_cast<MelSweepEntry*>(smem_raw + L.sweep);
    int* s_row_off = reinterpret_cast<int*>(smem_raw + L.rows);
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ int s_next_tile;
    __shared__ int s_norm_clip;
    __shared__ int s_norm_clip2;  // second slot, only used after the tile loop

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int tiles_per_clip = (a.n_frames + kTileFrames - 1) / kTileFrames;
    const int64_t total_tiles = a.batch * tiles_per_clip;
    const int64_t elems_per_clip = static_cast<int64_t>(a.n_mels) * a.n_frames;
    const DeviceTables* __restrict__ tab = a.tables;
    constexpr uint32_t kTileBytes = kAudioTile * sizeof(InT);

    // ---- one-time setup: constant operands -> shared memory / registers, mbarrier ----
    for (int i = tid; i < kNFFT; i += kThreads) s_tw[i] = tab->twiddle[i];
    for (int i = tid; i < kUsedBins; i += kThreads) s_sweep[i] = tab->sweep[i];
    for (int i = tid; i < 2 * kMaxMels; i += kThreads) s_row_off[i] = tab->row_off[i];
    if (tid < kSStride) s_S[(a.n_rows - 1) * kSStride + tid] = 0.f;  // the all-zero row (missing parts)
    float win_half[kRadix];
    {
        const int j = tid % kRadix;
        const float scale = sizeof(InT) == 2 ? (1.0f / 32768.0f) : 1.0f;  // int16 PCM: audio.py:62, exact
#pragma unroll
        for (int n1 = 0; n1 < kRadix; ++n1) win_half[n1] = tab->win_half[kRadix * n1 + j] * scale;
    }
    if (tid == 0) {
        mbar_init(&s_mbar, 1);
        s_norm_clip = -1;
        s_norm_clip2 = -1;
        const unsigned t = atomicAdd(a.tile_counter, 1u);
        s_next_tile = t < total_tiles ? static_cast<int>(t) : -1;
    }
    __syncthreads();

    int tile = s_next_tile;
    uint32_t parity = 0;
    int64_t prev_clip = -1;  // utterance of the tile whose completion this warp still has to signal
    const unsigned warps_per_clip = static_cast<unsigned>(tiles_per_clip) * kWarpsPerCta;
    TileSource src{};
    if (tile >= 0) {
        src = tile_source<InT>(a, tile, tiles_per_clip);
        if (src.bulk) {
            if (tid == 0) {
                mbar_arrive_expect_tx(&s_mbar, kTileBytes);
                bulk_copy_g2s(s_audio, static_cast<const InT*>(a.audio) + src.clip * a.stride_b + src.s0, kTileBytes, &s_mbar);
            }
        } else if (!src.all_zero) {
            stage_generic<InT>(a, src, s_audio, tid);
        }
    }
    __syncthreads();

    while (tile >= 0) {
        // thread 0 claims the tile after this one; the answer is needed only at barrier #1
        unsigned claimed = 0;
        if (tid == 0) claimed = atomicAdd(a.tile_counter, 1u);

        const int frames_valid = min(kTileFrames, a.n_frames - src.t0);
        float* const dst = a.out + src.clip * elems_per_clip + src.t0;
        uint32_t key = 0u;

        if (!src.all_zero) {
            if (src.bulk) { mbar_wait(&s_mbar, parity); parity ^= 1u; }
            phase_fft_first<InT>(tid, s_audio, win_half, s_tw, s_work);
        }
        if (tid == 0) s_next_tile = claimed < total_tiles ? static_cast<int>(claimed) : -1;
        __syncthreads();  // #1: s_audio is free, s_next_tile and s_norm_clip are published

        // stage the next tile while this one is being transformed
        const int next = s_next_tile;
        TileSource nsrc{};
        if (next >= 0) {
            nsrc = tile_source<InT>(a, next, tiles_per_clip);
            if (nsrc.bulk) {
                if (tid == 0) {
                    mbar_arrive_expect_tx(&s_mbar, kTileBytes);
                    bulk_copy_g2s(s_audio, static_cast<const InT*>(a.audio) + nsrc.clip * a.stride_b + nsrc.s0, kTileBytes, &s_mbar);
                }
            } else if (!nsrc.all_zero) {
                stage_generic<InT>(a, nsrc, s_audio, tid);
            }
        }
        // an utterance this CTA completed during the previous tile: normalise it in place
        const int norm_clip = s_norm_clip;
        if (norm_clip >= 0) {
            const float g = max_key_decode(__ldcg(a.max_keys + norm_clip));
            normalise_clip(a.out + static_cast<int64_t>(norm_clip) * elems_per_clip, elems_per_clip, g, tid);
        }

        if (!src.all_zero) {
            float2 r[kRadix];
            phase_fft_second_load(tid, s_work, r);
            __syncthreads();  // #2
            if (tid == 0 && norm_clip >= 0) s_norm_clip = -1;
            phase_fft_second_store(tid, r, s_work);
            __syncthreads();  // #3
            phase_power_load(tid, s_work, r);
            __syncthreads();  // #4
            phase_power_store(tid, r, s_P);
            __syncthreads();  // #5
            phase_mel_sweep(tid, s_P, s_sweep, s_S);
            __syncthreads();  // #6
        } else {
            __syncthreads();  // keeps s_norm_clip's reset ordered like the main path
            if (tid == 0 && norm_clip >= 0) s_norm_clip = -1;
            __syncthreads();
        }

        // Completion signal of the PREVIOUS tile, one pass late: its stores drained long ago, so the
        // fence is cheap, and the counter's old value is not needed until this tile's stores are out.
        unsigned done_before = 0;
        if (a.fused_norm && prev_clip >= 0) {
            __threadfence();  // the previous tile's rows and max are visible before it is counted
            if (lane == 0) done_before = atomicAdd(a.done_counters + prev_clip, 1u);
        }

        if (!src.all_zero) {
            key = phase_finish(tid, a.n_mels, s_S, s_row_off, frames_valid, dst, a.n_frames);
        } else if (lane < frames_valid) {
            // a tile of pure tail zeros: every value is log10(1e-10), computed by the same code
            const float floor_lg = log10_clamped(0.f);
            for (int m = warp; m < a.n_mels; m += kWarpsPerCta) dst[static_cast<int64_t>(m) * a.n_frames + lane] = floor_lg;
            key = max_key_encode(floor_lg);
        }

        // per-utterance max: warp REDUX + one atomic per warp
        key = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0) {
            atomicMax(a.max_keys + (a.global_max ? 0 : src.clip), key);
            if (a.fused_norm && prev_clip >= 0 && done_before == warps_per_clip - 1u) s_norm_clip = static_cast<int>(prev_clip);
        }
        prev_clip = src.clip;
        tile = next;
        src = nsrc;
    }

    // signal the last tile, then pick up an utterance that this CTA completed at the very end
    if (a.fused_norm && prev_clip >= 0) {
        __threadfence();
        if (lane == 0) {
            const unsigned done_before = atomicAdd(a.done_counters + prev_clip, 1u);
            if (done_before == warps_per_clip - 1u) s_norm_clip2 = static_cast<int>(prev_clip);
        }
    }
    __syncthreads();
    for (int k = 0; k < 2; ++k) {
        const int pending = k == 0 ? s_norm_clip : s_norm_clip2;
        if (pending < 0) continue;
        const float g = max_key_decode(__ldcg(a.max_keys + pending));
        normalise_clip(a.out + static_cast<int64_t>(pending) * elems_per_clip, elems_per_clip, g, tid);
    }
}

// Stand-alone pass 2: dynamic-range clamp + affine map, in place (audio.py:155-156).  Used when
// one max spans the whole call (the reference's 2-D semantics) or an utterance is too long for
// one CTA to normalise.
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

template <typename InT>
cudaError_t launch_fused(const LogmelArgs& a, cudaStream_t stream) {
    constexpr int kMaxDevices = 64;
    static int resident_by_device[kMaxDevices] = {0};  // benign race: idempotent per device
    const int smem_max = smem_layout(static_cast<int>(sizeof(InT)), kMaxSRows).total;
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err != cudaSuccess) return err;
    if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
    if (resident_by_device[device] == 0) {
        // function attributes are per device context
        err = cudaFuncSetAttribute(logmel_fft_fused_kernel<InT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (err != cudaSuccess) return err;
        int sms = 0, per_sm = 0;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return err;
        if ((err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, logmel_fft_fused_kernel<InT>, kThreads, smem_max)) != cudaSuccess) return err;
        resident_by_device[device] = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int resident_ctas = resident_by_device[device];
    const int tiles_per_clip = (a.n_frames + kTileFrames - 1) / kTileFrames;
    const int64_t tiles = a.batch * tiles_per_clip;
    const unsigned grid = static_cast<unsigned>(tiles < resident_ctas ? tiles : resident_ctas);
    const int smem = smem_layout(static_cast<int>(sizeof(InT)), a.n_rows).total;
    ProfileScope profile(0, stream);
    logmel_fft_fused_kernel<InT><<<grid, kThreads, smem, stream>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fft_fused(const LogmelArgs& a, int dtype, cudaStream_t stream) {
    const int tiles_per_clip = (a.n_frames + kTileFrames - 1) / kTileFrames;
    const int64_t tiles = a.batch * tiles_per_clip;
    if (tiles <= 0) return cudaSuccess;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    return dtype == 0 ? launch_fused<float>(a, stream) : launch_fused<int16_t>(a, stream);
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
