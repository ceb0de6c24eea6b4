// tcgen05 (tensor-core) variant of the fused log-mel front-end for sm_100a — DFT-as-GEMM with
// split-precision compensation (reference: whisper/audio.py:145-155; math in tc_core.cuh).
//
// One persistent CTA per SM, 128 frames (= 128 tensor-memory lanes = MMA M) per tile:
//   1. the tile's 20720 samples are staged in shared memory (rows of 160 at pitch 161, so the
//      frame-per-thread reads below are bank-conflict free);
//   2. 8 worker warps (two per TMEM lane quadrant) run stage 1 on the CUDA cores — windowed real
//      FFT-16 over the 25 strided sub-sequences of every frame, twiddle, fp16 hi/lo split — and
//      write the result straight into TENSOR MEMORY as the A operand (tcgen05.st, one 32-bit column
//      = one packed complex value; 8 blocks x [25 hi | 25 lo | 6 zero] columns);
//   3. one elected thread issues stage 2 on the tensor cores: per (block, N-half) unit 7 + 4
//      tcgen05.mma.kind::f16 (A from TMEM, B = the shared DFT-25 matrix [Bhi; Bhi] / [Blo] from shared
//      memory, fp32 accumulator in TMEM), i.e. hi*Bhi + lo*Bhi + hi*Blo; accumulators are double
//      buffered and handed over with tcgen05.commit -> mbarrier;
//   4. the workers read each accumulator (tcgen05.ld), form the power of 16 bins and add their mel
//      taps into a [n_mels, 128] tile in shared memory (two warps per quadrant own even / odd mels);
//   5. log10(max(.,1e-10)), 128-byte coalesced row stores, per-utterance max key.
// The (max-8, (x+4)/4) step runs as the shared pass-2 kernel.
#include <cuda_runtime.h>

#include "kernels.h"
#include "tc_core.cuh"

namespace b200mel {

namespace {

constexpr int kTcWorkerWarps = 8;
constexpr int kTcWorkers = kTcWorkerWarps * 32;   // 256
constexpr int kTcThreads = kTcWorkers + 32;       // + the MMA warp
constexpr int kTcTmemCols = 512;
constexpr int kTcStripBytes = kTcN * 16;          // one 8-half K strip of a 64-row operand

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&d)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle shared-memory operand descriptor (8 x 16 B core matrices):
// start >> 4 | (K-direction core-matrix stride >> 4) << 16 | (row-group stride >> 4) << 32 | version 1 << 46
__device__ __forceinline__ uint64_t operand_desc(uint32_t smem_addr) {
    return static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(kTcStripBytes >> 4) << 16) |
           (static_cast<uint64_t>(128 >> 4) << 32) | (1ull << 46);
}
// f16 x f16 -> f32, both operands K-major, N = 32, M = 128
constexpr uint32_t kTcIdesc = (1u << 4) | (static_cast<uint32_t>(kTcDCols >> 3) << 17) | (static_cast<uint32_t>(kTcTileFrames >> 4) << 24);

__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kTcIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct TcSmem {
    int audio, s_tile, b_main, b_corr, tap, win, tw, total;
};
__host__ __device__ constexpr int align128(int v) { return (v + 127) & ~127; }
__host__ __device__ constexpr TcSmem tc_smem(int n_mels) {
    TcSmem l{};
    l.audio = 0;
    l.s_tile = align128(kTcAudioFloats * 4);
    l.b_main = l.s_tile + align128((n_mels + 2) * kTcTileFrames * 4);
    l.b_corr = l.b_main + 2 * kTcBMainHalves * 2;
    l.tap = l.b_corr + 2 * kTcBCorrHalves * 2;
    l.win = l.tap + 2 * kTcUnits * 16 * static_cast<int>(sizeof(TcTap));
    l.tw = l.win + kTcN2 * 16 * 4;
    l.total = l.tw + kTcN2 * 8 * 8;
    return l;
}

template <typename InT> __device__ __forceinline__ float sample_to_float(InT v);
template <> __device__ __forceinline__ float sample_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float sample_to_float<int16_t>(int16_t v) { return static_cast<float>(v) * (1.0f / 32768.0f); }

template <typename InT>
__global__ void __launch_bounds__(kTcThreads, 1) logmel_tc_kernel(const LogmelArgs a, const TcTables* __restrict__ tt) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TcSmem L = tc_smem(a.n_mels);
    float* s_audio = reinterpret_cast<float*>(smem_raw + L.audio);
    float* s_S = reinterpret_cast<float*>(smem_raw + L.s_tile);
    __half* s_bmain = reinterpret_cast<__half*>(smem_raw + L.b_main);
    __half* s_bcorr = reinterpret_cast<__half*>(smem_raw + L.b_corr);
    TcTap (*s_tap)[kTcUnits][16] = reinterpret_cast<TcTap (*)[kTcUnits][16]>(smem_raw + L.tap);
    float (*s_win)[16] = reinterpret_cast<float (*)[16]>(smem_raw + L.win);
    float2 (*s_tw)[8] = reinterpret_cast<float2 (*)[8]>(smem_raw + L.tw);
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool worker = warp < kTcWorkerWarps;
    const int quad = warp & 3, parity_role = (warp >> 2) & 1;   // workers: TMEM lane quadrant, even/odd half
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    const int64_t total_tiles = a.batch * tiles_per_clip;

    // ---- one-time setup ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(static_cast<uint32_t>(kTcTmemCols)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        mbar_init(&s_empty[0], kTcWorkers); mbar_init(&s_empty[1], kTcWorkers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint32_t* src_main = reinterpret_cast<const uint32_t*>(tt->b_main);
        const uint32_t* src_corr = reinterpret_cast<const uint32_t*>(tt->b_corr);
        uint32_t* dst_main = reinterpret_cast<uint32_t*>(s_bmain);
        uint32_t* dst_corr = reinterpret_cast<uint32_t*>(s_bcorr);
        for (int i = tid; i < kTcBMainHalves; i += kTcThreads) dst_main[i] = src_main[i];   // 2 sets x halves / 2 words
        for (int i = tid; i < kTcBCorrHalves; i += kTcThreads) dst_corr[i] = src_corr[i];
        const TcTap* src_tap = &tt->tap[0][0][0];
        TcTap* dst_tap = &s_tap[0][0][0];
        for (int i = tid; i < 2 * kTcUnits * 16; i += kTcThreads) dst_tap[i] = src_tap[i];
        for (int i = tid; i < kTcN2 * 16; i += kTcThreads) (&s_win[0][0])[i] = (&tt->win[0][0])[i];
        for (int i = tid; i < kTcN2 * 8; i += kTcThreads) (&s_tw[0][0])[i] = (&tt->tw[0][0])[i];
        for (int i = tid; i < (a.n_mels + 2) * kTcTileFrames; i += kTcThreads) s_S[i] = 0.f;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // operand matrices: generic writes -> tensor-core reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);   // this warp's TMEM lane quadrant

    // zero every A column once: the 6 pad columns of each block are read (against zero B rows) and must stay finite
    if (worker && parity_role == 0) {
        for (int c = 0; c < kTcACols; ++c) tmem_st1(lane_addr + c, 0u);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

    uint32_t full_phase[2] = {0u, 0u}, empty_phase[2] = {0u, 0u};
    const uint64_t desc_main0 = operand_desc(smem_u32(s_bmain));
    const uint64_t desc_corr0 = operand_desc(smem_u32(s_bcorr));

    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int64_t clip = tile / tiles_per_clip;
        const int t0 = static_cast<int>(tile - clip * tiles_per_clip) * kTcTileFrames;

        // ---- 1. stage the audio tile (reflect padding at the clip ends, zeros beyond `valid`) ----
        {
            const InT* __restrict__ row = static_cast<const InT*>(a.audio) + clip * a.stride_b;
            int64_t valid = a.n_samples;
            if (a.lengths != nullptr) {
                const int64_t len = a.lengths[clip];
                valid = len < 0 ? 0 : (len < valid ? len : valid);
            }
            const int64_t s0 = static_cast<int64_t>(t0) * kHop - kHalfWin;
            if (sizeof(InT) == 4 && s0 >= 0 && s0 + kTcAudioSamples <= valid) {
                // interior tile: asynchronous 4-byte copies, one 160-sample row per warp pass (rows land at pitch 161)
                const InT* __restrict__ src = row + s0;
                for (int r = warp; r < kTcAudioRows; r += kTcThreads / 32) {
                    const int cols = r == kTcAudioRows - 1 ? kTcAudioSamples - kHop * (kTcAudioRows - 1) : kHop;
                    const uint32_t dst = smem_u32(s_audio + r * kTcRowPitch);
                    const InT* g = src + r * kHop;
#pragma unroll
                    for (int c = lane; c < kHop; c += 32)
                        if (c < cols) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * c), "l"(g + c) : "memory");
                }
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            } else {
                for (int r = warp; r < kTcAudioRows; r += kTcThreads / 32)
                    for (int c = lane; c < kHop; c += 32) {
                        const int i = r * kHop + c;
                        float v = 0.f;
                        if (i < kTcAudioSamples) {
                            const int64_t pos = s0 + i;
                            if (pos < a.total + kHalfWin) {
                                const int64_t idx = reflect_source_index(pos, a.total);
                                if (idx >= 0 && idx < valid) v = sample_to_float<InT>(__ldg(row + idx));
                            }
                        }
                        s_audio[r * kTcRowPitch + c] = v;
                    }
            }
        }
        __syncthreads();

        // ---- 2. stage 1 on the CUDA cores, A operand written to tensor memory ----
        if (worker) {
            const float* frame_audio = s_audio + kTcRowPitch * (quad * 32 + lane);
            const int n2_begin = parity_role == 0 ? 0 : 13, n2_end = parity_role == 0 ? 13 : kTcN2;
            for (int n2 = n2_begin; n2 < n2_end; ++n2) {
                uint32_t hi[kTcBlocks], lo[kTcBlocks];
                tc_stage1(frame_audio, n2, s_win[n2], s_tw[n2], hi, lo);
#pragma unroll
                for (int b = 0; b < kTcBlocks; ++b) {
                    tmem_st1(lane_addr + kTcBlockCols * b + n2, hi[b]);
                    tmem_st1(lane_addr + kTcBlockCols * b + kTcN2 + n2, lo[b]);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        if (!worker) {
            // ---- 3. stage 2 on the tensor cores: one elected thread issues, accumulators double buffered ----
            for (int u = 0; u < kTcUnits; ++u) {
                if (lane == 0) {
                    const int buf = u & 1, b = u >> 1, h = u & 1;
                    mbar_wait(&s_empty[buf], empty_phase[buf] ^ 1u);
                    empty_phase[buf] ^= 1u;
                    tc_fence_after();
                    const uint32_t d_tmem = tmem + kTcDBase + kTcDCols * buf;
                    const uint32_t a_tmem = tmem + kTcBlockCols * b;
                    const uint32_t set_off = (b == 0 ? 0u : 1u);
                    const uint64_t half_off = static_cast<uint64_t>((h * kTcDCols * 16) >> 4);
                    const uint64_t dm = desc_main0 + ((set_off * kTcBMainHalves * 2) >> 4) + half_off;
                    const uint64_t dc = desc_corr0 + ((set_off * kTcBCorrHalves * 2) >> 4) + half_off;
#pragma unroll
                    for (int s = 0; s < kTcKMain / 16; ++s)
                        mma_f16_ts(d_tmem, a_tmem + 8 * s, dm + static_cast<uint64_t>((2 * kTcStripBytes * s) >> 4), s > 0 ? 1u : 0u);
#pragma unroll
                    for (int s = 0; s < kTcKCorr / 16; ++s)
                        mma_f16_ts(d_tmem, a_tmem + 8 * s, dc + static_cast<uint64_t>((2 * kTcStripBytes * s) >> 4), 1u);
                    mma_commit(&s_full[buf]);
                }
                __syncwarp();
            }
        } else {
            // ---- 4. epilogue: power of 16 bins per unit -> this thread's mel taps ----
            char* s_col = reinterpret_cast<char*>(s_S + quad * 32 + lane);
            for (int u = 0; u < kTcUnits; ++u) {
                const int buf = u & 1;
                mbar_wait(&s_full[buf], full_phase[buf]);
                full_phase[buf] ^= 1u;
                tc_fence_after();
                float d[32];
                tmem_ld32(lane_addr + kTcDBase + kTcDCols * buf, d);
                tc_fence_before();
                mbar_arrive(&s_empty[buf]);
                tc_accumulate(d, s_tap[parity_role][u], s_col);
            }
            // ---- 5. log10 clamp, coalesced stores, max key; clear the S column for the next tile ----
            const int f = quad * 32 + lane, t = t0 + f;
            float mx = __uint_as_float(0xff800000u);
            float* out = a.out + (clip * a.n_mels + parity_role) * static_cast<int64_t>(a.n_frames) + t;
            for (int m = parity_role; m < a.n_mels; m += 2) {
                const float lg = log10_clamped(s_S[m * kTcTileFrames + f]);
                s_S[m * kTcTileFrames + f] = 0.f;
                if (t < a.n_frames) { *out = lg; mx = max_nan(mx, lg); }
                out += 2 * static_cast<int64_t>(a.n_frames);
            }
            uint32_t key = t < a.n_frames ? max_key_encode(mx) : 0u;
            key = __reduce_max_sync(0xffffffffu, key);
            if (lane == 0) atomicMax(a.max_keys + (a.global_max ? 0 : clip), key);
        }
        tc_fence_before();
        __syncthreads();   // every accumulator is drained: tensor memory and the audio tile are free again
        tc_fence_after();
    }

    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(static_cast<uint32_t>(kTcTmemCols)) : "memory");
}

template <typename InT>
cudaError_t launch_tc(const LogmelArgs& a, const TcTables* tables, cudaStream_t stream) {
    constexpr int kMaxDevices = 64;
    static int sms_by_device[kMaxDevices] = {0};
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err != cudaSuccess) return err;
    if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
    if (sms_by_device[device] == 0) {
        err = cudaFuncSetAttribute(logmel_tc_kernel<InT>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem(kMaxMels).total);
        if (err != cudaSuccess) return err;
        int sms = 0;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return err;
        sms_by_device[device] = sms;
    }
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    const int64_t tiles = a.batch * tiles_per_clip;
    const unsigned grid = static_cast<unsigned>(tiles < sms_by_device[device] ? tiles : sms_by_device[device]);
    ProfileScope profile(2, stream);
    logmel_tc_kernel<InT><<<grid, kTcThreads, tc_smem(a.n_mels).total, stream>>>(a, tables);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_tc_pass1(const LogmelArgs& a, const TcTables* tables, int dtype, cudaStream_t stream) {
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    if (a.batch * tiles_per_clip <= 0) return cudaSuccess;
    return dtype == 0 ? launch_tc<float>(a, tables, stream) : launch_tc<int16_t>(a, tables, stream);
}

}  // namespace b200mel
