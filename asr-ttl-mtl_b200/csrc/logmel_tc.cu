// tcgen05 (tensor-core) variant of the fused log-mel front-end for sm_100a: the folded DFT as four
// 128 x 104 x 112 GEMMs per 128-frame tile, 3-product fp16 split precision (math: tc_core.cuh;
// reference: whisper/audio.py:145-155).
//
// One persistent CTA per SM, warp-specialised, no __syncthreads in the steady state (mbarriers only):
//   producer warp   : stages the tile's 130 rows of 160 samples in shared memory, one bulk-TMA copy
//                     (cp.async.bulk -> mbarrier) per row at pitch 164 words; rows that touch a clip edge
//                     (reflect padding, zero tail, `lengths`) or are not 16-byte aligned are written by hand;
//   4 + 4 fold warps: one thread per frame (= TMEM lane).  The E warps compute ee / eo, the O warps oe / oo
//                     (window multiply and both folds fused: 5 flops per two values), split every value
//                     into fp16 hi + lo and write the packed pairs straight into TENSOR MEMORY as the
//                     A operand (tcgen05.st) - the data never touch shared memory again;
//   MMA warp        : one elected thread issues, per unit, 6 K-steps x 3 passes + 2 leftover steps of
//                     tcgen05.mma.kind::f16 (M 128, N 104, K 16; A from TMEM, B = the constant matrix from
//                     shared memory, fp32 accumulator in TMEM): hi Bh + lo Bh + hi Bl, then
//                     tcgen05.commit -> mbarrier hands the accumulator to the epilogue;
//   4 + 4 epilogue  : two warps per TMEM lane quadrant pull their half of the 104 accumulator columns
//     warps           into registers at once (tcgen05.ld), release the accumulator, and add w d^2 to the mels
//                     of each bin - mel structure and weights are compile-time constants (FFMA immediates),
//                     partial sums in registers; after the 4th unit: log10(max(., 1e-10)), 128-byte
//                     coalesced row stores and the utterance's max key (warp REDUX + one atomicMax).
// Tensor memory is exactly full: 408 operand columns (4 units x [hi | lo], tc_core.cuh) + 104 accumulator.
// The (max - 8, (x + 4) / 4) step runs as the shared pass-2 kernel.
#include <cuda_runtime.h>

#include <cstdlib>

#include "kernels.h"
#include "tc_core.cuh"

namespace b200mel {

namespace {

constexpr int kWarpO = 4, kWarpEpi0 = 8, kWarpEpi1 = 12, kWarpMma = 16, kWarpProducer = 17;
constexpr int kTcWarps = 20;
constexpr int kTcThreads = kTcWarps * 32;   // 640
constexpr uint32_t kSpinLimit = 1u << 24;    // a protocol bug traps instead of hanging the device

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) break;
        if (++spins > kSpinLimit) __trap();
    }
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory stores / loads (32 lanes x 32 bit per column, this warp's lane quadrant) ----
__device__ __forceinline__ void tmem_st1(uint32_t t, uint32_t a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(t), "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t t, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(t), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t t, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t t, float* d) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t t, float* d) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t t, float* d) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(t) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t t, float* d) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(t) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(r[i]);
}
// N columns starting at column address t, as 32 / 16 / 8 / 4-column pieces (N % 4 == 0); no wait
template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t t, float* d) {
    static_assert(N % 4 == 0 && N >= 0, "column count");
    if constexpr (N >= 32) { tmem_ld32(t, d); tmem_ld_cols<N - 32>(t + 32, d + 32); }
    else if constexpr (N >= 16) { tmem_ld16(t, d); tmem_ld_cols<N - 16>(t + 16, d + 16); }
    else if constexpr (N >= 8) { tmem_ld8(t, d); tmem_ld_cols<N - 8>(t + 8, d + 8); }
    else if constexpr (N >= 4) { tmem_ld4(t, d); tmem_ld_cols<N - 4>(t + 4, d + 4); }
}

// ---- tensor-core issue -------------------------------------------------------------------------
// K-major, no-swizzle shared-memory operand descriptor (8 x 16 B core matrices):
// start >> 4 | (K-direction core-matrix stride >> 4) << 16 | (8-row group stride >> 4) << 32 | version 1 << 46
__device__ __forceinline__ uint64_t operand_desc(uint32_t smem_addr, uint32_t k_stride_bytes) {
    return static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(k_stride_bytes >> 4) << 16) |
           (static_cast<uint64_t>(128 >> 4) << 32) | (1ull << 46);
}
// f16 x f16 -> f32, both operands K-major, M = 128, N = 104
constexpr uint32_t kTcIdesc = (1u << 4) | (static_cast<uint32_t>(kTcN >> 3) << 17) | (static_cast<uint32_t>(kTcTileFrames >> 4) << 24);

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

// ---- shared memory carve-up ----------------------------------------------------------------------
constexpr int kSmemOperands = 0;                                                  // 139776 B, 128-byte aligned
constexpr int kSmemAudio = kTcOperandBytes;                                       // 130 rows x 656 B
constexpr int kSmemStraddle = kSmemAudio + kTcAudioWords * 4;                     // [2 buffers][4 quadrants][3][32] floats
constexpr int kSmemBytes = kSmemStraddle + 2 * 4 * 3 * 32 * 4;
static_assert(kSmemAudio % 128 == 0 && kSmemBytes <= 227 * 1024, "shared memory budget");

struct TcBarriers {
    uint64_t audio_full, audio_empty;
    uint64_t a_full[2], a_empty[2];   // [E sweep, O sweep]
    uint64_t d_full, d_empty;
};

template <typename InT> __device__ __forceinline__ float sample_to_float(InT v);
template <> __device__ __forceinline__ float sample_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float sample_to_float<int16_t>(int16_t v) { return static_cast<float>(v) * (1.0f / 32768.0f); }

struct TileCoord { int64_t clip; int t0; };
__device__ __forceinline__ TileCoord tile_coord(int64_t tile, int tiles_per_clip) {
    TileCoord c;
    c.clip = tile / tiles_per_clip;
    c.t0 = static_cast<int>(tile - c.clip * tiles_per_clip) * kTcTileFrames;
    return c;
}

// ---- producer: one tile of audio into shared memory ----------------------------------------------
template <typename InT>
__device__ __forceinline__ void produce_tile(const LogmelArgs& a, const TileCoord& tc, float* s_audio, uint64_t* full, int lane) {
    const InT* __restrict__ row = static_cast<const InT*>(a.audio) + tc.clip * a.stride_b;
    int64_t valid = a.n_samples;
    if (a.lengths != nullptr) {
        const int64_t len = a.lengths[tc.clip];
        valid = len < 0 ? 0 : (len < valid ? len : valid);
    }
    const int64_t s0 = static_cast<int64_t>(tc.t0) * kHop - kHalfWin;
    const bool aligned = sizeof(InT) == 4 && (reinterpret_cast<uintptr_t>(row) & 15u) == 0;
    // rows [r_lo, r_hi) are whole, real, in-range samples: bulk copies.  The rest is written by hand.
    int r_lo = 0, r_hi = 0;
    if (aligned) {
        r_lo = s0 >= 0 ? 0 : static_cast<int>((-s0 + kHop - 1) / kHop);
        const int64_t room = valid - s0;   // samples available from the tile origin
        r_hi = room <= 0 ? 0 : static_cast<int>(room / kHop < kTcAudioRows ? room / kHop : kTcAudioRows);
        if (r_hi < r_lo) r_hi = r_lo;
    }
    for (int r = 0; r < kTcAudioRows; ++r) {
        if (r >= r_lo && r < r_hi) continue;
        for (int c = lane; c < kHop; c += 32) {
            const int64_t pos = s0 + static_cast<int64_t>(r) * kHop + c;
            float v = 0.f;
            if (pos < a.total + kHalfWin) {
                const int64_t idx = reflect_source_index(pos, a.total);
                if (idx >= 0 && idx < valid) v = sample_to_float<InT>(__ldg(row + idx));
            }
            s_audio[r * kTcRowPitch + c] = v;
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(full, static_cast<uint32_t>(r_hi - r_lo) * (kHop * 4));
    __syncwarp();
    for (int r = r_lo + lane; r < r_hi; r += 32)
        bulk_copy_g2s(smem_u32(s_audio + r * kTcRowPitch), row + s0 + static_cast<int64_t>(r) * kHop, kHop * 4, full);
}

// ---- fold warps: one sweep of one tile, A operand -> tensor memory -----------------------------------
template <int SWEEP, int J>
__device__ __forceinline__ void sweep_chunk_store(const float* fr, uint32_t lane_addr) {
    uint32_t hf[4], lf[4], hs[4], ls[4];
    tc_sweep_chunk<SWEEP, J>(fr, hf, lf, hs, ls);
    constexpr int u1 = SWEEP == 0 ? 0 : 2, u2 = u1 + 1;
    if constexpr (J < 2 * kTcMainSteps) {   // slots 8J..8J+7 of the main blocks
        tmem_st4(lane_addr + tc_hi_col(u1) + 4 * J, hf); tmem_st4(lane_addr + tc_lo_col(u1) + 4 * J, lf);
        tmem_st4(lane_addr + tc_hi_col(u2) + 4 * J, hs); tmem_st4(lane_addr + tc_lo_col(u2) + 4 * J, ls);
    } else {                                // slots 96..101: [hi x 3 | lo x 3] columns of the leftover area
        const uint32_t b1 = lane_addr + tc_left_col(u1), b2 = lane_addr + tc_left_col(u2);
        tmem_st2(b1, hf[0], hf[1]); tmem_st2(b1 + 2, hf[2], lf[0]); tmem_st2(b1 + 4, lf[1], lf[2]);
        tmem_st2(b2, hs[0], hs[1]); tmem_st2(b2 + 2, hs[2], ls[0]); tmem_st2(b2 + 4, ls[1], ls[2]);
    }
}
template <int SWEEP, int... J>
__device__ __forceinline__ void sweep_store(const float* fr, uint32_t lane_addr, std::integer_sequence<int, J...>) {
    (sweep_chunk_store<SWEEP, J>(fr, lane_addr), ...);
}

// ---- epilogue helpers -----------------------------------------------------------------------------
template <int NM, int HALF>
__device__ __forceinline__ void epilogue_unit(int unit, uint32_t d_addr, uint64_t* d_full, uint64_t* d_empty, uint32_t parity,
                                              int lane, float (&acc)[TcEpilogueLayout<NM>::acc_size(HALF)]) {
    using L = TcEpilogueLayout<NM>;
    float d[L::cols(HALF)];
    mbar_wait(d_full, parity);
    tc_fence_after();
    tmem_ld_cols<L::cols(HALF)>(d_addr + L::col0(HALF), d);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(d_empty);   // the accumulator is in registers: the next unit may overwrite it
    switch (unit) {
        case -1: acc[0] += d[0] + d[L::cols(HALF) - 1]; break;   // bring-up: loads only
        case 0: tc_epilogue_unit<NM, 0, HALF>(d, acc); break;
        case 1: tc_epilogue_unit<NM, 1, HALF>(d, acc); break;
        case 2: tc_epilogue_unit<NM, 2, HALF>(d, acc); break;
        default: tc_epilogue_unit<NM, 3, HALF>(d, acc); break;
    }
}

template <int NM, int HALF>
__device__ __forceinline__ void epilogue_role(const LogmelArgs& a, const int debug_stage, TcBarriers* bars, float* s_straddle,
                                              uint32_t tmem, int quad, int lane, int64_t total_tiles, int tiles_per_clip) {
    using L = TcEpilogueLayout<NM>;
    constexpr int ACC = L::acc_size(HALF);
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = 0.f;
    const uint32_t d_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16) + kTcDCol;
    uint32_t d_parity = 0, buf = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = tile_coord(tile, tiles_per_clip);
        // unit order on the tensor cores: 0, 1 (E sweep), 2, 3 (O sweep)
#pragma unroll 1
        for (int u = 0; u < kTcUnits; ++u) {
            epilogue_unit<NM, HALF>(debug_stage == 4 ? -1 : u, d_addr, &bars->d_full, &bars->d_empty, d_parity, lane, acc);
            d_parity ^= 1u;
        }
        // join the mels that straddle the split: half 1 hands its partial sums to half 0
        float* strad = s_straddle + ((buf * 4 + quad) * 3) * 32 + lane;
        if constexpr (HALF == 1) {
#pragma unroll
            for (int j = 0; j < L::straddle; ++j) strad[j * 32] = acc[j];
        }
        if constexpr (L::straddle > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        if constexpr (HALF == 0) {
#pragma unroll
            for (int j = 0; j < L::straddle; ++j) acc[L::high_base + j] += strad[j * 32];
        }
        buf ^= 1u;
        // log10 clamp, coalesced row stores (lane = frame), utterance max
        const int f = quad * 32 + lane, t = tc.t0 + f;
        const bool live = t < a.n_frames;
        constexpr int m_begin = HALF == 0 ? 0 : L::low_mels, m_end = HALF == 0 ? L::low_mels : NM;
        float* out = a.out + (tc.clip * NM + m_begin) * static_cast<int64_t>(a.n_frames) + t;
        float mx = __uint_as_float(0xff800000u);
#pragma unroll
        for (int m = m_begin; m < m_end; ++m) {
            const float lg = log10_clamped(acc[m - L::acc_base(HALF)]);
            if (live) { if (debug_stage != 5) *out = lg; mx = max_nan(mx, lg); }
            out += a.n_frames;
        }
#pragma unroll
        for (int i = 0; i < ACC; ++i) acc[i] = 0.f;
        uint32_t key = live ? max_key_encode(mx) : 0u;
        key = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0 && debug_stage != 5) atomicMax(a.max_keys + (a.global_max ? 0 : tc.clip), key);
    }
}

template <typename InT, int NM>
__global__ void __launch_bounds__(kTcThreads, 1)
logmel_tc_kernel(const __grid_constant__ LogmelArgs a, const unsigned char* __restrict__ operands, const int debug_stage) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_audio = reinterpret_cast<float*>(smem_raw + kSmemAudio);
    float* s_straddle = reinterpret_cast<float*>(smem_raw + kSmemStraddle);
    __shared__ __align__(8) TcBarriers bars;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, quad = warp & 3;
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    int64_t total_tiles = a.batch * tiles_per_clip;
    // bring-up aid (B200MEL_TC_DEBUG): 1 = setup only, 2 = + producer and folds of ONE tile,
    // 3 = + the tensor cores, 4 = + accumulator loads, 5 = + epilogue math, 6 = everything for one tile
    if (debug_stage > 0 && debug_stage != 7 && total_tiles > gridDim.x) total_tiles = gridDim.x;
    if (debug_stage == 1) total_tiles = 0;

    // ---- one-time setup: tensor memory, barriers, constant matrices -> shared memory ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&bars.audio_full, 1);
        mbar_init(&bars.audio_empty, 8);
        mbar_init(&bars.a_full[0], 4); mbar_init(&bars.a_full[1], 4);
        mbar_init(&bars.a_empty[0], 1); mbar_init(&bars.a_empty[1], 1);
        mbar_init(&bars.d_full, 1);
        mbar_init(&bars.d_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(operands);
        uint4* dst = reinterpret_cast<uint4*>(smem_raw + kSmemOperands);
        for (int i = tid; i < kTcOperandBytes / 16; i += kTcThreads) dst[i] = __ldg(src + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core (async proxy) reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);   // this warp's TMEM lane quadrant

    if (warp < kWarpO) {
        // zero every column once (every operand column is rewritten each tile; this only keeps idle lanes finite)
        for (int c = 0; c < 512; c += 4) { const uint32_t z[4] = {0u, 0u, 0u, 0u}; tmem_st4(lane_addr + c, z); }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // register budget per warpgroup: the CTA is launched with 5 x 96; the warpgroups trade inside that total
    // (a setmaxnreg.inc can only take what another warpgroup released): folds 80 + 80, epilogue 144 + 144, rest 32
    if (warp < kWarpEpi0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        // ===== fold warps: E sweep (warps 0-3) / O sweep (warps 4-7) =====
        const int sweep = warp < kWarpO ? 0 : 1;
        const float* fr = s_audio + (quad * 32 + lane) * kTcRowPitch;
        uint32_t parity = 0;
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            mbar_wait(&bars.audio_full, parity);
            mbar_wait(&bars.a_empty[sweep], parity ^ 1u);   // the tensor cores are done with the previous tile's operand
            tc_fence_after();
            if (sweep == 0) sweep_store<0>(fr, lane_addr, std::make_integer_sequence<int, kTcChunks>{});
            else sweep_store<1>(fr, lane_addr, std::make_integer_sequence<int, kTcChunks>{});
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&bars.audio_empty); mbar_arrive(&bars.a_full[sweep]); }
            parity ^= 1u;
        }
    } else if (warp < kWarpMma) {
        // ===== epilogue warps =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        if (debug_stage > 0 && debug_stage < 4) total_tiles = 0;
        if (warp < kWarpEpi1) epilogue_role<NM, 0>(a, debug_stage, &bars, s_straddle, tmem, quad, lane, total_tiles, tiles_per_clip);
        else epilogue_role<NM, 1>(a, debug_stage, &bars, s_straddle, tmem, quad, lane, total_tiles, tiles_per_clip);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp == kWarpMma && lane == 0 && (debug_stage == 0 || debug_stage >= 3)) {
            // ===== tensor-core issue: one elected thread =====
            const uint32_t op_base = smem_u32(smem_raw + kSmemOperands);
            const uint32_t d_tmem = tmem + kTcDCol;
            uint32_t a_parity = 0, d_parity = 1;   // d_empty: the first wait passes (accumulator starts free)
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                // u is a compile-time constant on purpose: with a run-time u, nvcc 12.9 folded &bars.a_empty[u >> 1]
                // into base + 4 u (right only for even u) and the commit hit a misaligned mbarrier
#pragma unroll
                for (int u = 0; u < kTcUnits; ++u) {
                    if ((u & 1) == 0) { mbar_wait(&bars.a_full[u >> 1], a_parity); }
                    if (u == 0 && debug_stage == 7) mbar_wait(&bars.a_full[1], a_parity);   // bring-up: no MMA while folds still store
                    if (debug_stage != 3) mbar_wait(&bars.d_empty, d_parity);
                    else if (u > 0) { mbar_wait(&bars.d_full, (u - 1) & 1); }
                    d_parity ^= 1u;
                    tc_fence_after();
                    const int m = tc_unit_matrix(u);
                    const uint32_t a_hi = tmem + tc_hi_col(u), a_lo = tmem + tc_lo_col(u);
                    const uint32_t b_hi = op_base + tc_matrix_offset(m, 0), b_lo = op_base + tc_matrix_offset(m, 1);
#pragma unroll
                    for (int s = 0; s < kTcMainSteps; ++s) {   // K step s = strips 2s, 2s+1 = slots 16s..16s+15
                        const uint64_t dh = operand_desc(b_hi + 2 * s * kTcStripBytes, kTcStripBytes);
                        const uint64_t dl = operand_desc(b_lo + 2 * s * kTcStripBytes, kTcStripBytes);
                        mma_f16_ts(d_tmem, a_hi + 8 * s, dh, s > 0 ? 1u : 0u);
                        mma_f16_ts(d_tmem, a_lo + 8 * s, dh, 1u);
                        mma_f16_ts(d_tmem, a_hi + 8 * s, dl, 1u);
                    }
                    // slots 96..101: one K step over the unit's [hi | lo] leftover columns, (hi + lo) Bh then hi Bl
                    const uint32_t a_left = tmem + tc_left_start(u);
                    mma_f16_ts(d_tmem, a_left, operand_desc(op_base + tc_left_offset(m, 0), kTcStripBytes), 1u);
                    mma_f16_ts(d_tmem, a_left, operand_desc(op_base + tc_left_offset(m, 1), kTcStripBytes), 1u);
                    mma_commit(&bars.d_full);
                    if (u & 1) mma_commit(&bars.a_empty[u >> 1]);   // both units of the sweep have consumed its operand
                }
                a_parity ^= 1u;
            }
            if (debug_stage == 3 && total_tiles > 0) mbar_wait(&bars.d_full, 1);   // nobody drains the accumulator in this stage
        } else if (warp == kWarpProducer) {
            // ===== audio producer =====
            uint32_t parity = 1;   // audio_empty: the first wait passes
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(&bars.audio_empty, parity);
                parity ^= 1u;
                produce_tile<InT>(a, tile_coord(tile, tiles_per_clip), s_audio, &bars.audio_full, lane);
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <typename InT, int NM>
cudaError_t launch_tc(const LogmelArgs& a, const TcTables* tables, cudaStream_t stream) {
    constexpr int kMaxDevices = 64;
    static int sms_by_device[kMaxDevices] = {0};
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err != cudaSuccess) return err;
    if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
    if (sms_by_device[device] == 0) {
        err = cudaFuncSetAttribute(logmel_tc_kernel<InT, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err != cudaSuccess) return err;
        int sms = 0;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return err;
        sms_by_device[device] = sms;
    }
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    const int64_t tiles = a.batch * tiles_per_clip;
    const unsigned grid = static_cast<unsigned>(tiles < sms_by_device[device] ? tiles : sms_by_device[device]);
    ProfileScope profile(2, stream);
    static const int debug_stage = std::getenv("B200MEL_TC_DEBUG") ? std::atoi(std::getenv("B200MEL_TC_DEBUG")) : 0;
    logmel_tc_kernel<InT, NM><<<grid, kTcThreads, kSmemBytes, stream>>>(a, tables->operands, debug_stage);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_tc_pass1(const LogmelArgs& a, const TcTables* tables, int dtype, cudaStream_t stream) {
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    if (a.batch * tiles_per_clip <= 0) return cudaSuccess;
    if (a.n_mels == 80) return dtype == 0 ? launch_tc<float, 80>(a, tables, stream) : launch_tc<int16_t, 80>(a, tables, stream);
    if (a.n_mels == 128) return dtype == 0 ? launch_tc<float, 128>(a, tables, stream) : launch_tc<int16_t, 128>(a, tables, stream);
    return cudaErrorInvalidValue;
}

}  // namespace b200mel
