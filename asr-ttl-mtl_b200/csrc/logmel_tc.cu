// tcgen05 (tensor-core) variant of the fused log-mel front-end for sm_100a: the folded DFT as four
// 128 x 104 x 112 GEMMs per 128-frame tile, 3-product fp16 split precision (math: tc_core.cuh;
// reference: whisper/audio.py:145-156).  DESIGN.md section 4.1 has the picture; in short:
//
// One persistent CTA per SM, 20 warps, warp-specialised, no __syncthreads in the steady state (mbarriers only):
//   loader warp     : ONE TMA tensor copy per tile brings the 130 rows of 160 samples into shared memory at pitch 164
//                     words (a 4-D tensor map whose rows overlap, see "loaders"); the next tile is prefetched into L2;
//                     a clip's first / last tile: the copy zero-fills what it cannot address and the warp rewrites the
//                     1-3 rows of real / reflected samples.  (`lengths` cuts, int16 PCM, unaligned rows: the fold warps
//                     stage the tile themselves with cp.async / converted samples.)
//   4 + 4 fold warps: one thread per frame (= TMEM lane), two warps per lane quadrant that split every sweep between
//                     them.  E sweep: ee / eo, then O sweep: oe / oo (window multiply and both folds fused, packed
//                     FMUL2 / FFMA2), every value split into fp16 hi + lo and written straight into TENSOR MEMORY as
//                     the A operand (tcgen05.st) - the data never touch shared memory again;
//   MMA warp        : one elected thread issues, per unit, 6 K-steps x 3 passes + 2 leftover steps of
//                     tcgen05.mma.kind::f16 (M 128, N 104, K 16; A from TMEM, B = the constant matrix from
//                     shared memory, fp32 accumulator in TMEM): hi Bh + lo Bh + hi Bl, then
//                     tcgen05.commit -> mbarrier hands the accumulator to the epilogue;
//   4 + 4 epilogue  : two warps per TMEM lane quadrant pull their half of the 104 accumulator columns
//     warps           into registers at once (tcgen05.ld), release the accumulator, and add w d^2 to the mels
//                     of each bin - mel structure and weights are compile-time constants (FFMA immediates),
//                     partial sums in registers; after the 4th unit: log10(max(., 1e-10)), (x + 4) / 4, 128-byte
//                     coalesced row stores, the utterance's and the tile's extremes (warp REDUX + atomicMax);
//   2 normaliser    : once an utterance is complete (per-clip counter) decide from its and the tile's extremes what the
//     warps           clamp at max - 8 does to each of this CTA's tiles: nothing (the usual case), a constant fill
//                     (digital silence, zero padding) or a clamp in place while the tile is still in L2 - so the
//                     front-end is one launch whose DRAM traffic is the algorithmic read + write.
// Tensor memory is exactly full: 408 operand columns (4 units x [hi | lo], tc_core.cuh) + 104 accumulator; so is shared
// memory (DFT matrices + one audio tile).  The hot code of all roles has to fit the 32 KB instruction cache: loops over
// table rows instead of unrolled code wherever the work is regular, and no bring-up code in the production build.
// (One max over a whole multi-utterance call: the shared pass-2 kernel normalises instead.)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_core.cuh"

namespace b200mel {

namespace {

constexpr int kWarpO = 4, kWarpEpi0 = 8, kWarpEpi1 = 12, kWarpMma = 16, kWarpLoad = 17, kWarpNorm = 18;   // normalisers: warps 18, 19
constexpr int kTcWarps = 20;
constexpr int kTcThreads = kTcWarps * 32;   // 640
constexpr uint32_t kSpinLimit = 1u << 17;    // x 20 us hint = 2.6 s: a protocol bug traps instead of hanging the device
constexpr uint32_t kWaitHintNs = 20000;      // suspend-time hint of one try_wait

// Bring-up timeline (B200MEL_TC_TRACE=1): CTA 0 stamps clock64() at the hand-over points of its first tiles.
constexpr int kTraceTiles = 8, kTraceEvents = 16, kTraceRoles = 6;
constexpr int kStampCtas = 256, kTileStamps = 64;   // (x 2: fold start and epilogue end of every tile)
#define TC_TILE_STAMPS (2 * kTileStamps + 32 + kStampCtas) //   // per-CTA start / end stamps, CTA 0's start of every tile
#define TC_TRACE(role, tile_index, event)                                                                        \
    do {                                                                                                         \
        if (trace != nullptr && blockIdx.x == 0 && (tile_index) >= trace_first && (tile_index) < trace_first + kTraceTiles &&      \
            (threadIdx.x & 31) == 0)                                                                             \
            trace[((role) * kTraceTiles + (tile_index) - trace_first) * kTraceEvents + (event)] = clock64();    \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waits are potentially-blocking try_waits with a suspend-time hint: the hardware parks the warp until the phase
// completes (or the hint expires), so a waiting role does not burn issue slots of the roles that are working.
// Bring-up aid: a wait that exceeds the spin limit records (CTA, barrier offset, parity, warp) in a host-mapped word
// before it traps, so the host can name the hand-over that hung (b200mel_last_cuda_error).
__device__ unsigned* g_tc_fault = nullptr;
__device__ __noinline__ void tc_fault(uint32_t code) {
    if (g_tc_fault != nullptr) {
        g_tc_fault[0] = code;
        g_tc_fault[1] = blockIdx.x;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait_addr(const uint32_t addr, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity), "r"(kWaitHintNs) : "memory");
        if (done) break;
        if (++spins > kSpinLimit) tc_fault(0x1000000u | ((addr & 0xfffu) << 12) | (parity << 8) | (threadIdx.x >> 5));
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_addr(smem_u32(bar), parity); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- tensor memory stores / loads (32 lanes x 32 bit per column, this warp's lane quadrant) ----
// The stores carry no "memory" clobber: they alias nothing the compiler can see, and their ordering against the
// tensor cores comes from tcgen05.wait::st + the fences, so shared-memory loads may be scheduled across them.
__device__ __forceinline__ void tmem_st2(uint32_t t, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(t), "r"(a), "r"(b));
}
__device__ __forceinline__ void tmem_st4(uint32_t t, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]));
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
// low word for the operand at shared address `smem_addr` (16-byte aligned, K strip stride = kTcStripBytes); operands
// further along are reached by adding (byte offset >> 4) to it
__device__ __forceinline__ uint32_t operand_desc_lo(uint32_t smem_addr) {
    return ((smem_addr & 0x3ffffu) >> 4) | (static_cast<uint32_t>(kTcStripBytes >> 4) << 16);
}
// f16 x f16 -> f32, both operands K-major, M = 128, N = 104
constexpr uint32_t kTcIdesc = (1u << 4) | (static_cast<uint32_t>(kTcN >> 3) << 17) | (static_cast<uint32_t>(kTcTileFrames >> 4) << 24);

// Issued by ONE elected lane of a converged warp (all operands warp-uniform): elect.sync inside the asm keeps the
// compiler from wrapping every MMA in a per-thread serialisation loop.
template <bool ACCUMULATE>
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo) {
    const uint64_t b_desc = (static_cast<uint64_t>(0x4008u) << 32) | b_desc_lo;   // 8-row group stride 128 B, version 1
    if constexpr (ACCUMULATE)
        asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\n@P tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;\n}\n"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kTcIdesc) : "memory");
    else
        asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\n@P tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 0;\n}\n"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kTcIdesc) : "memory");
}
__device__ __forceinline__ void mma_commit_addr(uint32_t bar_addr) {
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\n@P tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n"
                 ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) { mma_commit_addr(smem_u32(bar)); }

// what the issue loop needs per unit: operand columns and matrix offsets (descriptor units of 16 bytes); one tile's
// tensor-core work for a unit is 6 K steps x (hi Bh, lo Bh, hi Bl) + the leftover step twice
struct TcUnitIssue { uint32_t a_hi, a_lo, a_left, b_hi, b_lo, b_left0, b_left1, pad; };
struct TcUnitIssueTable { TcUnitIssue u[kTcUnits]; };
constexpr TcUnitIssueTable tc_make_unit_issue() {
    TcUnitIssueTable t{};
    for (int u = 0; u < kTcUnits; ++u) {
        const int m = tc_unit_matrix(u);
        t.u[u].a_hi = tc_hi_col(u); t.u[u].a_lo = tc_lo_col(u); t.u[u].a_left = tc_left_start(u);
        t.u[u].b_hi = tc_matrix_offset(m, 0) >> 4; t.u[u].b_lo = tc_matrix_offset(m, 1) >> 4;
        t.u[u].b_left0 = tc_left_offset(m, 0) >> 4; t.u[u].b_left1 = tc_left_offset(m, 1) >> 4;
        t.u[u].pad = 0;
    }
    return t;
}
__constant__ TcUnitIssue c_unit_issue[kTcUnits] = {tc_make_unit_issue().u[0], tc_make_unit_issue().u[1], tc_make_unit_issue().u[2],
                                                   tc_make_unit_issue().u[3]};

// ---- shared memory carve-up ----------------------------------------------------------------------
constexpr int kSmemOperands = 0;                                                  // 139776 B, 128-byte aligned
constexpr int kSmemAudio = kTcOperandBytes;                                       // 130 rows x 656 B
constexpr int kSmemStraddle = kSmemAudio + kTcAudioWords * 4;                     // [2 buffers][4 quadrants][3][32] floats
constexpr int kSmemBytes = kSmemStraddle + 2 * 4 * 3 * 32 * 4;
static_assert(kSmemAudio % 128 == 0 && kSmemBytes <= 227 * 1024, "shared memory budget");

struct TcBarriers {
    uint64_t audio_full, audio_empty;
    uint64_t tma_done;                // tiles whose edge rows the loader patches after the tensor copy has landed
    uint64_t a_full[2], a_empty[2];   // [E sweep, O sweep]
    uint64_t d_full, d_empty;
};

// ---- normaliser warps ------------------------------------------------------------------------------
// In-place dynamic-range clamp + affine map (audio.py:155-156) of ONE TILE of a finished utterance: the rows this
// CTA's epilogue wrote a few tile periods ago, read back through L2.  Every CTA normalises its own tiles, so the work
// is spread exactly like the tiles are; nothing is queued.

// clamp of an already rescaled value y = (lg + 4) / 4 at floor_y = ((g - 8) + 4) / 4 (NaN when the max is NaN, as in
// torch): identical to (max(lg, g - 8) + 4) / 4 because the rescaling is monotone
__device__ __forceinline__ float clamp_scaled(float y, float floor_y) { return (floor_y != floor_y) ? floor_y : (y < floor_y ? floor_y : y); }
__device__ __forceinline__ float4 normalise4(float4 x, float f) {
    x.x = clamp_scaled(x.x, f); x.y = clamp_scaled(x.y, f); x.z = clamp_scaled(x.z, f); x.w = clamp_scaled(x.w, f);
    return x;
}

// ---- output element type: float32 (the reference's dtype) or IEEE half (B200MEL_FLAG_OUT_F16) ----
__device__ __forceinline__ void out_store(float* p, float v) { *p = v; }
__device__ __forceinline__ void out_store(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ float out_load(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float out_load(const __half* p) { return __half2float(__ldcg(p)); }
__device__ __forceinline__ float4 out_load4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 out_load4(const __half* p) {
    const uint2 r = __ldcg(reinterpret_cast<const uint2*>(p));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void out_store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void out_store4(__half* p, float4 v) {
    uint2 r;
    r.x = tc_half2_bits(__floats2half2_rn(v.x, v.y));
    r.y = tc_half2_bits(__floats2half2_rn(v.z, v.w));
    *reinterpret_cast<uint2*>(p) = r;
}

// one warp overwrites one tile with the clamp value (every value of the tile was below it): stores only
template <int NM, typename OutT>
__device__ __forceinline__ void fill_tile_tc(OutT* __restrict__ tile_out, int64_t pitch, int frames, float v, int lane) {
    if ((pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(tile_out) & (4 * sizeof(OutT) - 1)) == 0 && (frames & 3) == 0) {
        if (lane < (frames >> 2)) {
            OutT* p = tile_out + 4 * lane;
            const float4 v4 = make_float4(v, v, v, v);
#pragma unroll 4
            for (int row = 0; row < NM; ++row, p += pitch) out_store4(p, v4);
        }
    } else {
        for (int i = lane; i < NM * frames; i += 32) out_store(tile_out + (i / frames) * pitch + i % frames, v);
    }
}

// one warp clamps one tile (NM rows of `frames` values at `pitch`) in place
template <int NM, typename OutT>
__device__ __forceinline__ void normalise_tile_tc(OutT* __restrict__ tile_out, int64_t pitch, int frames, float g /* the clamp in rescaled units */, int lane) {
    if ((pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(tile_out) & (4 * sizeof(OutT) - 1)) == 0) {
        // whole mel rows (lane = 4-value column), four rows in flight
        if (lane < (frames >> 2)) {
            OutT* p = tile_out + 4 * lane;
            int row = 0;
#pragma unroll 1
            for (; row + 3 < NM; row += 4, p += 4 * pitch) {
                const float4 x0 = out_load4(p), x1 = out_load4(p + pitch), x2 = out_load4(p + 2 * pitch), x3 = out_load4(p + 3 * pitch);
                out_store4(p, normalise4(x0, g));
                out_store4(p + pitch, normalise4(x1, g));
                out_store4(p + 2 * pitch, normalise4(x2, g));
                out_store4(p + 3 * pitch, normalise4(x3, g));
            }
#pragma unroll 1
            for (; row < NM; ++row, p += pitch) out_store4(p, normalise4(out_load4(p), g));
        }
        const int rest = frames & 3;                                           // 0 for whole clips (pitch % 4 == 0)
        if (rest != 0)
            for (int i = lane; i < NM * rest; i += 32) {
                OutT* q = tile_out + (i / rest) * pitch + (frames & ~3) + i % rest;
                out_store(q, clamp_scaled(out_load(q), g));
            }
    } else {
        for (int i = lane; i < NM * frames; i += 32) {
            OutT* q = tile_out + (i / frames) * pitch + i % frames;
            out_store(q, clamp_scaled(out_load(q), g));
        }
    }
}

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
// A CTA's tiles are blockIdx.x, + gridDim.x, ...: walking them needs one division up front, then adds only (the
// 64-bit division is ~100 instructions, and every role that walks the tiles would carry a copy in its hot loop).
struct TileCursor {
    TileCoord at;
    int step_clips, step_t0, frames_per_clip;   // gridDim.x tiles = step_clips whole utterances + step_t0 frames
    __device__ __forceinline__ TileCursor(int tiles_per_clip) {
        const unsigned first = blockIdx.x, tpc = static_cast<unsigned>(tiles_per_clip), grid = gridDim.x;
        at.clip = first / tpc;
        at.t0 = static_cast<int>(first % tpc) * kTcTileFrames;
        step_clips = static_cast<int>(grid / tpc);
        step_t0 = static_cast<int>(grid % tpc) * kTcTileFrames;
        frames_per_clip = tiles_per_clip * kTcTileFrames;
    }
    __device__ __forceinline__ TileCoord peek_next() const {
        TileCoord n = at;
        n.clip += step_clips; n.t0 += step_t0;
        if (n.t0 >= frames_per_clip) { n.t0 -= frames_per_clip; ++n.clip; }
        return n;
    }
    __device__ __forceinline__ void advance() { at = peek_next(); }
};

// ---- loaders: one tile of audio into shared memory -----------------------------------------------
// The tile is 130 rows of 160 samples (one contiguous span of the utterance) at pitch 164 words.
//   TMA mode (fp32, 16-byte aligned rows, tile wholly inside the utterance): ONE tensor copy per tile.  The batch is
//     described to the TMA unit as a 4-D tensor {164 samples, 4 quarter rows of 40, rows of 160, utterance} whose
//     innermost extent (164) overlaps the next row on purpose: a {164, 1, 130, 1} box then lands in shared memory as
//     130 rows at pitch 164 words - the padded, bank-conflict-free layout the fold reads - in a single instruction
//     that completes on the `full` mbarrier by byte count (the 4 pad words are never read);
//   cooperative mode (tiles at a clip edge - reflect padding, zero tail, `lengths` -, int16 PCM, unaligned rows): the
//     256 fold threads, which would idle until the tile is there anyway, move it as 16-byte cp.async chunks / converted
//     samples.
// Either way `full` takes 9 arrivals per tile: lane 0 of the loader warp (with the expected byte count) and lane 0 of
// each fold warp (at once in TMA mode, after its share of the copy in cooperative mode).
constexpr int kProducerThreads = 256;
constexpr int kChunksPerRow = kHop / 4;                       // 40
constexpr int kTileChunks = kTcAudioRows * kChunksPerRow;     // 5200
constexpr uint32_t kTmaTileBytes = kTcAudioWords * 4;         // 85280: the whole box, pad words included
constexpr int kTmaLeadRows = 2, kTmaQuarter = 3;              // tile start = 160 t0 - 200 = 160 (t0 - 2) + 3 * 40

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}

__device__ __forceinline__ int64_t valid_samples(const LogmelArgs& a, int64_t clip) {
    int64_t valid = a.n_samples;
    if (a.lengths != nullptr) {
        const int64_t len = a.lengths[clip];
        valid = len < 0 ? 0 : (len < valid ? len : valid);
    }
    return valid;
}

// Asks L2 for a later tile's samples (one bulk prefetch per tile, interior tiles only): the CTAs of a wave load in
// lock-step, so without it every staging phase waits on an HBM burst while HBM idles the rest of the time.
template <typename InT>
__device__ __forceinline__ void prefetch_tile_l2(const LogmelArgs& a, const TileCoord& tc) {
    const InT* row = static_cast<const InT*>(a.audio) + tc.clip * a.stride_b;
    const int64_t s0 = static_cast<int64_t>(tc.t0) * kHop - kHalfWin;
    int64_t first = s0 < 0 ? 0 : s0, last = s0 + kTcAudioSamples;
    if (last > a.n_samples) last = a.n_samples;
    const uintptr_t begin = (reinterpret_cast<uintptr_t>(row + first) + 15u) & ~static_cast<uintptr_t>(15u);
    const uintptr_t end = reinterpret_cast<uintptr_t>(row + last) & ~static_cast<uintptr_t>(15u);
    if (end > begin)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(begin), "r"(static_cast<uint32_t>(end - begin)) : "memory");
}

// whether the TMA unit brings this tile (same answer in the loader warp and in the fold warps):
//   - every row of the tile lies inside the utterance's valid samples, or
//   - the utterance is valid to its last sample in memory (no `lengths` cut): the TMA unit zero-fills the rows it cannot
//     address (before the first sample, past the last whole row) and the loader warp then rewrites the one to three of
//     them that hold real or reflected samples (patch_tile_edges) - so the tiles at a clip's two ends need no other path
template <typename InT>
__device__ __forceinline__ bool tile_uses_tma(const LogmelArgs& a, int tma_rows, const TileCoord& tc) {
    if constexpr (sizeof(InT) != 4) return false;
    if (tma_rows <= 0) return false;
    const int64_t valid = valid_samples(a, tc.clip);
    if (valid == a.n_samples) return true;
    if (tc.t0 < kTmaLeadRows || tc.t0 - kTmaLeadRows + kTcAudioRows > tma_rows) return false;
    return static_cast<int64_t>(tc.t0) * kHop - kHalfWin + kTcAudioSamples <= valid;
}
// tile row r holds positions p0 = 160 (t0 + r) - 200 ...; the TMA unit zero-filled it if its tensor row is out of range
__device__ __forceinline__ bool tile_row_needs_patch(const LogmelArgs& a, int tma_rows, const TileCoord& tc, int r) {
    const int c2 = tc.t0 - kTmaLeadRows + r;
    if (c2 >= 0 && c2 < tma_rows) return false;
    const int64_t p0 = static_cast<int64_t>(tc.t0 + r) * kHop - kHalfWin;
    if (p0 >= a.total + kHalfWin) return false;                                  // beyond the reflected tail: zeros
    if (p0 >= a.n_samples && a.n_samples + kHalfWin < a.total) return false;      // inside a long zero padding: zeros
    return true;
}
__device__ __forceinline__ bool tile_needs_patch(const LogmelArgs& a, int tma_rows, const TileCoord& tc) {
    const int c2_first = tc.t0 - kTmaLeadRows;
    return c2_first < 0 || c2_first + kTcAudioRows > tma_rows;     // (a superset test; the row test decides)
}
// the loader warp rewrites the zero-filled rows that hold real or reflected samples (5 samples per lane and row)
__device__ __forceinline__ void patch_tile_edges(const LogmelArgs& a, int tma_rows, const TileCoord& tc, float* s_audio, int lane) {
    const float* __restrict__ row = static_cast<const float*>(a.audio) + tc.clip * a.stride_b;
    // candidates: the rows before tensor row 0 (at most two, in a clip's first tile) and the rows from the first tensor
    // row past the end up to the end of the reflected tail (at most three)
    const int c2_first = tc.t0 - kTmaLeadRows;
    const int r_past = tma_rows - c2_first < 0 ? 0 : tma_rows - c2_first;
#pragma unroll 1
    for (int i = 0; i < kTmaLeadRows + 4; ++i) {
        const int r = i < kTmaLeadRows ? i : r_past + (i - kTmaLeadRows);
        if (r >= kTcAudioRows || (i >= kTmaLeadRows && r < kTmaLeadRows && c2_first < 0 && r + c2_first < 0)) continue;   // (no row twice)
        if (!tile_row_needs_patch(a, tma_rows, tc, r)) continue;
        const int64_t p0 = static_cast<int64_t>(tc.t0 + r) * kHop - kHalfWin;
        float* dst = s_audio + r * kTcRowPitch;
        // all five loads of the lane are issued before any is used (one L2 round trip per row, not five)
        float v[kHop / 32];
        bool real[kHop / 32];
#pragma unroll
        for (int k = 0; k < kHop / 32; ++k) {
            const int64_t pos = p0 + lane + 32 * k;
            int64_t idx = reflect_source_index(pos, a.total);
            real[k] = pos < a.total + kHalfWin && idx >= 0 && idx < a.n_samples;
            idx = real[k] ? idx : 0;
            v[k] = __ldg(row + idx);
        }
#pragma unroll
        for (int k = 0; k < kHop / 32; ++k) dst[lane + 32 * k] = real[k] ? v[k] : 0.f;
    }
}

__device__ __forceinline__ void tma_load_tile(const CUtensorMap* map, const TileCoord& tc, float* s_audio, uint64_t* full) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full)), "r"(kTmaTileBytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(s_audio)), "l"(map), "r"(0), "r"(kTmaQuarter), "r"(tc.t0 - kTmaLeadRows), "r"(static_cast<int>(tc.clip)),
                   "r"(smem_u32(full)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_tile(const CUtensorMap* map, const TileCoord& tc) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(map), "r"(0), "r"(kTmaQuarter), "r"(tc.t0 - kTmaLeadRows), "r"(static_cast<int>(tc.clip)) : "memory");
}

// cooperative mode, the 256 fold threads
template <typename InT>
__device__ __forceinline__ void produce_tile(const LogmelArgs& a, const TileCoord& tc, float* s_audio, uint64_t* full, int pt) {
    const InT* __restrict__ row = static_cast<const InT*>(a.audio) + tc.clip * a.stride_b;
    const int64_t valid = valid_samples(a, tc.clip);
    const int64_t s0 = static_cast<int64_t>(tc.t0) * kHop - kHalfWin;
    const bool aligned = sizeof(InT) == 4 && (reinterpret_cast<uintptr_t>(row) & 15u) == 0;
    if (s0 >= valid && valid + kHalfWin < a.total) {
        // the whole tile lies in the zero tail (`lengths`, right padding) and no reflection reaches a real sample
        float4* z = reinterpret_cast<float4*>(s_audio);
        for (int i = pt; i < kTcAudioWords / 4; i += kProducerThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        if ((pt & 31) == 0) mbar_arrive(full);
        return;
    }
    if constexpr (sizeof(InT) == 2) {
        // int16 PCM, tile wholly inside the utterance, 16-byte aligned: every thread pulls its ~10 chunks of 8 samples
        // (one 128-bit load each, all in flight at once), then scales them by 2^-15 into the fp32 rows (audio.py:62)
        if (s0 >= 0 && s0 + kTcAudioSamples <= valid && ((reinterpret_cast<uintptr_t>(row) + 2u * static_cast<uint64_t>(s0)) & 15u) == 0) {
            constexpr int kPcmChunks = kTcAudioSamples / 8;                      // 2590 chunks of 8 samples, 20 per row
            constexpr int kPerThread = (kPcmChunks + kProducerThreads - 1) / kProducerThreads;
            static_assert(kTcAudioSamples % 8 == 0 && kHop % 8 == 0, "the tile is a whole number of 8-sample chunks");
            const uint4* src = reinterpret_cast<const uint4*>(row + s0);
            uint4 raw[kPerThread];
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                const int c = pt + i * kProducerThreads;
                raw[i] = c < kPcmChunks ? __ldcg(src + c) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int i = 0; i < kPerThread; ++i) {
                const int c = pt + i * kProducerThreads;
                if (c < kPcmChunks) {
                    const int rr = c / (kHop / 8), col = (c - rr * (kHop / 8)) * 8;
                    const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[2 * j] = static_cast<float>(static_cast<int16_t>(w[j] & 0xffffu)) * (1.0f / 32768.0f);
                        v[2 * j + 1] = static_cast<float>(static_cast<int16_t>(w[j] >> 16)) * (1.0f / 32768.0f);
                    }
                    float4* dst = reinterpret_cast<float4*>(s_audio + rr * kTcRowPitch + col);
                    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            __syncwarp();
            if ((pt & 31) == 0) mbar_arrive(full);
            return;
        }
    }
    // chunk c = 40 r + k covers samples s0 + 4c .. + 3 and lands at word 164 r + 4 k
    int r = pt / kChunksPerRow, k = pt - r * kChunksPerRow;
    for (int c = pt; c < kTileChunks; c += kProducerThreads) {
        const int64_t pos = s0 + 4 * static_cast<int64_t>(c);
        if (pos < s0 + kTcAudioSamples) {                  // the second half of row 129 is never read
            float* dst = s_audio + r * kTcRowPitch + 4 * k;
            if (aligned && pos >= 0 && pos + 4 <= valid) {
                cp_async16(smem_u32(dst), reinterpret_cast<const float*>(row) + pos);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v = 0.f;
                    if (pos + i < a.total + kHalfWin) {
                        const int64_t idx = reflect_source_index(pos + i, a.total);
                        if (idx >= 0 && idx < valid) v = sample_to_float<InT>(__ldg(row + idx));
                    }
                    dst[i] = v;
                }
            }
        }
        r += 6; k += 16;                                   // 256 = 6 x 40 + 16
        if (k >= kChunksPerRow) { k -= kChunksPerRow; ++r; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if ((pt & 31) == 0) mbar_arrive(full);
}

// ---- fold warps: one sweep of one tile, A operand -> tensor memory -----------------------------------
__constant__ TcFoldRows c_fold_rows = tc_make_fold_rows();

// One compact loop for both sweeps (see TcFoldRows), run over the chunks [j0, j1): the two fold warps of a lane
// quadrant split every sweep between them (chunks 0..6 and 7..12), so a sweep's operand is complete in half the time
// and the tensor cores start on it that much earlier.  Chunk 12 is the leftover chunk with its own store pattern.
// `sweep` is warp-uniform, so the table rows arrive through the uniform datapath.
__device__ __forceinline__ void sweep_store(int sweep, int j0, int j1, const float* fr, uint32_t lane_addr) {
    const TcFoldRow* __restrict__ rows = c_fold_rows.row[sweep];
    const float sign = c_fold_rows.sign[sweep];
    float head[2];
    head[0] = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(fr) + rows[j0].head[0]);
    head[1] = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(fr) + rows[j0].head[1]);
    uint32_t c = lane_addr + (sweep == 0 ? tc_hi_col(0) : tc_hi_col(2)) + 4 * j0;   // hi block of the sweep's first unit
    uint32_t hf[4], lf[4], hs[4], ls[4];
#pragma unroll 1
    for (int j = j0; j < j1; ++j, c += 4) {
        tc_sweep_chunk_row(fr, rows[j], sign, head, hf, lf, hs, ls);
        if (j < 2 * kTcMainSteps) {                                        // slots 8j..8j+7 of the main blocks
            tmem_st4(c, hf); tmem_st4(c + 48, lf);                         // unit: [hi 48 | lo 48], next unit 96 columns on
            tmem_st4(c + 96, hs); tmem_st4(c + 144, ls);
        } else {
            // slots 96..101 (same loop body, its own store pattern): [hi x 3 | lo x 3] columns of the leftover area,
            // second unit 6 columns on
            const uint32_t b1 = lane_addr + (sweep == 0 ? tc_left_col(0) : tc_left_col(2)), b2 = b1 + 6;
            tmem_st2(b1, hf[0], hf[1]); tmem_st2(b1 + 2, hf[2], lf[0]); tmem_st2(b1 + 4, lf[1], lf[2]);
            tmem_st2(b2, hs[0], hs[1]); tmem_st2(b2 + 2, hs[2], ls[0]); tmem_st2(b2 + 4, ls[1], ls[2]);
        }
    }
}
constexpr int kFoldSplit = 7;   // chunks [0, 7) and [7, 13)
static_assert(tc_lo_col(0) - tc_hi_col(0) == 48 && tc_hi_col(1) - tc_hi_col(0) == 96 && tc_hi_col(3) - tc_hi_col(2) == 96 &&
              tc_left_col(1) - tc_left_col(0) == 6 && tc_left_col(3) - tc_left_col(2) == 6, "column arithmetic of sweep_store");

// ---- epilogue ---------------------------------------------------------------------------------------
// Hand-shake between the epilogue warps and the normaliser warps for the rare tiles that need the clamp: the rows an
// epilogue warp stored become visible to the normaliser only after a gpu-scope fence, which costs ~1000 cycles - so
// the epilogue fences only once a normaliser warp has asked for it (slow_mode), and publishes how far it has fenced.
struct TcNormState {
    uint32_t slow_mode;          // set by a normaliser warp that found a tile to clamp; never cleared
    uint32_t fenced_below[8];    // per epilogue warp: its rows of tiles with ordinal < this are visible gpu-wide
};

template <int NM, int HALF, typename OutT>
__device__ __forceinline__ void epilogue_role(const LogmelArgs& a, const int debug_stage, long long* trace, const int trace_first, TcBarriers* bars,
                                              TcNormState* norm, float* s_straddle,
                                              uint32_t tmem, int quad, int lane, int64_t total_tiles, int tiles_per_clip) {
    using L = TcEpilogueLayout<NM>;
    constexpr int ACC = L::acc_size(HALF);
    float acc[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = 0.f;
    const uint32_t d_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16) + kTcDCol + L::col0(HALF);
    const int me = 4 * HALF + quad;                                  // index among the 8 epilogue warps
    volatile uint32_t* const slow_mode = &norm->slow_mode;
    uint32_t d_parity = 0, buf = 0;
    // Fused normalisation: this warp's share of an utterance is counted (the normaliser warps wait for the count) one
    // tile late.  The count must follow the utterance's extremes: the two atomics return their old values, and the
    // count is issued only once those have come back (a register dependency instead of a fence).
    int64_t pending_clip = -1;
    uint32_t old_max = 0, old_min = 0, old_tile_max = 0, old_tile_min = 0;
    const int64_t my_tiles = static_cast<int64_t>(blockIdx.x) < total_tiles ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    TileCoord prev{0, 0};
    TileCursor cursor(tiles_per_clip);
    // Unit order on the tensor cores: 0, 1 (E sweep), 2, 3 (O sweep).  A tile is FINISHED (log10, stores, extremes) right
    // after its last unit: the tensor cores then wait for the next tile's E operand anyway, and unit 0 of the next tile
    // runs while the stores go out (tools/pipeline_model.py: better than finishing after the next tile's unit 0).
#pragma unroll 1
    for (int64_t k = 0; k < my_tiles; ++k) {
        constexpr bool more = true;
        const int ti = static_cast<int>(k);
        prev = cursor.at;
        cursor.advance();
#pragma unroll 1
        for (int u = 0; u < kTcUnits; ++u) {
            float d[L::cols(HALF)];
            if (more) {
                if (quad == 0) TC_TRACE(4 + HALF, ti, 3 * u);
                mbar_wait(&bars->d_full, d_parity);
                d_parity ^= 1u;
                tc_fence_after();
                tmem_ld_cols<L::cols(HALF)>(d_addr, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->d_empty);   // the accumulator is in registers: the next unit may overwrite it
                if (quad == 0) TC_TRACE(4 + HALF, ti, 3 * u + 1);
            }
            // Re and Im of a bin take the same weights: units 0 / 2 (even bins) share one body, units 1 / 3 (odd bins) the other
            if (debug_stage == 4) acc[0] += d[0] + d[L::cols(HALF) - 1];   // bring-up: loads only
            else if ((u & 1) == 0) tc_epilogue_unit<NM, 0, HALF>(d, acc);
            else tc_epilogue_unit<NM, 1, HALF>(d, acc);
            if (quad == 0) TC_TRACE(4 + HALF, ti, 3 * u + 2);
            if (u == kTcUnits - 1) {
                // ---- finish this tile ----
                if (pending_clip >= 0) {                       // count the tile before it
                    const bool fence = *slow_mode != 0;
                    if (fence) asm volatile("fence.acq_rel.gpu;" ::: "memory");
                    asm volatile("" ::"r"(old_max), "r"(old_min), "r"(old_tile_max), "r"(old_tile_min) : "memory");
                    __syncwarp();
                    if (lane == 0) {
                        atomicAdd(a.done_counters + pending_clip, 1u);
                        if (fence) *reinterpret_cast<volatile uint32_t*>(&norm->fenced_below[me]) = static_cast<uint32_t>(k);
                    }
                    pending_clip = -1;
                }
                if (quad == 0) TC_TRACE(4 + HALF, ti, 13);
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
                if (quad == 0) TC_TRACE(4 + HALF, ti, 14);
                // log10 clamp, coalesced row stores (lane = frame), utterance extremes
                const int f = quad * 32 + lane, t = prev.t0 + f;
                const bool live = t < a.n_frames;
                constexpr int m_begin = HALF == 0 ? 0 : L::low_mels, m_end = HALF == 0 ? L::low_mels : NM;
                const int64_t pitch = a.n_frames;
                OutT* const out = reinterpret_cast<OutT*>(a.out) + (prev.clip * NM + m_begin) * pitch + t;
                // With the normalisation fused, the affine half of it, (x + 4) / 4, is applied here (one FFMA, the same
                // single rounding as audio.py:156) and only the clamp at max - 8 is left for the normaliser warps - which
                // skip the utterance when its smallest value is not below max - 8 (tracked here as well).
                float mx = __uint_as_float(0xff800000u), mn = __uint_as_float(0x7f800000u);
                if (live) {
                    const float scale = a.fused_norm ? 0.25f : 1.0f, shift = a.fused_norm ? 1.0f : 0.0f;
                    const uint32_t pitch32 = static_cast<uint32_t>(a.n_frames);
                    // two mels per step: log2 on the MUFU, then log10 scaling and the affine map as packed FMUL2 / FFMA2
                    // (the same two roundings per value as the scalar form)
                    constexpr float kLog10Of2 = 0.30102999566398120f;
                    const float2 scale2 = make_float2(scale, scale), shift2 = make_float2(shift, shift);
#pragma unroll
                    for (int m = m_begin; m + 1 < m_end; m += 2) {
                        const float2 l2 = make_float2(log2_clamped(acc[m - L::acc_base(HALF)]), log2_clamped(acc[m + 1 - L::acc_base(HALF)]));
                        const float2 lg = __fmul2_rn(l2, make_float2(kLog10Of2, kLog10Of2));
                        const float2 y = __ffma2_rn(lg, scale2, shift2);
                        out_store(out + static_cast<uint64_t>(pitch32) * static_cast<uint32_t>(m - m_begin), y.x);
                        out_store(out + static_cast<uint64_t>(pitch32) * static_cast<uint32_t>(m + 1 - m_begin), y.y);
                        mx = max_nan(mx, max_nan(lg.x, lg.y));
                        mn = fminf(mn, fminf(lg.x, lg.y));
                    }
                    if constexpr ((m_end - m_begin) % 2 == 1) {
                        constexpr int m = m_end - 1;
                        const float lg = log10_clamped(acc[m - L::acc_base(HALF)]);
                        out_store(out + static_cast<uint64_t>(pitch32) * static_cast<uint32_t>(m - m_begin), fmaf(lg, scale, shift));
                        mx = max_nan(mx, lg);
                        mn = fminf(mn, lg);
                    }
                }
#pragma unroll
                for (int i = 0; i < ACC; ++i) acc[i] = 0.f;
                if (quad == 0) TC_TRACE(4 + HALF, ti, 15);
                uint32_t key = live ? max_key_encode(mx) : 0u;
                key = __reduce_max_sync(0xffffffffu, key);
                if (a.fused_norm) {
                    uint32_t inv = live ? ~max_key_encode(mn) : 0u;
                    inv = __reduce_max_sync(0xffffffffu, inv);
                    if (lane == 0) {
                        old_max = atomicMax(a.max_keys + (a.global_max ? 0 : prev.clip), key);
                        old_min = atomicMax(a.min_keys + prev.clip, inv);
                        if (a.tile_keys != nullptr) {   // the tile's own extremes: lets the clamp path skip or fill whole tiles
                            uint32_t* tk = a.tile_keys + 2 * (static_cast<int64_t>(blockIdx.x) + k * gridDim.x);
                            old_tile_max = atomicMax(tk, key);
                            old_tile_min = atomicMax(tk + 1, inv);
                        }
                    }
                    pending_clip = prev.clip;   // counted at the next finish
                } else if (lane == 0) {
                    atomicMax(a.max_keys + (a.global_max ? 0 : prev.clip), key);
                }
                if (quad == 0) TC_TRACE(4 + HALF, ti, 12);
                if (HALF == 0 && trace != nullptr && blockIdx.x == 0 && quad == 0 && lane == 0 && ti < kTileStamps)
                    trace[kTraceRoles * kTraceTiles * kTraceEvents + 6 * kStampCtas + kTileStamps + ti] = clock64();
            }
        }
    }
    // the last tile is counted behind an unconditional fence; after it every row of this warp is visible
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("" ::"r"(old_max), "r"(old_min), "r"(old_tile_max), "r"(old_tile_min) : "memory");
    __syncwarp();
    if (lane == 0) {
        if (pending_clip >= 0) atomicAdd(a.done_counters + pending_clip, 1u);
        *reinterpret_cast<volatile uint32_t*>(&norm->fenced_below[me]) = 0xffffffffu;
    }
}

// BRINGUP = false is the production build: the timeline stamps and the staged bring-up modes (B200MEL_TC_TRACE,
// B200MEL_TC_DEBUG) fold away, which also keeps the hot code inside the 32 KB instruction cache.
template <typename InT, int NM, bool BRINGUP, typename OutT>
__global__ void __launch_bounds__(kTcThreads, 1)
logmel_tc_kernel(const __grid_constant__ LogmelArgs a, const __grid_constant__ CUtensorMap audio_map, const int tma_rows,
                 const unsigned char* __restrict__ operands, const int debug_arg, long long* __restrict__ trace_arg, const int trace_first) {
    long long* const trace = BRINGUP ? trace_arg : nullptr;
    const int debug_stage = BRINGUP ? debug_arg : 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_audio = reinterpret_cast<float*>(smem_raw + kSmemAudio);
    float* s_straddle = reinterpret_cast<float*>(smem_raw + kSmemStraddle);
    __shared__ __align__(8) TcBarriers bars;
    __shared__ TcNormState norm_state;
    __shared__ uint32_t s_tmem;

    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from
    // it (role, TMEM lane quadrant, column addresses) in uniform registers
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31, quad = warp & 3;
    if (trace != nullptr && tid == 0) {   // bring-up: every CTA stamps its start and end (cycles and nanoseconds)
        long long* stamp = trace + kTraceRoles * kTraceTiles * kTraceEvents + 6 * blockIdx.x;
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        stamp[0] = clock64(); stamp[1] = static_cast<long long>(ns);
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        trace[kTraceRoles * kTraceTiles * kTraceEvents + 6 * kStampCtas + 2 * kTileStamps + 32 + blockIdx.x] = smid;
    }
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    int64_t total_tiles = a.batch * tiles_per_clip;
    // bring-up aid (B200MEL_TC_DEBUG): 1 = setup only, 2 = + producer and folds of ONE tile,
    // 3 = + the tensor cores, 4 = + accumulator loads, 6 = everything for one tile
    if (debug_stage > 0 && debug_stage != 7 && total_tiles > gridDim.x) total_tiles = gridDim.x;
    if (debug_stage == 1) total_tiles = 0;

    // ---- one-time setup: tensor memory, barriers, constant matrices -> shared memory ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&bars.audio_full, 9);
        mbar_init(&bars.audio_empty, 8);
        mbar_init(&bars.tma_done, 1);
        mbar_init(&bars.a_full[0], 8); mbar_init(&bars.a_full[1], 8);
        mbar_init(&bars.a_empty[0], 1); mbar_init(&bars.a_empty[1], 1);
        mbar_init(&bars.d_full, 1);
        mbar_init(&bars.d_empty, 8);
        norm_state.slow_mode = 0;
        for (int i = 0; i < 8; ++i) norm_state.fenced_below[i] = 0;
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
        // ===== fold warps: two per lane quadrant; both do half of the E sweep, then half of the O sweep =====
        const int part = warp < kWarpO ? 0 : 1;
        const float* fr = s_audio + (quad * 32 + lane) * kTcRowPitch;
        uint32_t parity = 0;
        int ti = 0;
        TileCursor cursor(tiles_per_clip);
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti, cursor.advance()) {
            if (quad == 0) TC_TRACE(1 + part, ti, 0);
            if (trace != nullptr && blockIdx.x == 0 && tid == 0 && ti < kTileStamps)
                trace[kTraceRoles * kTraceTiles * kTraceEvents + 6 * kStampCtas + ti] = clock64();
            const TileCoord tcl = cursor.at;
            if (tile_uses_tma<InT>(a, tma_rows, tcl)) {
                if (lane == 0) mbar_arrive(&bars.audio_full);   // the loader warp brings this tile
            } else {
                mbar_wait(&bars.audio_empty, parity ^ 1u);      // every fold warp has finished reading the previous tile
                produce_tile<InT>(a, tcl, s_audio, &bars.audio_full, tid);
            }
            mbar_wait(&bars.audio_full, parity);
            if (quad == 0) TC_TRACE(1 + part, ti, 1);
#pragma unroll 1
            for (int sweep = 0; sweep < 2; ++sweep) {
                mbar_wait(&bars.a_empty[sweep], parity ^ 1u);   // the tensor cores are done with the previous tile's operand
                if (quad == 0) TC_TRACE(1 + part, ti, 2 + 3 * sweep);
                tc_fence_after();
                // warp `part` takes the first chunks of the E sweep and the last ones of the O sweep (7 + 6 either way)
                const bool first_half = (part == 0) == (sweep == 0);
                sweep_store(sweep, first_half ? 0 : kFoldSplit, first_half ? kFoldSplit : kTcChunks, fr, lane_addr);
                if (quad == 0) TC_TRACE(1 + part, ti, 3 + 3 * sweep);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (sweep == 1) mbar_arrive(&bars.audio_empty);
                    mbar_arrive(&bars.a_full[sweep]);
                }
                if (quad == 0) TC_TRACE(1 + part, ti, 4 + 3 * sweep);
            }
            parity ^= 1u;
        }
    } else if (warp < kWarpMma) {
        // ===== epilogue warps =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
        if (debug_stage > 0 && debug_stage < 4) total_tiles = 0;
        if (warp < kWarpEpi1) epilogue_role<NM, 0, OutT>(a, debug_stage, trace, trace_first, &bars, &norm_state, s_straddle, tmem, quad, lane, total_tiles, tiles_per_clip);
        else epilogue_role<NM, 1, OutT>(a, debug_stage, trace, trace_first, &bars, &norm_state, s_straddle, tmem, quad, lane, total_tiles, tiles_per_clip);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp == kWarpMma && (debug_stage == 0 || debug_stage >= 3)) {
            // ===== tensor-core issue: the whole warp walks the loop, one elected lane issues =====
            // One compact loop over the units (per-unit columns and matrix offsets from a constant table): the issue
            // code stays small so it does not evict the fold and epilogue code from the instruction caches.
            const uint32_t desc0 = operand_desc_lo(smem_u32(smem_raw + kSmemOperands)), d_tmem = tmem + kTcDCol;
            const uint32_t a_full0 = smem_u32(&bars.a_full[0]), a_empty0 = smem_u32(&bars.a_empty[0]);
            constexpr uint32_t kStep = (2 * kTcStripBytes) >> 4;   // K step s = strips 2s, 2s+1 = slots 16s..16s+15
            uint32_t a_parity = 0, d_parity = 1;   // d_empty: the first wait passes (accumulator starts free)
            int ti = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
#pragma unroll 1
                for (int u = 0; u < kTcUnits; ++u) {
                    const uint32_t sweep_bar = static_cast<uint32_t>(u & 2) << 2;   // byte offset of the sweep's barrier (0 or 8)
                    if ((u & 1) == 0) mbar_wait_addr(a_full0 + sweep_bar, a_parity);
                    TC_TRACE(3, ti, 3 * u);
                    if (debug_stage != 3) mbar_wait(&bars.d_empty, d_parity);
                    else if (u > 0) { mbar_wait(&bars.d_full, (u - 1) & 1); }
                    d_parity ^= 1u;
                    TC_TRACE(3, ti, 3 * u + 1);
                    tc_fence_after();
                    const TcUnitIssue ui = c_unit_issue[u];
                    uint32_t a_hi = tmem + ui.a_hi, a_lo = tmem + ui.a_lo, b_hi = desc0 + ui.b_hi, b_lo = desc0 + ui.b_lo;
                    mma_f16_ts<false>(d_tmem, a_hi, b_hi);
                    mma_f16_ts<true>(d_tmem, a_lo, b_hi);
                    mma_f16_ts<true>(d_tmem, a_hi, b_lo);
#pragma unroll 1
                    for (int s = 1; s < kTcMainSteps; ++s) {
                        a_hi += 8; a_lo += 8; b_hi += kStep; b_lo += kStep;
                        mma_f16_ts<true>(d_tmem, a_hi, b_hi);
                        mma_f16_ts<true>(d_tmem, a_lo, b_hi);
                        mma_f16_ts<true>(d_tmem, a_hi, b_lo);
                    }
                    // slots 96..101: one K step over the unit's [hi | lo] leftover columns, (hi + lo) Bh then hi Bl
                    mma_f16_ts<true>(d_tmem, tmem + ui.a_left, desc0 + ui.b_left0);
                    mma_f16_ts<true>(d_tmem, tmem + ui.a_left, desc0 + ui.b_left1);
                    mma_commit(&bars.d_full);
                    TC_TRACE(3, ti, 3 * u + 2);
                    if (u & 1) mma_commit_addr(a_empty0 + sweep_bar);   // both units of the sweep have consumed its operand
                }
                a_parity ^= 1u;
            }
            if (debug_stage == 3 && total_tiles > 0) mbar_wait(&bars.d_full, 1);   // nobody drains the accumulator in this stage
        } else if (warp == kWarpLoad) {
            // ===== loader warp =====
            // Ask L2 for a tile one tile period before it is copied: the CTAs of a wave load in lock-step, so without
            // the prefetch every staging phase waits on an HBM burst while HBM idles the rest of the time.
            auto prefetch = [&](const TileCoord& tp) {
                if (tile_uses_tma<InT>(a, tma_rows, tp)) tma_prefetch_tile(&audio_map, tp);
                else prefetch_tile_l2<InT>(a, tp);
            };
            TileCursor cursor(tiles_per_clip);
            if (lane == 0 && static_cast<int64_t>(blockIdx.x) < total_tiles) prefetch(cursor.at);
            uint32_t parity = 1;   // audio_empty: the first wait passes
            uint32_t patch_parity = 0;
            int ti = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti, cursor.advance()) {
                TC_TRACE(0, ti, 0);
                if (lane == 0 && tile + gridDim.x < total_tiles) prefetch(cursor.peek_next());
                const TileCoord tcl = cursor.at;
                const bool tma = tile_uses_tma<InT>(a, tma_rows, tcl);
                mbar_wait(&bars.audio_empty, parity);           // every fold warp has finished reading the previous tile
                parity ^= 1u;
                TC_TRACE(0, ti, 1);
                if (tma && tile_needs_patch(a, tma_rows, tcl)) {
                    // a clip's first or last tile: the copy completes on a private barrier, then the edge rows are rewritten
                    if (lane == 0) tma_load_tile(&audio_map, tcl, s_audio, &bars.tma_done);
                    mbar_wait(&bars.tma_done, patch_parity);
                    patch_parity ^= 1u;
                    TC_TRACE(0, ti, 3);
                    patch_tile_edges(a, tma_rows, tcl, s_audio, lane);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.audio_full);
                } else if (lane == 0) {
                    if (tma) tma_load_tile(&audio_map, tcl, s_audio, &bars.audio_full);
                    else mbar_arrive(&bars.audio_full);         // cooperative mode: the fold warps bring the tile
                }
                TC_TRACE(0, ti, 2);
            }
        } else if (warp >= kWarpNorm && a.fused_norm && debug_stage == 0 && static_cast<int64_t>(blockIdx.x) < total_tiles) {
            // ===== normaliser warps =====
            // The two warps take this CTA's tiles in alternating groups of 32, one tile per lane: a lane waits until
            // every tile of its utterance has been counted (the max is final then), and decides from the utterance's
            // extremes whether the clamp at max - 8 touches it at all; tiles that need it (digital silence, zero
            // padding) are clamped in place by the whole warp while they are still in L2.
            const unsigned need = 8u * static_cast<unsigned>(tiles_per_clip);
            const int64_t my_tiles = (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
            for (int64_t g0 = 32 * (warp - kWarpNorm); g0 < my_tiles; g0 += 64) {
                const bool have = g0 + lane < my_tiles;
                const TileCoord tc = tile_coord(blockIdx.x + (have ? g0 + lane : 0) * gridDim.x, tiles_per_clip);
                if (have) {
                    const volatile uint32_t* counter = a.done_counters + tc.clip;
                    uint32_t polls = 0;
                    while (*counter < need) {
                        if (++polls > (1u << 21)) tc_fault(0x2000000u | (static_cast<uint32_t>(g0 + lane) << 8 & 0xffff00u) | (threadIdx.x >> 5));
                        __nanosleep(1000);
                    }
                }
                __syncwarp();
                __threadfence();   // the extremes (and the rows) are read after the counts
                float g = 0.f;
                int action = 0;                        // 0: leave the tile alone, 1: clamp it in place, 2: fill it with the clamp value
                if (have) {
                    g = max_key_decode(__ldcg(a.max_keys + tc.clip));
                    const float floor_lg = g - 8.0f;
                    const float smallest = max_key_decode(~__ldcg(a.min_keys + tc.clip));
                    if (!(smallest >= floor_lg)) {     // something in the utterance is below the clamp (or the max is NaN)
                        action = 1;
                        if (a.tile_keys != nullptr) {
                            const uint32_t* tk = a.tile_keys + 2 * (static_cast<int64_t>(blockIdx.x) + (g0 + lane) * gridDim.x);
                            const float tile_max = max_key_decode(__ldcg(tk)), tile_min = max_key_decode(~__ldcg(tk + 1));
                            if (tile_min >= floor_lg) action = 0;          // this tile is wholly above the clamp
                            else if (tile_max < floor_lg) action = 2;      // wholly below it (digital silence, zero padding)
                        }
                    }
                }
                unsigned todo = __ballot_sync(0xffffffffu, action != 0);
                if (todo != 0) {
                    // ask the epilogue warps to fence what they store from now on (see TcNormState) ...
                    if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&norm_state.slow_mode) = 1u;
                    __syncwarp();
                }
                while (todo != 0) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    // ... and wait until every one of them has fenced the rows of this tile (ordinal g0 + src)
                    if (lane < 8) {
                        const volatile uint32_t* fenced = &norm_state.fenced_below[lane];
                        uint32_t polls = 0;
                        while (*fenced <= static_cast<uint32_t>(g0 + src)) {
                            if (++polls > (1u << 21)) tc_fault(0x3000000u | (static_cast<uint32_t>(g0 + src) << 8 & 0xffff00u) | (threadIdx.x >> 5));
                            __nanosleep(500);
                        }
                    }
                    __syncwarp();
                    __threadfence();
                    const int64_t clip = __shfl_sync(0xffffffffu, static_cast<int>(tc.clip), src);
                    const int t0 = __shfl_sync(0xffffffffu, tc.t0, src);
                    const float gs = __shfl_sync(0xffffffffu, g, src);
                    const bool fill = __shfl_sync(0xffffffffu, action, src) == 2;
                    const int frames = a.n_frames - t0 < kTcTileFrames ? a.n_frames - t0 : kTcTileFrames;
                    const float floor_y = ((gs - 8.0f) + 4.0f) * 0.25f;
                    OutT* tile_out = reinterpret_cast<OutT*>(a.out) + clip * NM * static_cast<int64_t>(a.n_frames) + t0;
                    if (fill) fill_tile_tc<NM, OutT>(tile_out, a.n_frames, frames, floor_y, lane);
                    else normalise_tile_tc<NM, OutT>(tile_out, a.n_frames, frames, floor_y, lane);
                }
            }
        }
        __syncwarp();
    }

    if (trace != nullptr && lane == 0 && (warp == kWarpEpi0 || warp == kTcWarps - 1))   // pipeline / normaliser done, every CTA
        trace[kTraceRoles * kTraceTiles * kTraceEvents + 6 * blockIdx.x + (warp == kWarpEpi0 ? 4 : 5)] = clock64();
    if (trace != nullptr && blockIdx.x == 0 && lane == 0)   // when each warp of CTA 0 reaches the final barrier
        trace[kTraceRoles * kTraceTiles * kTraceEvents + 6 * kStampCtas + 2 * kTileStamps + warp] = clock64();
    tc_fence_before();
    __syncthreads();
    if (trace != nullptr && tid == 0) {
        long long* stamp = trace + kTraceRoles * kTraceTiles * kTraceEvents + 6 * blockIdx.x;
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        stamp[2] = clock64(); stamp[3] = static_cast<long long>(ns);
    }
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
        err = cudaFuncSetAttribute(logmel_tc_kernel<InT, NM, false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(logmel_tc_kernel<InT, NM, true, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err != cudaSuccess) return err;
        err = cudaFuncSetAttribute(logmel_tc_kernel<InT, NM, false, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err != cudaSuccess) return err;
        int sms = 0;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return err;
        sms_by_device[device] = sms;
    }
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    const int64_t tiles = a.batch * tiles_per_clip;
    const unsigned grid = static_cast<unsigned>(tiles < sms_by_device[device] ? tiles : sms_by_device[device]);
    // bring-up aid: host-mapped fault words a timed-out wait fills in before trapping; reported at exit
    static unsigned* fault_host = nullptr;
    if (fault_host == nullptr && cudaHostAlloc(&fault_host, 64, cudaHostAllocMapped) == cudaSuccess) {
        fault_host[0] = fault_host[1] = 0;
        unsigned* fault_dev = nullptr;
        if (cudaHostGetDevicePointer(&fault_dev, fault_host, 0) == cudaSuccess)
            cudaMemcpyToSymbolAsync(g_tc_fault, &fault_dev, sizeof(fault_dev), 0, cudaMemcpyHostToDevice, stream);
        static unsigned* fault_report = fault_host;
        std::atexit([] {
            if (fault_report[0] != 0)
                std::fprintf(stderr, "b200mel tcgen05 kernel: wait timed out, code 0x%08x in CTA %u\n", fault_report[0], fault_report[1]);
        });
    }
    // the batch as the TMA unit sees it (see "loaders" above); any reason it cannot be described leaves tma_rows = 0
    // and every tile in cooperative mode
    CUtensorMap audio_map;
    std::memset(&audio_map, 0, sizeof(audio_map));
    int tma_rows = 0;
    if (sizeof(InT) == 4 && (reinterpret_cast<uintptr_t>(a.audio) & 15u) == 0 && a.stride_b % 4 == 0 && a.n_samples >= 284 + kHop &&
        a.batch < (int64_t{1} << 31)) {
        using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = [] {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
            return reinterpret_cast<EncodeFn>(fn);
        }();
        // rows r with 160 r + 3 * 40 + 164 <= n_samples: every element of such a row lies inside the utterance's memory
        const int64_t rows = (a.n_samples - 284) / kHop + 1;
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(kTcRowPitch), 4, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(a.batch)};
        const cuuint64_t strides[3] = {kHop, kHop * 4, static_cast<cuuint64_t>(a.stride_b) * 4};   // bytes, dims 1..3
        const cuuint32_t box[4] = {static_cast<cuuint32_t>(kTcRowPitch), 1, static_cast<cuuint32_t>(kTcAudioRows), 1};
        const cuuint32_t elem[4] = {1, 1, 1, 1};
        if (encode != nullptr && rows < (int64_t{1} << 31) &&
            encode(&audio_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(a.audio), dims, strides, box, elem,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            tma_rows = static_cast<int>(rows);
    }
    static const bool no_tma = std::getenv("B200MEL_TC_NO_TMA") != nullptr;   // bring-up: force cooperative loading
    if (no_tma) tma_rows = 0;
    static const bool say_tma = std::getenv("B200MEL_TC_VERBOSE") != nullptr;
    if (say_tma) std::fprintf(stderr, "b200mel tcgen05: tma_rows = %d\n", tma_rows);
    ProfileScope profile(2, stream);
    static const int debug_stage = std::getenv("B200MEL_TC_DEBUG") ? std::atoi(std::getenv("B200MEL_TC_DEBUG")) : 0;
    static long long* trace = nullptr;
    static const bool want_trace = std::getenv("B200MEL_TC_TRACE") != nullptr;   // value: first of the 8 traced tiles of CTA 0
    static const int trace_first = want_trace ? std::atoi(std::getenv("B200MEL_TC_TRACE")) : 0;
    constexpr int kTraceWords = kTraceRoles * kTraceTiles * kTraceEvents;
    constexpr size_t kTraceBytes = sizeof(long long) * (kTraceWords + 6 * kStampCtas + TC_TILE_STAMPS);
    if (want_trace && trace == nullptr) { cudaMalloc(&trace, kTraceBytes); }
    if (want_trace) cudaMemsetAsync(trace, 0, kTraceBytes, stream);
    if (a.out_f16)
        logmel_tc_kernel<InT, NM, false, __half><<<grid, kTcThreads, kSmemBytes, stream>>>(a, audio_map, tma_rows, tables->operands, 0, nullptr, 0);
    else if (want_trace || debug_stage != 0)
        logmel_tc_kernel<InT, NM, true, float><<<grid, kTcThreads, kSmemBytes, stream>>>(a, audio_map, tma_rows, tables->operands, debug_stage,
                                                                                           want_trace ? trace : nullptr, trace_first);
    else
        logmel_tc_kernel<InT, NM, false, float><<<grid, kTcThreads, kSmemBytes, stream>>>(a, audio_map, tma_rows, tables->operands, 0, nullptr, 0);
    count_launch();
    err = cudaGetLastError();
    if (want_trace && err == cudaSuccess) {   // bring-up only: synchronises and prints CTA 0's timeline
        static long long host[kTraceWords + 6 * kStampCtas + TC_TILE_STAMPS];
        cudaStreamSynchronize(stream);
        cudaMemcpy(host, trace, kTraceBytes, cudaMemcpyDeviceToHost);
        {
            const long long* st = host + kTraceWords;
            long long ns0 = 0, ns1 = 0, cyc_min = 0, cyc_max = 0; double cyc_sum = 0, mhz_sum = 0;
            for (unsigned b = 0; b < grid && b < kStampCtas; ++b) {
                const long long cyc = st[6 * b + 2] - st[6 * b], ns = st[6 * b + 3] - st[6 * b + 1];
                if (b == 0 || st[6 * b + 1] < ns0) ns0 = st[6 * b + 1];
                if (b == 0 || st[6 * b + 3] > ns1) ns1 = st[6 * b + 3];
                if (b == 0 || cyc < cyc_min) cyc_min = cyc;
                if (b == 0 || cyc > cyc_max) cyc_max = cyc;
                cyc_sum += cyc; mhz_sum += ns > 0 ? 1e3 * cyc / ns : 0;
            }
            {
                long long pmin = 0, pmax = 0; double psum = 0;
                for (unsigned b = 0; b < grid && b < kStampCtas; ++b) {
                    const long long pc = st[6 * b + 4] - st[6 * b];
                    if (b == 0 || pc < pmin) pmin = pc;
                    if (b == 0 || pc > pmax) pmax = pc;
                    psum += pc;
                }
                std::fprintf(stderr, "trace CTAs: epilogue of the last tile done after min %lld avg %.0f max %lld cycles\n", pmin, psum / grid, pmax);
                std::fprintf(stderr, "trace CTA pipeline k-cycles by SM id (sm:kcycles@MHz):");
                const long long* sm = host + kTraceWords + 6 * kStampCtas + 2 * kTileStamps + 32;
                for (unsigned want = 0; want < 160; ++want)
                    for (unsigned b = 0; b < grid && b < kStampCtas; ++b)
                        if (sm[b] == want) {
                            const long long ns = st[6 * b + 3] - st[6 * b + 1];
                            std::fprintf(stderr, " %u:%lld@%lld", want, (st[6 * b + 4] - st[6 * b]) / 1000, ns > 0 ? 1000 * (st[6 * b + 2] - st[6 * b]) / ns : 0);
                        }
                std::fprintf(stderr, "\n");
            }
            std::fprintf(stderr, "trace CTAs: span %lld ns, cycles min %lld avg %.0f max %lld, SM clock %.0f MHz, %.1f tiles per CTA\n",
                         ns1 - ns0, cyc_min, cyc_sum / grid, cyc_max, mhz_sum / grid, static_cast<double>(tiles) / grid);
        }
        {
            const long long* ts = host + kTraceWords + 6 * kStampCtas;
            std::fprintf(stderr, "trace CTA 0 tile starts (cycles after the CTA's start; then deltas):");
            for (int i = 0; i < kTileStamps && ts[i] != 0; ++i)
                std::fprintf(stderr, " %lld", i == 0 ? ts[0] - host[kTraceWords] : ts[i] - ts[i - 1]);
            std::fprintf(stderr, "  | end after last start: %lld\n", host[kTraceWords + 2] - ts[(tiles + grid - 1) / grid - 1]);
            std::fprintf(stderr, "trace CTA 0 warps reach the final barrier at (k cycles):");
            for (int w = 0; w < kTcWarps; ++w) std::fprintf(stderr, " %lld", (ts[2 * kTileStamps + w] - host[kTraceWords]) / 1000);
            std::fprintf(stderr, "\n");
            std::fprintf(stderr, "trace CTA 0 epilogue ends minus fold starts:");
            for (int i = 0; i < kTileStamps && ts[i] != 0; ++i) std::fprintf(stderr, " %lld", ts[kTileStamps + i] - ts[i]);
            std::fprintf(stderr, "\n");
        }
        long long t0 = 0;
        for (int i = 0; i < kTraceWords; ++i) { const long long v = host[i]; if (v != 0 && (t0 == 0 || v < t0)) t0 = v; }
        static const char* names[kTraceRoles] = {"producer", "fold-E", "fold-O", "mma", "epi-0", "epi-1"};
        for (int r = 0; r < kTraceRoles; ++r)
            for (int t = 0; t < kTraceTiles; ++t) {
                std::fprintf(stderr, "trace %-8s tile %d:", names[r], t + trace_first);
                for (int e = 0; e < kTraceEvents; ++e) {
                    const long long v = host[(r * kTraceTiles + t) * kTraceEvents + e];
                    if (v) std::fprintf(stderr, " %d:%lld", e, v - t0);
                }
                std::fprintf(stderr, "\n");
            }
    }
    return err;
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
