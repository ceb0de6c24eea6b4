// tcgen05 (tensor-core) variant of the fused log-mel front-end for sm_100a: the folded DFT as four
// 128 x 104 x 112 GEMMs per 128-frame tile, 3-product fp16 split precision (math: tc_core.cuh;
// reference: whisper/audio.py:145-156).  DESIGN.md section 4.1 has the picture; in short:
//
// One persistent CTA per SM, 20 warps, warp-specialised, no __syncthreads in the steady state (mbarriers only):
//   3 x 4 fold warps: one thread per frame (= TMEM lane), three warps per lane quadrant that split every sweep between
//                     them.  The tile's 130 rows of 160 samples come as TWO half tiles of 66 rows (frames 0-63 / 64-127),
//                     each ONE TMA tensor copy into its own buffer at pitch 164 words (a 4-D tensor map whose rows overlap,
//                     see "loaders"): the last of a half's four warps to finish reading it issues the copy of the next
//                     tile's half itself (and the L2 prefetch of the one after) - no hand-over to a loader warp on the
//                     critical path.  A clip's first / last tile: the copy zero-fills what it cannot address and the warps
//                     rewrite the 1-3 rows of real / reflected samples.  int16 PCM: one bulk copy of the half's samples
//                     into the top of its buffer, expanded to float32 rows in place.  (`lengths` cuts, unaligned rows: the
//                     warps stage the half themselves.)  Then E sweep: ee / eo, O sweep: oe / oo (window multiply and both
//                     folds fused, packed FMUL2 / FFMA2), every value split into fp16 hi + lo and written straight into
//                     TENSOR MEMORY as the A operand (tcgen05.st) - the data never touch shared memory again.  The
//                     quadrant's scale step (largest |sample| -> power of two, tc_core.cuh) costs no pass of its own: the
//                     E sweep runs with the previous tile's step while it tracks the largest |sample| it reads, and is
//                     repeated in the rare case that the step it should have used is another one;
//   MMA warp        : one elected thread issues, per unit, 6 K-steps x 3 products + 2 leftover steps of
//                     tcgen05.mma.kind::f16 (M 128, N 104, K 16; A from TMEM, B = the constant matrix from
//                     shared memory, fp32 accumulator in TMEM): lo Bh + hi Bl first, hi Bh last, then
//                     tcgen05.commit -> mbarrier hands the accumulator to the epilogue;
//   4 epilogue      : one warp per TMEM lane quadrant (a thread = a frame with ALL its mels in registers - the four warps
//     warps           run the same code, which is what the instruction cache wants) pulls the 104 accumulator columns
//                     into registers (tcgen05.ld), releases the accumulator, and adds w d^2 to the mels
//                     of each bin - mel structure and weights are compile-time constants (FFMA immediates),
//                     partial sums in registers.  A tile is finished - log2 on the MUFU, ONE fused multiply-add for
//                     (log10 + 4) / 4 with the data scale 2^-2k in its addend, the 1e-10 clamp as max(., -1.5), 128-byte
//                     coalesced row stores (immediate offsets when the output pitch is the usual 3000 frames), the
//                     utterance's and the tile's extremes (warp REDUX + atomicMax) - after the NEXT tile's first unit has
//                     been pulled (80 mels), so the tensor cores never wait for the stores.
// Digital silence (zero padding, `lengths`) is cheap: a lane quadrant whose rows are all zeros gets a zero operand instead
// of the sweeps and hands its rows back at once; a tile of nothing but silence is neither multiplied nor pulled nor stored.
// No CTA ever waits for another one - there is no cross-CTA hand-over at all - so the kernel makes progress with any
// number of co-resident CTAs.  What is left of the normalisation, the clamp at max - 8, needs every tile of an utterance:
// the FINISH kernel right behind (tc_finish_kernel, a warp per tile) decides from the utterance's and the tile's extremes
// what the clamp does to the tile: nothing (the usual case: it reads three words and moves on), a constant fill (digital
// silence / zero padding - tiles whose samples are all zero are not even stored by the epilogue) or a clamp in place while
// the tile is still in L2.  For speech-like input the DRAM traffic stays the algorithmic read + write.
// Tensor memory is exactly full: 408 operand columns (4 units x [hi | lo], tc_core.cuh) + 104 accumulator; so is shared
// memory (DFT matrices + the two half tiles).  The hot code of all roles has to fit the 32 KB instruction cache: loops
// over table rows instead of unrolled code wherever the work is regular.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_core.cuh"

namespace b200mel {

namespace {

constexpr int kFoldParts = 3;                 // fold warps per lane quadrant
constexpr int kWarpEpi0 = 4 * kFoldParts, kWarpMma = kWarpEpi0 + 4;   // fold 0-11, epilogue 12-15, MMA 16 (17-19 complete the warpgroup and idle)
constexpr int kTcWarps = 20;
constexpr int kTcFixedPitch = 3000;          // N_FRAMES of a 30 s clip (audio.py:21): the kernels specialised for this output pitch
constexpr int kTcThreads = kTcWarps * 32;   // 640
constexpr uint32_t kSpinLimit = 1u << 23;    // polls of >= 64 ns: a protocol bug ends the kernel after a second or so instead of hanging the device
constexpr uint32_t kPollNs = 64;             // sleep between two polls of a barrier

// Optional timeline (compile with -DB200MEL_TC_TRACE, tools/tc_trace.py): CTA 0 stamps clock64() at the hand-over
// points of 8 of its tiles.  Compiled out of the production library.
constexpr int kTraceTiles = 8, kTraceEvents = 16, kTraceRoles = 6;
constexpr int kTraceWords = kTraceRoles * kTraceTiles * kTraceEvents + 8;
#if defined(B200MEL_TC_TRACE) || defined(B200MEL_TC_SWITCHES)
// (bring-up switches ride in the upper bits of trace_first: 0x400 = no L2 prefetch, 0x800 = bulk instead of tensor L2 prefetch, 0x1000 = no finish in the epilogue, 0x2000 = no mel sums, 0x4000 = never repeat the E sweep, 0x8000 = one chunk per fold warp and sweep, 0x10000 = no short cut for digital silence - measurement only, wrong results)
#define TC_DEBUG_FLAG(bit) ((trace_first_arg & (bit)) != 0)
#else
#define TC_DEBUG_FLAG(bit) false
#endif
#if defined(B200MEL_TC_TRACE)
#define TC_TRACE(role, tile_index, event)                                                                        \
    do {                                                                                                         \
        if (trace != nullptr && blockIdx.x == 0 && (tile_index) >= trace_first && (tile_index) < trace_first + kTraceTiles &&      \
            (threadIdx.x & 31) == 0)                                                                             \
            trace[((role) * kTraceTiles + (tile_index) - trace_first) * kTraceEvents + (event)] = clock64();    \
    } while (0)
#else
#define TC_TRACE(role, tile_index, event) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A wait that exceeds the spin limit (a protocol bug, never the data) records what hung in g_tc_fault (b200mel_kernel_fault
// reads it), raises the CTA's abort flag and returns: every later wait of the CTA falls through, so the kernel runs to its end
// with garbage in this launch's output instead of hanging - or trapping and taking the caller's CUDA context with it.
__device__ unsigned g_tc_fault[2] = {0u, 0u};
struct TcAbort { volatile uint32_t* flag; };

__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity, volatile uint32_t* abort) {
    uint32_t spins = 0;
    do {
        __nanosleep(kPollNs);
        if (*abort != 0) return;
        if (++spins > kSpinLimit) {
            g_tc_fault[0] = 0x1000000u | ((addr & 0xfffu) << 12) | (parity << 8) | (threadIdx.x >> 5);
            g_tc_fault[1] = blockIdx.x;
            *abort = 1u;
            return;
        }
    } while (!mbar_try(addr, parity));
}
// A wait is a try_wait without a suspend-time hint (it blocks for the hardware's own short time limit), then such tries a
// short sleep apart: a waiting role takes next to no issue slots from the roles that are working.  Measured on one box
// (tools/ab_libs.sh, medians of 5 x 200 launches): 1 - 3 % faster for the kernel as a whole than parking the warp with a
// 20 us hint, whose wake-up is the slower one; a pure test_wait between the sleeps is 3 % slower.
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity, TcAbort ab) {
    if (!mbar_try(addr, parity)) mbar_wait_slow(addr, parity, ab.flag);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, TcAbort ab) { mbar_wait_addr(smem_u32(bar), parity, ab); }
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
// tensor-core work for a unit is 6 K steps x (lo Bh, hi Bl), the leftover correction, 6 K steps of hi Bh and the leftover step
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
constexpr int kSmemAudio = kTcOperandBytes;                                       // two half tiles: 66 rows x 656 B each
constexpr int kSmemBytes = kSmemAudio + kTcHalfStride + kTcHalfBytes;
static_assert(kSmemAudio % 128 == 0 && kTcHalfStride % 128 == 0 && kSmemBytes + 1024 <= 227 * 1024, "shared memory budget");

struct TcBarriers {
    uint64_t audio_full[2];           // per half tile: the tensor / bulk copy has landed (1 arrival + its bytes)
    uint64_t a_full[2], a_empty[2];   // [E sweep, O sweep]
    uint64_t d_full, d_empty;
};

// what the folds tell the epilogue about a tile: the scale step of each lane quadrant and whether all its samples are
// zero, [tile parity][quadrant]
struct TcTileInfo { uint32_t scale[2][4]; uint32_t silent[2][4]; uint32_t quad_max[4][4]; uint32_t probe[4][4]; uint32_t released[2][2]; };

// ---- normaliser -------------------------------------------------------------------------------------
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

// The finish kernel's blocks (8 warps) work on one tile at a time together: warp w takes mel rows w, w + 8, ...; a lane is a
// column of four frames.  All of a warp's loads are issued before its first store, so a tile costs ONE round trip to L2 -
// a warp walking the rows of a tile alone would pay one per row pair, and a batch in which only a few tiles need the clamp
// (128 mels: a handful per 6144; zero-padded clips: the one tile per clip where the sound stops) would wait for it.
constexpr int kFinishWarps = 8;

// fill: every value of the tile was below the clamp (or never stored) - stores only
template <int NM, typename OutT>
__device__ __forceinline__ void fill_tile_rows(OutT* __restrict__ tile_out, int64_t pitch, int frames, float v, int warp, int lane) {
    if ((pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(tile_out) & (4 * sizeof(OutT) - 1)) == 0 && (frames & 3) == 0) {
        if (lane < (frames >> 2)) {
            const float4 v4 = make_float4(v, v, v, v);
#pragma unroll 4
            for (int row = warp; row < NM; row += kFinishWarps) out_store4(tile_out + row * pitch + 4 * lane, v4);
        }
    } else {
        for (int row = warp; row < NM; row += kFinishWarps)
            for (int i = lane; i < frames; i += 32) out_store(tile_out + row * pitch + i, v);
    }
}

// clamp in place (NM rows of `frames` values at `pitch`); g: the clamp in rescaled units
template <int NM, typename OutT>
__device__ __forceinline__ void normalise_tile_rows(OutT* __restrict__ tile_out, int64_t pitch, int frames, float g, int warp, int lane) {
    constexpr int kRows = NM / kFinishWarps;                                    // 10 or 16 rows per warp
    static_assert(NM % kFinishWarps == 0, "rows per warp");
    if ((pitch & 3) == 0 && (reinterpret_cast<uintptr_t>(tile_out) & (4 * sizeof(OutT) - 1)) == 0) {
        if (lane < (frames >> 2)) {
            OutT* p = tile_out + warp * pitch + 4 * lane;
            float4 x[kRows];
#pragma unroll
            for (int i = 0; i < kRows; ++i) x[i] = out_load4(p + i * kFinishWarps * pitch);
#pragma unroll
            for (int i = 0; i < kRows; ++i) out_store4(p + i * kFinishWarps * pitch, normalise4(x[i], g));
        }
        const int rest = frames & 3;                                           // 0 for whole clips (pitch % 4 == 0)
        if (rest != 0 && lane < rest)
            for (int row = warp; row < NM; row += kFinishWarps) {
                OutT* q = tile_out + row * pitch + (frames & ~3) + lane;
                out_store(q, clamp_scaled(out_load(q), g));
            }
    } else {
        for (int row = warp; row < NM; row += kFinishWarps)
            for (int i = lane; i < frames; i += 32) {
                OutT* q = tile_out + row * pitch + i;
                out_store(q, clamp_scaled(out_load(q), g));
            }
    }
}

template <typename InT> __device__ __forceinline__ float sample_to_float(InT v);
template <> __device__ __forceinline__ float sample_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float sample_to_float<int16_t>(int16_t v) { return static_cast<float>(v) * (1.0f / 32768.0f); }

struct TileCoord { int64_t clip; int t0; };
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
    __device__ __forceinline__ TileCoord next_of(TileCoord n) const {
        n.clip += step_clips; n.t0 += step_t0;
        if (n.t0 >= frames_per_clip) { n.t0 -= frames_per_clip; ++n.clip; }
        return n;
    }
    __device__ __forceinline__ TileCoord peek_next() const { return next_of(at); }
    __device__ __forceinline__ void advance() { at = peek_next(); }
};

// ---- loaders: one HALF tile of audio into shared memory ----------------------------------------------
// Half h of a tile is 66 rows of 160 samples (one contiguous span of the utterance: frames t0 + 64 h ... + 63 and their
// overlap) at pitch 164 words.
//   TMA mode (fp32, 16-byte aligned rows): ONE tensor copy per half.  The batch is described to the TMA unit as a 4-D
//     tensor {164 samples, 4 quarter rows of 40, rows of 160, utterance} whose innermost extent (164) overlaps the next
//     row on purpose: a {164, 1, 66, 1} box then lands in shared memory as 66 rows at pitch 164 words - the padded,
//     bank-conflict-free layout the fold reads - in a single instruction that completes on the half's `full` mbarrier by
//     byte count (the 4 pad words are never read).  The TMA unit zero-fills the rows it cannot address (before the
//     utterance's first sample, past its last whole row); the loader warp then rewrites the one to three of them that
//     hold real or reflected samples, so a clip's two ends need no other path;
//   PCM mode (int16, half wholly inside the utterance, 16-byte aligned): one bulk copy of the half's 10560 samples into
//     the top of the half's buffer; the half's four fold warps pull them into registers and expand them to float32 rows
//     (x 2^-15, audio.py:62) over the same buffer;
//   cooperative mode (`lengths` cuts, unaligned rows, int16 at a clip's ends): the half's 128 fold threads, which would
//     idle until the rows are there anyway, move them as 16-byte cp.async chunks / converted samples.
// The copies complete on the half's `full` barrier by byte count; whoever issued them is its one arrival.
constexpr int kHalfThreads = 2 * kFoldParts * 32;                // the fold threads of a half tile: 192
constexpr int kChunksPerRow = kHop / 4;                       // 40
constexpr int kHalfChunks = kTcHalfRows * kChunksPerRow;      // 2640
constexpr uint32_t kTmaHalfBytes = kTcHalfBytes;              // the whole box, pad words included
constexpr int kTmaLeadRows = 2, kTmaQuarter = 3;              // tile start = 160 t0 - 200 = 160 (t0 - 2) + 3 * 40
constexpr int kHalfSamples = kTcHalfRows * kHop;              // 10560
constexpr int kPcmStageOffset = kTcHalfBytes - 2 * kHalfSamples;   // 22176: the int16 samples sit at the top of the buffer
static_assert(kPcmStageOffset % 16 == 0 && (2 * kHalfSamples) % 16 == 0, "bulk copy alignment");
enum HalfMode { kModeTma = 0, kModePcm = 1, kModeCoop = 2 };

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src_gmem) : "memory");
}

// The real samples of an utterance: `lengths[clip]` clamped to the row, or the whole row.  The fold warps fetch the raw
// length a few tiles ahead (length_of) - a load from global memory at the top of a tile would sit on their critical path.
__device__ __forceinline__ int32_t length_of(const LogmelArgs& a, int64_t clip) {
    return (a.lengths != nullptr && clip < a.batch) ? __ldg(a.lengths + clip) : 0;
}
__device__ __forceinline__ int64_t valid_from(const LogmelArgs& a, int32_t length) {
    if (a.lengths == nullptr) return a.n_samples;
    const int64_t len = length;
    return len < 0 ? 0 : (len < a.n_samples ? len : a.n_samples);
}

// how half h of the tile at `tc` reaches shared memory (same answer in the loader warp and in the fold warps)
template <typename InT>
__device__ __forceinline__ int half_mode(const LogmelArgs& a, int tma_rows, const TileCoord& tc, int h, int64_t valid) {
    const int64_t s0 = static_cast<int64_t>(tc.t0 + kTcHalfFrames * h) * kHop - kHalfWin;     // first sample of the half
    if constexpr (sizeof(InT) == 4) {
        if (tma_rows <= 0) return kModeCoop;
        if (valid == a.n_samples) return kModeTma;
        // `lengths`: a half of real samples, or one the utterance ends in (copied whole, the rows behind the end are zeroed in
        // shared memory: zero_cut_rows), comes by tensor copy like any other; the zero tail and the rare cut inside the last
        // rows of the memory row are the warps' business (produce_half)
        const int first = tc.t0 - kTmaLeadRows + kTcHalfFrames * h;
        if (first + kTcHalfRows > tma_rows || s0 >= valid) return kModeCoop;
        return kModeTma;
    } else {
        const uintptr_t src = reinterpret_cast<uintptr_t>(static_cast<const InT*>(a.audio) + tc.clip * a.stride_b) + 2u * static_cast<uint64_t>(s0 < 0 ? 0 : s0);
        return (s0 >= 0 && s0 + kHalfSamples <= valid && (src & 15u) == 0) ? kModePcm : kModeCoop;
    }
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
__device__ __forceinline__ bool half_needs_patch(int tma_rows, const TileCoord& tc, int h) {
    const int c2_first = tc.t0 - kTmaLeadRows + kTcHalfFrames * h;
    return c2_first < 0 || c2_first + kTcHalfRows > tma_rows;     // (a superset test; the row test decides)
}
// The half's fold threads rewrite the zero-filled rows of half h that hold real or reflected samples: thread pt takes
// sample pt of every such row (pt < 160).  In two steps, so that the samples travel while the tensor copy is still in
// flight: patch_fetch (before waiting for the copy) -> registers, patch_store (after it has landed) -> shared memory.
// Candidates (tile rows): the rows before tensor row 0 (at most two, in a clip's first tile) and the rows from the first
// tensor row past the end up to the end of the reflected tail (at most three).
constexpr int kPatchCand = kTmaLeadRows + 4;
struct PatchRows { float v[kPatchCand]; int row[kPatchCand]; };
__device__ __forceinline__ void patch_fetch(const LogmelArgs& a, int tma_rows, const TileCoord& tc, int h, int pt, PatchRows& p, int64_t valid) {
    const float* __restrict__ src = static_cast<const float*>(a.audio) + tc.clip * a.stride_b;
    const int c2_first = tc.t0 - kTmaLeadRows;
    const int r_past = tma_rows - c2_first < 0 ? 0 : tma_rows - c2_first;
#pragma unroll
    for (int i = 0; i < kPatchCand; ++i) {
        const int r = i < kTmaLeadRows ? i : r_past + (i - kTmaLeadRows);
        const int rl = r - kTcHalfFrames * h;                                   // row inside the half
        const bool twice = i >= kTmaLeadRows && r < kTmaLeadRows && c2_first < 0 && r + c2_first < 0;   // (no row twice)
        p.row[i] = (!twice && rl >= 0 && rl < kTcHalfRows && tile_row_needs_patch(a, tma_rows, tc, r)) ? rl : -1;
        p.v[i] = 0.f;
        if (p.row[i] >= 0 && pt < kHop) {
            const int64_t pos = static_cast<int64_t>(tc.t0 + r) * kHop - kHalfWin + pt;
            const int64_t idx = reflect_source_index(pos, a.total);
            if (pos < a.total + kHalfWin && idx >= 0 && idx < valid) p.v[i] = __ldg(src + idx);
        }
    }
}
__device__ __forceinline__ void patch_store(const PatchRows& p, float* s_half, int pt) {
#pragma unroll
    for (int i = 0; i < kPatchCand; ++i)
        if (p.row[i] >= 0 && pt < kHop) s_half[p.row[i] * kTcRowPitch + pt] = p.v[i];
}

// `lengths`: the utterance ends inside this half (copied whole by the TMA unit).  The half's fold threads zero what lies
// behind its last real sample: thread pt takes column pt of every row.
__device__ __forceinline__ void zero_cut_rows(float* s_half, int64_t s0, int64_t valid, int pt) {
    if (pt >= kHop) return;
    int64_t pos = s0 + pt;
#pragma unroll 1
    for (int r = 0; r < kTcHalfRows; ++r, pos += kHop)
        if (pos >= valid) s_half[r * kTcRowPitch + pt] = 0.f;
}

__device__ __forceinline__ void tma_load_half(const CUtensorMap* map, const TileCoord& tc, int h, float* s_half, uint64_t* full) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full)), "r"(kTmaHalfBytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(s_half)), "l"(map), "r"(0), "r"(kTmaQuarter), "r"(tc.t0 - kTmaLeadRows + kTcHalfFrames * h), "r"(static_cast<int>(tc.clip)),
                   "r"(smem_u32(full)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_half(const CUtensorMap* map, const TileCoord& tc, int h) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(map), "r"(0), "r"(kTmaQuarter), "r"(tc.t0 - kTmaLeadRows + kTcHalfFrames * h), "r"(static_cast<int>(tc.clip)) : "memory");
}
// int16: the half's samples as one bulk copy that completes on the half's `full` barrier
__device__ __forceinline__ void pcm_load_half(const LogmelArgs& a, const TileCoord& tc, int h, float* s_half, uint64_t* done) {
    const int64_t s0 = static_cast<int64_t>(tc.t0 + kTcHalfFrames * h) * kHop - kHalfWin;
    const int16_t* src = static_cast<const int16_t*>(a.audio) + tc.clip * a.stride_b + s0;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(done)), "r"(2 * kHalfSamples) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(s_half) + kPcmStageOffset), "l"(src), "r"(2 * kHalfSamples), "r"(smem_u32(done)) : "memory");
}
// Asks L2 for a later half's samples (non-tensor form: interior, aligned part only): the CTAs of a wave load in lock-step,
// so without it every staging phase waits on an HBM burst while HBM idles the rest of the time.
template <typename InT>
__device__ __forceinline__ void prefetch_half_l2(const LogmelArgs& a, const TileCoord& tc, int h) {
    const InT* row = static_cast<const InT*>(a.audio) + tc.clip * a.stride_b;
    const int64_t s0 = static_cast<int64_t>(tc.t0 + kTcHalfFrames * h) * kHop - kHalfWin;
    int64_t first = s0 < 0 ? 0 : s0, last = s0 + kHalfSamples;
    if (last > a.n_samples) last = a.n_samples;
    const uintptr_t begin = (reinterpret_cast<uintptr_t>(row + first) + 15u) & ~static_cast<uintptr_t>(15u);
    const uintptr_t end = reinterpret_cast<uintptr_t>(row + last) & ~static_cast<uintptr_t>(15u);
    if (last > first && end > begin)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(begin), "r"(static_cast<uint32_t>(end - begin)) : "memory");
}

// PCM mode, the half's fold threads: int16 samples at the top of the buffer -> float32 rows over the whole buffer.
// Every thread pulls its chunks of 8 samples into registers; only when all of them have (named barrier) may the rows be
// written, because they overwrite the staged samples.
__device__ __forceinline__ void expand_pcm_half(float* s_half, int pt, int bar_id) {
    constexpr int kChunks = kHalfSamples / 8;                                 // 1320 chunks of 8 samples, 20 per row
    constexpr int kPerThread = (kChunks + kHalfThreads - 1) / kHalfThreads;  // 7
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(s_half) + kPcmStageOffset);
    uint4 raw[kPerThread];
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const int c = pt + i * kHalfThreads;
        raw[i] = c < kChunks ? src[c] : make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kHalfThreads) : "memory");
#pragma unroll
    for (int i = 0; i < kPerThread; ++i) {
        const int c = pt + i * kHalfThreads;
        if (c < kChunks) {
            const int rr = c / (kHop / 8), col = (c - rr * (kHop / 8)) * 8;
            const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
            float v[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // int16 s -> s / 32768 without the conversion unit (I2F is 8 issue cycles a warp): with the sign bit flipped
                // the sample is u = s + 32768 in 0..65535; the float whose bits are 0x4B00uuuu is 2^23 + u exactly, and
                // (2^23 + u) 2^-15 - 257 = s / 32768 - exact, one fused multiply-add (audio.py:62's arithmetic)
                const uint32_t biased = w[j] ^ 0x80008000u;
                v[2 * j] = fmaf(__uint_as_float(__byte_perm(biased, 0x4b000000u, 0x7610)), 1.0f / 32768.0f, -257.0f);
                v[2 * j + 1] = fmaf(__uint_as_float(__byte_perm(biased, 0x4b000000u, 0x7632)), 1.0f / 32768.0f, -257.0f);
            }
            float4* dst = reinterpret_cast<float4*>(s_half + rr * kTcRowPitch + col);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kHalfThreads) : "memory");   // every row is written before anybody folds
}

// cooperative mode, the half's fold threads
template <typename InT>
// returns true if nothing was staged because the whole half is zeros (the caller then skips the arithmetic as well)
__device__ __forceinline__ bool produce_half(const LogmelArgs& a, const TileCoord& tc, int h, float* s_half, int pt, int bar_id, int64_t valid) {
    const InT* __restrict__ row = static_cast<const InT*>(a.audio) + tc.clip * a.stride_b;
    const int64_t s0 = static_cast<int64_t>(tc.t0 + kTcHalfFrames * h) * kHop - kHalfWin;
    const bool aligned = sizeof(InT) == 4 && (reinterpret_cast<uintptr_t>(row) & 15u) == 0;
    // the whole half lies in the zero tail (`lengths`, right padding) and no reflection reaches a real sample: nothing to stage,
    // nothing to read (the same answer in all of the half's threads)
    if (s0 >= valid && valid + kHalfWin < a.total) return true;
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kHalfThreads) : "memory");   // all four warps are done reading the previous tile's rows
    // chunk c = 40 r + k covers samples s0 + 4c .. + 3 and lands at word 164 r + 4 k
    int r = pt / kChunksPerRow, k = pt - r * kChunksPerRow;
    for (int c = pt; c < kHalfChunks; c += kHalfThreads) {
        const int64_t pos = s0 + 4 * static_cast<int64_t>(c);
        float* dst = s_half + r * kTcRowPitch + 4 * k;
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
        r += kHalfThreads / kChunksPerRow; k += kHalfThreads % kChunksPerRow;   // 192 = 4 x 40 + 32
        if (k >= kChunksPerRow) { k -= kChunksPerRow; ++r; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kHalfThreads) : "memory");   // every row is written before anybody folds
    return false;
}

// ---- fold warps ----------------------------------------------------------------------------------------
__constant__ TcFoldTables c_fold = tc_make_fold_tables();

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// One compact loop for both sweeps (see tc_core.cuh), run over the chunks [j0, j1): the two fold warps of a lane
// quadrant split every sweep between them (chunks 0..6 and 7..12), so a sweep's operand is complete in half the time
// and the tensor cores start on it that much earlier.  Chunk 12 is the leftover chunk with its own store pattern.
// `sweep`, `scale` are warp-uniform, so the table rows arrive through the uniform datapath.  fr: shared address of the
// frame's first sample.  TRACK: also returns the largest |sample| the chunks read (the sweeps are bound by the fp16
// conversions, which have their own pipe - the extra maxima ride in issue slots that are free anyway).
template <bool TRACK>
__device__ __forceinline__ float sweep_store(int sweep, uint32_t scale, int j0, int j1, uint32_t fr, uint32_t lane_addr) {
    const TcFoldOffsets* __restrict__ off = c_fold.off[sweep];
    const TcFoldWeights* __restrict__ wts = c_fold.w[scale][sweep];
    const float sign = c_fold.sign[sweep];
    float head[2];
    head[0] = lds32(fr + off[j0].head[0]);
    head[1] = lds32(fr + off[j0].head[1]);
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t c = lane_addr + (sweep == 0 ? tc_hi_col(0) : tc_hi_col(2)) + 4 * j0;   // hi block of the sweep's first unit
#pragma unroll 1
    for (int j = j0; j < j1; ++j, c += 4) {
        const TcFoldOffsets& o = off[j];
        float up[2][8], down[2][8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 u0 = lds128(fr + o.group[4 * q]), u1 = lds128(fr + o.group[4 * q + 1]);
            const float4 d0 = lds128(fr + o.group[4 * q + 2]), d1 = lds128(fr + o.group[4 * q + 3]);
            up[q][0] = u0.x; up[q][1] = u0.y; up[q][2] = u0.z; up[q][3] = u0.w;
            up[q][4] = u1.x; up[q][5] = u1.y; up[q][6] = u1.z; up[q][7] = u1.w;
            down[q][0] = head[q];
            down[q][1] = d0.w; down[q][2] = d0.z; down[q][3] = d0.y; down[q][4] = d0.x;
            down[q][5] = d1.w; down[q][6] = d1.z; down[q][7] = d1.y;
            head[q] = d1.x;
            if constexpr (TRACK) {
                // (every sample of the frame passes through one E chunk; the heads are counted by the chunk that loads them)
                m[2 * q] = fmaxf(fmaxf(m[2 * q], fabsf(u0.x)), fabsf(u0.y));
                m[2 * q] = fmaxf(fmaxf(m[2 * q], fabsf(u0.z)), fabsf(u0.w));
                m[2 * q] = fmaxf(fmaxf(m[2 * q], fabsf(u1.x)), fabsf(u1.y));
                m[2 * q] = fmaxf(fmaxf(m[2 * q], fabsf(u1.z)), fabsf(u1.w));
                m[2 * q + 1] = fmaxf(fmaxf(m[2 * q + 1], fabsf(d0.x)), fabsf(d0.y));
                m[2 * q + 1] = fmaxf(fmaxf(m[2 * q + 1], fabsf(d0.z)), fabsf(d0.w));
                m[2 * q + 1] = fmaxf(fmaxf(m[2 * q + 1], fabsf(d1.x)), fabsf(d1.y));
                m[2 * q + 1] = fmaxf(fmaxf(m[2 * q + 1], fabsf(d1.z)), fabsf(d1.w));
            }
        }
        uint32_t hf[4], lf[4], hs[4], ls[4];
        tc_chunk_math(up, down, wts[j], sign, hf, lf, hs, ls);
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
    return fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
}
static_assert(tc_lo_col(0) - tc_hi_col(0) == 48 && tc_hi_col(1) - tc_hi_col(0) == 96 && tc_hi_col(3) - tc_hi_col(2) == 96 &&
              tc_left_col(1) - tc_left_col(0) == 6 && tc_left_col(3) - tc_left_col(2) == 6, "column arithmetic of sweep_store");

// ---- digital silence ---------------------------------------------------------------------------------
// A lane quadrant whose samples are ALL zero (the zero tail pad_or_trim appends, the 30 s of padding transcribe asks for,
// `lengths`) needs no arithmetic: its operand is zeros.  The test reads the quadrant's 34 rows once, shared between its
// three warps (lane = row, a warp takes every third 16-byte chunk of it; rows 32 and 33 chunk by chunk) - but only after a
// look at four samples per lane has found nothing but zeros, so a tile of sound pays one load and a vote.  -0.0 counts as
// zero (its power is 0 as well); NaN / Inf do not.  `probe`: the quadrant's three exchange words.
__device__ __forceinline__ bool quadrant_is_silent(uint32_t quad_rows, uint32_t fr, int part, int lane, int quad, volatile uint32_t* probe) {
    const float4 q = lds128(fr);
    const uint32_t first = (__float_as_uint(q.x) | __float_as_uint(q.y) | __float_as_uint(q.z) | __float_as_uint(q.w)) & 0x7fffffffu;
    if (__any_sync(0xffffffffu, first != 0u)) return false;      // (the three warps see the same samples: the same branch)
    uint32_t bits = 0u;
    const uint32_t row = quad_rows + lane * (kTcRowPitch * 4);
#pragma unroll 2
    for (int c = part; c < kHop / 4; c += kFoldParts) {
        const float4 v = lds128(row + 16 * c);
        bits |= __float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w);
    }
    {   // rows 32 and 33: 80 chunks over the quadrant's 96 threads
        const int c = part * 32 + lane;
        if (c < 2 * (kHop / 4)) {
            const float4 v = lds128(quad_rows + (32 + c / (kHop / 4)) * (kTcRowPitch * 4) + 16 * (c % (kHop / 4)));
            bits |= __float_as_uint(v.x) | __float_as_uint(v.y) | __float_as_uint(v.z) | __float_as_uint(v.w);
        }
    }
    const bool sound = __any_sync(0xffffffffu, (bits & 0x7fffffffu) != 0u);
    if (lane == 0) probe[part] = sound ? 1u : 0u;
    asm volatile("bar.sync %0, %1;" ::"r"(5 + quad), "n"(32 * kFoldParts) : "memory");
    const uint32_t any_sound = probe[0] | probe[1] | probe[2];
    asm volatile("bar.sync %0, %1;" ::"r"(5 + quad), "n"(32 * kFoldParts) : "memory");   // (read before the next tile's probe is written)
    return any_sound == 0u;
}
// the chunks [j0, j1) of a sweep's two units as zeros (same columns as sweep_store)
__device__ __forceinline__ void sweep_store_zeros(int sweep, int j0, int j1, uint32_t lane_addr) {
    const uint32_t z[4] = {0u, 0u, 0u, 0u};
    uint32_t c = lane_addr + (sweep == 0 ? tc_hi_col(0) : tc_hi_col(2)) + 4 * j0;
#pragma unroll 1
    for (int j = j0; j < j1; ++j, c += 4) {
        if (j < 2 * kTcMainSteps) {
            tmem_st4(c, z); tmem_st4(c + 48, z);
            tmem_st4(c + 96, z); tmem_st4(c + 144, z);
        } else {
            const uint32_t b1 = lane_addr + (sweep == 0 ? tc_left_col(0) : tc_left_col(2));
            tmem_st4(b1, z); tmem_st4(b1 + 4, z); tmem_st4(b1 + 8, z);     // 12 columns: both units' [hi x 3 | lo x 3]
        }
    }
}

// ---- epilogue ---------------------------------------------------------------------------------------
// ---- finish a tile: log10 clamp, coalesced row stores (lane = frame), extremes; leaves acc zeroed ----
// `at`: the tile; k: its index among the CTA's tiles; y_offset / silent: what the folds said about it (see epilogue_role)
template <int NM, typename OutT, int PITCH>
__device__ __forceinline__ void epilogue_finish_tile(const LogmelArgs& a, long long* trace, const int trace_first_arg, float (&acc)[NM],
                                                     const TileCoord& at, int64_t k, float y_offset, bool silent, int quad, int lane) {
    [[maybe_unused]] const int trace_first = trace_first_arg & 0xff;
    [[maybe_unused]] const int ti = static_cast<int>(k);
    const int f = quad * 32 + lane, t = at.t0 + f;
    const bool live = t < a.n_frames && !silent && !TC_DEBUG_FLAG(0x1000);   // (0x1000, measurement only: no finish)
    const int64_t pitch = a.n_frames;
    OutT* const out = reinterpret_cast<OutT*>(a.out) + at.clip * NM * pitch + t;
    // y = (log10(max(P, 1e-10)) + 4) / 4 of the mel power P = acc 2^-2k, as ONE fused multiply-add behind the MUFU:
    //   y = log2(acc) (log10(2) / 4) + (1 - 2k log10(2) / 4),  then max(y, -1.5)   [-1.5 = (log10(1e-10) + 4) / 4]
    // (the data scale leaves as part of the addend - exact in real arithmetic, a power of two; the clamp at 1e-10
    // moves behind the logarithm, where log2(0) = -inf comes out as exactly -1.5, like the reference's clamp).
    // Only the clamp at max - 8 is left for the finish kernel - which skips the tile when its smallest value is not
    // below max - 8 (tracked here as well, in y; converted to log10 units once per lane).
    float mx = __uint_as_float(0xff800000u), mn = __uint_as_float(0x7f800000u);
    if (live) {
        constexpr float kLog10Of2Quarter = 0.30102999566398120f * 0.25f;
        const float2 mul2 = make_float2(kLog10Of2Quarter, kLog10Of2Quarter), add2 = make_float2(y_offset, y_offset);
#pragma unroll
        for (int m = 0; m < NM; m += 2) {
            float2 l2;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2.x) : "f"(acc[m]));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2.y) : "f"(acc[m + 1]));
            float2 y = __ffma2_rn(l2, mul2, add2);
            y.x = max_nan(y.x, -1.5f);
            y.y = max_nan(y.y, -1.5f);
            if constexpr (PITCH > 0) {
                out_store(out + PITCH * m, y.x);
                out_store(out + PITCH * (m + 1), y.y);
            } else {
                const uint32_t pitch32 = static_cast<uint32_t>(a.n_frames);
                out_store(out + static_cast<uint64_t>(pitch32) * static_cast<uint32_t>(m), y.x);
                out_store(out + static_cast<uint64_t>(pitch32) * static_cast<uint32_t>(m + 1), y.y);
            }
            mx = max_nan(mx, max_nan(y.x, y.y));
            mn = fminf(mn, fminf(y.x, y.y));
        }
        mx = fmaf(mx, 4.0f, -4.0f);     // back to log10 units: what the keys hold
        mn = fmaf(mn, 4.0f, -4.0f);
    }
#pragma unroll
    for (int i = 0; i < NM; ++i) acc[i] = 0.f;
    if (quad == 0) TC_TRACE(4, ti, 15);
    uint32_t key = live ? max_key_encode(mx) : 0u;
    key = __reduce_max_sync(0xffffffffu, key);
    uint32_t inv = live ? ~max_key_encode(mn) : 0u;
    inv = __reduce_max_sync(0xffffffffu, inv);
    if (lane == 0 && key != 0u) {       // (key 0: nothing stored - frames past the end, or a silent tile)
        atomicMax(a.max_keys + (a.global_max ? 0 : at.clip), key);
        atomicMax(a.min_keys + at.clip, inv);
        if (a.tile_keys != nullptr) {   // the tile's own extremes: lets the finish kernel skip or fill whole tiles
            uint32_t* tk = a.tile_keys + 2 * (static_cast<int64_t>(blockIdx.x) + k * gridDim.x);
            atomicMax(tk, key);
            atomicMax(tk + 1, inv);
        }
    }
    if (quad == 0) TC_TRACE(4, ti, 12);
}

template <int NM, typename OutT, int PITCH>
__device__ __forceinline__ void epilogue_role(const LogmelArgs& a, long long* trace, const int trace_first_arg, TcBarriers* bars,
                                              const TcTileInfo* info, TcAbort ab,
                                              uint32_t tmem, int quad, int lane, int64_t total_tiles, int tiles_per_clip) {
    using L = TcEpilogueLayout<NM>;
    [[maybe_unused]] const int trace_first = trace_first_arg & 0xff;
    float acc[NM];
#pragma unroll
    for (int i = 0; i < NM; ++i) acc[i] = 0.f;
    const uint32_t d_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16) + kTcDCol;
    uint32_t d_parity = 0;
    const int64_t my_tiles = static_cast<int64_t>(blockIdx.x) < total_tiles ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    TileCursor cursor(tiles_per_clip);
    // Unit order on the tensor cores: 0, 1 (E sweep), 2, 3 (O sweep).  The accumulator is single: the tensor cores start a
    // unit only when the previous one has been pulled into registers.  So a tile is FINISHED (log10, stores, extremes)
    //   - 80 mels (the unit's 104 columns fit the registers beside the 80 sums): after unit 0 of the NEXT tile has been
    //     pulled - the accumulator is free again while the stores go out, unit 1's MMAs (and with them the release of the
    //     E operand to the folds) do not wait for the finish;
    //   - 128 mels (the columns come in two pieces): right after the tile's last unit.
    constexpr bool kDeferFinish = L::pieces == 1;
    TileCoord done{0, 0};
    float y_done = 1.0f;
    bool silent_done = false;
#pragma unroll 1
    for (int64_t k = 0; k < my_tiles; ++k) {
        const int ti = static_cast<int>(k);
        const TileCoord cur = cursor.at;
        cursor.advance();
        float y_offset = 1.0f;
        bool silent = false;
#pragma unroll 1
        for (int u = 0; u < kTcUnits; ++u) {
            if (quad == 0) TC_TRACE(4, ti, 3 * u);
            mbar_wait(&bars->d_full, d_parity, ab);
            d_parity ^= 1u;
            tc_fence_after();
            if (u == 0) {
                // the quadrant's scale step was written before the folds released the E operand, i.e. before this unit's MMAs
                y_offset = c_fold.y_offset[*reinterpret_cast<const volatile uint32_t*>(&info->scale[k & 1][quad])];
                // a tile whose samples are ALL zero (zero padding) is not stored: the finish kernel fills it (it needs per-tile keys for that)
                const volatile uint32_t* z = info->silent[k & 1];
                silent = a.tile_keys != nullptr && (z[0] & z[1] & z[2] & z[3]) != 0u;
            }
            // Re and Im of a bin take the same weights: units 0 / 2 (even bins) share one body, units 1 / 3 (odd bins) the other.
            // The columns come in L::pieces pieces; the accumulator is released once the last piece is in registers.
            // (a tile of digital silence: the tensor cores were not asked anything - see the issue warp - and the sums stay zero)
            if constexpr (kDeferFinish) {
                float d[L::piece_cols];
                if (!silent) {
                    tmem_ld_cols<L::piece_cols>(d_addr, d);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->d_empty);   // the next unit may overwrite the accumulator
                if (quad == 0) TC_TRACE(4, ti, 3 * u + 1);
                if (u == 0 && k > 0) epilogue_finish_tile<NM, OutT, PITCH>(a, trace, trace_first_arg, acc, done, k - 1, y_done, silent_done, quad, lane);
                if (silent || TC_DEBUG_FLAG(0x2000)) continue;   // (0x2000, measurement only: pull and release, no mel sums)
                if ((u & 1) == 0) tc_epilogue_unit<NM, 0, 0, L::piece_cols>(d, acc);
                else tc_epilogue_unit<NM, 1, 0, L::piece_cols>(d, acc);
            } else {
                if (silent) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->d_empty);
                    continue;
                }
#pragma unroll
                for (int piece = 0; piece < L::pieces; ++piece) {
                    float d[L::piece_cols];
                    tmem_ld_cols<L::piece_cols>(d_addr + piece * L::piece_cols, d);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (piece == L::pieces - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->d_empty);   // the next unit may overwrite the accumulator
                        if (quad == 0) TC_TRACE(4, ti, 3 * u + 1);
                    }
                    if (TC_DEBUG_FLAG(0x2000)) continue;   // (measurement only: pull and release, no mel sums)
                    if ((u & 1) == 0) {
                        if (piece == 0) tc_epilogue_unit<NM, 0, 0, L::piece_cols>(d, acc);
                        else tc_epilogue_unit<NM, 0, (L::pieces - 1) * L::piece_cols, L::piece_cols>(d, acc);
                    } else {
                        if (piece == 0) tc_epilogue_unit<NM, 1, 0, L::piece_cols>(d, acc);
                        else tc_epilogue_unit<NM, 1, (L::pieces - 1) * L::piece_cols, L::piece_cols>(d, acc);
                    }
                }
            }
            if (quad == 0) TC_TRACE(4, ti, 3 * u + 2);
            if constexpr (!kDeferFinish) {
                if (u == kTcUnits - 1) epilogue_finish_tile<NM, OutT, PITCH>(a, trace, trace_first_arg, acc, cur, k, y_offset, silent, quad, lane);
            }
        }
        done = cur;
        y_done = y_offset;
        silent_done = silent;
    }
    if constexpr (kDeferFinish) {
        if (my_tiles > 0) epilogue_finish_tile<NM, OutT, PITCH>(a, trace, trace_first_arg, acc, done, my_tiles - 1, y_done, silent_done, quad, lane);
    }
}

// PITCH: the output's row pitch n_frames as a compile-time constant (3000: the 30 s clips of every consumer - the row
// stores then take immediate offsets instead of a 64-bit multiply-add each), or 0: any
template <typename InT, int NM, typename OutT, int PITCH>
__global__ void __launch_bounds__(kTcThreads, 1)
logmel_tc_kernel(const __grid_constant__ LogmelArgs a, const __grid_constant__ CUtensorMap audio_map, const int tma_rows,
                 const unsigned char* __restrict__ operands, long long* __restrict__ trace, const int trace_first_arg) {
    const int trace_first = trace_first_arg & 0xff;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_audio = reinterpret_cast<float*>(smem_raw + kSmemAudio);
    __shared__ __align__(8) TcBarriers bars;
    __shared__ TcTileInfo info;
    __shared__ uint32_t s_tmem, s_abort;
    const TcAbort ab{&s_abort};

    // the warp index through a shuffle: the compiler then knows it is warp-uniform and keeps everything derived from
    // it (role, TMEM lane quadrant, column addresses) in uniform registers
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31, quad = warp & 3;
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    const int64_t total_tiles = a.batch * tiles_per_clip;

    // ---- one-time setup: tensor memory, barriers, constant matrices -> shared memory ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&bars.audio_full[0], 1); mbar_init(&bars.audio_full[1], 1);
        info.released[0][0] = info.released[0][1] = info.released[1][0] = info.released[1][1] = 0;
        mbar_init(&bars.a_full[0], 4 * kFoldParts); mbar_init(&bars.a_full[1], 4 * kFoldParts);
        mbar_init(&bars.a_empty[0], 1); mbar_init(&bars.a_empty[1], 1);
        mbar_init(&bars.d_full, 1);
        mbar_init(&bars.d_empty, 4);
        s_abort = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the first tile's two halves are on their way while the CTA fetches its matrices (nothing has touched the audio
        // buffers yet; their pad words are never read, so they need no initialisation)
        if (static_cast<int64_t>(blockIdx.x) < total_tiles) {
            const TileCursor first(tiles_per_clip);
            const int64_t valid = valid_from(a, length_of(a, first.at.clip));
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                float* s_half = s_audio + h * (kTcHalfStride / 4);
                const int mode = half_mode<InT>(a, tma_rows, first.at, h, valid);
                if (mode == kModeTma) tma_load_half(&audio_map, first.at, h, s_half, &bars.audio_full[h]);
                else if (mode == kModePcm) pcm_load_half(a, first.at, h, s_half, &bars.audio_full[h]);
            }
        }
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(operands);
        uint4* dst = reinterpret_cast<uint4*>(smem_raw + kSmemOperands);
        for (int i = tid; i < kTcOperandBytes / 16; i += kTcThreads) dst[i] = __ldg(src + i);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core / TMA (async proxy) accesses
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(quad * 32) << 16);   // this warp's TMEM lane quadrant

    if (warp < 4) {
        // zero every column once (every operand column is rewritten each tile; this only keeps idle lanes finite)
        for (int c = 0; c < 512; c += 4) { const uint32_t z[4] = {0u, 0u, 0u, 0u}; tmem_st4(lane_addr + c, z); }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // register budget per warpgroup: the CTA is launched with 5 x 96; the warpgroups trade inside that total
    // (a setmaxnreg.inc can only take what another warpgroup released): folds 3 x 72, epilogue 232, rest 32
    if (warp < kWarpEpi0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // ===== fold warps: two per lane quadrant; both do half of the E sweep, then half of the O sweep =====
        const int part = warp >> 2;                                           // 0..2
        const int half = quad >> 1;
        float* s_half = s_audio + half * (kTcHalfStride / 4);
        const int half_thread = (quad & 1) * 32 + lane + 64 * part;          // 0..191 among the half's fold threads
        const uint32_t quad_rows = smem_u32(s_half) + (quad & 1) * 32 * (kTcRowPitch * 4);
        const uint32_t fr = quad_rows + lane * (kTcRowPitch * 4);
        const int half_bar = 9 + half;                                        // named barrier of the half's four warps
        // the copy of half h of the tile at `tp` (one thread): tensor copy, bulk copy of PCM samples, or nothing
        auto issue_half = [&](const TileCoord& tp, int64_t valid) {
            const int mode = half_mode<InT>(a, tma_rows, tp, half, valid);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the rows were read through the generic proxy
            if (mode == kModeTma) tma_load_half(&audio_map, tp, half, s_half, &bars.audio_full[half]);
            else if (mode == kModePcm) pcm_load_half(a, tp, half, s_half, &bars.audio_full[half]);
        };
        auto prefetch_half = [&](const TileCoord& tp, int64_t valid) {
            if (TC_DEBUG_FLAG(0x400)) return;
            if (half_mode<InT>(a, tma_rows, tp, half, valid) == kModeTma && !TC_DEBUG_FLAG(0x800)) tma_prefetch_half(&audio_map, tp, half);
            else prefetch_half_l2<InT>(a, tp, half);
        };
        uint32_t parity = 0, full_parity = 0;
        uint32_t last_scale = 5u;                                             // (2^12: right for samples of order 1)
        int ti = 0;
        TileCursor cursor(tiles_per_clip);
        // the utterance lengths of this tile, the next one and the one after (the tiles whose copy / L2 prefetch this one issues)
        int32_t len_cur = length_of(a, cursor.at.clip), len_next = length_of(a, cursor.peek_next().clip),
                len_after = length_of(a, cursor.next_of(cursor.peek_next()).clip);
        // (the first tile's copy was issued at the top of the kernel; one thread per half asks L2 for the second tile)
        if (part == 0 && (quad & 1) == 0 && lane == 0 && static_cast<int64_t>(blockIdx.x) + gridDim.x < total_tiles)   // warps 0 and 2
            prefetch_half(cursor.peek_next(), valid_from(a, len_next));
        for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti, cursor.advance()) {
            if (quad == 0) TC_TRACE(part == 2 ? 5 : 1 + part, ti, 0);
            const TileCoord tcl = cursor.at;
            // (in flight while this tile is folded)
            const int32_t len_far = length_of(a, cursor.next_of(cursor.next_of(cursor.peek_next())).clip);
            const int64_t valid_cur = valid_from(a, len_cur);
            const int mode = half_mode<InT>(a, tma_rows, tcl, half, valid_cur);
            bool known_zero = false;
            if (mode == kModeTma) {
                if constexpr (sizeof(InT) == 4) {
                    const int64_t s0 = static_cast<int64_t>(tcl.t0 + kTcHalfFrames * half) * kHop - kHalfWin;
                    const bool cut = s0 + kHalfSamples > valid_cur && valid_cur < a.n_samples;   // (`lengths`: the utterance ends in this half)
                    if (half_needs_patch(tma_rows, tcl, half)) {
                        // a clip's first or last rows: fetch what the zero-filled rows should hold while the copy is in flight
                        PatchRows rows;
                        patch_fetch(a, tma_rows, tcl, half, half_thread, rows, valid_cur);
                        mbar_wait(&bars.audio_full[half], full_parity, ab);
                        if (cut) zero_cut_rows(s_half, s0, valid_cur, half_thread);
                        patch_store(rows, s_half, half_thread);
                        asm volatile("bar.sync %0, %1;" ::"r"(half_bar), "n"(kHalfThreads) : "memory");
                    } else {
                        mbar_wait(&bars.audio_full[half], full_parity, ab);
                        if (cut) {
                            zero_cut_rows(s_half, s0, valid_cur, half_thread);
                            asm volatile("bar.sync %0, %1;" ::"r"(half_bar), "n"(kHalfThreads) : "memory");
                        }
                    }
                    full_parity ^= 1u;
                }
            } else if (mode == kModePcm) {
                mbar_wait(&bars.audio_full[half], full_parity, ab);   // the bulk copy of the int16 samples has landed
                full_parity ^= 1u;
                expand_pcm_half(s_half, half_thread, half_bar);
            } else {
                known_zero = produce_half<InT>(a, tcl, half, s_half, half_thread, half_bar, valid_cur);
            }
            if (quad == 0) TC_TRACE(part == 2 ? 5 : 1 + part, ti, 1);
            // (the quadrant's three warps read the same rows: the same answer, no exchange)
            const bool quad_zero = (known_zero || quadrant_is_silent(quad_rows, fr, part, lane, quad, info.probe[quad])) && !TC_DEBUG_FLAG(0x10000);
            if (quad == 0) TC_TRACE(part == 2 ? 5 : 1 + part, ti, 8);
            const int release_sweep = quad_zero ? 0 : 1;   // the pass after which this warp no longer reads the half's rows
            uint32_t scale = last_scale;
#pragma unroll 1
            for (int sweep = 0; sweep < 2; ++sweep) {
                mbar_wait(&bars.a_empty[sweep], parity ^ 1u, ab);   // the tensor cores are done with the previous tile's operand
                if (quad == 0) TC_TRACE(part == 2 ? 5 : 1 + part, ti, 2 + 3 * sweep);
                tc_fence_after();
                // the quadrant's three warps take 5 + 4 + 4 chunks of a sweep; the O sweep hands the 5 to the other end, so
                // that the warps come out at 9 / 8 / 9 chunks per tile
                const int slot = sweep == 0 ? part : kFoldParts - 1 - part;
                const int j0 = slot == 0 ? 0 : 1 + 4 * slot, j1 = TC_DEBUG_FLAG(0x8000) ? j0 + 1 : 5 + 4 * slot;   // (0x8000, measurement only: one chunk per warp and sweep)
                if (quad_zero) {
                    // nothing but zeros: the operand is zeros, the scale step stays, the tile info says so
                    sweep_store_zeros(sweep, j0, j1, lane_addr);
                    if (sweep == 0 && part == 0 && lane == 0) {
                        *reinterpret_cast<volatile uint32_t*>(&info.scale[ti & 1][quad]) = scale;
                        *reinterpret_cast<volatile uint32_t*>(&info.silent[ti & 1][quad]) = 1u;
                    }
                } else if (sweep == 0) {
                    // the E sweep with the previous tile's scale step, tracking the largest |sample|; the quadrant's two warps
                    // then agree on the step this tile calls for and repeat their chunks if it is another one
                    const float m = sweep_store<true>(0, scale, j0, j1, fr, lane_addr);
                    const uint32_t mine = __reduce_max_sync(0xffffffffu, __float_as_uint(m));   // non-negative floats order like their bits
                    if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&info.quad_max[quad][part]) = mine;
                    asm volatile("bar.sync %0, %1;" ::"r"(5 + quad), "n"(32 * kFoldParts) : "memory");
                    // (the partners read this tile's values before any of them can write the next tile's: the O sweep is between)
                    const volatile uint32_t* qm = info.quad_max[quad];
                    uint32_t other = qm[0] > qm[1] ? qm[0] : qm[1];
                    other = other > qm[2] ? other : qm[2];
                    const uint32_t want = static_cast<uint32_t>(tc_scale_index(other));
                    if (want != scale && !TC_DEBUG_FLAG(0x4000)) {
                        scale = want;
                        sweep_store<false>(0, scale, j0, j1, fr, lane_addr);
                    }
                    last_scale = scale;
                    if (part == 0 && lane == 0) {
                        *reinterpret_cast<volatile uint32_t*>(&info.scale[ti & 1][quad]) = scale;
                        *reinterpret_cast<volatile uint32_t*>(&info.silent[ti & 1][quad]) = other == 0u ? 1u : 0u;
                    }
                } else {
                    sweep_store<false>(1, scale, j0, j1, fr, lane_addr);
                }
                if (quad == 0) TC_TRACE(part == 2 ? 5 : 1 + part, ti, 3 + 3 * sweep);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bars.a_full[sweep]);
                    if (sweep == release_sweep) {
                        // this warp has read the half's rows for the last time (at once, if there is nothing but zeros in them);
                        // the last of the half's warps to say so brings the next tile's rows (and asks L2 for the ones after)
                        // (one count per tile parity: a warp whose rows are zeros for two tiles in a row may say so for the next
                        // tile while another one still reads this tile's rows)
                        __threadfence_block();
                        if (atomicAdd(&info.released[half][ti & 1], 1u) == 2u * kFoldParts - 1u) {
                            *reinterpret_cast<volatile uint32_t*>(&info.released[half][ti & 1]) = 0u;
                            if (tile + gridDim.x < total_tiles) {
                                const TileCoord next = cursor.peek_next();
                                issue_half(next, valid_from(a, len_next));
                                if (tile + 2 * static_cast<int64_t>(gridDim.x) < total_tiles) prefetch_half(cursor.next_of(next), valid_from(a, len_after));
                            }
                        }
                    }
                }
                if (quad == 0) TC_TRACE(part == 2 ? 5 : 1 + part, ti, 4 + 3 * sweep);
            }
            parity ^= 1u;
            len_cur = len_next;
            len_next = len_after;
            len_after = len_far;
        }
    } else if (warp < kWarpMma) {
        // ===== epilogue warps =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        epilogue_role<NM, OutT, PITCH>(a, trace, trace_first_arg, &bars, &info, ab, tmem, quad, lane, total_tiles, tiles_per_clip);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
        if (warp == kWarpMma) {
            // ===== tensor-core issue: the whole warp walks the loop, one elected lane issues =====
            // One compact loop over the units (per-unit columns and matrix offsets from a constant table): the issue
            // code stays small so it does not evict the fold and epilogue code from the instruction caches.
            const uint32_t desc0 = operand_desc_lo(smem_u32(smem_raw + kSmemOperands)), d_tmem = tmem + kTcDCol;
            const uint32_t a_full0 = smem_u32(&bars.a_full[0]), a_empty0 = smem_u32(&bars.a_empty[0]);
            constexpr uint32_t kStep = (2 * kTcStripBytes) >> 4;   // K step s = strips 2s, 2s+1 = slots 16s..16s+15
            uint32_t a_parity = 0, d_parity = 1;   // d_empty: the first wait passes (accumulator starts free)
            int ti = 0;
            for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
                bool skip_tile = false;
#pragma unroll 1
                for (int u = 0; u < kTcUnits; ++u) {
                    const uint32_t sweep_bar = static_cast<uint32_t>(u & 2) << 2;   // byte offset of the sweep's barrier (0 or 8)
                    if ((u & 1) == 0) mbar_wait_addr(a_full0 + sweep_bar, a_parity, ab);
                    TC_TRACE(3, ti, 3 * u);
                    if (u == 0) {
                        // a tile of digital silence that the finish kernel will fill (the folds said so before they released
                        // the E operand): nothing to multiply - the hand-overs below still happen, the commits arrive at once
                        const volatile uint32_t* z = info.silent[ti & 1];
                        skip_tile = a.tile_keys != nullptr && (z[0] & z[1] & z[2] & z[3]) != 0u;
                    }
                    mbar_wait(&bars.d_empty, d_parity, ab);
                    d_parity ^= 1u;
                    TC_TRACE(3, ti, 3 * u + 1);
                    tc_fence_after();
                    if (skip_tile) {
                        mma_commit(&bars.d_full);
                        if (u & 1) mma_commit_addr(a_empty0 + sweep_bar);
                        continue;
                    }
                    const TcUnitIssue ui = c_unit_issue[u];
                    // Issue order: the two small products (lo Bh, hi Bl: 2^-11 of the result) first, the main product hi Bh
                    // last.  The tensor cores truncate the fp32 accumulator at every MMA; added in this order only the 7 main
                    // steps truncate at the result's own magnitude instead of all 20 (measured against the CPU emulator:
                    // tests/tools/parity_full.py).
                    const uint32_t a_hi = tmem + ui.a_hi, a_lo = tmem + ui.a_lo, b_hi = desc0 + ui.b_hi, b_lo = desc0 + ui.b_lo;
                    mma_f16_ts<false>(d_tmem, a_lo, b_hi);
                    mma_f16_ts<true>(d_tmem, a_hi, b_lo);
#pragma unroll 1
                    for (uint32_t s = 1; s < kTcMainSteps; ++s) {
                        mma_f16_ts<true>(d_tmem, a_lo + 8 * s, b_hi + kStep * s);
                        mma_f16_ts<true>(d_tmem, a_hi + 8 * s, b_lo + kStep * s);
                    }
                    // slots 96..101: one K step over the unit's [hi | lo] leftover columns: hi Bl here, (hi + lo) Bh at the end
                    mma_f16_ts<true>(d_tmem, tmem + ui.a_left, desc0 + ui.b_left1);
#pragma unroll 1
                    for (uint32_t s = 0; s < kTcMainSteps; ++s) mma_f16_ts<true>(d_tmem, a_hi + 8 * s, b_hi + kStep * s);
                    mma_f16_ts<true>(d_tmem, tmem + ui.a_left, desc0 + ui.b_left0);
                    mma_commit(&bars.d_full);
                    TC_TRACE(3, ti, 3 * u + 2);
                    if (u & 1) mma_commit_addr(a_empty0 + sweep_bar);   // both units of the sweep have consumed its operand
                }
                a_parity ^= 1u;
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
        err = cudaFuncSetAttribute(logmel_tc_kernel<InT, NM, float, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(logmel_tc_kernel<InT, NM, __half, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(logmel_tc_kernel<InT, NM, float, kTcFixedPitch>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(logmel_tc_kernel<InT, NM, __half, kTcFixedPitch>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (err != cudaSuccess) return err;
        int sms = 0;
        if ((err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return err;
        sms_by_device[device] = sms;
    }
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    const int64_t tiles = a.batch * tiles_per_clip;
    // one CTA per SM at most; CTAs never wait for one another, so any number of them may actually be resident
    unsigned grid = static_cast<unsigned>(tiles < sms_by_device[device] ? tiles : sms_by_device[device]);
#if defined(B200MEL_TC_TRACE) || defined(B200MEL_TC_SWITCHES)
    if (std::getenv("B200MEL_TC_GRID") != nullptr) {   // measurement only: fewer CTAs (what a tile costs without the other SMs' traffic)
        const unsigned want = static_cast<unsigned>(std::atoi(std::getenv("B200MEL_TC_GRID")));
        if (want >= 1 && want < grid) grid = want;
    }
#endif
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
        const cuuint32_t box[4] = {static_cast<cuuint32_t>(kTcRowPitch), 1, static_cast<cuuint32_t>(kTcHalfRows), 1};
        const cuuint32_t elem[4] = {1, 1, 1, 1};
        if (encode != nullptr && rows < (int64_t{1} << 31) &&
            encode(&audio_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(a.audio), dims, strides, box, elem,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            tma_rows = static_cast<int>(rows);
    }
#if defined(B200MEL_TC_TRACE) || defined(B200MEL_TC_SWITCHES)
    if (std::getenv("B200MEL_TC_NO_TMA") != nullptr) tma_rows = 0;   // measurement only: every tile staged by the fold warps (cp.async)
#endif
    ProfileScope profile(2, stream);
    long long* trace = nullptr;
    int trace_first = 0;
#if defined(B200MEL_TC_TRACE)
    static long long* trace_dev = nullptr;
    static const bool want_trace = std::getenv("B200MEL_TC_TRACE") != nullptr;   // value: first of the 8 traced tiles of CTA 0
    if (want_trace) {
        trace_first = std::atoi(std::getenv("B200MEL_TC_TRACE")) & 0xff;
        if (trace_dev == nullptr) cudaMalloc(&trace_dev, sizeof(long long) * kTraceWords);
        cudaMemsetAsync(trace_dev, 0, sizeof(long long) * kTraceWords, stream);
        trace = trace_dev;
    }
#endif
#if defined(B200MEL_TC_TRACE) || defined(B200MEL_TC_SWITCHES)
    if (std::getenv("B200MEL_TC_FLAGS") != nullptr) trace_first |= std::atoi(std::getenv("B200MEL_TC_FLAGS")) << 8;
#endif
    const bool fixed = a.n_frames == kTcFixedPitch;
    if (a.out_f16) {
        if (fixed) logmel_tc_kernel<InT, NM, __half, kTcFixedPitch><<<grid, kTcThreads, kSmemBytes, stream>>>(a, audio_map, tma_rows, tables->operands, trace, trace_first);
        else logmel_tc_kernel<InT, NM, __half, 0><<<grid, kTcThreads, kSmemBytes, stream>>>(a, audio_map, tma_rows, tables->operands, trace, trace_first);
    } else {
        if (fixed) logmel_tc_kernel<InT, NM, float, kTcFixedPitch><<<grid, kTcThreads, kSmemBytes, stream>>>(a, audio_map, tma_rows, tables->operands, trace, trace_first);
        else logmel_tc_kernel<InT, NM, float, 0><<<grid, kTcThreads, kSmemBytes, stream>>>(a, audio_map, tma_rows, tables->operands, trace, trace_first);
    }
    count_launch();
    err = cudaGetLastError();
#if defined(B200MEL_TC_TRACE)
    if (trace != nullptr && err == cudaSuccess) {   // bring-up only: synchronises and prints CTA 0's timeline
        static long long host[kTraceWords];
        cudaStreamSynchronize(stream);
        cudaMemcpy(host, trace, sizeof(host), cudaMemcpyDeviceToHost);
        long long t0 = 0;
        for (int i = 0; i < kTraceRoles * kTraceTiles * kTraceEvents; ++i) { const long long v = host[i]; if (v != 0 && (t0 == 0 || v < t0)) t0 = v; }
        static const char* names[kTraceRoles] = {"-", "fold-0", "fold-1", "mma", "epi", "fold-2"};
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
#endif
    return err;
}


// ---- finish kernel: what the clamp at max - 8 (audio.py:155) does to each tile ------------------------------------
// The front-end kernel leaves y = (log10 + 4) / 4 and the extremes of every utterance and of every 128-frame tile; the
// clamp needs the utterance's (or, for the reference's 2-D semantics, the whole call's) final max g: out = max(y, floor),
// floor = ((g - 8) + 4) / 4 - identical to (max(lg, g - 8) + 4) / 4, and, rounding being monotone, also in half precision.
// One warp per tile, a few words per tile: most tiles need nothing (their smallest value is not below g - 8); a tile that
// lies wholly below the clamp - or was never stored because its samples were all zero - is filled with the constant; the
// rest (an utterance's tile where the sound stops) is clamped value by value while it is still in L2.
template <int NM, typename OutT>
__global__ void __launch_bounds__(32 * kFinishWarps) tc_finish_kernel(const LogmelArgs a, int tiles_per_clip) {
    // launched with programmatic stream serialisation: the grid is set up while the front-end kernel drains, and waits
    // here until that kernel has completed and its writes are visible
    asm volatile("griddepcontrol.wait;" ::: "memory");
    struct Work { OutT* out; float value; int frames; int action; };
    __shared__ Work work[kFinishWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t total_tiles = a.batch * tiles_per_clip;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kFinishWarps;
    // every warp looks at one tile; then the block works through the tiles that need something, all warps on each
    for (int64_t base = static_cast<int64_t>(blockIdx.x) * kFinishWarps; base < total_tiles; base += stride) {
        const int64_t tile = base + warp;
        int action = 0;                      // 0: leave the tile alone, 1: clamp it in place, 2: fill it
        float value = 0.f;
        int frames = 0;
        OutT* tile_out = nullptr;
        if (tile < total_tiles) {
            const int64_t clip = tile / tiles_per_clip;
            const int t0 = static_cast<int>(tile - clip * tiles_per_clip) * kTcTileFrames;
            const uint32_t gkey = __ldg(a.max_keys + (a.global_max ? 0 : clip));
            const float g = gkey == 0u ? -10.0f : max_key_decode(gkey);     // (key 0: nothing but silent tiles)
            const float floor_lg = g - 8.0f;
            const float floor_y = ((g - 8.0f) + 4.0f) * 0.25f;
            value = floor_y;
            if (a.tile_keys != nullptr) {
                const uint32_t kmax = __ldg(a.tile_keys + 2 * tile), kmin = __ldg(a.tile_keys + 2 * tile + 1);
                if (kmax == 0u) {                // never stored: all its samples were zero, every value is log10(1e-10) = -10
                    action = 2;
                    value = floor_lg > -10.0f ? floor_y : (floor_lg != floor_lg ? floor_y : -1.5f);
                } else {
                    const float tile_max = max_key_decode(kmax), tile_min = max_key_decode(~kmin);
                    if (!(tile_min >= floor_lg)) action = tile_max < floor_lg ? 2 : 1;
                }
            } else {
                const float smallest = max_key_decode(~__ldg(a.min_keys + clip));
                if (!(smallest >= floor_lg)) action = 1;
            }
            frames = a.n_frames - t0 < kTcTileFrames ? a.n_frames - t0 : kTcTileFrames;
            tile_out = reinterpret_cast<OutT*>(a.out) + clip * NM * static_cast<int64_t>(a.n_frames) + t0;
        }
        const bool any = __syncthreads_or(action != 0);
        if (!any) continue;                  // (block-uniform)
        if (lane == 0) work[warp] = Work{tile_out, value, frames, action};
        __syncthreads();
#pragma unroll 1
        for (int w = 0; w < kFinishWarps; ++w) {
            const Work job = work[w];
            if (job.action == 2) fill_tile_rows<NM, OutT>(job.out, a.n_frames, job.frames, job.value, warp, lane);
            else if (job.action == 1) normalise_tile_rows<NM, OutT>(job.out, a.n_frames, job.frames, job.value, warp, lane);
        }
        __syncthreads();                     // the list is rewritten in the next round
    }
}

}  // namespace

cudaError_t launch_tc_finish(const LogmelArgs& a, cudaStream_t stream) {
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    const int64_t tiles = a.batch * tiles_per_clip;
    if (tiles <= 0) return cudaSuccess;
    int device = 0, sms = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (err != cudaSuccess) return err;
    const int64_t want = (tiles + 7) / 8;
    const unsigned grid = static_cast<unsigned>(want < 8 * static_cast<int64_t>(sms) ? want : 8 * static_cast<int64_t>(sms));
    ProfileScope profile(1, stream);
    cudaLaunchConfig_t config = {};
    config.gridDim = dim3(grid);
    config.blockDim = dim3(32 * kFinishWarps);
    config.dynamicSmemBytes = 0;
    config.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    config.attrs = attr;
    config.numAttrs = 1;
    if (a.n_mels == 80) {
        err = a.out_f16 ? cudaLaunchKernelEx(&config, tc_finish_kernel<80, __half>, a, tiles_per_clip)
                        : cudaLaunchKernelEx(&config, tc_finish_kernel<80, float>, a, tiles_per_clip);
    } else if (a.n_mels == 128) {
        err = a.out_f16 ? cudaLaunchKernelEx(&config, tc_finish_kernel<128, __half>, a, tiles_per_clip)
                        : cudaLaunchKernelEx(&config, tc_finish_kernel<128, float>, a, tiles_per_clip);
    } else {
        return cudaErrorInvalidValue;
    }
    if (err != cudaSuccess) return err;
    count_launch();
    return cudaGetLastError();
}

unsigned tc_kernel_fault(unsigned* cta) {
    unsigned host[2] = {0u, 0u};
    if (cudaMemcpyFromSymbol(host, g_tc_fault, sizeof(host)) != cudaSuccess) { cudaGetLastError(); return 0u; }
    if (cta != nullptr) *cta = host[1];
    return host[0];
}

cudaError_t launch_tc_pass1(const LogmelArgs& a, const TcTables* tables, int dtype, cudaStream_t stream) {
    const int tiles_per_clip = (a.n_frames + kTcTileFrames - 1) / kTcTileFrames;
    if (a.batch * tiles_per_clip <= 0) return cudaSuccess;
    if (a.n_mels == 80) return dtype == 0 ? launch_tc<float, 80>(a, tables, stream) : launch_tc<int16_t, 80>(a, tables, stream);
    if (a.n_mels == 128) return dtype == 0 ? launch_tc<float, 128>(a, tables, stream) : launch_tc<int16_t, 128>(a, tables, stream);
    return cudaErrorInvalidValue;
}

}  // namespace b200mel
