// Encoder stem behind the front-end: conv1 (kernel 3, padding 1) + GELU of Whisper's AudioEncoder
// (reference whisper/model.py:179, :193: `x = F.gelu(self.conv1(x))`) as an implicit GEMM on the tcgen05 tensor cores,
// with the last step of the log-mel normalisation - the clamp at max - 8 (whisper/audio.py:155) - folded into its
// input load, so the front-end can hand over its un-clamped output and one max per utterance.
//
//   out[b, n, t] = gelu(bias[n] + sum_{c, k} W[n, c, k] x[b, c, t + k - 1]),   x = max(y, floor_b), x[.., -1] = x[.., T] = 0
//
// As a GEMM per 128-frame tile: D[128 channels, 128 frames] = sum over the three taps k of W_k[128 channels, n_mels]
// X_k[n_mels, 128 frames], kind::tf32 (operands rounded to TF32 - what cudnn's convolution does with torch's default
// allow_tf32 - fp32 accumulation in tensor memory).
//   A = the CTA's 128-channel slice of the weights, all three taps: 240 columns of TENSOR MEMORY, written once per CTA;
//   B = the input tile from shared memory, staged ONCE, transposed to K-major - [n_mels / 4][130 frames][4 mels]: every
//       frame a 16-byte row, rows contiguous - so the operand of tap k is the same tile starting k rows (16 k bytes)
//       further on: no im2col copy;
//   D = two accumulators of 128 columns: the epilogue of a tile overlaps the MMAs of the next.
// One persistent CTA per SM; CTAs with the other slices walk the same tiles at the same time, so the input comes from HBM
// once.  17 warps: 2 x 4 epilogue (thread = channel: bias, GELU, the 32 x 32 piece through a swizzled shared-memory buffer
// and out with a TMA tensor store - full lines, no L1 tag traffic), 1 MMA issue, 8 loaders (coalesced 16-byte loads,
// clamp, TF32 rounding, 4 x 4 register transpose, 16-byte shared stores; a tile's loads are issued before the wait for
// its stage, the next tile's lines are asked from L2).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace b200mel {

namespace {

constexpr int kStemTile = 128;                  // frames per tile = MMA N
constexpr int kStemRows = kStemTile + 2;        // + one frame either side (kernel 3, padding 1)
constexpr int kStemN = 128;                     // channels per CTA = MMA M = TMEM lanes
constexpr int kStemMels = 80;
constexpr int kStemKChunks = kStemMels / 4;     // 16-byte K chunks (4 tf32)
constexpr int kStemChunkBytes = kStemRows * 16 + 16;   // 2096: one K chunk of the tile + 16 bytes, so that the loaders' 16-byte stores
                                                       // (4 mel quads x 2 frame groups per quarter warp: chunk stride 524 words = 12 banks) hit 8 different bank groups
constexpr int kStemXBytes = (kStemKChunks * kStemChunkBytes + 127) / 128 * 128;   // 41984 per input buffer (a multiple of 128)
constexpr int kStemStages = 3;
constexpr int kStemWarpMma = 8, kStemLoaderWarps = 8, kStemLoaderGroups = 1;   // warps 0-7 epilogue, 8 MMA issue, 9-16 loaders (one group of eight: a second group taking every other tile measured 2.5 % slower)
constexpr int kStemWarps = kStemWarpMma + 1 + kStemLoaderGroups * kStemLoaderWarps, kStemThreads = kStemWarps * 32;
constexpr int kStemPieceBytes = 32 * 32 * 4;    // an epilogue warp's 32 channels x 32 frames on their way out
constexpr int kStemOutOffset = (kStemStages * kStemXBytes + 1023) / 1024 * 1024;   // (128-byte swizzle: 1024-byte aligned)
constexpr int kStemSmem = kStemOutOffset + kStemWarpMma * 2 * kStemPieceBytes + 1024;   // + slack to align the base
constexpr int kStemWCols = 3 * kStemMels;       // tensor-memory columns of the weights: column = tap * 80 + mel
constexpr int kStemDCol = 256;                  // the two accumulators: columns 256-383, 384-511
constexpr uint32_t kStemIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kStemTile >> 3) << 17) |
                                (static_cast<uint32_t>(kStemN >> 4) << 24);   // tf32 x tf32 -> f32, K-major, M 128, N 128
static_assert(kStemXBytes % 128 == 0 && kStemSmem <= 227 * 1024, "shared memory layout");
static_assert(kStemWCols <= kStemDCol, "tensor memory layout");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
        if (done || ++spins > (1u << 17)) break;   // (a protocol bug ends the kernel with garbage instead of hanging the device)
    }
}
// for the roles that run several tiles ahead: back off between polls instead of taking issue slots from the epilogue
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, unsigned sleep_ns) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done || ++spins > (1u << 22)) break;
        __nanosleep(sleep_ns);
    }
}
// a loader group's wait for its stage: one warp polls the mbarrier, the other seven block on a named barrier - a blocked
// warp takes no issue slots, a polling one does (measured: sixteen polling warps took a third of the SM's issue slots)
__device__ __forceinline__ void stage_wait(uint64_t* bar, uint32_t parity, bool leader, int bar_id) {
    if (leader) mbar_wait_relaxed(bar, parity, 500);
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(kStemLoaderWarps * 32) : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float x) {      // round to nearest (ties away), as cudnn / cuBLAS do
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// K-major, no-swizzle operand: rows of 16 bytes, 8-row groups 128 bytes apart (i.e. rows contiguous), K chunks `lbo` bytes apart
__device__ __forceinline__ uint64_t stem_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (static_cast<uint64_t>(0x4008u) << 32) | ((smem_addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, bool accumulate) {
    asm volatile("{\n.reg .pred P, Q;\nelect.sync _|P, 0xffffffff;\nsetp.ne.u32 Q, %4, 0;\n"
                 "@P tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, Q;\n}\n"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(kStemIdesc), "r"(accumulate ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\n@P tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t t, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(t), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
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

// GELU(v) = v Phi(v) = h + |h| erf(|h| sqrt 2), h = v / 2, with erfc(u) = 2^(-u q(u)) on [0, 4.3] (beyond: < 2e-9):
// q a degree-6 fit (tools/fit_gelu.py: |erfc error| 4.5e-7, |GELU error| <= 8.2e-8 for every v - below float32's spacing
// at 1) - one MUFU and nine FMA-pipe instructions instead of erff's two dozen.  `h` = half the pre-activation.
__device__ __forceinline__ float gelu_from_half(float h) {
    const float a = fabsf(h);
    const float u = fminf(a * 1.41421356237309515f, 4.3f);
    float q = -5.393774335971102e-05f;
    q = fmaf(q, u, 0.00014493041089735925f);
    q = fmaf(q, u, 0.0031301151029765606f);
    q = fmaf(q, u, -0.03049657866358757f);
    q = fmaf(q, u, 0.14962077140808105f);
    q = fmaf(q, u, 0.918138325214386f);
    q = fmaf(q, u, 1.6279326677322388f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-u * q));
    return fmaf(-a, e, h + a);       // h + |h| (1 - erfc): NaN stays NaN
}

// two values at once: the same operations (and roundings) as gelu_from_half on each, as packed FMUL2 / FFMA2 / FADD2 -
// half the issue slots of the epilogue's arithmetic
__device__ __forceinline__ float2 gelu2_from_half(float2 h) {
    const float2 a = make_float2(fabsf(h.x), fabsf(h.y));
    float2 u = __fmul2_rn(a, make_float2(1.41421356237309515f, 1.41421356237309515f));
    u.x = fminf(u.x, 4.3f);
    u.y = fminf(u.y, 4.3f);
    float2 q = make_float2(-5.393774335971102e-05f, -5.393774335971102e-05f);
    q = __ffma2_rn(q, u, make_float2(0.00014493041089735925f, 0.00014493041089735925f));
    q = __ffma2_rn(q, u, make_float2(0.0031301151029765606f, 0.0031301151029765606f));
    q = __ffma2_rn(q, u, make_float2(-0.03049657866358757f, -0.03049657866358757f));
    q = __ffma2_rn(q, u, make_float2(0.14962077140808105f, 0.14962077140808105f));
    q = __ffma2_rn(q, u, make_float2(0.918138325214386f, 0.918138325214386f));
    q = __ffma2_rn(q, u, make_float2(1.6279326677322388f, 1.6279326677322388f));
    const float2 t = __fmul2_rn(make_float2(-u.x, -u.y), q);
    float2 e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(t.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(t.y));
    return __ffma2_rn(make_float2(-a.x, -a.y), e, __fadd2_rn(h, a));
}

// (clip, tile inside the clip) of a CTA's k-th tile, advanced by the grid's stride without a division
struct TileStep { int clips, tiles, tiles_per_clip; };
struct TileWalk {
    int clip, tile;
    __device__ __forceinline__ void advance(const TileStep& s) {
        clip += s.clips;
        tile += s.tiles;
        if (tile >= s.tiles_per_clip) { tile -= s.tiles_per_clip; ++clip; }
    }
};

struct StemBarriers { uint64_t x_full[kStemStages], x_empty[kStemStages], d_full[2], d_empty[2]; };

struct StemArgs {
    const float* mel;          // [batch, 80, n_frames]: y = (log10 + 4) / 4, clamped or not
    const uint32_t* max_keys;  // [batch] ([1] with global_max) order-preserving keys of the max log10 (the front-end's workspace), or nullptr: `mel` is final
    const uint32_t* tile_keys; // [batch * tiles][2] or nullptr: a tile whose max key is 0 was never stored by the front-end (all its samples were zero)
    int global_max;
    const float* weight;       // [n_state, 80, 3] (torch Conv1d layout)
    const float* bias;         // [n_state]
    float* out;                // [batch, n_state, n_frames]
    __half* out_fm;            // frame-major variant: half [batch, fm_frames, n_state], the operand layout of conv2 (stem_conv2.cu)
    int fm_frames;             // its frames per clip (n_frames rounded up to even)
    int fm_tma;                // out_map describes out_fm as [batch * fm_frames rows, n_state]: whole 32-frame pieces leave by TMA tensor store
    int64_t batch;
    int n_frames, n_state;
    int vector_io;             // n_frames % 4 == 0 and 16-byte aligned pointers: 16-byte loads, TMA tensor stores
    int debug;                 // measurement switches (switches build only, B200MEL_STEM_FLAGS): 1 no GELU, 2 no stores, 4 no loads
};
#if defined(B200MEL_TC_TRACE) || defined(B200MEL_TC_SWITCHES)
#define STEM_DEBUG(bit) ((a.debug & (bit)) != 0)
#else
#define STEM_DEBUG(bit) false
#endif

// clamp of audio.py:155 in the (x + 4) / 4 domain, then TF32
// (max.NaN keeps a NaN of either side, like torch.maximum; the tensor core drops the low 13 bits of what it reads, so
// adding half a TF32 ulp to the bit pattern of a finite value is round-to-nearest, ties away - what cvt.rna.tf32.f32 does,
// without its final mask.  Not for NaN: max.NaN returns 0x7fffffff, which the add would wrap to a denormal.)
__device__ __forceinline__ uint32_t stem_input(float y, float floor_y) {
    const float m = max_nan(y, floor_y);
    return __float_as_uint(m) + (fabsf(m) < __uint_as_float(0x7f800000u) ? 0x1000u : 0u);
}

template <bool kFm>
__global__ void __launch_bounds__(kStemThreads, 1) stem_conv1_gelu_kernel(const StemArgs a, const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* const smem_raw = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
    __shared__ __align__(8) StemBarriers bars;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int slices = a.n_state / kStemN;
    const int slice = blockIdx.x % slices;                       // this CTA's 128 channels
    const int walkers = gridDim.x / slices;                      // CTAs that share the tiles of a slice
    const int walker = blockIdx.x / slices;
    const int tiles_per_clip = (a.n_frames + kStemTile - 1) / kStemTile;
    const int64_t total_tiles = a.batch * tiles_per_clip;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < kStemStages; ++i) {
            mbar_init(&bars.x_full[i], kStemLoaderWarps);
            mbar_init(&bars.x_empty[i], 1);      // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars.d_full[i], 1);       // tcgen05.commit
            mbar_init(&bars.d_empty[i], 4);      // the accumulator's four epilogue warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp < 4) {
        // the slice's weights into tensor memory: lane = channel, column = tap * 80 + mel
        const float* w = a.weight + (static_cast<int64_t>(slice) * kStemN + warp * 32 + lane) * kStemWCols;
        const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll 1
        for (int m = 0; m < kStemWCols / 8; ++m) {
            const int tap = (8 * m) / kStemMels, c0 = (8 * m) % kStemMels;
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = to_tf32(__ldg(w + (c0 + i) * 3 + tap));
            tmem_st8(lane_addr + 8 * m, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // this CTA's tiles: walker, walker + walkers, ... - (clip, tile inside the clip) advance without a division per tile
    const int my_tiles = walker < total_tiles ? static_cast<int>((total_tiles - walker + walkers - 1) / walkers) : 0;
    const TileWalk first{walker / tiles_per_clip, walker % tiles_per_clip};
    const TileStep step{walkers / tiles_per_clip, walkers % tiles_per_clip, tiles_per_clip};

    if (warp > kStemWarpMma) {
        // ===== loaders: one tile [80 mels x 130 frames], clamped, rounded and transposed to [mel / 4][frame][mel % 4];
        // kStemLoaderGroups groups of 8 warps, group g takes the CTA's tiles g, g + groups, ... =====
        const int group = (warp - kStemWarpMma - 1) / kStemLoaderWarps;
        const int lw = (warp - kStemWarpMma - 1) % kStemLoaderWarps, lt = lw * 32 + lane;
        const int q_in = lane & 3, g_in = lane >> 2;             // 4 mel quads x 8 frame groups per warp item
        // vector path: 20 warp items (5 blocks of 4 mel quads x 4 blocks of 8 frame groups of 4) over 8 warps
        int item_f[3];
        int64_t item_src[3];
        uint32_t item_dst[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int item = lw + kStemLoaderWarps * r;
            const int q = (item % 5) * 4 + q_in;
            item_f[r] = 4 * ((item / 5) * 8 + g_in);             // frame inside the tile; row = f + 1
            item_src[r] = static_cast<int64_t>(4 * q) * a.n_frames + item_f[r];
            item_dst[r] = q * kStemChunkBytes + (item_f[r] + 1) * 16;
        }
        const bool third = lw + 2 * kStemLoaderWarps < 20;       // warps 0-3 have a third item
        const int halo_c = lt % kStemMels, halo_side = lt / kStemMels;   // threads 0-159: the frame before / behind the tile
        const uint32_t halo_dst = (halo_c >> 2) * kStemChunkBytes + (halo_side ? kStemRows - 1 : 0) * 16 + (halo_c & 3) * 4;
        TileWalk w = first, ahead = first;
        for (int g = 0; g < group; ++g) w.advance(step);
        ahead = w;
        auto advance_group = [&](TileWalk& t) {
#pragma unroll
            for (int g = 0; g < kStemLoaderGroups; ++g) t.advance(step);
        };
        advance_group(ahead);
        for (int k = group; k < my_tiles; k += kStemLoaderGroups, w = ahead, advance_group(ahead)) {
            const int stage = k % kStemStages;
            const uint32_t parity = ((k / kStemStages) & 1) ^ 1u;     // x_empty: the first use of a stage passes
            const int t0 = w.tile * kStemTile;
            const int64_t tile = static_cast<int64_t>(w.clip) * tiles_per_clip + w.tile;
            float floor_y = __uint_as_float(0xff800000u);        // -inf: no clamp
            if (a.max_keys != nullptr) {
                const uint32_t key = __ldg(a.max_keys + (a.global_max ? 0 : w.clip));
                const float g = key == 0u ? -10.0f : max_key_decode(key);
                floor_y = ((g - 8.0f) + 4.0f) * 0.25f;
            }
            const uint32_t* tk = a.tile_keys != nullptr ? a.tile_keys + 2 * tile : nullptr;
            const bool silent = tk != nullptr && __ldg(tk) == 0u;                 // never written: (log10(1e-10) + 4) / 4 everywhere
            const float* src = a.mel + (static_cast<int64_t>(w.clip) * kStemMels * a.n_frames + t0);
            unsigned char* xs = smem_raw + stage * kStemXBytes;
            // the frame before and the frame behind the tile (the convolution's zero padding at the ends of the clip)
            float halo = 0.f;
            bool halo_inside = false;
            if (lt < 2 * kStemMels) {
                const int t = halo_side ? t0 + kStemTile : t0 - 1;
                halo_inside = t >= 0 && t < a.n_frames;
                if (halo_inside) {
                    const bool halo_silent = tk != nullptr && __ldg(tk + (halo_side ? 2 : -2)) == 0u;
                    halo = halo_silent ? -1.5f : __ldg(src + static_cast<int64_t>(halo_c) * a.n_frames + (t - t0));
                }
            }
            if (a.vector_io) {
                // all loads first.  n_frames % 4 == 0: a group of 4 frames is inside the clip or outside as a whole.
                float4 v[3][4];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float* p = src + item_src[r];
                    const bool load = (r < 2 || third) && !silent && t0 + item_f[r] < a.n_frames && !STEM_DEBUG(4);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        v[r][m] = silent ? make_float4(-1.5f, -1.5f, -1.5f, -1.5f) : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (load) v[r][m] = __ldg(reinterpret_cast<const float4*>(p + static_cast<int64_t>(m) * a.n_frames));
                    }
                }
                // ask L2 for this thread's part of the group's next tile
                if (k + kStemLoaderGroups < my_tiles) {
                    const bool asilent = a.tile_keys != nullptr && __ldg(a.tile_keys + 2 * (static_cast<int64_t>(ahead.clip) * tiles_per_clip + ahead.tile)) == 0u;
                    const float* asrc = a.mel + (static_cast<int64_t>(ahead.clip) * kStemMels * a.n_frames + ahead.tile * kStemTile);
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        if ((r < 2 || third) && !asilent && ahead.tile * kStemTile + item_f[r] < a.n_frames) {
#pragma unroll
                            for (int m = 0; m < 4; ++m)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(asrc + item_src[r] + static_cast<int64_t>(m) * a.n_frames));
                        }
                    }
                }
                stage_wait(&bars.x_empty[stage], parity, lw == 0, 1 + group);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    if (r < 2 || third) {
                        const bool inside = t0 + item_f[r] < a.n_frames;
                        uint4* dst = reinterpret_cast<uint4*>(xs + item_dst[r]);
                        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
                        dst[0] = inside ? make_uint4(stem_input(v[r][0].x, floor_y), stem_input(v[r][1].x, floor_y), stem_input(v[r][2].x, floor_y), stem_input(v[r][3].x, floor_y)) : zero;
                        dst[1] = inside ? make_uint4(stem_input(v[r][0].y, floor_y), stem_input(v[r][1].y, floor_y), stem_input(v[r][2].y, floor_y), stem_input(v[r][3].y, floor_y)) : zero;
                        dst[2] = inside ? make_uint4(stem_input(v[r][0].z, floor_y), stem_input(v[r][1].z, floor_y), stem_input(v[r][2].z, floor_y), stem_input(v[r][3].z, floor_y)) : zero;
                        dst[3] = inside ? make_uint4(stem_input(v[r][0].w, floor_y), stem_input(v[r][1].w, floor_y), stem_input(v[r][2].w, floor_y), stem_input(v[r][3].w, floor_y)) : zero;
                    }
                }
            } else {
                // any n_frames / alignment: element by element (consecutive threads along the frame axis)
                stage_wait(&bars.x_empty[stage], parity, lw == 0, 1 + group);
#pragma unroll 1
                for (int i = lt; i < kStemMels * kStemTile; i += kStemLoaderWarps * 32) {
                    const int c = i / kStemTile, f = i % kStemTile;
                    uint32_t x = 0u;
                    if (t0 + f < a.n_frames) x = stem_input(silent ? -1.5f : __ldg(src + static_cast<int64_t>(c) * a.n_frames + f), floor_y);
                    *reinterpret_cast<uint32_t*>(xs + (c >> 2) * kStemChunkBytes + (f + 1) * 16 + (c & 3) * 4) = x;
                }
            }
            if (lt < 2 * kStemMels) *reinterpret_cast<uint32_t*>(xs + halo_dst) = halo_inside ? stem_input(halo, floor_y) : 0u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.x_full[stage]);
        }
    } else if (warp == kStemWarpMma) {
        // ===== MMA issue: 3 taps x 10 K steps of 8 per tile =====
        uint32_t x_parity = 0, d_parity = 1;                     // d_empty: the first two waits pass
        int stage = 0, buf = 0;
        const uint32_t x_base = smem_u32(smem_raw);
        for (int k = 0; k < my_tiles; ++k) {
            mbar_wait_relaxed(&bars.x_full[stage], x_parity, 32);
            mbar_wait_relaxed(&bars.d_empty[buf], d_parity, 32);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t desc0 = stem_desc(x_base + stage * kStemXBytes, kStemChunkBytes);
            const uint32_t d_tmem = tmem + kStemDCol + buf * kStemTile;
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                for (int j = 0; j < kStemMels / 8; ++j)          // K = 8 tf32 per MMA = two 16-byte chunks; the tap moves the start by one row
                    mma_tf32_ts(d_tmem, tmem + kk * kStemMels + 8 * j, desc0 + ((16 * kk + 2 * j * kStemChunkBytes) >> 4), kk + j > 0);
            mma_commit(&bars.d_full[buf]);
            mma_commit(&bars.x_empty[stage]);
            if (++stage == kStemStages) { stage = 0; x_parity ^= 1u; }
            buf ^= 1;
            if (buf == 0) d_parity ^= 1u;
        }
    } else {
        // ===== epilogue: two warps per TMEM lane quadrant, one per accumulator - they take the CTA's tiles in turn.
        // thread = channel, 128 frames in pieces of 32: bias, GELU, 32-byte sector stores =====
        const int buf = warp >> 2;
        uint32_t parity = 0;
        const int n = (warp & 3) * 32 + lane;
        const float half_bias = 0.5f * __ldg(a.bias + slice * kStemN + n);
        const uint32_t d_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + kStemDCol + buf * kStemTile;
        float* const out_n = a.out + (static_cast<int64_t>(slice) * kStemN + n) * a.n_frames;
        // the warp's two 32 x 32 staging pieces: rows of 128 bytes = 32 frames of one channel, 16-byte chunks XOR-swizzled by
        // the row (CU_TENSOR_MAP_SWIZZLE_128B), so that the 8 lanes of a store phase hit 8 different chunks
        const uint32_t piece_base = smem_u32(smem_raw + kStemOutOffset + warp * 2 * kStemPieceBytes) + lane * 128;
        const uint32_t swizzle = (lane & 7) << 4;
        const int row0 = slice * kStemN + (warp & 3) * 32;       // + clip * n_state: the piece's first row of out viewed as [batch * n_state, n_frames]
        int pieces_out = 0;
        TileWalk w = first;
        if (buf) w.advance(step);
        for (int k = buf; k < my_tiles; k += 2, w.advance(step), w.advance(step)) {
            const int t0 = w.tile * kStemTile;
            float* out = out_n + (static_cast<int64_t>(w.clip) * a.n_state * a.n_frames + t0);
            mbar_wait_relaxed(&bars.d_full[buf], parity, 100);
            parity ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // The accumulator comes in half pieces of 16 frames through two sets of registers: the next half piece's tcgen05.ld is
            // in flight while this one goes through the GELU; two half pieces make one 32 x 32 store.
            auto activate = [&](float (&d)[16]) {
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float2 h = __ffma2_rn(make_float2(d[i], d[i + 1]), make_float2(0.5f, 0.5f), make_float2(half_bias, half_bias));
                    const float2 g = STEM_DEBUG(1) ? h : gelu2_from_half(h);
                    d[i] = g.x;
                    d[i + 1] = g.y;
                }
            };
            auto stage = [&](const float (&d)[16], uint32_t dst, int half_piece) {   // 16 frames = chunks 4 h .. 4 h + 3 of the 128-byte row
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((4 * half_piece + j) << 4) ^ swizzle)), "f"(d[4 * j]),
                                 "f"(d[4 * j + 1]), "f"(d[4 * j + 2]), "f"(d[4 * j + 3]) : "memory");
            };
            float da[16], db[16];
            tmem_ld16(d_addr, da);
#pragma unroll 1
            for (int piece = 0; piece < kStemTile / 32; ++piece) {
                const int t = t0 + piece * 32;
                const bool store = t < a.n_frames && !STEM_DEBUG(2);   // (warp-uniform)
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");            // da: frames 32 p .. + 15
                tmem_ld16(d_addr + piece * 32 + 16, db);
                activate(da);
                uint32_t dst = 0;
                const bool fm_staged = kFm && a.fm_tma && t + 32 <= a.n_frames;   // (warp-uniform) a whole piece inside the clip
                if constexpr (kFm) {
                    // frame-major half: a frame's 32 channels of this warp are 64 contiguous bytes.  A whole piece is staged as
                    // [32 frames][32 channels] and leaves with one TMA tensor store; the clip's last, partial piece by itself.
                    if (store && fm_staged) {
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        __syncwarp();
                        dst = piece_base - lane * 128 + (pieces_out & 1) * kStemPieceBytes;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst + i * 64 + lane * 2), "h"(__half_as_ushort(__float2half_rn(da[i]))) : "memory");
                    } else if (store) {
                        __half* fm = a.out_fm + ((static_cast<int64_t>(w.clip) * a.fm_frames + t) * a.n_state + slice * kStemN + n);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (t + i < a.n_frames) fm[static_cast<int64_t>(i) * a.n_state] = __float2half_rn(da[i]);
                    }
                } else if (store && a.vector_io) {
                    // the staging piece used two stores ago must have been read by the TMA unit
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    dst = piece_base + (pieces_out & 1) * kStemPieceBytes;
                    stage(da, dst, 0);
                } else if (store) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (t + i < a.n_frames) out[piece * 32 + i] = da[i];
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");            // db: frames 32 p + 16 .. + 31
                if (piece + 1 < kStemTile / 32) {
                    tmem_ld16(d_addr + (piece + 1) * 32, da);
                } else {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.d_empty[buf]);   // the whole accumulator is in registers
                }
                activate(db);
                if constexpr (kFm) {
                    if (store && fm_staged) {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst + (16 + i) * 64 + lane * 2), "h"(__half_as_ushort(__float2half_rn(db[i]))) : "memory");
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                         ::"l"(&out_map), "r"(slice * kStemN + (warp & 3) * 32), "r"(w.clip * a.fm_frames + t), "r"(dst) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        ++pieces_out;
                    } else if (store) {
                        __half* fm = a.out_fm + ((static_cast<int64_t>(w.clip) * a.fm_frames + t + 16) * a.n_state + slice * kStemN + n);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (t + 16 + i < a.n_frames) fm[static_cast<int64_t>(i) * a.n_state] = __float2half_rn(db[i]);
                    }
                } else if (store && a.vector_io) {
                    stage(db, dst, 1);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        // frames beyond n_frames are clipped by the tensor map
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                     ::"l"(&out_map), "r"(t), "r"(w.clip * a.n_state + row0), "r"(dst - lane * 128) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++pieces_out;
                } else if (store) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (t + 16 + i < a.n_frames) out[piece * 32 + 16 + i] = db[i];
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the staging pieces live until the stores have read them
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace

cudaError_t launch_stem_conv1_gelu(const float* mel, const uint32_t* max_keys, const uint32_t* tile_keys, int global_max, int64_t batch, int n_frames, const float* weight,
                                   const float* bias, int n_state, float* out, void* out_fm16, cudaStream_t stream) {
    if (batch <= 0 || n_frames <= 0) return cudaSuccess;
    constexpr int kMaxDevices = 64;
    static int sms_by_device[kMaxDevices] = {0};                      // (also: the kernel's attributes are set on this device)
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err != cudaSuccess) return err;
    if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
    if (sms_by_device[device] == 0) {
        err = cudaFuncSetAttribute(stem_conv1_gelu_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(stem_conv1_gelu_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem);
        int count = 0;
        if (err == cudaSuccess) err = cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device);
        if (err != cudaSuccess) return err;
        sms_by_device[device] = count;
    }
    const int sms = sms_by_device[device];
    const int slices = n_state / kStemN;
    const int64_t tiles = batch * ((n_frames + kStemTile - 1) / kStemTile);
    int64_t walkers = sms / slices;                                   // CTAs per slice; every CTA of the grid is resident
    if (walkers < 1) walkers = 1;
    if (walkers > tiles) walkers = tiles;
    // the output as the TMA unit sees it: [batch * n_state rows, n_frames], boxes of 32 rows x 32 frames, 128-byte swizzle.
    // Anything that keeps it from being described (odd n_frames, unaligned pointers, an absurd size) leaves vector_io = 0:
    // scalar loads and stores.
    CUtensorMap out_map;
    std::memset(&out_map, 0, sizeof(out_map));
    int vector_io = 0;
    int fm_tma = 0;
    const int fm_frames = n_frames + (n_frames & 1);
    if (out_fm16 != nullptr) {
        vector_io = n_frames % 4 == 0 && reinterpret_cast<uintptr_t>(mel) % 16 == 0;   // (the loads only)
        // the frame-major result as the TMA unit sees it: [batch * fm_frames rows, n_state] half, boxes of 32 frames x 32 channels
        if (reinterpret_cast<uintptr_t>(out_fm16) % 16 == 0 && batch * fm_frames < (int64_t{1} << 31)) {
            using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            static EncodeFn encode = [] {
                void* fn = nullptr;
                cudaDriverEntryPointQueryResult q;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
                return reinterpret_cast<EncodeFn>(fn);
            }();
            const cuuint64_t dims[2] = {static_cast<cuuint64_t>(n_state), static_cast<cuuint64_t>(batch * fm_frames)};
            const cuuint64_t strides[1] = {static_cast<cuuint64_t>(n_state) * 2};
            const cuuint32_t box[2] = {32, 32};
            const cuuint32_t elem[2] = {1, 1};
            if (encode != nullptr && encode(&out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out_fm16, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                fm_tma = 1;
        }
    } else if (n_frames % 4 == 0 && reinterpret_cast<uintptr_t>(mel) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
        batch * n_state < (int64_t{1} << 31)) {
        using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = [] {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
            return reinterpret_cast<EncodeFn>(fn);
        }();
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(n_frames), static_cast<cuuint64_t>(batch * n_state)};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(n_frames) * 4};
        const cuuint32_t box[2] = {32, 32};
        const cuuint32_t elem[2] = {1, 1};
        if (encode != nullptr && encode(&out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            vector_io = 1;
    }
    StemArgs a{mel, max_keys, tile_keys, global_max, weight, bias, out, static_cast<__half*>(out_fm16), fm_frames, fm_tma, batch, n_frames, n_state, vector_io, 0};
#if defined(B200MEL_TC_TRACE) || defined(B200MEL_TC_SWITCHES)
    if (std::getenv("B200MEL_STEM_FLAGS") != nullptr) a.debug = std::atoi(std::getenv("B200MEL_STEM_FLAGS"));
#endif
    ProfileScope profile(3, stream);
    if (out_fm16 != nullptr) stem_conv1_gelu_kernel<true><<<static_cast<unsigned>(walkers * slices), kStemThreads, kStemSmem, stream>>>(a, out_map);
    else stem_conv1_gelu_kernel<false><<<static_cast<unsigned>(walkers * slices), kStemThreads, kStemSmem, stream>>>(a, out_map);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200mel
