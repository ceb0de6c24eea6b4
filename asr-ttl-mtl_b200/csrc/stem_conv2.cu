// Encoder stem, second layer: conv2 (kernel 3, stride 2, padding 1) + GELU, the move to [batch, frames, channels] and the
// positional embedding of Whisper's AudioEncoder (reference whisper/model.py:180, :194-197):
//
//   out[b, t, n] = gelu(bias[n] + sum_{c, k} W[n, c, k] h[b, c, 2 t + k - 1]) (+ pos[t, n]),   h[.., -1] = h[.., T] = 0
//
// as a GEMM on the tcgen05 tensor cores, D[128 channels, 256 frames] += W_k[128, 64 c] H_k[64 c, 256 frames] over the three
// taps and n_state / 64 channel chunks (K = 3 n_state), kind::f16: the operands are IEEE half - the same 11-bit
// significand TF32 has, i.e. the operand precision of torch's own convolution on this GPU (allow_tf32, cudnn's default)
// and of the reference's fp16 inference (transcribe.py:127) - with float32 accumulation in tensor memory.
//
// The first layer (stem_conv.cu, frame-major variant) leaves h as half [batch, frames, channels]: channels contiguous =
// K-major, so a tap's operand is ONE TMA tensor copy.  Viewed as [batch, frames / 2, parity, channels], tap k of output
// frame t is row (t - 1, odd), (t, even), (t, odd): the stride-2 convolution needs no im2col and no strided descriptor, and
// the zero padding in front of the clip is the copy's out-of-bounds fill.  The weights come as half [3, n_state, n_state]
// (tap, out channel, in channel: packed once by the host mirror), also one tensor copy per stage.
//
// One persistent CTA per SM, 18 warps: 16 epilogue (thread = channel; four warps per tensor-memory lane quadrant, 64 frames
// each: bias, GELU, positional embedding, staged pieces out by TMA tensor store - a frame's 32 channels are one line), 1 MMA issue, 1 TMA
// producer.  Operands in two rings (128-byte swizzle; see kC2AStages), two accumulators of 256 columns: the epilogue of a
// tile overlaps the MMAs of the next.  CTAs with the other 128-channel slices walk the same tiles at the same
// time, so h comes from HBM once.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace b200mel {

namespace {

constexpr int kC2M = 128;                        // out channels per CTA = MMA M = TMEM lanes
constexpr int kC2N = 256;                        // output frames per tile = MMA N
constexpr int kC2K = 64;                         // in channels per stage: 128 bytes of half = one swizzle row
// Two rings.  Weights: five slots of 16 KB, one (tap, 64-channel chunk) box each.  Input frames: three slots of 33 KB, in turn
// the ODD frames of a chunk - rows t0 - 1 .. t0 + 262, read by tap 0 from row 0 and by tap 2 from row 1 (a descriptor that
// starts one 128-byte row into the buffer: the swizzle is a function of the shared-memory address, so the rows the tensor
// copy wrote are the rows the MMA reads) - and its EVEN frames, rows t0 .. t0 + 255, read by tap 1.  The odd frames come
// from L2 once for both taps: 113 KB per chunk instead of 144 through L2 and into shared memory.
constexpr int kC2AStages = 5, kC2BStages = 3;
constexpr int kC2ABytes = kC2M * 128, kC2BBytes = kC2N * 128;        // 16 KB, 32 KB
constexpr int kC2ExtraRows = 8, kC2BSlotBytes = kC2BBytes + kC2ExtraRows * 128;   // 33 KB
constexpr int kC2BOffset = kC2AStages * kC2ABytes;
constexpr int kC2OutOffset = kC2BOffset + kC2BStages * kC2BSlotBytes;             // the epilogue warps' staging pieces: [16 frames][32 channels] float
constexpr int kC2OutPieceBytes = 16 * 128;
constexpr int kC2Smem = kC2OutOffset + 16 * kC2OutPieceBytes + 1024;              // + slack to align the base
static_assert(kC2BSlotBytes % 1024 == 0 && kC2BOffset % 1024 == 0, "128-byte swizzle: operands 1024-byte aligned");
constexpr int kC2EpiWarps = 16, kC2WarpMma = 16, kC2WarpTma = 17, kC2Threads = 18 * 32;
constexpr int kC2PartCols = kC2N / (kC2EpiWarps / 4), kC2Piece = 16;   // frames of the tile per epilogue warp (64), per pull (16)
constexpr uint32_t kC2Idesc = (1u << 4) | (static_cast<uint32_t>(kC2N >> 3) << 17) | (static_cast<uint32_t>(kC2M >> 4) << 24);   // f16 x f16 -> f32, K-major
static_assert(kC2Smem <= 227 * 1024, "shared memory layout");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// (a protocol bug ends the kernel with garbage instead of hanging the device)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned sleep_ns) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done || ++spins > (1u << 22)) break;
        if (sleep_ns) __nanosleep(sleep_ns);
    }
}
// K-major operand with 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr, uint32_t base_offset = 0) {
    return (uint64_t{2} << 61) | (static_cast<uint64_t>(base_offset) << 49) | (uint64_t{1} << 46) | (uint64_t{1024 >> 4} << 32) | (uint64_t{1} << 16) |
           ((smem_addr & 0x3ffffu) >> 4);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, bool accumulate) {
    asm volatile("{\n.reg .pred Q;\nsetp.ne.u32 Q, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, Q;\n}\n"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kC2Idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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

// GELU(v) = v Phi(v) = h + |h| erf(|h| sqrt 2), h = v / 2: the degree-6 fit of stem_conv.cu (tools/fit_gelu.py,
// |GELU error| <= 8.2e-8), two values at once as packed FMUL2 / FFMA2 / FADD2
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

struct C2Barriers { uint64_t a_full[kC2AStages], a_empty[kC2AStages], b_full[kC2BStages], b_empty[kC2BStages], d_full[2], d_empty[2]; };

struct C2Args {
    const float* bias;     // [n_state]
    const float* pos;      // [frames_out, n_state] or nullptr
    void* out;             // [batch, frames_out, n_state], float or half (the kernel's OutT)
    int64_t batch;
    int frames_out, n_state;
    int out_tma;           // out_map describes `out` as [batch * frames_out rows, n_state]: whole pieces leave by TMA tensor store
    int debug;             // measurement switches (switches build only, B200MEL_C2_FLAGS): 1 no GELU / stores, 2 no MMAs, 4 no input copies
};
#if defined(B200MEL_TC_TRACE) || defined(B200MEL_TC_SWITCHES)
#define C2_DEBUG(bit) ((a.debug & (bit)) != 0)
#else
#define C2_DEBUG(bit) false
#endif

template <typename OutT> __device__ __forceinline__ OutT c2_out(float v);
template <> __device__ __forceinline__ float c2_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half c2_out<__half>(float v) { return __float2half_rn(v); }

template <typename OutT>
__global__ void __launch_bounds__(kC2Threads, 1) stem_conv2_gelu_kernel(const C2Args a, const __grid_constant__ CUtensorMap w_map,
                                                                        const __grid_constant__ CUtensorMap h_map,
                                                                        const __grid_constant__ CUtensorMap h8_map,
                                                                        const __grid_constant__ CUtensorMap out_map) {
    extern __shared__ unsigned char smem_unaligned[];
    unsigned char* const smem_raw = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
    __shared__ __align__(8) C2Barriers bars;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int slices = a.n_state / kC2M;
    const int slice = blockIdx.x % slices;                       // this CTA's 128 out channels
    const int walkers = gridDim.x / slices;                      // CTAs that share the tiles of a slice
    const int walker = blockIdx.x / slices;
    const int tiles_per_clip = (a.frames_out + kC2N - 1) / kC2N;
    const int64_t total_tiles = a.batch * tiles_per_clip;
    const int my_tiles = walker < total_tiles ? static_cast<int>((total_tiles - walker + walkers - 1) / walkers) : 0;
    const int k_chunks = a.n_state / kC2K;                       // stages per tile = 3 k_chunks

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < kC2AStages; ++i) {
            mbar_init(&bars.a_full[i], 1);       // the producer's arrive.expect_tx
            mbar_init(&bars.a_empty[i], 1);      // tcgen05.commit
        }
        for (int i = 0; i < kC2BStages; ++i) {
            mbar_init(&bars.b_full[i], 1);
            mbar_init(&bars.b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars.d_full[i], 1);       // tcgen05.commit
            mbar_init(&bars.d_empty[i], kC2EpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == kC2WarpTma) {
        // ===== producer: per stage the weights' box (tap, slice, 64 in channels) and the tap's 256 rows of h =====
        if (lane == 0) {
            uint32_t a_parity = 1, b_parity = 1;                 // empty: the first use of a slot passes
            int a_slot = 0, b_slot = 0;
            for (int k = 0; k < my_tiles; ++k) {
                const int64_t tile = walker + static_cast<int64_t>(k) * walkers;
                const int clip = static_cast<int>(tile / tiles_per_clip);
                const int t0 = static_cast<int>(tile % tiles_per_clip) * kC2N;
                for (int j = 0; j < k_chunks; ++j) {
                    auto weights = [&](int tap) {
                        mbar_wait(&bars.a_empty[a_slot], a_parity, 32);
                        const uint32_t bar = smem_u32(&bars.a_full[a_slot]);
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kC2ABytes) : "memory");
                        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                     ::"r"(smem_u32(smem_raw + a_slot * kC2ABytes)), "l"(&w_map), "r"(j * kC2K), "r"(slice * kC2M), "r"(tap), "r"(bar) : "memory");
                        if (++a_slot == kC2AStages) { a_slot = 0; a_parity ^= 1u; }
                    };
                    // input frame 2 t + tap - 1 = (row t - 1, odd), (row t, even), (row t, odd); rows outside the clip are filled with zeros
                    auto frames = [&](int odd) {
                        mbar_wait(&bars.b_empty[b_slot], b_parity, 32);
                        unsigned char* dst = smem_raw + kC2BOffset + b_slot * kC2BSlotBytes;
                        const uint32_t bar = smem_u32(&bars.b_full[b_slot]);
                        if (C2_DEBUG(4)) {
                            mbar_arrive(&bars.b_full[b_slot]);
                            if (++b_slot == kC2BStages) { b_slot = 0; b_parity ^= 1u; }
                            return;
                        }
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(odd ? kC2BSlotBytes : kC2BBytes) : "memory");
                        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                     ::"r"(smem_u32(dst)), "l"(&h_map), "r"(j * kC2K), "r"(odd), "r"(t0 - odd), "r"(clip), "r"(bar) : "memory");
                        if (odd)                                     // the rows behind the 256th: tap 2 reads one row further
                            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                         ::"r"(smem_u32(dst + kC2BBytes)), "l"(&h8_map), "r"(j * kC2K), "r"(1), "r"(t0 - 1 + kC2N), "r"(clip), "r"(bar) : "memory");
                        if (++b_slot == kC2BStages) { b_slot = 0; b_parity ^= 1u; }
                    };
                    frames(1);
                    weights(0);
                    weights(2);
                    frames(0);
                    weights(1);
                }
            }
        }
    } else if (warp == kC2WarpMma) {
        // ===== MMA issue: 4 MMAs of K 16 per stage =====
        if (lane == 0) {
            uint32_t a_parity = 0, b_parity = 0, d_parity = 1;   // d_empty: the first two waits pass
            int a_slot = 0, b_slot = 0, buf = 0;
            const uint32_t base = smem_u32(smem_raw);
            for (int k = 0; k < my_tiles; ++k) {
                mbar_wait(&bars.d_empty[buf], d_parity, 32);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + buf * kC2N;
                for (int j = 0; j < k_chunks; ++j) {
                    // one tap: wait for its weights, 4 MMAs of K 16 (32 bytes further inside the 128-byte row each), release the weights
                    auto tap = [&](uint64_t b_desc, bool first) {
                        mbar_wait(&bars.a_full[a_slot], a_parity, 0);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t a_desc = sw128_desc(base + a_slot * kC2ABytes);
#pragma unroll
                        for (int i = 0; i < kC2K / 16; ++i)
                            if (!C2_DEBUG(2)) mma_f16_ss(d_tmem, a_desc + 2 * i, b_desc + 2 * i, !(first && i == 0));
                        mma_commit(&bars.a_empty[a_slot]);
                        if (++a_slot == kC2AStages) { a_slot = 0; a_parity ^= 1u; }
                    };
                    auto frames_ready = [&]() {
                        mbar_wait(&bars.b_full[b_slot], b_parity, 0);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        return base + kC2BOffset + b_slot * kC2BSlotBytes;
                    };
                    auto frames_done = [&]() {
                        mma_commit(&bars.b_empty[b_slot]);
                        if (++b_slot == kC2BStages) { b_slot = 0; b_parity ^= 1u; }
                    };
                    const uint32_t odd = frames_ready();
                    tap(sw128_desc(odd), j == 0);                                      // tap 0: rows t - 1
                    tap(sw128_desc(odd + 128), false);                            // tap 2: rows t, the same buffer one row in
                    frames_done();
                    const uint32_t even = frames_ready();
                    tap(sw128_desc(even), false);                                      // tap 1
                    frames_done();
                }
                mma_commit(&bars.d_full[buf]);
                buf ^= 1;
                if (buf == 0) d_parity ^= 1u;
            }
        }
    } else {
        // ===== epilogue: thread = channel; warp w takes lane quadrant w % 4 and frames 64 (w / 4) .. + 63 of the tile, in pieces
        // of 16.  Four warps per scheduler: the GELU is a chain of dependent packed FMAs, two warps left 60 % of the issue slots idle
        // and the epilogue, not the tensor cores, set the pace. =====
        const int quadrant = warp & 3, part = warp >> 2;
        const int n = slice * kC2M + quadrant * 32 + lane;
        const float half_bias = 0.5f * __ldg(a.bias + n);
        constexpr int kPieces = kC2PartCols / kC2Piece;
        const uint32_t staging = smem_u32(smem_raw + kC2OutOffset + warp * kC2OutPieceBytes);
        uint32_t parities = 0;                                   // bit b: the parity of d_full[b] to wait for
        for (int k = 0; k < my_tiles; ++k) {
            const int buf = k & 1;
            const int64_t tile = walker + static_cast<int64_t>(k) * walkers;
            const int clip = static_cast<int>(tile / tiles_per_clip);
            const int t0 = static_cast<int>(tile % tiles_per_clip) * kC2N + part * kC2PartCols;
            const uint32_t d_addr = tmem + (static_cast<uint32_t>(quadrant * 32) << 16) + buf * kC2N + part * kC2PartCols;
            OutT* out = static_cast<OutT*>(a.out) + ((static_cast<int64_t>(clip) * a.frames_out + t0) * a.n_state + n);
            const float* pos = a.pos != nullptr ? a.pos + (static_cast<int64_t>(t0) * a.n_state + n) : nullptr;
            const int last = a.frames_out - 1 - t0;              // last frame of the clip, relative to this warp's first
            // a piece's 16 rows of the positional embedding: independent loads, asked for one piece ahead (the first
            // piece's before the wait for the accumulator), so their latency is never on the epilogue's chain
            float p[kC2Piece];
#pragma unroll
            for (int i = 0; i < kC2Piece; ++i) p[i] = 0.f;
            auto load_pos = [&](int piece) {
                if (pos != nullptr && piece * kC2Piece <= last) {
                    const float* q = pos + static_cast<int64_t>(piece * kC2Piece) * a.n_state;
                    if (piece * kC2Piece + kC2Piece - 1 <= last) {
#pragma unroll
                        for (int i = 0; i < kC2Piece; ++i) p[i] = __ldg(q + static_cast<int64_t>(i) * a.n_state);
                    } else {
#pragma unroll
                        for (int i = 0; i < kC2Piece; ++i) p[i] = __ldg(q + static_cast<int64_t>(min(i, last - piece * kC2Piece)) * a.n_state);
                    }
                }
            };
            load_pos(0);
            mbar_wait(&bars.d_full[buf], (parities >> buf) & 1u, 100);
            parities ^= 1u << buf;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int piece = 0; piece < kPieces; ++piece) {
                if (piece * kC2Piece > last) {                   // (warp-uniform) nothing left of the clip: only the hand-over
                    if (piece == kPieces - 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.d_empty[buf]);
                    }
                    continue;
                }
                float d[kC2Piece];
                tmem_ld16(d_addr + piece * kC2Piece, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (piece == kPieces - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.d_empty[buf]);   // the warp's part of the accumulator is in registers
                }
                if (C2_DEBUG(1)) continue;
#pragma unroll
                for (int i = 0; i < kC2Piece; i += 2) {
                    const float2 h = __ffma2_rn(make_float2(d[i], d[i + 1]), make_float2(0.5f, 0.5f), make_float2(half_bias, half_bias));
                    const float2 g = gelu2_from_half(h);
                    d[i] = g.x + p[i];
                    d[i + 1] = g.y + p[i + 1];
                }
                if (piece + 1 < kPieces) load_pos(piece + 1);
                OutT* o = out + static_cast<int64_t>(piece * kC2Piece) * a.n_state;
                if (piece * kC2Piece + kC2Piece - 1 <= last && a.out_tma) {
                    // a whole piece inside the clip: staged as [16 frames][32 channels] (a row = 128 / 64 contiguous bytes of the result)
                    // and out with one TMA tensor store - no L1 tag traffic, one instruction instead of sixteen stores
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous piece has been read
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < kC2Piece; ++i) {
                        if constexpr (sizeof(OutT) == 4)
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(staging + i * 128 + lane * 4), "f"(d[i]) : "memory");
                        else
                            asm volatile("st.shared.b16 [%0], %1;" ::"r"(staging + i * 64 + lane * 2), "h"(__half_as_ushort(__float2half_rn(d[i]))) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                                     ::"l"(&out_map), "r"(slice * kC2M + quadrant * 32), "r"(clip * a.frames_out + t0 + piece * kC2Piece), "r"(staging) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else if (piece * kC2Piece + kC2Piece - 1 <= last) {
#pragma unroll
                    for (int i = 0; i < kC2Piece; ++i) o[static_cast<int64_t>(i) * a.n_state] = c2_out<OutT>(d[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < kC2Piece; ++i)
                        if (piece * kC2Piece + i <= last) o[static_cast<int64_t>(i) * a.n_state] = c2_out<OutT>(d[i]);
                }
            }
        }
    }
    if (warp < kC2EpiWarps && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the staging pieces live until the stores have read them
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn tensor_map_encoder() {
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        return reinterpret_cast<EncodeFn>(fn);
    }();
    return encode;
}

}  // namespace

cudaError_t launch_stem_conv2_gelu(const void* h_fm16, int64_t batch, int frames_padded, const void* weight_f16, const float* bias, const float* pos,
                                   int n_state, void* out, int out_f16, cudaStream_t stream) {
    const int frames_out = frames_padded / 2;
    if (batch <= 0 || frames_out <= 0) return cudaSuccess;
    constexpr int kMaxDevices = 64;
    static int sms_by_device[kMaxDevices] = {0};                      // (also: the kernel's attributes are set on this device)
    int device = 0;
    cudaError_t err = cudaGetDevice(&device);
    if (err != cudaSuccess) return err;
    if (device < 0 || device >= kMaxDevices) return cudaErrorInvalidDevice;
    if (sms_by_device[device] == 0) {
        err = cudaFuncSetAttribute(stem_conv2_gelu_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC2Smem);
        if (err == cudaSuccess) err = cudaFuncSetAttribute(stem_conv2_gelu_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, kC2Smem);
        int count = 0;
        if (err == cudaSuccess) err = cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device);
        if (err != cudaSuccess) return err;
        sms_by_device[device] = count;
    }
    const EncodeFn encode = tensor_map_encoder();
    if (encode == nullptr) return cudaErrorNotSupported;
    const uint64_t c = static_cast<uint64_t>(n_state);
    CUtensorMap w_map, h_map, h8_map, out_map;
    std::memset(&out_map, 0, sizeof(out_map));
    int out_tma = 0;
    std::memset(&w_map, 0, sizeof(w_map));
    std::memset(&h_map, 0, sizeof(h_map));
    std::memset(&h8_map, 0, sizeof(h8_map));
    {   // weights: half [3 taps][n_state out][n_state in]; box = 64 in channels x 128 out channels of one tap
        const cuuint64_t dims[3] = {c, c, 3};
        const cuuint64_t strides[2] = {c * 2, c * c * 2};
        const cuuint32_t box[3] = {kC2K, kC2M, 1};
        const cuuint32_t elem[3] = {1, 1, 1};
        if (encode(&w_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(weight_f16), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    {   // h: half [batch][frames_padded / 2][parity][n_state]; box = 64 channels x 256 rows of one parity
        const cuuint64_t dims[4] = {c, 2, static_cast<cuuint64_t>(frames_out), static_cast<cuuint64_t>(batch)};
        const cuuint64_t strides[3] = {c * 2, c * 4, static_cast<cuuint64_t>(frames_padded) * c * 2};
        const cuuint32_t box[4] = {kC2K, 1, kC2N, 1};
        const cuuint32_t elem[4] = {1, 1, 1, 1};
        if (encode(&h_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(h_fm16), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
        const cuuint32_t box8[4] = {kC2K, 1, kC2ExtraRows, 1};   // the odd frames behind the 256th row
        if (encode(&h8_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(h_fm16), dims, strides, box8, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return cudaErrorInvalidValue;
    }
    if (reinterpret_cast<uintptr_t>(out) % 16 == 0 && batch * frames_out < (int64_t{1} << 31)) {
        // the result as the TMA unit sees it: [batch * frames_out rows, n_state] float / half, boxes of 16 frames x 32 channels
        const cuuint64_t dims[2] = {c, static_cast<cuuint64_t>(batch * frames_out)};
        const cuuint64_t strides[1] = {c * (out_f16 ? 2 : 4)};
        const cuuint32_t box[2] = {32, kC2Piece};
        const cuuint32_t elem[2] = {1, 1};
        if (encode(&out_map, out_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
            out_tma = 1;
    }
    const int sms = sms_by_device[device];
    const int slices = n_state / kC2M;
    const int64_t tiles = batch * ((frames_out + kC2N - 1) / kC2N);
    int64_t walkers = sms / slices;                                   // CTAs per slice; every CTA of the grid is resident
    if (walkers < 1) walkers = 1;
    if (walkers > tiles) walkers = tiles;
    C2Args a{bias, pos, out, batch, frames_out, n_state, out_tma, 0};
#if defined(B200MEL_TC_TRACE) || defined(B200MEL_TC_SWITCHES)
    if (std::getenv("B200MEL_C2_FLAGS") != nullptr) a.debug = std::atoi(std::getenv("B200MEL_C2_FLAGS"));
#endif
    ProfileScope profile(3, stream);
    if (out_f16) stem_conv2_gelu_kernel<__half><<<static_cast<unsigned>(walkers * slices), kC2Threads, kC2Smem, stream>>>(a, w_map, h_map, h8_map, out_map);
    else stem_conv2_gelu_kernel<float><<<static_cast<unsigned>(walkers * slices), kC2Threads, kC2Smem, stream>>>(a, w_map, h_map, h8_map, out_map);
    count_launch();
    return cudaGetLastError();
}

}  // namespace b200mel
