// Math core of the fused log-mel front-end (FFT variant), shared by the sm_100a
// kernel (logmel_fft.cu) and the CPU choreography emulator used by the CPU test
// suite (tests/emul/emul_fft.cpp).  Everything here is __host__ __device__ so the
// exact index maps and butterflies the GPU runs can be checked without a GPU.
//
// What it computes, per frame (reference whisper/audio.py:147-154):
//   X_t[k] = sum_n hann[n] * p[160 t + n] * exp(-2 pi i k n / 400),  k = 0..200
//   P_t[k] = |X_t[k]|^2 ;  S[m,t] = sum_k F[m,k] P_t[k] ;  log10(max(S, 1e-10))
//
// Two real frames (a, b) ride one complex 400-point FFT of z = a + i b:
//   Z = A + i B  =>  2 A[k] = Z[k] + conj Z[400-k],  2 i B[k] = Z[k] - conj Z[400-k].
// The 1/2 is folded into the window table, so |A|^2 and |B|^2 come out directly.
// The 400-point FFT is Cooley-Tukey 20 x 20 (n = 20 n1 + n2, k = k1 + 20 k2); each
// 20-point DFT is a Good-Thomas 4 x 5 prime-factor butterfly held in registers.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD inline
#if defined(B200_HOST_HAS_CUDA_HEADERS)  // host build that also pulls in CUDA's own vector types
#include <vector_functions.h>
#include <vector_types.h>
#else
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
#endif
#endif

namespace b200mel {

constexpr int kNFFT = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;          // rfft bins; bins 0 and 200 carry zero mel weight
constexpr int kUsedBins = 200;      // bins 0..199 are materialised
constexpr int kRadix = 20;          // 400 = 20 x 20
constexpr int kHalfWin = 200;       // reflect pad on each side (center=True)

// ---- tile geometry of the FFT-variant kernel --------------------------------
constexpr int kTileFrames = 32;                       // frames per CTA pass
constexpr int kGroups = kTileFrames / 2;              // frame pairs per pass
constexpr int kThreads = kGroups * kRadix;            // 320 threads
constexpr int kAudioTile = kHop * kTileFrames + (kNFFT - kHop);  // 5360 samples staged
constexpr int kYStride = 21;                          // padded row of the 20x20 transpose (float2)
constexpr int kGroupStride = kRadix * kYStride;       // 420 float2 of scratch per frame pair
constexpr int kOutStride = kTileFrames + 1;           // padded row of the output staging tile
constexpr int kMaxMelWeights = 512;                   // packed non-zero filter weights (<= 394 used)
constexpr int kMaxMels = 128;
constexpr int kPStride = 34;                          // padded row of the power tile P[bin][frame]
constexpr int kSegBins = 20;                          // bins swept by one warp (10 warps x 20 = 200 bins)
constexpr int kSegments = kUsedBins / kSegBins;       // == warps per CTA
constexpr int kSStride = kTileFrames + 1;             // padded row of the partial-sum tile S[row][frame]
constexpr int kMaxSRows = kMaxMels + 32;              // mel rows + second parts of segment-straddling mels
constexpr int kWarpsPerCta = kThreads / 32;
static_assert(kSegments == kWarpsPerCta, "one 20-bin segment per warp");

// One bin of the warp-uniform mel sweep.  w0/w1: weight of the active even/odd mel at this
// bin (0 if none).  emit0/emit1: byte offset of the S row that receives slot 0/1 after this
// bin, negative = keep accumulating.
struct MelSweepEntry { float w0, w1; int emit0, emit1; };
B200_HD int s_row_offset(int row) { return row * kSStride * static_cast<int>(sizeof(float)); }

// Packed per-mel band descriptor: first bin | count << 8 | weight offset << 16.
B200_HD int mel_band_pack(int first, int count, int offset) { return first | (count << 8) | (offset << 16); }
B200_HD int mel_band_first(int p) { return p & 0xff; }
B200_HD int mel_band_count(int p) { return (p >> 8) & 0xff; }
B200_HD int mel_band_offset(int p) { return (p >> 16) & 0xffff; }

// ---- complex helpers ---------------------------------------------------------
// A complex number is a float2 (re, im) living in an aligned register pair, so complex
// add / subtract / scale map onto Blackwell's packed fp32 instructions (FADD2 / FFMA2 / FMUL2,
// PTX add/fma/mul.f32x2, sm_100+): one issue slot per complex operation instead of two.
// The host build (CPU emulator) uses the scalar equivalents; both round to nearest, and the
// subtraction is an FMA with -1, so the two builds compute the same values.
B200_HD float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    return __fadd2_rn(a, b);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
B200_HD float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a * (s, s): complex times real
B200_HD float2 cscale(float2 a, float s) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    return __fmul2_rn(a, make_float2(s, s));
#else
    return make_float2(a.x * s, a.y * s);
#endif
}
// a * (s, s) + c
B200_HD float2 cfma(float2 a, float s, float2 c) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
    return __ffma2_rn(a, make_float2(s, s), c);
#else
    return make_float2(a.x * s + c.x, a.y * s + c.y);
#endif
}
B200_HD float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// 5-point forward DFT, in place on u[0..4].
B200_HD void dft5(float2 (&u)[5]) {
    const float c1 = 0.30901699437494742f;   // cos(2 pi / 5)
    const float c2 = -0.80901699437494742f;  // cos(4 pi / 5)
    const float s1 = 0.95105651629515357f;   // sin(2 pi / 5)
    const float s2 = 0.58778525229247313f;   // sin(4 pi / 5)
    const float2 a1 = cadd(u[1], u[4]), b1 = csub(u[1], u[4]);
    const float2 a2 = cadd(u[2], u[3]), b2 = csub(u[2], u[3]);
    const float2 p1 = cfma(a2, c2, cfma(a1, c1, u[0]));
    const float2 p2 = cfma(a2, c1, cfma(a1, c2, u[0]));
    const float2 q1 = cfma(b2, s2, cscale(b1, s1));
    const float2 q2 = cfma(b2, -s1, cscale(b1, s2));
    u[0] = cadd(cadd(u[0], a1), a2);
    // X1 = p1 - i q1, X4 = p1 + i q1, X2 = p2 - i q2, X3 = p2 + i q2
    u[1] = make_float2(p1.x + q1.y, p1.y - q1.x);
    u[4] = make_float2(p1.x - q1.y, p1.y + q1.x);
    u[2] = make_float2(p2.x + q2.y, p2.y - q2.x);
    u[3] = make_float2(p2.x - q2.y, p2.y + q2.x);
}

// 20-point forward DFT: natural-order in, natural-order out, all indices static.
// Good-Thomas maps: n = (5 n1 + 4 n2) mod 20, k = (5 k1 + 16 k2) mod 20.
B200_HD void dft20(const float2 (&x)[kRadix], float2 (&X)[kRadix]) {
    float2 a[4][5];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2) {
        const float2 u0 = x[(4 * n2) % 20], u1 = x[(5 + 4 * n2) % 20];
        const float2 u2 = x[(10 + 4 * n2) % 20], u3 = x[(15 + 4 * n2) % 20];
        const float2 t0 = cadd(u0, u2), t1 = csub(u0, u2), t2 = cadd(u1, u3), t3 = csub(u1, u3);
        a[0][n2] = cadd(t0, t2);
        a[2][n2] = csub(t0, t2);
        a[1][n2] = make_float2(t1.x + t3.y, t1.y - t3.x);  // t1 - i t3
        a[3][n2] = make_float2(t1.x - t3.y, t1.y + t3.x);  // t1 + i t3
    }
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        dft5(a[k1]);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) X[(5 * k1 + 16 * k2) % 20] = a[k1][k2];
    }
}

// ---- reflect-padded, zero-extended sample fetch ---------------------------------
// s: position in the centre-padded signal minus 200, i.e. an index into the
// zero-extended waveform x' of length total = n_samples + right_zero_pad
// (audio.py:145-146), reflected at both ends as torch.stft(center=True) does.
// valid: samples actually present in memory (<= n_samples; `lengths` fast path).
B200_HD int64_t reflect_source_index(int64_t s, int64_t total) {
    if (s < 0) s = -s;
    if (s >= total) s = 2 * (total - 1) - s;
    return s;
}

// ---- max-reduction key: order-preserving float -> uint32, NaN sorts highest ------
// (torch.max propagates NaN, audio.py:155, so a NaN anywhere must win the reduction)
B200_HD uint32_t float_bits(float v) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(v);
#else
    union { float f; uint32_t u; } c; c.f = v; return c.u;
#endif
}
B200_HD float bits_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
B200_HD uint32_t max_key_encode(float v) {
    if (v != v) return 0xffffffffu;
    const uint32_t b = float_bits(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
B200_HD float max_key_decode(uint32_t k) {
    if (k == 0xffffffffu) return bits_float(0x7fc00000u);
    return bits_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// max that propagates NaN (PTX max.NaN.f32), like torch.max / torch.maximum.
B200_HD float max_nan(float a, float b) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
#else
    return (a != a || b != b) ? bits_float(0x7fc00000u) : (a > b ? a : b);
#endif
}

// log10(max(S, 1e-10)) of audio.py:154; NaN passes through like torch.clamp.
B200_HD float log10_clamped(float s) {
    s = max_nan(s, 1e-10f);
#if defined(__CUDA_ARCH__)
    float l2;  // s >= 1e-10 is a normal number: the flush-to-zero form needs no denormal fix-up
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(s));
    return l2 * 0.30102999566398120f;
#else
    return __builtin_log2f(s) * 0.30102999566398120f;
#endif
}

// log2(max(S, 1e-10)): log10_clamped before its final scaling (for callers that scale two values with one FMUL2).
B200_HD float log2_clamped(float s) {
    s = max_nan(s, 1e-10f);
#if defined(__CUDA_ARCH__)
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(s));
    return l2;
#else
    return __builtin_log2f(s);
#endif
}

// Final dynamic-range clamp and affine map of audio.py:155-156.
B200_HD float normalise(float lg, float gmax) {
    if (gmax != gmax) return gmax;  // NaN max poisons the whole call, as in torch
    const float floor_v = gmax - 8.0f;
    const float v = (lg < floor_v) ? floor_v : lg;
    return (v + 4.0f) * 0.25f;
}

// ---- the phases of one 32-frame pass, one call per thread -------------------------
// tid in [0, 320): group g = tid / 20 owns local frames (2g, 2g+1); j = tid % 20.

// Phase 1: window, first 20-point DFT over n1 (samples 20 n1 + j), twiddle, transpose.
//   s_audio: kAudioTile samples (float, or raw int16 PCM), s_audio[i] = padded sample at tile origin + i
//   win_half: this thread's 20 window taps, 0.5 * hann[20 n1 + j] (times 1/32768 for int16 PCM)
//   s_tw: [j][k1] float2 twiddles exp(-2 pi i j k1 / 400)
//   s_work: per-group scratch, receives Y[k1][j]
template <typename InT>
B200_HD void phase_fft_first(int tid, const InT* s_audio, const float (&win_half)[kRadix],
                             const float2* s_tw, float2* s_work) {
    const int g = tid / kRadix, j = tid % kRadix;
    const InT* fa = s_audio + (2 * g) * kHop + j;
    float2 x[kRadix], y[kRadix];
#pragma unroll
    for (int n1 = 0; n1 < kRadix; ++n1)
        x[n1] = make_float2(static_cast<float>(fa[kRadix * n1]) * win_half[n1],
                            static_cast<float>(fa[kRadix * n1 + kHop]) * win_half[n1]);
    dft20(x, y);
    float2* dst = s_work + g * kGroupStride + j;
    dst[0] = y[0];
    // the twiddle table is symmetric (W^(j k1)); index it [k1][j] so a warp reads consecutive words
#pragma unroll
    for (int k1 = 1; k1 < kRadix; ++k1) dst[k1 * kYStride] = cmul(y[k1], s_tw[k1 * kRadix + j]);
}

// Phase 2a: thread k1 = j gathers row k1 of the transpose into registers.
B200_HD void phase_fft_second_load(int tid, const float2* s_work, float2 (&r)[kRadix]) {
    const int g = tid / kRadix, j = tid % kRadix;
    const float2* src = s_work + g * kGroupStride + j * kYStride;
#pragma unroll
    for (int n2 = 0; n2 < kRadix; ++n2) r[n2] = src[n2];
}

// Phase 2b: second 20-point DFT over n2; Z[k1 + 20 k2] stored linearly (400 float2 per group).
B200_HD void phase_fft_second_store(int tid, const float2 (&r)[kRadix], float2* s_work) {
    const int g = tid / kRadix, j = tid % kRadix;
    float2 z[kRadix];
    dft20(r, z);
    float2* dst = s_work + g * kGroupStride + j;
#pragma unroll
    for (int k2 = 0; k2 < kRadix; ++k2) dst[kRadix * k2] = z[k2];
}

// Phase 3a: thread j gathers Z[k], Z[400-k] for its bins k = j + 20 i, i = 0..9.
B200_HD void phase_power_load(int tid, const float2* s_work, float2 (&r)[kRadix]) {
    const int g = tid / kRadix, j = tid % kRadix;
    const float2* z = s_work + g * kGroupStride;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const int k = j + kRadix * i;
        r[2 * i] = z[k];
        r[2 * i + 1] = z[(kNFFT - k) % kNFFT];
    }
}

// Phase 3b: split the packed spectrum and store both frames' power into the CTA-wide
// tile P[bin][frame] (bins 0..199, frames 2g and 2g+1 side by side).
B200_HD void phase_power_store(int tid, const float2 (&r)[kRadix], float* s_P) {
    const int g = tid / kRadix, j = tid % kRadix;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        // A[k] = Z[k] + conj Z[400-k] = (u.x, v.y);  i B[k] = Z[k] - conj Z[400-k] = (v.x, u.y)
        // (the window table carries the 1/2), so |A|^2 = u.x^2 + v.y^2 and |B|^2 = v.x^2 + u.y^2
        const float2 u = cadd(r[2 * i], r[2 * i + 1]), v = csub(r[2 * i], r[2 * i + 1]);
        const int k = j + kRadix * i;
        *reinterpret_cast<float2*>(s_P + k * kPStride + 2 * g) =
            make_float2(u.x * u.x + v.y * v.y, v.x * v.x + u.y * u.y);
    }
}

// Phase 4: mel projection as a warp-uniform sweep.  Warp q owns bins [20q, 20q+20), lane f
// owns frame f, so every lane executes the same program (weights and emit rows are
// broadcast loads) and the power tile is read along its contiguous frame axis.
B200_HD void phase_mel_sweep(int tid, const float* s_P, const MelSweepEntry* s_sweep, float* s_S) {
    const int q = tid >> 5, f = tid & 31;
    const MelSweepEntry* e = s_sweep + q * kSegBins;
    const float* p = s_P + q * kSegBins * kPStride + f;
    char* s_lane = reinterpret_cast<char*>(s_S + f);
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int b = 0; b < kSegBins; ++b) {
        const MelSweepEntry en = e[b];
        const float pv = p[b * kPStride];
        acc0 += en.w0 * pv;
        acc1 += en.w1 * pv;
        if (en.emit0 >= 0) { *reinterpret_cast<float*>(s_lane + en.emit0) = acc0; acc0 = 0.f; }
        if (en.emit1 >= 0) { *reinterpret_cast<float*>(s_lane + en.emit1) = acc1; acc1 = 0.f; }
    }
}

// Phase 5: join the partial sums, log10 with the 1e-10 clamp, write the [n_mels, 32] tile
// (warp = mel row, lane = frame: one 128-byte row segment per store) and return the max key
// of this thread's valid values.  dst points at out[clip][0][t0]; row_pitch = n_frames.
// s_row_off[m] / s_row_off[kMaxMels + m]: byte offsets of the (up to) two partial-sum rows of
// mel m; a missing part points at the all-zero row.
B200_HD uint32_t phase_finish(int tid, int n_mels, const float* s_S, const int* s_row_off,
                              int frames_valid, float* dst, int64_t row_pitch) {
    const int w = tid >> 5, f = tid & 31;
    if (f >= frames_valid) return 0u;
    const char* s_lane = reinterpret_cast<const char*>(s_S + f);
    float* out = dst + static_cast<int64_t>(w) * row_pitch + f;
    const int64_t step = row_pitch * kWarpsPerCta;
    float mx = bits_float(0xff800000u);  // -inf
    for (int m = w; m < n_mels; m += kWarpsPerCta) {
        const float s = *reinterpret_cast<const float*>(s_lane + s_row_off[m]) +
                        *reinterpret_cast<const float*>(s_lane + s_row_off[kMaxMels + m]);
        const float lg = log10_clamped(s);
        *out = lg;
        out += step;
        mx = max_nan(mx, lg);
    }
    return max_key_encode(mx);
}

}  // namespace b200mel
