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
struct float2 { float x, y; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
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
// bin (0 if none).  emit: low/high 16 bits = S row that receives slot 0/1 after this bin,
// 0xffff = keep accumulating.
struct MelSweepEntry { float w0, w1; unsigned emit; int pad; };

// Packed per-mel band descriptor: first bin | count << 8 | weight offset << 16.
B200_HD int mel_band_pack(int first, int count, int offset) { return first | (count << 8) | (offset << 16); }
B200_HD int mel_band_first(int p) { return p & 0xff; }
B200_HD int mel_band_count(int p) { return (p >> 8) & 0xff; }
B200_HD int mel_band_offset(int p) { return (p >> 16) & 0xffff; }

// ---- complex helpers ---------------------------------------------------------
B200_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
B200_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
B200_HD float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// 5-point forward DFT, in place on u[0..4].
B200_HD void dft5(float2 (&u)[5]) {
    const float c1 = 0.30901699437494742f;   // cos(2 pi / 5)
    const float c2 = -0.80901699437494742f;  // cos(4 pi / 5)
    const float s1 = 0.95105651629515357f;   // sin(2 pi / 5)
    const float s2 = 0.58778525229247313f;   // sin(4 pi / 5)
    const float2 a1 = cadd(u[1], u[4]), b1 = csub(u[1], u[4]);
    const float2 a2 = cadd(u[2], u[3]), b2 = csub(u[2], u[3]);
    const float2 p1 = make_float2(u[0].x + c1 * a1.x + c2 * a2.x, u[0].y + c1 * a1.y + c2 * a2.y);
    const float2 p2 = make_float2(u[0].x + c2 * a1.x + c1 * a2.x, u[0].y + c2 * a1.y + c1 * a2.y);
    const float2 q1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    const float2 q2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    u[0] = make_float2(u[0].x + a1.x + a2.x, u[0].y + a1.y + a2.y);
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

// log10(max(S, 1e-10)) of audio.py:154; NaN passes through like torch.clamp.
B200_HD float log10_clamped(float s) {
    s = (s < 1e-10f) ? 1e-10f : s;
#if defined(__CUDA_ARCH__)
    return __log2f(s) * 0.30102999566398120f;
#else
    return __builtin_log2f(s) * 0.30102999566398120f;
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
        const float2 zk = r[2 * i], zc = r[2 * i + 1];
        const float are = zk.x + zc.x, aim = zk.y - zc.y;   // 2A[k]/2 (window carries the 1/2)
        const float bre = zk.x - zc.x, bim = zk.y + zc.y;   // |2iB[k]/2| = |B[k]|
        const int k = j + kRadix * i;
        *reinterpret_cast<float2*>(s_P + k * kPStride + 2 * g) =
            make_float2(are * are + aim * aim, bre * bre + bim * bim);
    }
}

// Phase 4: mel projection as a warp-uniform sweep.  Warp q owns bins [20q, 20q+20), lane f
// owns frame f, so every lane executes the same program (weights and emit rows are
// broadcast loads) and the power tile is read along its contiguous frame axis.
B200_HD void phase_mel_sweep(int tid, const float* s_P, const MelSweepEntry* s_sweep, float* s_S) {
    const int q = tid >> 5, f = tid & 31;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
    for (int b = 0; b < kSegBins; ++b) {
        const int k = q * kSegBins + b;
        const MelSweepEntry e = s_sweep[k];
        const float p = s_P[k * kPStride + f];
        acc0 += e.w0 * p;
        acc1 += e.w1 * p;
        const unsigned e0 = e.emit & 0xffffu, e1 = e.emit >> 16;
        if (e0 != 0xffffu) { s_S[e0 * kSStride + f] = acc0; acc0 = 0.f; }
        if (e1 != 0xffffu) { s_S[e1 * kSStride + f] = acc1; acc1 = 0.f; }
    }
}

// Phase 5: join the partial sums, log10 with the 1e-10 clamp, write the [n_mels, 32] tile
// (warp = mel row, lane = frame: one 128-byte row segment per store) and return the max key
// of this thread's valid values.  dst points at out[clip][0][t0]; row_pitch = n_frames.
B200_HD uint32_t phase_finish(int tid, int n_mels, const float* s_S, const short* s_row_a, const short* s_row_b,
                              int frames_valid, float* dst, int64_t row_pitch) {
    const int w = tid >> 5, f = tid & 31;
    uint32_t key = 0u;
    if (f >= frames_valid) return key;
    for (int m = w; m < n_mels; m += kWarpsPerCta) {
        const int ra = s_row_a[m], rb = s_row_b[m];
        float s = 0.f;
        if (ra >= 0) s = s_S[ra * kSStride + f];
        if (rb >= 0) s += s_S[rb * kSStride + f];
        const float lg = log10_clamped(s);
        dst[static_cast<int64_t>(m) * row_pitch + f] = lg;
        const uint32_t k = max_key_encode(lg);
        key = k > key ? k : key;
    }
    return key;
}

}  // namespace b200mel
