// Math core of the tcgen05 (tensor-core) variant of the log-mel front-end, shared by the sm_100a
// kernel (logmel_tc.cu) and the CPU emulator (tests/emul/emul_tc.cpp).
//
// The 400-point real DFT of a frame is factored 16 x 25 (n = 25 n1 + n2, k = k1 + 16 k2):
//   stage 1, CUDA cores : Y[k1, n2] = sum_n1 w[25 n1 + n2] x[25 n1 + n2] W16^(n1 k1)      (real FFT-16, k1 = 0..8)
//                         Yt[b, n2] = Y[b, n2] * W400^(n2 b)                                (twiddle, b = 1..7)
//   stage 2, tcgen05    : X[b + 16 k2] = sum_n2 Yt[b, n2] W25^(n2 k2)                       (DFT-25 as a GEMM)
// Stage 2 is a dense contraction whose matrix is THE SAME for b = 1..7 (the twiddles were applied
// on the CUDA cores), so one 50 x 50 real matrix serves seven "blocks"; block 0 packs the two real
// rows (Y[0, n2], Y[8, n2]) and has its own matrix (bins 16 k2 and 8 + 16 k2).
// Precision: fp32 values are split x = hi + lo with hi, lo in fp16 (22 mantissa bits together) and
// the product is formed as hi*Bhi + lo*Bhi + hi*Blo with fp32 accumulation in tensor memory — the
// 3-product compensation of "3xTF32", at the f16 MMA rate and half the operand bytes.  Inputs are
// pre-scaled by 2^8 (through the window table) so the lo parts stay normal fp16 numbers; the mel
// weights carry the 2^-16 that undoes it (both exact).
#pragma once

#include <stdint.h>

#include <cuda_fp16.h>

#include "logmel_core.cuh"

namespace b200mel {

constexpr int kTcTileFrames = 128;                    // frames per tile = TMEM lanes = MMA M
constexpr int kTcRowPitch = kHop + 1;                 // audio tile rows of 160 samples stored at pitch 161
constexpr int kTcAudioSamples = kHop * kTcTileFrames + (kNFFT - kHop);          // 20720
constexpr int kTcAudioRows = (kTcAudioSamples + kHop - 1) / kHop;               // 130
constexpr int kTcAudioFloats = kTcAudioRows * kTcRowPitch;                      // 20930
constexpr int kTcBlocks = 8;                          // block 0: (Y0, Y8) pairs; blocks 1..7: twiddled Y[b]
constexpr int kTcN2 = 25;
constexpr int kTcBlockCols = 56;                      // TMEM columns per block: 25 hi | 25 lo | 6 zero pad
constexpr int kTcACols = kTcBlocks * kTcBlockCols;    // 448
constexpr int kTcDCols = 32;                          // accumulator columns of one unit (16 complex outputs)
constexpr int kTcDBase = kTcACols;                    // two accumulator buffers at columns 448 and 480
constexpr int kTcUnits = kTcBlocks * 2;               // (block, N-half) units per tile
constexpr int kTcKMain = 112;                         // K of the hi|lo pass: 50 + 50 (+12 zero rows) halves, 7 k-steps
constexpr int kTcKCorr = 64;                          // K of the hi * Blo pass: 50 (+14 zero rows) halves, 4 k-steps
constexpr int kTcN = 64;                              // N of a block: 50 real outputs padded to 2 x 32
constexpr float kTcInputScale = 256.0f;               // exact power of two, see header
constexpr float kTcPowerUnscale = 1.0f / 65536.0f;

// Operand matrices in the tcgen05 shared-memory layout (K-major, no swizzle: 8 x 16 B core matrices,
// strips [k / 8][n][8 halves]).  Two sets: [0] block 0, [1] blocks 1..7.
constexpr int kTcBMainHalves = kTcKMain * kTcN;       // 7168
constexpr int kTcBCorrHalves = kTcKCorr * kTcN;       // 4096

// One output of a unit as seen by an epilogue thread: weight and byte offset (into the S tile row of
// this thread) of the mel that this thread's parity owns at that bin; w == 0 when there is none.
struct TcTap { float w; int s_off; };

struct TcTables {
    float win[kTcN2][16];               // win[n2][n1] = 128 * hann[25 n1 + n2]  (FFT-16 below yields 2 X)
    float2 tw[kTcN2][8];                // tw[n2][b] = W400^(n2 b), b = 1..7 ([0] unused)
    __half b_main[2][kTcBMainHalves];   // [Bhi; Bhi; 0] in smem operand layout
    __half b_corr[2][kTcBCorrHalves];   // [Blo; 0]
    TcTap tap[2][kTcUnits][16];         // [mel parity][unit][complex output j]
    int n_mels;
};

// ---- real FFT-16, returns 2 * X[k] for k = 0..8 (X[0], X[8] real) ---------------------------------
// x[n1] are the 16 windowed samples.  Packed as z[m] = x[2m] + i x[2m+1], an 8-point complex FFT
// (radix-2, decimation in time) and the usual real-FFT split.
B200_HD void fft16_real_x2(const float (&x)[16], float2 (&X)[9]) {
    const float r = 0.70710678118654752f;
    float2 z[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) z[m] = make_float2(x[2 * m], x[2 * m + 1]);
    // 8-point FFT: bit-reversed pairing (0,4)(2,6)(1,5)(3,7)
    const float2 a0 = cadd(z[0], z[4]), a1 = csub(z[0], z[4]);
    const float2 a2 = cadd(z[2], z[6]), a3 = csub(z[2], z[6]);
    const float2 a4 = cadd(z[1], z[5]), a5 = csub(z[1], z[5]);
    const float2 a6 = cadd(z[3], z[7]), a7 = csub(z[3], z[7]);
    // 4-point combines: (a0,a1,a2,a3) -> even-index FFT4 E[0..3]; (a4..a7) -> odd-index FFT4 O[0..3]
    const float2 e0 = cadd(a0, a2), e2 = csub(a0, a2);
    const float2 e1 = make_float2(a1.x + a3.y, a1.y - a3.x);   // a1 - i a3
    const float2 e3 = make_float2(a1.x - a3.y, a1.y + a3.x);   // a1 + i a3
    const float2 o0 = cadd(a4, a6), o2 = csub(a4, a6);
    const float2 o1 = make_float2(a5.x + a7.y, a5.y - a7.x);
    const float2 o3 = make_float2(a5.x - a7.y, a5.y + a7.x);
    // Z[k] = E[k] + W8^k O[k], Z[k+4] = E[k] - W8^k O[k];  W8 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
    const float2 t1 = make_float2(r * (o1.x + o1.y), r * (o1.y - o1.x));
    const float2 t2 = make_float2(o2.y, -o2.x);
    const float2 t3 = make_float2(r * (o3.y - o3.x), -r * (o3.x + o3.y));
    float2 Z[8];
    Z[0] = cadd(e0, o0); Z[4] = csub(e0, o0);
    Z[1] = cadd(e1, t1); Z[5] = csub(e1, t1);
    Z[2] = cadd(e2, t2); Z[6] = csub(e2, t2);
    Z[3] = cadd(e3, t3); Z[7] = csub(e3, t3);
    // real split: 2 X[k] = E'[k] + W16^k O'[k], 2 X[8-k] = conj(E'[k] - W16^k O'[k])
    //   E'[k] = Z[k] + conj Z[8-k],  O'[k] = -i (Z[k] - conj Z[8-k])
    X[0] = make_float2(2.0f * (Z[0].x + Z[0].y), 0.f);
    X[8] = make_float2(2.0f * (Z[0].x - Z[0].y), 0.f);
    X[4] = make_float2(2.0f * Z[4].x, -2.0f * Z[4].y);
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;   // cos, sin(pi/8)
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const float wr = k == 1 ? c1 : (k == 2 ? r : s1);    // W16^k = wr - i wi
        const float wi = k == 1 ? s1 : (k == 2 ? r : c1);
        const float2 zk = Z[k], zc = Z[8 - k];
        const float2 E = make_float2(zk.x + zc.x, zk.y - zc.y);
        const float2 O = make_float2(zk.y + zc.y, zc.x - zk.x);           // -i (zk - conj zc)
        const float2 T = make_float2(O.x * wr + O.y * wi, O.y * wr - O.x * wi);   // O * (wr - i wi)
        X[k] = make_float2(E.x + T.x, E.y + T.y);
        X[8 - k] = make_float2(E.x - T.x, T.y - E.y);
    }
}

// fp32 pair -> packed fp16 hi pair and packed fp16 lo pair (lo = x - float(hi), both round-to-nearest)
B200_HD void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - back.x, b - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Stage 1 for one (frame, n2): 16 strided samples -> the 8 packed hi words and 8 packed lo words that
// go to TMEM columns (56 b + n2) and (56 b + 25 + n2), b = 0..7.
//   frame_audio: this frame's first sample inside the padded audio tile (row pitch 161)
B200_HD void tc_stage1(const float* frame_audio, int n2, const float (&win)[16], const float2 (&tw)[8],
                       uint32_t (&hi)[kTcBlocks], uint32_t (&lo)[kTcBlocks]) {
    float x[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        const int n = 25 * n1 + n2;
        x[n1] = frame_audio[n + (n >= kHop) + (n >= 2 * kHop)] * win[n1];
    }
    float2 X[9];
    fft16_real_x2(x, X);
    split_pair(X[0].x, X[8].x, hi[0], lo[0]);
#pragma unroll
    for (int b = 1; b < kTcBlocks; ++b) {
        const float2 t = cmul(X[b], tw[b]);
        split_pair(t.x, t.y, hi[b], lo[b]);
    }
}

// Epilogue of one unit for one thread: 16 complex outputs -> power -> accumulate this thread's
// parity taps into its column of the S tile (s_col points at S[0][frame]).  The 16 taps of a unit
// never alias (tc_tables.h), so all loads are issued before the stores.
B200_HD void tc_accumulate(const float (&d)[32], const TcTap* taps, char* s_col) {
    float acc[16];
    int off[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const TcTap t = taps[j];
        off[j] = t.s_off;
        const float p = d[2 * j] * d[2 * j] + d[2 * j + 1] * d[2 * j + 1];
        acc[j] = *reinterpret_cast<const float*>(s_col + t.s_off) + t.w * p;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) *reinterpret_cast<float*>(s_col + off[j]) = acc[j];
}

}  // namespace b200mel
