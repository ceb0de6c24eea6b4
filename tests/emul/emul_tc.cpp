// CPU emulator of the tcgen05 variant (TEST INFRASTRUCTURE, not a fallback): runs the sweep / epilogue
// functions of csrc/tc_core.cuh with the kernel's tile geometry, tensor-memory column map and operand
// tables (csrc/tc_tables.h); the tensor-core MMAs are replaced by the exact fp16 x fp16 products they stand
// for, read through the same operand layouts - including the neighbour's columns a leftover K step reads.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../asr-ttl-mtl_b200/csrc/tables.h"
#include "../../asr-ttl-mtl_b200/csrc/tc_tables.h"

using namespace b200mel;

namespace {

float half_bits_to_float(uint16_t bits) {
    __half_raw raw;
    raw.x = bits;
    return __half2float(__half(raw));
}

// fp16 element `slot` of a 7-step pass that starts at tensor-memory column `col0` of `frame`
float a_at(const std::vector<uint32_t>& tmem, int frame, int col0, int slot) {
    const uint32_t word = tmem[frame * 512 + col0 + slot / 2];
    return half_bits_to_float(static_cast<uint16_t>((slot & 1) ? (word >> 16) : (word & 0xffffu)));
}

float b_bits(const TcTables& tab, int offset, int r, int kp) {
    uint16_t bits;
    std::memcpy(&bits, tab.operands + offset + tc_operand_offset(r, kp), 2);
    return half_bits_to_float(bits);
}

// How the fp32 accumulator is modelled: 0 = exact sum of the exact fp16 x fp16 products, rounded once (the arithmetic the
// three-product scheme stands for); 1 = a model of the hardware: after every MMA (one K step of 16 products) the
// accumulator is truncated towards zero at the fp32 ulp of the largest addend (Fasi et al., "Numerical behavior of NVIDIA
// tensor cores"), in the kernel's issue order - the small products first, the main product last.
int g_accumulate_model = 0;

double mma_step(double acc, const double* prod, int n) {
    double s = acc, mag = std::fabs(acc);
    for (int i = 0; i < n; ++i) { s += prod[i]; mag = std::max(mag, std::fabs(prod[i])); }
    if (g_accumulate_model == 0 || mag == 0.0) return s;
    int e;
    std::frexp(mag, &e);                                  // mag = m 2^e, m in [0.5, 1): ulp of a 24-bit significand = 2^(e-24)
    const double ulp = std::ldexp(1.0, e - 24);
    return std::trunc(s / ulp) * ulp;
}

// accumulator column kp of unit u for one frame: the 6 x 2 small MMAs, the leftover correction, the 6 main MMAs and the
// leftover main step, as issued
float unit_column(const TcTables& tab, const std::vector<uint32_t>& tmem, int f, int u, int kp) {
    const int m = tc_unit_matrix(u);
    double acc = 0.0, prod[16];
    for (int s = 0; s < kTcMainSteps; ++s) {
        for (int i = 0; i < 16; ++i) prod[i] = static_cast<double>(a_at(tmem, f, tc_lo_col(u), 16 * s + i)) * b_bits(tab, tc_matrix_offset(m, 0), 16 * s + i, kp);
        acc = mma_step(acc, prod, 16);
        for (int i = 0; i < 16; ++i) prod[i] = static_cast<double>(a_at(tmem, f, tc_hi_col(u), 16 * s + i)) * b_bits(tab, tc_matrix_offset(m, 1), 16 * s + i, kp);
        acc = mma_step(acc, prod, 16);
    }
    for (int i = 0; i < 16; ++i) prod[i] = static_cast<double>(a_at(tmem, f, tc_left_start(u), i)) * b_bits(tab, tc_left_offset(m, 1), i, kp);
    acc = mma_step(acc, prod, 16);
    for (int s = 0; s < kTcMainSteps; ++s) {
        for (int i = 0; i < 16; ++i) prod[i] = static_cast<double>(a_at(tmem, f, tc_hi_col(u), 16 * s + i)) * b_bits(tab, tc_matrix_offset(m, 0), 16 * s + i, kp);
        acc = mma_step(acc, prod, 16);
    }
    for (int i = 0; i < 16; ++i) prod[i] = static_cast<double>(a_at(tmem, f, tc_left_start(u), i)) * b_bits(tab, tc_left_offset(m, 0), i, kp);   // the leftover K step reads 8 columns from tc_left_start(u)
    acc = mma_step(acc, prod, 16);
    return static_cast<float>(acc);
}

constexpr TcFoldTables kFold = tc_make_fold_tables();

// one sweep of one frame at scale step `scale`, stored to the emulated tensor-memory lane exactly like the kernel's fold
// warps do; a warp may start a sweep at any chunk (the kernel's two warps per quadrant start at chunks 0 and 7), so the
// head carry is checked against the row's own head offsets chunk by chunk
int g_chunk_mismatch = 0;
void sweep_to_tmem(int sweep, int scale, const float* fr, uint32_t* lane) {
    const int u1 = 2 * sweep, u2 = u1 + 1;
    float head[2] = {fr[kFold.off[sweep][0].head[0] / 4], fr[kFold.off[sweep][0].head[1] / 4]};
    for (int j = 0; j < kTcChunks; ++j) {
        if (j > 0 && j < kTcChunks - 1 &&
            (head[0] != fr[kFold.off[sweep][j].head[0] / 4] || head[1] != fr[kFold.off[sweep][j].head[1] / 4]))
            g_chunk_mismatch = 1;
        uint32_t hf[4], lf[4], hs[4], ls[4];
        tc_sweep_chunk_host(fr, kFold.off[sweep][j], kFold.w[scale][sweep][j], kFold.sign[sweep], head, hf, lf, hs, ls);
        if (j < 2 * kTcMainSteps) {
            for (int q = 0; q < 4; ++q) {
                lane[tc_hi_col(u1) + 4 * j + q] = hf[q]; lane[tc_lo_col(u1) + 4 * j + q] = lf[q];
                lane[tc_hi_col(u2) + 4 * j + q] = hs[q]; lane[tc_lo_col(u2) + 4 * j + q] = ls[q];
            }
        } else {
            for (int q = 0; q < 3; ++q) {
                lane[tc_left_col(u1) + q] = hf[q]; lane[tc_left_col(u1) + 3 + q] = lf[q];
                lane[tc_left_col(u2) + q] = hs[q]; lane[tc_left_col(u2) + 3 + q] = ls[q];
            }
        }
    }
}

// scale step of lane quadrant q of the staged tile: the largest |sample| its 32 frames read (samples 0 .. 5359 from the
// quadrant's first row - what the kernel's E sweep tracks)
int quadrant_scale(const float* s_audio, int q) {
    float m = 0.f;
    for (int i = 0; i < 31 * kHop + kNFFT; ++i) m = tc_abs_max(m, s_audio[(32 * q + i / kHop) * kTcRowPitch + i % kHop]);
    return tc_scale_index(float_bits(m));
}

template <int NM, int U>
void epilogue_unit(const float* d, float (&acc)[NM]) {
    using L = TcEpilogueLayout<NM>;
    // the kernel pulls a unit's columns as L::pieces pieces, in column order
    float piece[L::piece_cols];
    for (int c = 0; c < L::piece_cols; ++c) piece[c] = d[c];
    tc_epilogue_unit<NM, U, 0, L::piece_cols>(piece, acc);
    if constexpr (L::pieces == 2) {
        for (int c = 0; c < L::piece_cols; ++c) piece[c] = d[L::piece_cols + c];
        tc_epilogue_unit<NM, U, L::piece_cols, L::piece_cols>(piece, acc);
    }
}

template <int NM>
int run(const float* audio, int64_t n_samples, int64_t valid, int64_t right_pad, const float* filters, float* out,
        int do_normalise) {
    static TcTables tab;
    if (build_tc_tables(NM, filters, &tab) != kTablesOk) return 5;
    const int64_t total = n_samples + (right_pad > 0 ? right_pad : 0);
    if (total <= kHalfWin) return 3;
    const int n_frames = static_cast<int>(total / kHop);
    const int tiles = (n_frames + kTcTileFrames - 1) / kTcTileFrames;
    if (valid > n_samples) valid = n_samples;

    std::vector<float> s_audio((kTcAudioRows + 1) * kTcRowPitch + 8, 0.f);
    std::vector<uint32_t> tmem(kTcTileFrames * 512, 0u);   // tensor memory: [lane][column], zero-initialised like the kernel
    uint32_t clip_key = 0;

    for (int tile = 0; tile < tiles; ++tile) {
        const int t0 = tile * kTcTileFrames;
        const int64_t s0 = static_cast<int64_t>(t0) * kHop - kHalfWin;
        for (int i = 0; i < kTcAudioSamples; ++i) {
            const int64_t s = s0 + i;
            float v = 0.f;
            if (s < total + kHalfWin) {
                const int64_t idx = reflect_source_index(s, total);
                if (idx >= 0 && idx < valid) v = audio[idx];
            }
            s_audio[(i / kHop) * kTcRowPitch + i % kHop] = v;
        }
        int scale[4];
        for (int q = 0; q < 4; ++q) scale[q] = quadrant_scale(s_audio.data(), q);
        for (int f = 0; f < kTcTileFrames; ++f) {
            const float* fr = s_audio.data() + f * kTcRowPitch;
            uint32_t* lane = tmem.data() + f * 512;
            sweep_to_tmem(0, scale[f / 32], fr, lane);
            sweep_to_tmem(1, scale[f / 32], fr, lane);
        }
        for (int f = 0; f < kTcTileFrames && t0 + f < n_frames; ++f) {
            float acc[NM] = {0.f};
            for (int u = 0; u < kTcUnits; ++u) {
                float d[kTcN];
                for (int kp = 0; kp < kTcN; ++kp) d[kp] = unit_column(tab, tmem, f, u, kp);
                switch (u) {
                    case 0: epilogue_unit<NM, 0>(d, acc); break;
                    case 1: epilogue_unit<NM, 1>(d, acc); break;
                    case 2: epilogue_unit<NM, 2>(d, acc); break;
                    default: epilogue_unit<NM, 3>(d, acc); break;
                }
            }
            for (int m = 0; m < NM; ++m) {
                const float s = acc[m];
                const float lg = log10_clamped(s * kFold.unscale[scale[f / 32]]);
                out[static_cast<int64_t>(m) * n_frames + t0 + f] = lg;
                clip_key = std::max(clip_key, max_key_encode(lg));
            }
        }
    }
    if (do_normalise) {
        const float g = max_key_decode(clip_key);
        for (int64_t i = 0; i < static_cast<int64_t>(NM) * n_frames; ++i) out[i] = normalise(out[i], g);
    }
    return 0;
}

}  // namespace

extern "C" void emul_tc_accumulate_model(int model) { g_accumulate_model = model; }

extern "C" int emul_tc_logmel(const float* audio, int64_t n_samples, int64_t valid, int64_t right_pad,
                              int n_mels, const float* filters, float* out, int do_normalise) {
    if (n_mels == 80) return run<80>(audio, n_samples, valid, right_pad, filters, out, do_normalise);
    if (n_mels == 128) return run<128>(audio, n_samples, valid, right_pad, filters, out, do_normalise);
    return 2;
}

// the folded spectrum of ONE frame (400 samples): out[2k], out[2k+1] = Re, |Im| partial products D / 4096
// for bins k = 0..199, straight from the emulated accumulators (checks folds, matrices and slot maps)
extern "C" int emul_tc_frame_spectrum(const float* frame400, double* re, double* im_abs) {
    static TcTables tab;
    static bool built = false;
    if (!built) {
        std::vector<float> dummy(80 * kBins, 0.f);
        build_tc_tables(80, dummy.data(), &tab);   // the matrices do not depend on the filters (status ignored)
        built = true;
    }
    std::vector<float> s_audio(3 * kTcRowPitch + 8, 0.f);
    for (int n = 0; n < kNFFT; ++n) s_audio[tc_off(n)] = frame400[n];
    std::vector<uint32_t> tmem(512, 0u);
    float m = 0.f;
    for (int n = 0; n < kNFFT; ++n) m = tc_abs_max(m, frame400[n]);
    const int scale = tc_scale_index(float_bits(m));
    sweep_to_tmem(0, scale, s_audio.data(), tmem.data());
    sweep_to_tmem(1, scale, s_audio.data(), tmem.data());
    for (int u = 0; u < kTcUnits; ++u) {
        for (int kp = 0; kp < kTcBinsPerUnit; ++kp) {
            const int bin = tc_unit_bin(u, kp);
            const double v = static_cast<double>(unit_column(tab, tmem, 0, u, kp)) / (static_cast<double>(tc_pow2(tc_scale_exponent(scale))) * kTcMatrixScale);
            if (u < 2) re[bin] = v; else im_abs[bin] = v < 0 ? -v : v;
        }
    }
    return 0;
}

// 1 when the sample a sweep chunk hands on to the next one ever differed from the head sample the fold table names for
// that chunk (a warp may start a sweep at any chunk)
extern "C" int emul_tc_chunk_mismatch() { return g_chunk_mismatch; }
