// CPU emulator of the tcgen05 variant (TEST INFRASTRUCTURE, not a fallback): runs the sweep / epilogue
// functions of csrc/tc_core.cuh with the kernel's tile geometry, tensor-memory column map and operand
// tables (csrc/tc_tables.h); the tensor-core MMAs are replaced by the exact fp16 x fp16 products they stand
// for, read through the same operand layouts - including the neighbour's columns a leftover K step reads.
#include <algorithm>
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

// accumulator column kp of unit u for one frame: the 6 x 3 main MMAs and the 2 leftover MMAs, as issued
float unit_column(const TcTables& tab, const std::vector<uint32_t>& tmem, int f, int u, int kp) {
    const int m = tc_unit_matrix(u);
    double acc = 0.0;
    for (int r = 0; r < 16 * kTcMainSteps; ++r) {
        const double bh = b_bits(tab, tc_matrix_offset(m, 0), r, kp), bl = b_bits(tab, tc_matrix_offset(m, 1), r, kp);
        const double ah = a_at(tmem, f, tc_hi_col(u), r), al = a_at(tmem, f, tc_lo_col(u), r);
        acc += ah * bh + al * bh + ah * bl;
    }
    for (int r = 0; r < 16; ++r) {   // the leftover K step reads 8 columns from tc_left_start(u)
        const double av = a_at(tmem, f, tc_left_start(u), r);
        acc += av * b_bits(tab, tc_left_offset(m, 0), r, kp) + av * b_bits(tab, tc_left_offset(m, 1), r, kp);
    }
    return static_cast<float>(acc);
}

constexpr TcFoldTable kFold = tc_make_fold_table();
constexpr TcFoldWeights kFoldWeights = tc_make_fold_weights();
constexpr TcFoldRows kFoldRows = tc_make_fold_rows();
float head_row[2][2];     // head carry of the row-driven form, per sweep
int g_chunk_mismatch = 0;   // set when the compile-time chunk and the table-driven chunk disagree

// one sweep of one frame, stored to the emulated tensor-memory lane exactly like the kernel's fold warps do
template <int SWEEP, int J>
void chunk_to_tmem(const float* fr, float (&head)[2], uint32_t* lane) {
    constexpr int u1 = 2 * SWEEP, u2 = u1 + 1;
    uint32_t hf[4], lf[4], hs[4], ls[4];
    tc_sweep_chunk_ct<SWEEP, J>(fr, kFoldWeights, head, hf, lf, hs, ls);
    {   // the table-driven form of the same chunk must agree bit for bit
        uint32_t hf2[4], lf2[4], hs2[4], ls2[4];
        tc_sweep_chunk<SWEEP>(fr, kFold.c[SWEEP][J], hf2, lf2, hs2, ls2);
        const int live = J < 2 * kTcMainSteps ? 4 : 3;
        for (int q = 0; q < live; ++q)
            if (hf[q] != hf2[q] || lf[q] != lf2[q] || hs[q] != hs2[q] || ls[q] != ls2[q]) g_chunk_mismatch = 1;
    }
    {   // ... and so must the row-driven form the kernel runs (its own head carry, checked chunk by chunk)
        // a warp may start a sweep at any chunk (the kernel's two warps per quadrant start at chunks 0 and 7): the row's
        // own head offsets must give what the carry would have
        if (J == 0) { head_row[SWEEP][0] = fr[kFoldRows.row[SWEEP][0].head[0] / 4]; head_row[SWEEP][1] = fr[kFoldRows.row[SWEEP][0].head[1] / 4]; }
        else if (J < kTcChunks - 1 &&
                 (head_row[SWEEP][0] != fr[kFoldRows.row[SWEEP][J].head[0] / 4] || head_row[SWEEP][1] != fr[kFoldRows.row[SWEEP][J].head[1] / 4]))
            g_chunk_mismatch = 1;
        uint32_t hf3[4], lf3[4], hs3[4], ls3[4];
        tc_sweep_chunk_row(fr, kFoldRows.row[SWEEP][J], kFoldRows.sign[SWEEP], head_row[SWEEP], hf3, lf3, hs3, ls3);
        const int live = J < 2 * kTcMainSteps ? 4 : 3;
        for (int q = 0; q < live; ++q)
            if (hf[q] != hf3[q] || lf[q] != lf3[q] || hs[q] != hs3[q] || ls[q] != ls3[q]) g_chunk_mismatch = 1;
    }
    if (J < 2 * kTcMainSteps) {
        for (int q = 0; q < 4; ++q) {
            lane[tc_hi_col(u1) + 4 * J + q] = hf[q]; lane[tc_lo_col(u1) + 4 * J + q] = lf[q];
            lane[tc_hi_col(u2) + 4 * J + q] = hs[q]; lane[tc_lo_col(u2) + 4 * J + q] = ls[q];
        }
    } else {
        for (int q = 0; q < 3; ++q) {
            lane[tc_left_col(u1) + q] = hf[q]; lane[tc_left_col(u1) + 3 + q] = lf[q];
            lane[tc_left_col(u2) + q] = hs[q]; lane[tc_left_col(u2) + 3 + q] = ls[q];
        }
    }
}
template <int SWEEP, int... J>
void sweep_seq(const float* fr, uint32_t* lane, std::integer_sequence<int, J...>) {
    float head[2];
    tc_sweep_heads<SWEEP>(fr, head);
    (chunk_to_tmem<SWEEP, J>(fr, head, lane), ...);
}
void sweep_to_tmem(int sweep, const float* fr, uint32_t* lane) {
    if (sweep == 0) sweep_seq<0>(fr, lane, std::make_integer_sequence<int, kTcChunks>{});
    else sweep_seq<1>(fr, lane, std::make_integer_sequence<int, kTcChunks>{});
}

template <int NM, int U>
void epilogue_unit(const float* d, float* acc0, float* acc1) {
    using L = TcEpilogueLayout<NM>;
    float d0[L::cols(0)], d1[L::cols(1)];
    float (&a0)[L::acc_size(0)] = *reinterpret_cast<float (*)[L::acc_size(0)]>(acc0);
    float (&a1)[L::acc_size(1)] = *reinterpret_cast<float (*)[L::acc_size(1)]>(acc1);
    for (int c = 0; c < L::cols(0); ++c) d0[c] = d[c];
    for (int c = 0; c < L::cols(1); ++c) d1[c] = d[L::split + c];
    tc_epilogue_unit<NM, U, 0>(d0, a0);
    tc_epilogue_unit<NM, U, 1>(d1, a1);
}

template <int NM>
int run(const float* audio, int64_t n_samples, int64_t valid, int64_t right_pad, const float* filters, float* out,
        int do_normalise) {
    using L = TcEpilogueLayout<NM>;
    static TcTables tab;
    if (build_tc_tables(NM, filters, &tab) != kTablesOk) return 5;
    const int64_t total = n_samples + (right_pad > 0 ? right_pad : 0);
    if (total <= kHalfWin) return 3;
    const int n_frames = static_cast<int>(total / kHop);
    const int tiles = (n_frames + kTcTileFrames - 1) / kTcTileFrames;
    if (valid > n_samples) valid = n_samples;

    std::vector<float> s_audio(kTcAudioWords + 8, 0.f);
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
        for (int f = 0; f < kTcTileFrames; ++f) {
            const float* fr = s_audio.data() + f * kTcRowPitch;
            uint32_t* lane = tmem.data() + f * 512;
            sweep_to_tmem(0, fr, lane);
            sweep_to_tmem(1, fr, lane);
        }
        for (int f = 0; f < kTcTileFrames && t0 + f < n_frames; ++f) {
            float acc0[L::acc_size(0)] = {0.f}, acc1[L::acc_size(1)] = {0.f};
            for (int u = 0; u < kTcUnits; ++u) {
                float d[kTcN];
                for (int kp = 0; kp < kTcN; ++kp) d[kp] = unit_column(tab, tmem, f, u, kp);
                switch (u) {
                    case 0: epilogue_unit<NM, 0>(d, acc0, acc1); break;
                    case 1: epilogue_unit<NM, 1>(d, acc0, acc1); break;
                    case 2: epilogue_unit<NM, 2>(d, acc0, acc1); break;
                    default: epilogue_unit<NM, 3>(d, acc0, acc1); break;
                }
            }
            for (int m = 0; m < NM; ++m) {
                float s = 0.f;
                if (m < L::low_mels) s += acc0[m];
                if (m >= L::high_base) s += acc1[m - L::high_base];
                const float lg = log10_clamped(s);
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
    std::vector<float> s_audio(3 * kTcRowPitch, 0.f);
    for (int n = 0; n < kNFFT; ++n) s_audio[tc_off(n)] = frame400[n];
    std::vector<uint32_t> tmem(512, 0u);
    sweep_to_tmem(0, s_audio.data(), tmem.data());
    sweep_to_tmem(1, s_audio.data(), tmem.data());
    for (int u = 0; u < kTcUnits; ++u) {
        for (int kp = 0; kp < kTcBinsPerUnit; ++kp) {
            const int bin = tc_unit_bin(u, kp);
            const double v = static_cast<double>(unit_column(tab, tmem, 0, u, kp)) / (kTcDataScale * kTcMatrixScale);
            if (u < 2) re[bin] = v; else im_abs[bin] = v < 0 ? -v : v;
        }
    }
    return 0;
}

// 1 when any chunk computed so far differed between the compile-time form (what the kernel runs) and the
// table-driven form of the fold (tc_core.cuh)
extern "C" int emul_tc_chunk_mismatch() { return g_chunk_mismatch; }
