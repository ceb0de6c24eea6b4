// CPU emulator of the tcgen05 variant (TEST INFRASTRUCTURE, not a fallback): runs the stage-1 /
// epilogue functions of csrc/tc_core.cuh with the kernel's tile geometry, tensor-memory column
// layout and operand tables; the tensor-core MMAs are replaced by the exact fp16 x fp16 products
// they stand for, read through the same operand layouts.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../asr-ttl-mtl_b200/csrc/tables.h"
#include "../../asr-ttl-mtl_b200/csrc/tc_tables.h"

using namespace b200mel;

static float half_at(const std::vector<uint32_t>& A, int frame, int half_index) {
    const uint32_t word = A[frame * 512 + half_index / 2];
    __half_raw raw;
    raw.x = static_cast<unsigned short>((half_index & 1) ? (word >> 16) : (word & 0xffffu));
    return __half2float(__half(raw));
}

extern "C" int emul_tc_logmel(const float* audio, int64_t n_samples, int64_t valid, int64_t right_pad,
                              int n_mels, const float* filters, float* out, int do_normalise) {
    static TcTables tab;
    if (build_tc_tables(n_mels, filters, &tab) != kTablesOk) return 5;
    const int64_t total = n_samples + (right_pad > 0 ? right_pad : 0);
    if (total <= kHalfWin) return 3;
    const int n_frames = static_cast<int>(total / kHop);
    const int tiles = (n_frames + kTcTileFrames - 1) / kTcTileFrames;
    if (valid > n_samples) valid = n_samples;

    std::vector<float> s_audio(kTcAudioFloats);
    std::vector<uint32_t> A(kTcTileFrames * 512, 0u);  // tensor memory: [lane][column]
    std::vector<float> S((n_mels + 2) * kTcTileFrames, 0.f);
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
            s_audio[i + i / kHop] = v;
        }
        for (int f = 0; f < kTcTileFrames; ++f)
            for (int n2 = 0; n2 < kTcN2; ++n2) {
                uint32_t hi[kTcBlocks], lo[kTcBlocks];
                tc_stage1(s_audio.data() + kTcRowPitch * f, n2, tab.win[n2], tab.tw[n2], hi, lo);
                for (int b = 0; b < kTcBlocks; ++b) {
                    A[f * 512 + kTcBlockCols * b + n2] = hi[b];
                    A[f * 512 + kTcBlockCols * b + kTcN2 + n2] = lo[b];
                }
            }
        for (int u = 0; u < kTcUnits; ++u) {
            const int b = u / 2, h = u % 2, set = b == 0 ? 0 : 1;
            for (int f = 0; f < kTcTileFrames; ++f) {
                float d[32];
                for (int n = 0; n < 32; ++n) {
                    double acc = 0.0;
                    for (int k = 0; k < kTcKMain; ++k)
                        acc += static_cast<double>(half_at(A, f, 2 * kTcBlockCols * b + k)) *
                               static_cast<double>(__half2float(tab.b_main[set][tc_operand_index(32 * h + n, k)]));
                    for (int k = 0; k < kTcKCorr; ++k)
                        acc += static_cast<double>(half_at(A, f, 2 * kTcBlockCols * b + k)) *
                               static_cast<double>(__half2float(tab.b_corr[set][tc_operand_index(32 * h + n, k)]));
                    d[n] = static_cast<float>(acc);
                }
                for (int p = 0; p < 2; ++p)
                    tc_accumulate(d, tab.tap[p][u], reinterpret_cast<char*>(S.data() + f));
            }
        }
        for (int f = 0; f < kTcTileFrames && t0 + f < n_frames; ++f)
            for (int m = 0; m < n_mels; ++m) {
                const float lg = log10_clamped(S[m * kTcTileFrames + f]);
                out[static_cast<int64_t>(m) * n_frames + t0 + f] = lg;
                const uint32_t k = max_key_encode(lg);
                if (k > clip_key) clip_key = k;
            }
        std::fill(S.begin(), S.end(), 0.f);
    }
    if (do_normalise) {
        const float g = max_key_decode(clip_key);
        for (int64_t i = 0; i < static_cast<int64_t>(n_mels) * n_frames; ++i) out[i] = normalise(out[i], g);
    }
    return 0;
}

extern "C" void emul_fft16_real_x2(const float* x16, float* out18) {
    float x[16];
    float2 X[9];
    for (int i = 0; i < 16; ++i) x[i] = x16[i];
    fft16_real_x2(x, X);
    for (int k = 0; k < 9; ++k) { out18[2 * k] = X[k].x; out18[2 * k + 1] = X[k].y; }
}
