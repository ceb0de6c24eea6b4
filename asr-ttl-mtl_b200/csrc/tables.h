// Host-side construction of the kernels' constant operands (window, twiddles, mel bands).
// Plain C++ (no CUDA headers) so the CPU choreography emulator builds it with g++ too.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstring>

#include "logmel_core.cuh"

namespace b200mel {

// Constant operands of the front-end, built once per plan on the host in float64
// and kept in device global memory (each CTA stages what it needs in shared memory).
struct DeviceTables {
    float win_half[kNFFT];          // 0.5 * periodic Hann (torch.hann_window(400), audio.py:147)
    float2 twiddle[kNFFT];          // [j][k1] exp(-2 pi i j k1 / 400)
    float mel_weights[kMaxMelWeights];  // non-zero filter taps, band after band
    int mel_band[kMaxMels];         // mel_band_pack(first bin, taps, offset)
    int n_mels;
    int n_weights;
};

constexpr int kTablesOk = 0;
constexpr int kTablesBadFilters = 5;  // == B200MEL_ERR_BAD_FILTERS

// filters: float32 [n_mels, 201] row-major (mel_filters(), reference whisper/audio.py:91-107)
inline int build_tables(int n_mels, const float* filters, DeviceTables* t) {
    std::memset(t, 0, sizeof(*t));
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < kNFFT; ++n)
        t->win_half[n] = static_cast<float>(0.5 * (0.5 - 0.5 * std::cos(two_pi * n / kNFFT)));
    for (int j = 0; j < kRadix; ++j)
        for (int k1 = 0; k1 < kRadix; ++k1) {
            const double ang = -two_pi * static_cast<double>(j * k1) / kNFFT;
            t->twiddle[j * kRadix + k1] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
        }
    int offset = 0;
    for (int m = 0; m < n_mels; ++m) {
        const float* row = filters + static_cast<size_t>(m) * kBins;
        int first = -1, last = -1;
        for (int k = 0; k < kBins; ++k)
            if (row[k] != 0.0f) { if (first < 0) first = k; last = k; }
        if (first < 0) { t->mel_band[m] = mel_band_pack(0, 0, 0); continue; }
        // bin 200 (Nyquist) is never materialised by the kernels; both Whisper banks leave it at zero
        if (last >= kUsedBins) return kTablesBadFilters;
        const int count = last - first + 1;
        if (count > 255 || offset + count > kMaxMelWeights) return kTablesBadFilters;
        for (int i = 0; i < count; ++i) t->mel_weights[offset + i] = row[first + i];
        t->mel_band[m] = mel_band_pack(first, count, offset);
        offset += count;
    }
    t->n_mels = n_mels;
    t->n_weights = offset;
    return kTablesOk;
}

}  // namespace b200mel
