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
    // Warp-uniform mel sweep (FFT variant v2): one entry per bin, see phase_mel_sweep.
    MelSweepEntry sweep[kUsedBins];
    int row_off[2 * kMaxMels];      // byte offsets of the two partial-sum rows of mel m ([m], [kMaxMels + m]);
                                    // a missing part points at the all-zero row (index n_rows - 1)
    int n_mels;
    int n_weights;
    int n_rows;                     // rows of the partial-sum tile S: n_mels + straddling mels + 1 zero row
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

    // ---- sweep program: bins are walked in 10 segments of 20; at every bin at most two
    // mels are active and they are consecutive, so mel m accumulates in slot (m & 1).  A mel
    // whose band crosses a segment boundary is summed in two parts (rows row_a / row_b).
    for (int k = 0; k < kUsedBins; ++k) {
        t->sweep[k].w0 = 0.f; t->sweep[k].w1 = 0.f; t->sweep[k].emit0 = -1; t->sweep[k].emit1 = -1;
    }
    int n_rows = n_mels;
    int row_a[kMaxMels], row_b[kMaxMels];
    for (int m = 0; m < n_mels; ++m) {
        row_a[m] = -1; row_b[m] = -1;
        const int band = t->mel_band[m];
        const int first = mel_band_first(band), count = mel_band_count(band);
        if (count == 0) continue;
        const int last = first + count - 1;
        const float* w = t->mel_weights + mel_band_offset(band);
        const int slot = m & 1;
        for (int i = 0; i < count; ++i) {
            float& dst = slot ? t->sweep[first + i].w1 : t->sweep[first + i].w0;
            if (dst != 0.f) return kTablesBadFilters;  // two active mels of the same parity at one bin
            dst = w[i];
        }
        // the slot must be free again before mel m+2 starts
        if (m + 2 < n_mels && mel_band_count(t->mel_band[m + 2]) &&
            mel_band_first(t->mel_band[m + 2]) <= last) return kTablesBadFilters;
        const int seg_first = first / kSegBins, seg_last = last / kSegBins;
        if (seg_last - seg_first > 1) return kTablesBadFilters;
        auto set_emit = [&](int bin, int row) {
            (slot ? t->sweep[bin].emit1 : t->sweep[bin].emit0) = s_row_offset(row);
        };
        row_a[m] = m;
        if (seg_first == seg_last) {
            set_emit(last, m);
        } else {
            if (n_rows >= kMaxSRows - 1) return kTablesBadFilters;
            set_emit(seg_first * kSegBins + kSegBins - 1, m);
            row_b[m] = n_rows;
            set_emit(last, n_rows);
            ++n_rows;
        }
    }
    const int zero_row = n_rows++;
    for (int m = 0; m < kMaxMels; ++m) {
        t->row_off[m] = s_row_offset(m < n_mels && row_a[m] >= 0 ? row_a[m] : zero_row);
        t->row_off[kMaxMels + m] = s_row_offset(m < n_mels && row_b[m] >= 0 ? row_b[m] : zero_row);
    }
    t->n_rows = n_rows;
    return kTablesOk;
}

}  // namespace b200mel
