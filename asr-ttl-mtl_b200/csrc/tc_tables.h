// Host-side construction of the tcgen05 variant's constant operands (tc_core.cuh: TcTables).
// Plain C++ (+ cuda_fp16.h host conversions) so the CPU emulator builds it with g++ too.
#pragma once

#include <cmath>
#include <cstring>

#include "tc_core.cuh"

namespace b200mel {

// index of element (n, k) in a K-major, no-swizzle tcgen05 operand of kTcN rows: strips [k/8][n][8]
inline int tc_operand_index(int n, int k) { return (k / 8) * (kTcN * 8) + n * 8 + (k % 8); }

// Output order of a block: N-half h, slot j (0..15).  Within one half no two bins share a mel (they are
// 16 bins apart and no filter is that wide), so an epilogue thread can pipeline its 16 tap updates.
//   block 0 : half 0 = X[16 (j+1)] (j < 12, from Y[0,.]);  half 1 = X[8 + 16 j] (j < 12, from Y[8,.])
//   block b : half 0 = X[b + 16 j] (j < 13, bins <= 199);  half 1 = X[b + 16 (13 + j)] (j < 12), mirrored to 400 - k
// Returns the DFT index k (0..399) or -1 for a padding slot.
inline int tc_output_k(int b, int h, int j) {
    if (b == 0) return j < 12 ? (h == 0 ? 16 * (j + 1) : 8 + 16 * j) : -1;
    if (h == 0) return j < 13 ? b + 16 * j : -1;
    return j < 12 ? b + 16 * (13 + j) : -1;
}
// bin (1..199) that slot feeds: |X[400-k]| = |X[k]| for real input
inline int tc_output_bin(int b, int h, int j) {
    const int k = tc_output_k(b, h, j);
    return k < 0 ? -1 : (k <= 199 ? k : kNFFT - k);
}

// filters: float32 [n_mels, 201] row-major.  Returns kTablesOk or kTablesBadFilters.
inline int build_tc_tables(int n_mels, const float* filters, TcTables* t) {
    std::memset(static_cast<void*>(t), 0, sizeof(*t));
    const double two_pi = 6.283185307179586476925286766559;
    t->n_mels = n_mels;
    for (int n2 = 0; n2 < kTcN2; ++n2) {
        for (int n1 = 0; n1 < 16; ++n1) {
            const int n = 25 * n1 + n2;
            t->win[n2][n1] = static_cast<float>(0.5 * kTcInputScale * (0.5 - 0.5 * std::cos(two_pi * n / kNFFT)));
        }
        for (int b = 0; b < 8; ++b) {
            const double ang = -two_pi * n2 * b / kNFFT;
            t->tw[n2][b] = make_float2(static_cast<float>(std::cos(ang)), static_cast<float>(std::sin(ang)));
        }
    }
    // dense stage-2 matrices B[k = (n2, re|im)][n = 32 h + 2 j + (re|im)], float64.
    // X[k] = sum_n2 In[n2] exp(-2 pi i n2 k / 400) * (twiddle already applied for blocks 1..7, so the
    // remaining factor there is W25^(n2 k2) = exp(-2 pi i n2 (k - b) / 400)).
    static double B[2][50][kTcN];
    std::memset(B, 0, sizeof(B));
    for (int n2 = 0; n2 < kTcN2; ++n2)
        for (int h = 0; h < 2; ++h)
            for (int j = 0; j < 16; ++j) {
                const int col = 32 * h + 2 * j;
                int k = tc_output_k(1, h, j);            // blocks 1..7 share one matrix: use b = 1, k2 = (k - 1) / 16
                if (k >= 0) {
                    const double th = two_pi * n2 * ((k - 1) / 16) / 25.0, cs = std::cos(th), sn = std::sin(th);
                    B[1][2 * n2][col] = cs;      B[1][2 * n2 + 1][col] = sn;       // (a + i b)(cos - i sin)
                    B[1][2 * n2][col + 1] = -sn; B[1][2 * n2 + 1][col + 1] = cs;
                }
                k = tc_output_k(0, h, j);                // block 0: real inputs Y0 (row 2 n2) / Y8 (row 2 n2 + 1)
                if (k >= 0) {
                    const double th = two_pi * n2 * k / kNFFT;
                    const int row = h == 0 ? 2 * n2 : 2 * n2 + 1;
                    B[0][row][col] = std::cos(th);
                    B[0][row][col + 1] = -std::sin(th);
                }
            }
    for (int set = 0; set < 2; ++set)
        for (int n = 0; n < kTcN; ++n)
            for (int k = 0; k < 50; ++k) {
                const __half hi = __float2half_rn(static_cast<float>(B[set][k][n]));
                const __half lo = __float2half_rn(static_cast<float>(B[set][k][n] - static_cast<double>(__half2float(hi))));
                t->b_main[set][tc_operand_index(n, k)] = hi;       // rows 0..49  : Bhi (times A hi)
                t->b_main[set][tc_operand_index(n, 50 + k)] = hi;  // rows 50..99 : Bhi (times A lo)
                t->b_corr[set][tc_operand_index(n, k)] = lo;       // rows 0..49  : Blo (times A hi)
            }
    // epilogue taps
    const int row_bytes = kTcTileFrames * static_cast<int>(sizeof(float));
    for (int p = 0; p < 2; ++p)
        for (int u = 0; u < kTcUnits; ++u)
            for (int j = 0; j < 16; ++j) {
                TcTap tap;
                tap.w = 0.f;
                tap.s_off = (n_mels + p) * row_bytes;  // scratch row of this parity
                const int bin = tc_output_bin(u / 2, u % 2, j);
                if (bin >= 0) {
                    int found = 0;
                    for (int m = p; m < n_mels; m += 2) {
                        const float w = filters[static_cast<size_t>(m) * kBins + bin];
                        if (w != 0.f) {
                            if (found++) return kTablesBadFilters;  // two active mels of one parity at a bin
                            tap.w = w * kTcPowerUnscale;
                            tap.s_off = m * row_bytes;
                        }
                    }
                }
                t->tap[p][u][j] = tap;
            }
    // bins the variant never computes (0 and 200) must carry no weight
    for (int m = 0; m < n_mels; ++m)
        if (filters[static_cast<size_t>(m) * kBins] != 0.f || filters[static_cast<size_t>(m) * kBins + 200] != 0.f)
            return kTablesBadFilters;
    return kTablesOk;
}

}  // namespace b200mel
