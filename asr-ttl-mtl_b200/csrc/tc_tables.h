// Host-side construction of the tcgen05 variant's constant operands (tc_core.cuh: TcTables).
// Plain C++ (+ cuda_fp16.h host conversions) so the CPU emulator builds it with g++ too.
#pragma once

#include <cmath>
#include <cstring>

#include "tc_core.cuh"

namespace b200mel {

// index of element (n, k) in a K-major, no-swizzle tcgen05 operand of kTcN rows: strips [k/8][n][8]
inline int tc_operand_index(int n, int k) { return (k / 8) * (kTcN * 8) + n * 8 + (k % 8); }

// bins produced by block 0, in output order: X[16 j] (j = 1..12) from Y[0,.], X[8 + 16 j] (j = 0..11) from Y[8,.]
inline int tc_block0_bin(int c) { return c < 12 ? 16 * (c + 1) : 8 + 16 * (c - 12); }

// bin (0..199) of complex output c of block b, or -1 if that output is padding
inline int tc_output_bin(int b, int c) {
    if (b == 0) return c < 24 ? tc_block0_bin(c) : -1;
    if (c >= kTcN2) return -1;
    const int k = b + 16 * c;
    return k <= 199 ? k : kNFFT - k;  // |X[400-k]| = |X[k]| for real input
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
    // dense stage-2 matrices B[k = (n2, re|im)][n = (output c, re|im)], float64
    static double B[2][50][kTcN];
    std::memset(B, 0, sizeof(B));
    for (int n2 = 0; n2 < kTcN2; ++n2) {
        for (int c = 0; c < kTcN2; ++c) {  // blocks 1..7: (a + i b)(cos - i sin)
            const double th = two_pi * n2 * c / 25.0, cs = std::cos(th), sn = std::sin(th);
            B[1][2 * n2][2 * c] = cs;      B[1][2 * n2 + 1][2 * c] = sn;
            B[1][2 * n2][2 * c + 1] = -sn; B[1][2 * n2 + 1][2 * c + 1] = cs;
        }
        for (int c = 0; c < 24; ++c) {     // block 0: real inputs Y0 (row 2 n2) and Y8 (row 2 n2 + 1)
            const int k = tc_block0_bin(c);
            const double th = two_pi * n2 * k / kNFFT;  // W25^(n2 k2) for k = 16 k2; carries W400^(8 n2) for k = 8 + 16 k2
            const int row = c < 12 ? 2 * n2 : 2 * n2 + 1;
            B[0][row][2 * c] = std::cos(th);
            B[0][row][2 * c + 1] = -std::sin(th);
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
                const int bin = tc_output_bin(u / 2, 16 * (u % 2) + j);
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
