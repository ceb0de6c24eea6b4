// Host-side construction of the tcgen05 variant's constant operands: the three folded-DFT matrices
// (fp16 hi / lo parts, already in the tensor-core shared-memory operand layout) and the run-time mel
// weights.  Plain C++ plus cuda_fp16.h, so the CPU emulator builds it with g++ as well.
#pragma once

#include <cmath>
#include <cstring>

#include "tables.h"
#include "tc_core.cuh"

namespace b200mel {

// byte offset of element (row r = K slot, column kp = bin index k') inside one matrix:
// K-major, no swizzle - 8 x 16-byte core matrices, strips [r / 8][kp][r % 8] of fp16
inline int tc_operand_offset(int r, int kp) { return (r / 8) * kTcStripBytes + kp * 16 + (r % 8) * 2; }

// Operand blob, copied verbatim into shared memory by every CTA:
//   [matrix 0 even-cos, 1 odd-cos, 2 even-sin][part 0 hi, 1 lo][12 strips]   rows (slots) 0..95
//   [matrix][0 main, 1 correction][2 strips]                                 the leftover K step (slots 96..101):
//        main       rows = [zero x 4p | Bh[96..101] | Bh[96..101] | zero]  multiplies the step's [hi | lo] columns
//        correction rows = [zero x 4p | Bl[96..101] | zero ...]            multiplies the same columns (hi Bl)
//   with p = tc_matrix_left_pos(matrix).
constexpr int kTcOperandBytes = kTcMatrices * kTcMatrixBytes + 3 * 2 * kTcLeftStepBytes;   // 139776
struct TcTables {
    unsigned char operands[kTcOperandBytes];
    int n_mels;
};
B200_HD constexpr int tc_matrix_offset(int matrix, int part) { return (2 * matrix + part) * kTcMatrixBytes; }
B200_HD constexpr int tc_left_offset(int matrix, int which) { return kTcMatrices * kTcMatrixBytes + (2 * matrix + which) * kTcLeftStepBytes; }

// exact-phase cosine / sine of 2 pi p / 400 for an integer p
inline double tc_cos400(long p) {
    p %= 400; if (p < 0) p += 400;
    if (p == 100 || p == 300) return 0.0;
    return std::cos(6.283185307179586476925286766559 * static_cast<double>(p) / 400.0);
}
inline double tc_sin400(long p) {
    p %= 400; if (p < 0) p += 400;
    if (p == 0 || p == 200) return 0.0;
    return std::sin(6.283185307179586476925286766559 * static_cast<double>(p) / 400.0);
}

// value of matrix `matrix` at (slot r, bin index kp), before the 2^4 scale (see tc_core.cuh header)
inline double tc_matrix_value(int matrix, int r, int kp) {
    if (r > 100 || kp >= kTcBinsPerUnit) return 0.0;
    switch (matrix) {
        case 0:  // Re X[2kp]: slot r is n = r; the centres r = 0 (2 y[200]) and r = 100 (2 e[100]) count half
            return tc_cos400(2L * kp * r) * ((r == 0 || r == 100) ? 0.5 : 1.0);
        case 1:  // Re X[2kp+1] (n = r) and Im X[2kp+1] (n = 100 - r): the r = 0 centre counts half, r = 100 is cos(odd pi/2) = 0
            return r == 100 ? 0.0 : tc_cos400(static_cast<long>(2 * kp + 1) * r) * (r == 0 ? 0.5 : 1.0);
        default:  // -Im X[2kp]: slot r is n = 100 - r
            return tc_sin400(2L * kp * (100 - r));
    }
}

// The epilogue's mel weights are compile-time constants (mel_bands.h): a plan can use the tcgen05 variant
// only if its filter matrix is bit-equal to them.
template <int NM>
inline int tc_check_filters(const float* filters) {
    for (int m = 0; m < NM; ++m)
        for (int k = 0; k < kBins; ++k) {
            float expect = 0.0f;
            if (k < kUsedBins) {
                const int j = m - MelBands<NM>::bin_mel0[k];
                if (MelBands<NM>::bin_mel0[k] >= 0 && j >= 0 && j < MelBands<NM>::bin_count[k]) expect = MelBands<NM>::bin_weight[k][j];
            }
            if (std::memcmp(&expect, filters + m * kBins + k, sizeof(float)) != 0 && !(expect == 0.0f && filters[m * kBins + k] == 0.0f))
                return kTablesBadFilters;
        }
    return kTablesOk;
}

// filters: float32 [n_mels, 201] row-major.  kTablesBadFilters: this is not the Whisper filterbank of
// mel_bands.h, so only the FFT variant can serve the plan.
inline int build_tc_tables(int n_mels, const float* filters, TcTables* t) {
    std::memset(t, 0, sizeof(*t));
    t->n_mels = n_mels;
    for (int matrix = 0; matrix < 3; ++matrix)
        for (int r = 0; r < 96 + kTcLeftSlots; ++r)
            for (int kp = 0; kp < kTcN; ++kp) {
                const double v = tc_matrix_value(matrix, r, kp) * kTcMatrixScale;
                const __half hi = __float2half_rn(static_cast<float>(v));
                const __half lo = __float2half_rn(static_cast<float>(v - static_cast<double>(__half2float(hi))));
                if (r < 96) {
                    std::memcpy(t->operands + tc_matrix_offset(matrix, 0) + tc_operand_offset(r, kp), &hi, 2);
                    std::memcpy(t->operands + tc_matrix_offset(matrix, 1) + tc_operand_offset(r, kp), &lo, 2);
                } else {
                    const int row = 4 * tc_matrix_left_pos(matrix) + (r - 96);   // row of the hi slot inside the leftover step
                    std::memcpy(t->operands + tc_left_offset(matrix, 0) + tc_operand_offset(row, kp), &hi, 2);
                    std::memcpy(t->operands + tc_left_offset(matrix, 0) + tc_operand_offset(row + kTcLeftSlots, kp), &hi, 2);
                    std::memcpy(t->operands + tc_left_offset(matrix, 1) + tc_operand_offset(row, kp), &lo, 2);
                }
            }
    if (n_mels == 80) return tc_check_filters<80>(filters);
    if (n_mels == 128) return tc_check_filters<128>(filters);
    return kTablesBadFilters;
}

}  // namespace b200mel
