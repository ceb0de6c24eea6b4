// Issue rate of the instructions the fold sweep is made of, per SM sub-partition (sm_100a):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
// Each test runs W warps per sub-partition (blockDim = 128 W), every warp issuing 8 independent chains of one
// instruction; prints cycles per warp-instruction per sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int kIters = 2048;

template <int OP>
__global__ void rate_kernel(float* out, long long* cycles, float seed) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = seed + i + threadIdx.x * 0.001f;
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if constexpr (OP == 0) {            // F2FP.F16.F32.PACK_AB
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "+r"(h[i]) : "f"(__uint_as_float(h[i])), "f"(v[2 * i + 1]));
            } else if constexpr (OP == 1) {     // FHADD (f32 + f16 -> f32)
                asm volatile("{\n.reg .b16 l, u;\nmov.b32 {l, u}, %1;\nadd.rn.f32.f16 %0, l, %0;\n}\n" : "+f"(v[i]) : "r"(h[i]));
            } else if constexpr (OP == 2) {     // FFMA2
                asm volatile("{\n.reg .b64 a, b;\nmov.b64 a, {%0, %1};\nmov.b64 b, {%2, %3};\nfma.rn.f32x2 a, a, b, b;\nmov.b64 {%0, %1}, a;\n}\n"
                             : "+f"(v[2 * i]), "+f"(v[2 * i + 1]) : "f"(0.999f), "f"(1.001f));
            } else if constexpr (OP == 3) {     // FFMA
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(0.999f), "f"(0.5f));
            } else if constexpr (OP == 4) {     // FMNMX3 |a| |b|
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(v[8 + i]), "f"(seed));
            } else if constexpr (OP == 5) {     // MUFU.LG2
                asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            } else if constexpr (OP == 6) {     // F2F.F16.F32 (scalar convert)
                unsigned short s;
                asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(s) : "f"(__uint_as_float(h[i])));
                h[i] = s;
            } else if constexpr (OP == 7) {     // PRMT
                asm volatile("prmt.b32 %0, %0, %1, 0x7610;" : "+r"(h[i]) : "r"(h[(i + 1) & 7]));
            } else if constexpr (OP == 8) {     // cvt.rz packed
                asm volatile("cvt.rz.f16x2.f32 %0, %1, %2;" : "+r"(h[i]) : "f"(__uint_as_float(h[i])), "f"(v[2 * i + 1]));
            } else if constexpr (OP == 9) {     // bf16x2 pack
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "+r"(h[i]) : "f"(__uint_as_float(h[i])), "f"(v[2 * i + 1]));
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += v[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
        rate_kernel<OP><<<148, 128 * warps_per_smsp>>>(out, cyc, 1.5f);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double per = static_cast<double>(h[0]) / (kIters * 8.0 * warps_per_smsp);
        printf("%-28s %d warp(s) per sub-partition: %6.2f cycles per warp-instruction (%s)\n", name, warps_per_smsp, per, cudaGetErrorString(cudaGetLastError()));
    }
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("F2FP.F16.F32.PACK_AB (rn)");
    run<8>("F2FP.F16.F32.PACK_AB (rz)");
    run<9>("F2FP.BF16.F32.PACK_AB");
    run<6>("F2F.F16.F32 (scalar)");
    run<1>("FHADD f32+f16");
    run<2>("FFMA2");
    run<3>("FFMA");
    run<4>("FMNMX3");
    run<5>("MUFU.LG2");
    run<7>("PRMT");
    return 0;
}
