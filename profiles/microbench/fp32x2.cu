// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <bool kPacked>
__global__ void __launch_bounds__(256) fma_loop(float2* out, int iters, float seed) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x, seed - i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (kPacked) a[i] = __ffma2_rn(a[i], m, c);
            else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
        }
    }
    float2 s = a[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, iters = 20000;
    float2* out; cudaMalloc(&out, sizeof(float2) * blocks * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (packed) fma_loop<true><<<blocks, 256>>>(out, iters, 1.f); else fma_loop<false><<<blocks, 256>>>(out, iters, 1.f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fma = double(blocks) * 256 * iters * 16;  // scalar FMAs executed
            if (rep == 2) printf("%s: %.3f ms, %.1f TFLOP/s fp32, %.1f fma/clk/SM @1.965GHz\n", packed ? "FFMA2 (f32x2)" : "FFMA  (scalar)",
                                 ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.965e9);
        }
    }
    return 0;
}
