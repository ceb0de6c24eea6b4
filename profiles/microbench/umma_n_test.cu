// Mechanics test for the tcgen05 path: A operand written to TMEM with tcgen05.st (fp16 pairs packed in
// 32-bit columns), B operand in shared memory (K-major, no swizzle, 8x16B core matrices), fp32
// accumulator in TMEM, tcgen05.commit -> mbarrier, tcgen05.ld epilogue.  D[128,32] = A[128,64] * B[32,64]^T.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ts_test umma_ts_test.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int M = 128, N = NVAL, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(160) umma_ts_kernel(const float* A, const float* B, float* D) {
    __shared__ __align__(128) __half s_B[N * K];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B -> shared, canonical K-major no-swizzle: strips [k/8][n][8 halves]
    for (int i = tid; i < N * K; i += blockDim.x) {
        const int n = i / K, k = i % K;
        s_B[(k / 8) * (N * 8) + n * 8 + (k % 8)] = __float2half_rn(B[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp < 4) {  // A row = TMEM lane = tid; 64 halves = 32 columns at [0, 32)
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < 32; c0 += 8) {
            uint32_t v[8];
            for (int c = 0; c < 8; ++c) {
                const __half2 h = __floats2half2_rn(A[tid * K + 2 * (c0 + c)], A[tid * K + 2 * (c0 + c) + 1]);
                v[c] = *reinterpret_cast<const uint32_t*>(&h);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                         ::"r"(taddr + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
        }
        for (int c0 = 0; c0 < N + 8; c0 += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + 64 + c0), "r"(__float_as_uint(12345.0f)) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic smem writes -> async proxy (UMMA) reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
            // smem descriptor: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout NONE
            const uint32_t lbo = N * 16, sbo = 128;
            const uint64_t desc0 = (uint64_t)((smem_u32(s_B) & 0x3ffff) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
                                   ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // f16 x f16 -> f32, K-major both
            for (int s = 0; s < K / 16; ++s) {
                const uint64_t desc = desc0 + (uint64_t)((2 * lbo * s) >> 4);
                const uint32_t a_t = tmem + 8 * s, d_t = tmem + 64, acc = s > 0;
                asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                             ::"r"(d_t), "r"(a_t), "l"(desc), "r"(idesc), "r"(acc) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
        }
    }
    if (warp < 4) {
        asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DN;\n bra W;\n DN:\n}"
                     ::"r"(smem_u32(&s_bar)), "r"(0u) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 64;
        for (int c0 = 0; c0 < N + 8; c0 += 8) {
            uint32_t r[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr + c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int c = 0; c < 8; ++c) D[tid * (N + 8) + c0 + c] = __uint_as_float(r[c]);
        }

    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * (N + 8)), ref(M * N);
    for (int mode = 0; mode < 2; ++mode) {
        srand(1);
        for (int i = 0; i < M * K; ++i) A[i] = mode == 0 ? (float)(i / K) + (float)(i % K) / 64.0f : (float)(rand() % 2001 - 1000) / 1000.0f;
        for (int i = 0; i < N * K; ++i) B[i] = mode == 0 ? ((i % K) == ((i / K) % 32) * 2 ? 1.0f : 0.0f) : (float)(rand() % 2001 - 1000) / 1000.0f;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double s = 0;
                for (int k = 0; k < K; ++k) s += (double)__half2float(__float2half_rn(A[m * K + k])) * (double)__half2float(__float2half_rn(B[n * K + k]));
                ref[m * N + n] = (float)s;
            }
        float *dA, *dB, *dD;
        cudaMalloc(&dA, sizeof(float) * M * K); cudaMalloc(&dB, sizeof(float) * N * K); cudaMalloc(&dD, sizeof(float) * M * (N + 8));
        cudaMemcpy(dA, A.data(), sizeof(float) * M * K, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), sizeof(float) * N * K, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, sizeof(float) * M * (N + 8));
        umma_ts_kernel<<<1, 160>>>(dA, dB, dD);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, sizeof(float) * M * (N + 8), cudaMemcpyDeviceToHost);
        double err = 0;
        double tail = 0; for (int m = 0; m < M; ++m) { for (int n = 0; n < N; ++n) err = fmax(err, fabs((double)D[m * (N + 8) + n] - ref[m * N + n])); for (int n = N; n < N + 8; ++n) tail = fmax(tail, fabs((double)D[m * (N + 8) + n] - 12345.0)); }
        printf("N=%d tail_dev=%.3e mode %d: cuda=%s max_abs_err=%.3e  D[0][0..3]=%g %g %g %g  D[5][0..3]=%g %g %g %g (ref %g %g %g %g)\n", N, tail, mode, cudaGetErrorString(e), err,
               D[0], D[1], D[2], D[3], D[5 * N], D[5 * N + 1], D[5 * N + 2], D[5 * N + 3], ref[5 * N], ref[5 * N + 1], ref[5 * N + 2], ref[5 * N + 3]);
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    return 0;
}
