// Microbenchmark: tcgen05.mma.kind::f16 issue rate as a function of N (M = 128, K = 16 per instruction),
// A operand in tensor memory, B operand in shared memory (K-major, no swizzle), plus tcgen05.ld rate.
// Also checks which N are legal at M = 128 (an illegal shape traps -> the launch reports an error).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// grid = 148 CTAs so every SM runs the same loop (shared clocks / power state as in the real kernel)
template <int N, int nd, int ss>
__global__ void __launch_bounds__(160) mma_rate_kernel(int reps, int ld_reps, long long* out) {
    constexpr int ksteps = 7;
    extern __shared__ __align__(128) unsigned char smem[];
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
    // B: ksteps * 16 rows of K, N columns, all zero (finite)
    for (int i = tid; i < ksteps * 16 * 256 / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    for (int i = tid; i < ksteps * 4096 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + 64 * 1024)[i] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (warp < 4) {  // zero all of tensor memory (A columns must be finite)
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < 512; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    long long t_mma = 0, t_ld = 0;
    if (warp == 4 && lane == 0) {
        const uint32_t lbo = N * 16, sbo = 128;
        const uint64_t desc0 = (uint64_t)((smem_u32(smem) & 0x3ffff) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t a_desc0 = (uint64_t)(((smem_u32(smem) + 64 * 1024) & 0x3ffff) >> 4) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        uint64_t bd[ksteps], ad[ksteps];
#pragma unroll
        for (int s = 0; s < ksteps; ++s) { bd[s] = desc0 + (uint64_t)((2 * lbo * s) >> 4); ad[s] = a_desc0 + (uint64_t)((2 * 2048 * s) >> 4); }
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int s = 0; s < ksteps; ++s) {
                const uint32_t a_t = tmem + 8 * s, d_t = tmem + 256 + (uint32_t)((s % nd) * (256 / nd));
                if (ss) {
                    asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(d_t), "l"(ad[s]), "l"(bd[s]), "r"(idesc) : "memory");
                } else {
                    asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(d_t), "r"(a_t), "l"(bd[s]), "r"(idesc) : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
        asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DN;\n bra W;\n DN:\n}"
                     ::"r"(smem_u32(&s_bar)), "r"(0u) : "memory");
        t_mma = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 256;
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int r = 0; r < ld_reps; ++r) {
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                           "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                           "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr + 32 * (r & 3)) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= v[i];
        }
        t_ld = clock64() - t0;
        if (acc == 0x12345678u) out[3] = 1;
    }
    if (blockIdx.x == 0) {
        if (warp == 4 && lane == 0) out[0] = t_mma;
        if (tid == 0) out[1] = t_ld;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
    long long* d_out;
    cudaMalloc(&d_out, 4 * sizeof(long long));
    const int reps = 400, ld_reps = 2000;
#define RUN(N, ND, SS) do { \
        cudaFuncSetAttribute(mma_rate_kernel<N, ND, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); \
        cudaMemset(d_out, 0, 4 * sizeof(long long)); \
        mma_rate_kernel<N, ND, SS><<<148, 160, 96 * 1024>>>(reps, ld_reps, d_out); \
        cudaError_t e = cudaDeviceSynchronize(); \
        long long h[4] = {0, 0, 0, 0}; \
        if (e == cudaSuccess) cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost); \
        printf("N=%3d nd=%d ss=%d: %s  cycles/mma=%.1f (floor 128*N/256=%.1f)  cycles/ld.x32=%.1f\n", N, ND, SS, cudaGetErrorString(e), \
               (double)h[0] / (7 * reps), 128.0 * N / 256.0, (double)h[1] / ld_reps); \
        if (e != cudaSuccess) return 1; } while (0)
    const int sel = argc > 1 ? atoi(argv[1]) : 0;
    if (sel == 0) { RUN(16, 1, 0); RUN(32, 1, 0); RUN(48, 1, 0); RUN(64, 1, 0); RUN(96, 1, 0); RUN(112, 1, 0); RUN(128, 1, 0); RUN(256, 1, 0); }
    if (sel == 1) { RUN(16, 4, 0); RUN(32, 4, 0); RUN(48, 4, 0); RUN(64, 4, 0); RUN(112, 2, 0); RUN(32, 7, 0); }
    if (sel == 2) { RUN(16, 1, 1); RUN(32, 1, 1); RUN(64, 1, 1); RUN(112, 1, 1); RUN(128, 1, 1); RUN(256, 1, 1); RUN(32, 4, 1); RUN(64, 4, 1); }
    if (sel == 3) { RUN(104, 1, 0); }
    if (sel == 4) { RUN(56, 1, 0); }
    return 0;
}
