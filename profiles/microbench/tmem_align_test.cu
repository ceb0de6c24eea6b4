// Which tensor-memory column alignments do tcgen05.st / tcgen05.ld (32x32b.xN) and the A operand of
// tcgen05.mma (A in TMEM) need?  One test per process (argv[1]); a failing one reports a sticky CUDA error.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(160) k(int test, int col, int dcol, uint32_t N, int boff, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char dyn[];
    __half* s_B = reinterpret_cast<__half*>(dyn + boff);
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0 && test == 3) out[3] = smem_u32(dyn);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 104 * 16; i += blockDim.x) s_B[i] = __float2half_rn(1.0f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    if (warp < 4) {
        for (int c = 0; c < 512; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(lane_addr + c), "r"(0u) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (test == 1) {  // st.x4 at an arbitrary column, read back with ld.x4 at the same column
            uint32_t a = 0x11110000u + tid, r0, r1, r2, r3;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lane_addr + col), "r"(a), "r"(a + 1), "r"(a + 2), "r"(a + 3) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(lane_addr + col) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (tid == 5) { out[0] = r0; out[1] = r3; out[2] = a; }
        }
        if (test == 2) {  // ld.x32 at an arbitrary column
            uint32_t a = 0x22220000u + tid;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_addr + col + 31), "r"(a) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                           "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                           "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                           "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(lane_addr + col) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (tid == 5) { out[0] = v[31]; out[1] = v[0]; out[2] = a; }
        }
        if (test == 3) {  // A operand (K = 16 halves = 8 columns) at an arbitrary column: ones in those 8 columns
            const __half2 one = __floats2half2_rn(1.0f, 1.0f);
            const uint32_t o = *reinterpret_cast<const uint32_t*>(&one);
            for (int c = 0; c < 8; ++c) asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(lane_addr + col + c), "r"(o) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (test == 3) {
        if (tid == 128) {
            const uint32_t lbo = N * 16, sbo = 128;
            const uint64_t desc = (uint64_t)((smem_u32(s_B) & 0x3ffff) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 0;" ::"r"(tmem + dcol), "r"(tmem + col), "l"(desc), "r"(idesc) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
        }
        if (warp < 4) {
            asm volatile("{\n .reg .pred p;\n W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DN;\n bra W;\n DN:\n}" ::"r"(smem_u32(&s_bar)), "r"(0u) : "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r0, r1, r2, r3;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(lane_addr + dcol + N - 4) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (tid == 5) { out[0] = r0; out[1] = r3; out[2] = __float_as_uint(16.0f); }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
    const int test = atoi(argv[1]), col = atoi(argv[2]), dcol = argc > 3 ? atoi(argv[3]) : 448, N = argc > 4 ? atoi(argv[4]) : 32, boff = argc > 5 ? atoi(argv[5]) : 0;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    uint32_t* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    k<<<1, 160, 200 * 1024>>>(test, col, dcol, (uint32_t)N, boff, d);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t h[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("test %d col %d dcol %d N %d boff %d: %s  got %08x %08x expect %08x dynbase %u\n", test, col, dcol, N, boff, cudaGetErrorString(e), h[0], h[1], h[2], h[3]);
    return 0;
}
