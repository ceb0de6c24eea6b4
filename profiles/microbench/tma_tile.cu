// How long does one 128-frame audio tile take to reach shared memory?  Issue-to-arrival time of the tile copy of the
// tcgen05 log-mel kernel (two half tiles of 66 rows x 164 words from the overlapping-row 4-D tensor map) against other ways
// to move the same samples: one box for the whole tile, quarter boxes, a dense 2-D box, plain bulk copies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_tile tma_tile.cu -lcuda && ./tma_tile
// Every CTA walks tiles blockIdx.x, + gridDim.x, ... of a [clips, 480000] float32 batch like the real kernel, one elected
// thread issues the copies of a tile and waits for them (optionally after asking L2 for the next tile), clock64 around it.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

constexpr int kHop = 160, kPitch = 164, kHalfRows = 66, kTileRows = 130, kFrames = 128;
constexpr int kHalfBytes = kHalfRows * kPitch * 4, kHalfStride = (kHalfBytes + 127) / 128 * 128;
constexpr int kSamples = 480000, kTilesPerClip = 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma4(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma4_prefetch(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma3(const CUtensorMap* map, void* dst, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk(const void* src, void* dst, uint64_t* bar, uint32_t bytes) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// MODE 0: two half boxes {164,1,66,1} (the kernel's)   1: one box {164,1,130,1}   2: four quarter boxes {164,1,34,1} (34-row quarters)
//      3: dense 3-D box {160, 66, 1} x 2 (pitch 160: bank conflicts in the real kernel - speed of the copy only)
//      4: two bulk copies of 66 x 640 contiguous bytes   5: 130 bulk copies of one row (640 B) each to pitch-164 rows
template <int MODE>
__global__ void __launch_bounds__(128) tile_kernel(const __grid_constant__ CUtensorMap map4, const __grid_constant__ CUtensorMap map3, const float* audio,
                                                   int clips, int prefetch, long long* cycles, int* counts) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int total = clips * kTilesPerClip;
    uint32_t parity = 0;
    long long sum = 0;
    int n = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int clip = tile / kTilesPerClip, t0 = (tile % kTilesPerClip) * kFrames;
        const int next = tile + gridDim.x, nclip = next / kTilesPerClip, nt0 = (next % kTilesPerClip) * kFrames;
        // interior tiles only (the kernel patches a clip's ends): rows t0 - 2 .. of the tensor map must exist
        if (t0 == 0 || t0 + kFrames + 2 > 2998) continue;
        const long long t_start = clock64();
        if (MODE == 0) {
            expect(&bar, 2 * kHalfBytes);
            tma4(&map4, smem, &bar, 0, 3, t0 - 2, clip);
            tma4(&map4, smem + kHalfStride, &bar, 0, 3, t0 - 2 + 64, clip);
        } else if (MODE == 1) {
            expect(&bar, kTileRows * kPitch * 4);
            tma4(&map4, smem, &bar, 0, 3, t0 - 2, clip);
        } else if (MODE == 2) {
            expect(&bar, 4 * 34 * kPitch * 4);
            for (int q = 0; q < 4; ++q) tma4(&map4, smem + q * 22400, &bar, 0, 3, t0 - 2 + 32 * q, clip);
        } else if (MODE == 3) {
            expect(&bar, 2 * kHalfRows * kHop * 4);
            tma3(&map3, smem, &bar, 0, t0 - 2, clip);   // (row-aligned start: the copy speed does not depend on the 120-sample offset much)
            tma3(&map3, smem + kHalfStride, &bar, 0, t0 - 2 + 64, clip);
        } else if (MODE == 4) {
            const float* src = audio + static_cast<size_t>(clip) * kSamples + static_cast<size_t>(t0) * kHop - 200;
            expect(&bar, 2 * kHalfRows * kHop * 4);
            bulk(src, smem, &bar, kHalfRows * kHop * 4);
            bulk(src + 64 * kHop, smem + kHalfStride, &bar, kHalfRows * kHop * 4);
        } else {
            const float* src = audio + static_cast<size_t>(clip) * kSamples + static_cast<size_t>(t0) * kHop - 200;
            expect(&bar, kTileRows * kHop * 4);
            for (int r = 0; r < kTileRows; ++r) bulk(src + r * kHop, smem + r * kPitch * 4, &bar, kHop * 4);
        }
        if (prefetch && next < total && nt0 != 0 && nt0 + kFrames + 2 <= 2998) {
            if (MODE <= 2) { tma4_prefetch(&map4, 0, 3, nt0 - 2, nclip); tma4_prefetch(&map4, 0, 3, nt0 - 2 + 64, nclip); }
            else bulk_prefetch(audio + static_cast<size_t>(nclip) * kSamples + static_cast<size_t>(nt0) * kHop - 200, 132 * kHop * 4);
        }
        mbar_wait(&bar, parity);
        parity ^= 1u;
        sum += clock64() - t_start;
        ++n;
        if (prefetch > 1) __nanosleep(prefetch);   // leave the prefetch time to land (the real kernel computes for ~8000 cycles here)
    }
    cycles[blockIdx.x] = sum;
    counts[blockIdx.x] = n;
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE>
void run(const char* name, const CUtensorMap& m4, const CUtensorMap& m3, const float* audio, int clips, long long* d_cyc, int* d_cnt) {
    const int smem_bytes = 4 * 22400 + 1024;
    cudaFuncSetAttribute(tile_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    for (int grid : {4, 148})
        for (int prefetch : {0, 1, 4000}) {
            tile_kernel<MODE><<<grid, 128, smem_bytes>>>(m4, m3, audio, clips, prefetch, d_cyc, d_cnt);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            int c[148];
            cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
            cudaMemcpy(c, d_cnt, sizeof(int) * grid, cudaMemcpyDeviceToHost);
            double s = 0, n = 0;
            for (int i = 0; i < grid; ++i) { s += h[i]; n += c[i]; }
            printf("%-44s clips %4d grid %3d prefetch %-4d: %7.0f cycles per tile (%s)\n", name, clips, grid, prefetch, n > 0 ? s / n : 0.0, cudaGetErrorString(e));
        }
}

int main() {
    for (int clips : {8, 256}) {   // 15 MB (stays in L2) / 491 MB (streams from HBM)
        float* audio;
        cudaMalloc(&audio, static_cast<size_t>(clips) * kSamples * 4);
        cudaMemset(audio, 0, static_cast<size_t>(clips) * kSamples * 4);
        long long* d_cyc;
        int* d_cnt;
        cudaMalloc(&d_cyc, 148 * sizeof(long long));
        cudaMalloc(&d_cnt, 148 * sizeof(int));
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
        CUtensorMap m4, m3;
        {
            const cuuint64_t rows = (kSamples - 284) / kHop + 1;
            const cuuint64_t dims[4] = {kPitch, 4, rows, static_cast<cuuint64_t>(clips)};
            const cuuint64_t strides[3] = {kHop, kHop * 4, static_cast<cuuint64_t>(kSamples) * 4};
            const cuuint32_t box[4] = {kPitch, 1, kHalfRows, 1};
            const cuuint32_t elem[4] = {1, 1, 1, 1};
            // (one map serves the 66-row halves; the whole-tile and quarter modes need their own box: encode per mode below)
            if (encode(&m4, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, audio, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode 4d failed\n"); return 1; }
            const cuuint64_t dims3[3] = {kHop, kSamples / kHop, static_cast<cuuint64_t>(clips)};
            const cuuint64_t strides3[2] = {kHop * 4, static_cast<cuuint64_t>(kSamples) * 4};
            const cuuint32_t box3[3] = {kHop, kHalfRows, 1};
            const cuuint32_t elem3[3] = {1, 1, 1};
            if (encode(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, audio, dims3, strides3, box3, elem3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode 3d failed\n"); return 1; }
        }
        run<0>("two half boxes {164,1,66,1} (kernel)", m4, m3, audio, clips, d_cyc, d_cnt);
        {
            CUtensorMap m = m4;
            const cuuint64_t rows = (kSamples - 284) / kHop + 1;
            const cuuint64_t dims[4] = {kPitch, 4, rows, static_cast<cuuint64_t>(clips)};
            const cuuint64_t strides[3] = {kHop, kHop * 4, static_cast<cuuint64_t>(kSamples) * 4};
            const cuuint32_t elem[4] = {1, 1, 1, 1};
            const cuuint32_t box1[4] = {kPitch, 1, kTileRows, 1};
            encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, audio, dims, strides, box1, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            run<1>("one box {164,1,130,1}", m, m3, audio, clips, d_cyc, d_cnt);
            const cuuint32_t box2[4] = {kPitch, 1, 34, 1};
            encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, audio, dims, strides, box2, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            run<2>("four quarter boxes {164,1,34,1}", m, m3, audio, clips, d_cyc, d_cnt);
        }
        run<3>("two dense 3-D boxes {160,66,1}", m4, m3, audio, clips, d_cyc, d_cnt);
        run<4>("two bulk copies of 42,240 contiguous bytes", m4, m3, audio, clips, d_cyc, d_cnt);
        run<5>("130 bulk copies of one row to pitch 164", m4, m3, audio, clips, d_cyc, d_cnt);
        cudaFree(audio);
        cudaFree(d_cyc);
        cudaFree(d_cnt);
    }
    return 0;
}
