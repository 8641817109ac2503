// How fast can one SM write a bf16 output tile?  The GEMM epilogue's global stores cost ~2 us per 128 x 128 tile
// (tools/gemm_fixed.py parts: 14.5 us with them, 8.6 us without, 3 tiles per CTA); this probe times store patterns alone.
// Every CTA writes `tiles` tiles of 128 rows x 128 bf16 columns of a [M, N] row-major matrix (N = 1280), from registers /
// shared memory, with:   0: 64 B per row x 8 rows per warp instruction (the epilogue's pattern)
//                        1: 128 B per row x 4 rows        2: 256 B per row x 2 rows (one full tile row per 16 lanes)
//                        3: 1-D TMA bulk stores, one 256 B row segment per lane      4: pattern 0 with st.global.cs (streaming)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/store_probe tools/store_probe.cu && gpurun_out/store_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(256) probe(__nv_bfloat16* C, int M, int N, int tiles_total) {
    extern __shared__ __align__(128) uint8_t sm[];          // one 128 x 128 bf16 tile (32 KB), row = 256 B
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = N / 128;
    if (MODE == 3) {
        for (int i = threadIdx.x; i < 128 * 128 / 8; i += 256) reinterpret_cast<uint4*>(sm)[i] = make_uint4(i, i, i, i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
        const int m0 = (t / n_tiles) * 128, n0 = (t % n_tiles) * 128;
        const uint4 v = make_uint4(t, lane, warp, 7);
        if (MODE == 0 || MODE == 4) {
            // warp w: lane quarter q = w & 3 (32 rows), column chunks (w >> 2), (w >> 2) + 2 of 32 columns; 4 lanes per row
#pragma unroll 1
            for (int c = (warp >> 2); c < 4; c += 2)
#pragma unroll
                for (int st = 0; st < 4; ++st) {
                    const int row = m0 + (warp & 3) * 32 + st * 8 + (lane >> 2), col = n0 + c * 32 + (lane & 3) * 8;
                    uint4* dst = reinterpret_cast<uint4*>(C + (size_t)row * N + col);
                    if (MODE == 4) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    else *dst = v;
                }
        } else if (MODE == 1) {
            // 8 lanes per row (128 B), 4 rows per instruction; warp w: rows (w & 3) * 32 .., column half w >> 2
#pragma unroll
            for (int st = 0; st < 8; ++st) {
                const int row = m0 + (warp & 3) * 32 + st * 4 + (lane >> 3), col = n0 + (warp >> 2) * 64 + (lane & 7) * 8;
                *reinterpret_cast<uint4*>(C + (size_t)row * N + col) = v;
            }
        } else if (MODE == 2) {
            // 16 lanes per row (256 B = the whole tile row), 2 rows per instruction; warp w: rows w * 16 ..
#pragma unroll
            for (int st = 0; st < 8; ++st) {
                const int row = m0 + warp * 16 + st * 2 + (lane >> 4), col = n0 + (lane & 15) * 8;
                *reinterpret_cast<uint4*>(C + (size_t)row * N + col) = v;
            }
        } else {
            // one lane per row: 1-D bulk store of the 256 B row segment from shared memory; 128 rows = 4 warps x 32 lanes
            if (warp < 4) {
                const int row = warp * 32 + lane;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;"
                             :: "l"(C + (size_t)(m0 + row) * N + n0), "r"(smem_u32(sm + row * 256)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (MODE == 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE>
static void run(const char* name, __nv_bfloat16* C, int M, int N) {
    const int tiles = (M / 128) * (N / 128);
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148, 296}) {
        for (int i = 0; i < 3; ++i) probe<MODE><<<grid, 256, 32768>>>(C, M, N, tiles);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) probe<MODE><<<grid, 256, 32768>>>(C, M, N, tiles);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / 20, bytes = (double)M * N * 2;
        printf("%-28s M=%5d grid %3d: %7.2f us per launch  %7.1f GB/s  (%s)\n", name, M, grid, us, bytes / us / 1e3, cudaGetErrorString(cudaGetLastError()));
    }
}

int main() {
    const int N = 1280;
    for (int M : {4096, 32768}) {
        __nv_bfloat16* C;
        cudaMalloc(&C, (size_t)M * N * 2);
        run<0>("64B x 8 rows (epilogue)", C, M, N);
        run<4>("64B x 8 rows, st.cs", C, M, N);
        run<1>("128B x 4 rows", C, M, N);
        run<2>("256B x 2 rows", C, M, N);
        run<3>("TMA bulk 256B rows", C, M, N);
        cudaFree(C);
    }
    return 0;
}
