// TMEM / MUFU micro-probes for the attention redesign (sm_100a).  Build + run:  tools/run_tmem_probe.sh
//   1. tcgen05.mma with the A operand in TENSOR MEMORY (P V with P written by tcgen05.st): layout check against a host reference
//   2. tcgen05.ld / tcgen05.st throughput per SM for 4 / 8 / 16 warps (bytes per clock)
//   3. ex2.approx throughput per SM, alone and mixed with FMA-pipe work
#include "../aozora_sdxl_training_b200/csrc/common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

using namespace aoz;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// ---- 1. P (TMEM) x V (smem, [keys, 64] rows of 128 B, 128B swizzle) -> O (TMEM) --------------------------------
__global__ void __launch_bounds__(128) ts_mma_kernel(const __nv_bfloat16* __restrict__ Pm, const __nv_bfloat16* __restrict__ V,
                                                      float* __restrict__ O, int a_col, int d_col) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 256);
    // V tile into smem with the TMA 128B-swizzle layout: row = key (128 rows), 8 chunks of 16 bytes
    for (int i = threadIdx.x; i < 128 * 8; i += 128) {
        const int row = i >> 3, ch = i & 7;
        *reinterpret_cast<uint4*>(smem + sw128_offset(row, ch)) = *reinterpret_cast<const uint4*>(V + row * 64 + ch * 8);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    // P row r: 128 bf16 = 64 packed words -> TMEM columns [a_col, a_col + 64) of lane r
    uint32_t w[32];
    for (int half = 0; half < 2; ++half) {
        for (int c = 0; c < 32; ++c) w[c] = reinterpret_cast<const uint32_t*>(Pm + r * 128)[half * 32 + c];
        tmem_st32(tmem + lane_off + a_col + half * 32, w);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);            // A K-major (TMEM), B MN-major ([k rows, 64 n])
        const uint32_t sV = smem_u32(smem);
        for (int k = 0; k < 8; ++k)
            umma_bf16_ts(tmem + d_col, tmem + a_col + k * 8, make_sdesc_sw128(sV + k * 2048, 8192, 1024), idesc, k > 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int c = 0; c < 2; ++c) {
        tmem_ld32(tmem + lane_off + d_col + c * 32, v);
        tc_wait_ld();
        for (int e = 0; e < 32; ++e) O[r * 64 + c * 32 + e] = __uint_as_float(v[e]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// ---- 2. TMEM load / store throughput ---------------------------------------------------------------------------
// mode 0: x32 loads, wait after each; 1: x32 loads, wait after every 4; 2: x16 loads wait each; 3: x32 stores (wait every 4)
template <int MODE>
__global__ void __launch_bounds__(512) tmem_bw_kernel(int iters, long long* __restrict__ cycles, uint32_t* __restrict__ sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128) % 512;
    uint32_t v[32];
    for (int e = 0; e < 32; ++e) v[e] = threadIdx.x + e;
    tmem_st32(base, v); tmem_st32(base + 32, v); tmem_st32(base + 64, v); tmem_st32(base + 96, v);
    tc_wait_st();
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) { tmem_ld32(base + c * 32, v); tc_wait_ld(); acc ^= v[0] ^ v[31]; }
        } else if (MODE == 1) {
            uint32_t a[32], b[32], c2[32];
            tmem_ld32(base, v); tmem_ld32(base + 32, a); tmem_ld32(base + 64, b); tmem_ld32(base + 96, c2);
            tc_wait_ld();
            acc ^= v[0] ^ a[1] ^ b[2] ^ c2[3];
        } else if (MODE == 2) {
#pragma unroll
            for (int c = 0; c < 8; ++c) { tmem_ld16(base + c * 16, v); tc_wait_ld(); acc ^= v[0] ^ v[15]; }
        } else {
            v[0] = acc + i;
            tmem_st32(base, v); tmem_st32(base + 32, v); tmem_st32(base + 64, v); tmem_st32(base + 96, v);
            tc_wait_st();
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---- 3. ex2 / FMA mix ------------------------------------------------------------------------------------------
// per iteration and thread: NEX ex2.approx + NFMA dependent-free FMAs (16 independent chains)
template <int NEX, int NFMA>
__global__ void __launch_bounds__(512) mufu_kernel(int iters, long long* __restrict__ cycles, float* __restrict__ sink) {
    float x[16];
    for (int e = 0; e < 16; ++e) x[e] = -0.001f * (threadIdx.x + e);
    float f = 1.0f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int e = 0; e < NEX; ++e) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[e & 15]));
#pragma unroll
        for (int e = 0; e < NFMA; ++e) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[(e + 7) & 15]) : "f"(f));
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0.f;
    for (int e = 0; e < 16; ++e) s += x[e];
    if (s == 1234.5f) sink[0] = s;
}

static double avg_cycles(long long* d, int n) {
    std::vector<long long> h(n);
    CK(cudaMemcpy(h.data(), d, n * sizeof(long long), cudaMemcpyDeviceToHost));
    double s = 0;
    for (auto c : h) s += (double)c;
    return s / n;
}

int main() {
    // ---------- 1 ----------
    {
        std::vector<__nv_bfloat16> P(128 * 128), V(128 * 64);
        std::vector<float> Pf(128 * 128), Vf(128 * 64);
        srand(1);
        for (int i = 0; i < 128 * 128; ++i) { float x = (rand() % 2001 - 1000) / 1000.0f; P[i] = __float2bfloat16(x); Pf[i] = __bfloat162float(P[i]); }
        for (int i = 0; i < 128 * 64; ++i) { float x = (rand() % 2001 - 1000) / 1000.0f; V[i] = __float2bfloat16(x); Vf[i] = __bfloat162float(V[i]); }
        __nv_bfloat16 *dP, *dV; float* dO;
        CK(cudaMalloc(&dP, P.size() * 2)); CK(cudaMalloc(&dV, V.size() * 2)); CK(cudaMalloc(&dO, 128 * 64 * 4));
        CK(cudaMemcpy(dP, P.data(), P.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dV, V.data(), V.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(ts_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 18 * 1024));
        for (int trial = 0; trial < 2; ++trial) {
            const int a_col = trial == 0 ? 0 : 64, d_col = trial == 0 ? 128 : 192;
            CK(cudaMemset(dO, 0, 128 * 64 * 4));
            ts_mma_kernel<<<1, 128, 18 * 1024>>>(dP, dV, dO, a_col, d_col);
            CK(cudaDeviceSynchronize());
            std::vector<float> O(128 * 64);
            CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < 64; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 128; ++k) ref += (double)Pf[r * 128 + k] * Vf[k * 64 + n];
                    maxerr = fmax(maxerr, fabs(ref - O[r * 64 + n]));
                }
            printf("ts_mma (A in TMEM cols %d.., D cols %d..): max abs err vs host %.3e  -> %s\n", a_col, d_col, maxerr, maxerr < 1e-3 ? "LAYOUT OK" : "MISMATCH");
        }
    }
    // ---------- 2 ----------
    long long* dc; uint32_t* dsink; float* fsink;
    CK(cudaMalloc(&dc, 148 * 8)); CK(cudaMalloc(&dsink, 4)); CK(cudaMalloc(&fsink, 4));
    const int iters = 2000;
    for (int warps : {4, 8, 16}) {
        const double bytes = (double)iters * warps * 4 * 4096;          // per CTA: 4 x (32 lanes x 32 cols x 4 B) per warp-iteration
        tmem_bw_kernel<0><<<148, warps * 32>>>(iters, dc, dsink); CK(cudaDeviceSynchronize());
        double c0 = avg_cycles(dc, 148);
        tmem_bw_kernel<1><<<148, warps * 32>>>(iters, dc, dsink); CK(cudaDeviceSynchronize());
        double c1 = avg_cycles(dc, 148);
        tmem_bw_kernel<2><<<148, warps * 32>>>(iters, dc, dsink); CK(cudaDeviceSynchronize());
        double c2 = avg_cycles(dc, 148);
        tmem_bw_kernel<3><<<148, warps * 32>>>(iters, dc, dsink); CK(cudaDeviceSynchronize());
        double c3 = avg_cycles(dc, 148);
        printf("tmem %2d warps/SM: ld.x32 wait-each %.1f B/clk | ld.x32 4-deep %.1f B/clk | ld.x16 wait-each %.1f B/clk | st.x32 4-deep %.1f B/clk\n",
               warps, bytes / c0, bytes / c1, bytes / c2, bytes / c3);
    }
    // ---------- 3 ----------
    for (int warps : {4, 8, 16}) {
        const int it = 4000;
        mufu_kernel<16, 0><<<148, warps * 32>>>(it, dc, fsink); CK(cudaDeviceSynchronize());
        double a = avg_cycles(dc, 148);
        mufu_kernel<0, 64><<<148, warps * 32>>>(it, dc, fsink); CK(cudaDeviceSynchronize());
        double b = avg_cycles(dc, 148);
        mufu_kernel<16, 64><<<148, warps * 32>>>(it, dc, fsink); CK(cudaDeviceSynchronize());
        double c = avg_cycles(dc, 148);
        mufu_kernel<16, 96><<<148, warps * 32>>>(it, dc, fsink); CK(cudaDeviceSynchronize());
        double d = avg_cycles(dc, 148);
        printf("alu %2d warps/SM: ex2 only %.2f /clk/SM | fma only %.1f /clk/SM | 16 ex2 + 64 fma: %.2f ex2/clk (%.1f fma/clk) | 16 ex2 + 96 fma: %.2f ex2/clk (%.1f fma/clk)\n",
               warps, (double)it * 16 * warps * 32 / a, (double)it * 64 * warps * 32 / b, (double)it * 16 * warps * 32 / c,
               (double)it * 64 * warps * 32 / c, (double)it * 16 * warps * 32 / d, (double)it * 96 * warps * 32 / d);
    }
    return 0;
}
