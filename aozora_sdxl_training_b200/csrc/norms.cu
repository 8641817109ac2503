// GroupNorm(32)(+SiLU) and LayerNorm, forward and backward, over channels-last bf16 activations (sm_100a).
//
// Replaces the ATen group_norm / silu / layer_norm calls inside diffusers' ResnetBlock2D, Transformer2DModel
// and BasicTransformerBlock that the reference reaches through train.py:2760.
// Numerics contract: statistics and the normalise(+SiLU) math in fp32, ONE rounding to bf16
// on store -- what CUDA autocast does (fp32 norm output rounded when the next conv/linear casts its input).
// All kernels are HBM-streaming: 16-byte vector accesses along the contiguous channel dimension, fp32
// partial sums reduced with warp shuffles, deterministic two-level reductions (no atomics).
//   GroupNorm fwd: 4 B/element (stats pass re-read is served from L2 for SDXL sizes, <= 84 MB per tensor)
//   LayerNorm fwd: 4 B/element; bwd 6 B/element.
#include "common.cuh"

namespace aoz {

constexpr int GN_GROUPS = 32;
constexpr int GN_THREADS = 256;
constexpr int GN_MAX_CHUNKS = 64;

__device__ __forceinline__ void unpack8(const uint4& a, float* f) {
    f[0] = bf16lo(a.x); f[1] = bf16hi(a.x); f[2] = bf16lo(a.y); f[3] = bf16hi(a.y);
    f[4] = bf16lo(a.z); f[5] = bf16hi(a.z); f[6] = bf16lo(a.w); f[7] = bf16hi(a.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------
// GroupNorm statistics: x [NB, HW, C] -> partial [NB][chunks][GROUPS][2] (sum, sumsq)
// Row-lane partials are added to the per-channel shared-memory sums ONE LANE AT A TIME (plain adds between barriers): float
// atomics would make the summation order, and with it the low bits of every statistic and gradient, differ run to run.
// When there is more than one row lane every thread owns exactly one vector column (cols == vec_per_row).
__device__ __forceinline__ void gn_ordered_accumulate(float* sm, int C, int v, int tr, int row_lanes, bool active, const float* s, const float* q) {
    for (int l = 0; l < row_lanes; ++l) {
        if (active && tr == l) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { sm[v * 8 + e] += s[e]; sm[C + v * 8 + e] += q[e]; }
        }
        __syncthreads();
    }
}

// grid (chunks, NB).  Each thread owns one 8-channel vector column and walks pixels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int HW, int C, float* __restrict__ partial) {
    pdl_enter();
    extern __shared__ float sm[];          // [2][C] per-channel sums
    const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const int vec_per_row = C / 8;
    const int rows_per_chunk = (HW + chunks - 1) / chunks;
    const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
    for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) sm[i] = 0.f;
    __syncthreads();
    const __nv_bfloat16* xb = x + (size_t)n * HW * C;
    // threads are laid out as (row lane, vector column): column = tid % vec_cols_per_pass
    const int cols = min(vec_per_row, GN_THREADS);
    const int row_lanes = GN_THREADS / cols;
    const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
    float s[8], q[8];
    if (tr < row_lanes) {
        for (int v = tc; v < vec_per_row; v += cols) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
            int r = r0 + tr;
            for (; r + 3 * row_lanes < r1; r += 4 * row_lanes) {           // four rows in flight
                uint4 px[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) px[u] = ld_stream(xb + (size_t)(r + u * row_lanes) * C + v * 8);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float f[8];
                    unpack8(px[u], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
                }
            }
            for (; r < r1; r += row_lanes) {
                float f[8];
                unpack8(ld_stream(xb + (size_t)r * C + v * 8), f);
#pragma unroll
                for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
            }
            if (row_lanes == 1) {                                  // single writer per channel
#pragma unroll
                for (int e = 0; e < 8; ++e) { sm[v * 8 + e] = s[e]; sm[C + v * 8 + e] = q[e]; }
            }
        }
    }
    if (row_lanes > 1) gn_ordered_accumulate(sm, C, tc, tr, row_lanes, tr < row_lanes && tc < vec_per_row, s, q);
    __syncthreads();
    // channels -> groups (fixed order)
    const int cpg = C / GN_GROUPS;
    if (threadIdx.x < 2 * GN_GROUPS) {
        const int g = threadIdx.x % GN_GROUPS, which = threadIdx.x / GN_GROUPS;
        float a = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) a += sm[which * C + c];
        partial[(((size_t)n * chunks + chunk) * GN_GROUPS + g) * 2 + which] = a;
    }
}

// GroupNorm apply: y = silu?((x - mean) * rstd * gamma + beta); also writes mean/rstd [NB][GROUPS] (chunk 0 block)
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                const __nv_bfloat16* __restrict__ beta, const float* __restrict__ partial, int stat_chunks, int HW, int C,
                float eps, int silu, __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                float* __restrict__ rstd_out) {
    pdl_enter();
    __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    const int n = blockIdx.y;
    const int cpg = C / GN_GROUPS;
    if (threadIdx.x < GN_GROUPS) {
        float s = 0.f, q = 0.f;
        for (int k = 0; k < stat_chunks; ++k) {
            const float* p = partial + (((size_t)n * stat_chunks + k) * GN_GROUPS + threadIdx.x) * 2;
            s += p[0]; q += p[1];
        }
        const float cnt = (float)HW * (float)cpg;
        const float mean = s / cnt;
        const float var = fmaxf(q / cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        s_mean[threadIdx.x] = mean; s_rstd[threadIdx.x] = rstd;
        if (blockIdx.x == 0) { mean_out[n * GN_GROUPS + threadIdx.x] = mean; rstd_out[n * GN_GROUPS + threadIdx.x] = rstd; }
    }
    __syncthreads();
    // A thread owns one 8-channel vector column (its gamma / beta / mean / rstd stay in registers: no per-element group
    // division or shared-memory lookup in the loop) and walks its row lane of this block's pixel chunk, four rows in flight.
    const int vec_per_row = C / 8;
    const int cols = min(vec_per_row, GN_THREADS);
    const int row_lanes = GN_THREADS / cols;
    const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
    const int rows_per_chunk = (HW + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
    const __nv_bfloat16* xb = x + (size_t)n * HW * C;
    __nv_bfloat16* yb = y + (size_t)n * HW * C;
    if (tr >= row_lanes) return;
    for (int v = tc; v < vec_per_row; v += cols) {
        float gm[8], bt[8], sc[8], sh[8];
        unpack8(*reinterpret_cast<const uint4*>(gamma + v * 8), gm);
        unpack8(*reinterpret_cast<const uint4*>(beta + v * 8), bt);
#pragma unroll
        for (int e = 0; e < 8; ++e) {                          // z = x * sc + sh
            const int g = (v * 8 + e) / cpg;
            sc[e] = s_rstd[g] * gm[e];
            sh[e] = bt[e] - s_mean[g] * sc[e];
        }
        auto one = [&](const uint4& px, int r) {
            float f[8];
            unpack8(px, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float z = fmaf(f[e], sc[e], sh[e]);
                if (silu) z = z * sigmoidf_(z);
                f[e] = z;
            }
            st_stream(yb + (size_t)r * C + v * 8, pack8(f));
        };
        int r = r0 + tr;
        for (; r + 3 * row_lanes < r1; r += 4 * row_lanes) {
            uint4 px[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) px[u] = ld_stream(xb + (size_t)(r + u * row_lanes) * C + v * 8);
#pragma unroll
            for (int u = 0; u < 4; ++u) one(px[u], r + u * row_lanes);
        }
        for (; r < r1; r += row_lanes) one(ld_stream(xb + (size_t)r * C + v * 8), r);
    }
}

// GroupNorm backward pass 1: per-channel sums of dz and dz*xhat over a pixel chunk.
// partial [NB][chunks][2][C]
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_stats_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                    const float* __restrict__ mean, const float* __restrict__ rstd, int HW, int C, int silu,
                    float* __restrict__ partial) {
    pdl_enter();
    extern __shared__ float sm[];          // [2][C]
    __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS];
    const int n = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const int cpg = C / GN_GROUPS;
    if (threadIdx.x < GN_GROUPS) { s_mean[threadIdx.x] = mean[n * GN_GROUPS + threadIdx.x]; s_rstd[threadIdx.x] = rstd[n * GN_GROUPS + threadIdx.x]; }
    for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) sm[i] = 0.f;
    __syncthreads();
    const int vec_per_row = C / 8;
    const int rows_per_chunk = (HW + chunks - 1) / chunks;
    const int r0 = chunk * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
    const size_t base = (size_t)n * HW * C;
    const int cols = min(vec_per_row, GN_THREADS);
    const int row_lanes = GN_THREADS / cols;
    const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
    float a[8], b[8];
    if (tr < row_lanes) {
        for (int v = tc; v < vec_per_row; v += cols) {
            float gm[8], bt[8], mu[8], rs[8];
            unpack8(*reinterpret_cast<const uint4*>(gamma + v * 8), gm);
            unpack8(*reinterpret_cast<const uint4*>(beta + v * 8), bt);
#pragma unroll
            for (int e = 0; e < 8; ++e) { a[e] = 0.f; b[e] = 0.f; const int g = (v * 8 + e) / cpg; mu[e] = s_mean[g]; rs[e] = s_rstd[g]; }
            auto acc_row = [&](const uint4& px, const uint4& pd) {
                float fx[8], fd[8];
                unpack8(px, fx);
                unpack8(pd, fd);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = (fx[e] - mu[e]) * rs[e];
                    float dz = fd[e];
                    if (silu) {
                        const float z = xh * gm[e] + bt[e];
                        const float sg = sigmoidf_(z);
                        dz *= sg * (1.0f + z * (1.0f - sg));
                    }
                    a[e] += dz; b[e] = fmaf(dz, xh, b[e]);
                }
            };
            int r = r0 + tr;
            for (; r + row_lanes < r1; r += 2 * row_lanes) {               // two rows (four loads) in flight
                const size_t o0 = base + (size_t)r * C + v * 8, o1 = base + (size_t)(r + row_lanes) * C + v * 8;
                const uint4 x0 = ld_stream(x + o0), d0 = ld_stream(dy + o0), x1 = ld_stream(x + o1), d1 = ld_stream(dy + o1);
                acc_row(x0, d0);
                acc_row(x1, d1);
            }
            for (; r < r1; r += row_lanes) acc_row(ld_stream(x + base + (size_t)r * C + v * 8), ld_stream(dy + base + (size_t)r * C + v * 8));
            if (row_lanes == 1) {                                  // single writer per channel
#pragma unroll
                for (int e = 0; e < 8; ++e) { sm[v * 8 + e] = a[e]; sm[C + v * 8 + e] = b[e]; }
            }
        }
    }
    if (row_lanes > 1) gn_ordered_accumulate(sm, C, tc, tr, row_lanes, tr < row_lanes && tc < vec_per_row, a, b);
    __syncthreads();
    float* out = partial + ((size_t)n * chunks + chunk) * 2 * C;
    for (int i = threadIdx.x; i < 2 * C; i += GN_THREADS) out[i] = sm[i];
}

// pass 2a: one block per (group, image): reduce the chunk partials of the group's channels, write the per-channel
// sums chansum[n][2][C] and the group terms db = sum_c gamma*A, ds = sum_c gamma*B  -> group_terms[n][g][2]
__global__ void __launch_bounds__(128)
gn_bwd_group_kernel(const float* __restrict__ partial, const __nv_bfloat16* __restrict__ gamma, int chunks, int C,
                    float* __restrict__ chansum, float* __restrict__ group_terms) {
    pdl_enter();
    __shared__ float s_db[128], s_ds[128], s_a[128], s_b[128];
    const int g = blockIdx.x, n = blockIdx.y;
    const int cpg = C / GN_GROUPS;
    float db = 0.f, ds = 0.f;
    // thread t handles channel (t % cpg) of the group and chunk lanes (t / cpg) when cpg < 128
    const int lanes = max(1, 128 / cpg);
    const int cl = threadIdx.x % cpg, kl = threadIdx.x / cpg;
    if (kl < lanes) {
        for (int c0 = cl; c0 < cpg; c0 += (lanes > 1 ? cpg : 128)) {
            const int c = g * cpg + c0;
            float a = 0.f, b = 0.f;
            for (int k = kl; k < chunks; k += lanes) {
                const float* p = partial + ((size_t)n * chunks + k) * 2 * C;
                a += p[c]; b += p[C + c];
            }
            if (lanes == 1) {
                chansum[((size_t)n * 2 + 0) * C + c] = a;
                chansum[((size_t)n * 2 + 1) * C + c] = b;
                const float gm = __bfloat162float(gamma[c]);
                db = fmaf(gm, a, db); ds = fmaf(gm, b, ds);
            } else {                                   // cpg < 128: one channel per (cl), chunk lanes combined in fixed order below
                s_a[kl * cpg + cl] = a;
                s_b[kl * cpg + cl] = b;
            }
        }
    }
    if (lanes > 1) {
        __syncthreads();
        if (threadIdx.x < cpg) {
            const int c = g * cpg + threadIdx.x;
            float a = 0.f, b = 0.f;
            for (int k = 0; k < lanes; ++k) { a += s_a[k * cpg + threadIdx.x]; b += s_b[k * cpg + threadIdx.x]; }
            chansum[((size_t)n * 2 + 0) * C + c] = a;
            chansum[((size_t)n * 2 + 1) * C + c] = b;
            const float gm = __bfloat162float(gamma[c]);
            db = gm * a; ds = gm * b;
        }
    }
    s_db[threadIdx.x] = db; s_ds[threadIdx.x] = ds;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_db[threadIdx.x] += s_db[threadIdx.x + o]; s_ds[threadIdx.x] += s_ds[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        group_terms[((size_t)n * GN_GROUPS + g) * 2 + 0] = s_db[0];
        group_terms[((size_t)n * GN_GROUPS + g) * 2 + 1] = s_ds[0];
    }
}
// pass 2b: dgamma / dbeta = sum over images of the per-channel sums
__global__ void gn_bwd_param_kernel(const float* __restrict__ chansum, int NB, int C, __nv_bfloat16* __restrict__ dgamma,
                                    __nv_bfloat16* __restrict__ dbeta, int accumulate) {
    pdl_enter();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, b = 0.f;
    for (int n = 0; n < NB; ++n) { a += chansum[((size_t)n * 2 + 0) * C + c]; b += chansum[((size_t)n * 2 + 1) * C + c]; }
    if (accumulate) { a = round_bf16(a) + __bfloat162float(dbeta[c]); b = round_bf16(b) + __bfloat162float(dgamma[c]); }
    dbeta[c] = __float2bfloat16_rn(a);
    dgamma[c] = __float2bfloat16_rn(b);
}

// pass 3: dx = rstd * (dz*gamma - db/cnt - xhat * ds/cnt)
__global__ void __launch_bounds__(GN_THREADS)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ group_terms,
                    int HW, int C, int silu, const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx) {
    pdl_enter();
    __shared__ float s_mean[GN_GROUPS], s_rstd[GN_GROUPS], s_db[GN_GROUPS], s_ds[GN_GROUPS];
    const int n = blockIdx.y;
    const int cpg = C / GN_GROUPS;
    if (threadIdx.x < GN_GROUPS) {
        const float inv = 1.0f / ((float)HW * (float)cpg);
        s_mean[threadIdx.x] = mean[n * GN_GROUPS + threadIdx.x];
        s_rstd[threadIdx.x] = rstd[n * GN_GROUPS + threadIdx.x];
        s_db[threadIdx.x] = group_terms[(n * GN_GROUPS + threadIdx.x) * 2 + 0] * inv;
        s_ds[threadIdx.x] = group_terms[(n * GN_GROUPS + threadIdx.x) * 2 + 1] * inv;
    }
    __syncthreads();
    const int vec_per_row = C / 8;
    const int cols = min(vec_per_row, GN_THREADS);
    const int row_lanes = GN_THREADS / cols;
    const int tc = threadIdx.x % cols, tr = threadIdx.x / cols;
    const int rows_per_chunk = (HW + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * rows_per_chunk, r1 = min(HW, r0 + rows_per_chunk);
    const size_t base = (size_t)n * HW * C;
    if (tr >= row_lanes) return;
    for (int v = tc; v < vec_per_row; v += cols) {             // column-owner loop, see gn_apply_kernel
        float gm[8], bt[8], mu[8], rs[8], gdb[8], gds[8];
        unpack8(*reinterpret_cast<const uint4*>(gamma + v * 8), gm);
        unpack8(*reinterpret_cast<const uint4*>(beta + v * 8), bt);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int g = (v * 8 + e) / cpg;
            mu[e] = s_mean[g]; rs[e] = s_rstd[g]; gdb[e] = s_db[g]; gds[e] = s_ds[g];
        }
        auto one = [&](const uint4& px, const uint4& pd, const uint4& pr, int r) {
            float fx[8], fd[8];
            unpack8(px, fx);
            unpack8(pd, fd);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float xh = (fx[e] - mu[e]) * rs[e];
                float dz = fd[e];
                if (silu) {
                    const float z = xh * gm[e] + bt[e];
                    const float sg = sigmoidf_(z);
                    dz *= sg * (1.0f + z * (1.0f - sg));
                }
                fx[e] = rs[e] * (dz * gm[e] - gdb[e] - xh * gds[e]);
            }
            if (dres) {
                float fr[8];
                unpack8(pr, fr);
#pragma unroll
                for (int e = 0; e < 8; ++e) fx[e] = round_bf16(fx[e]) + fr[e];
            }
            st_stream(dx + base + (size_t)r * C + v * 8, pack8(fx));
        };
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        int r = r0 + tr;
        for (; r + row_lanes < r1; r += 2 * row_lanes) {        // two rows (up to six 16-byte loads) in flight
            uint4 px[2], pd[2], pr[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const size_t off = base + (size_t)(r + u * row_lanes) * C + v * 8;
                px[u] = ld_stream(x + off);
                pd[u] = ld_stream(dy + off);
                pr[u] = dres ? ld_stream(dres + off) : zero4;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) one(px[u], pd[u], pr[u], r + u * row_lanes);
        }
        for (; r < r1; r += row_lanes) {
            const size_t off = base + (size_t)r * C + v * 8;
            one(ld_stream(x + off), ld_stream(dy + off), dres ? ld_stream(dres + off) : zero4, r);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, C % 8 == 0, C <= 2048
// ---------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;     // vectors of 8 per lane -> C <= 2048
constexpr int LN_WARPS = 8;

__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
              const __nv_bfloat16* __restrict__ beta, long long rows, int C, float eps, __nv_bfloat16* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    pdl_enter();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = C / 8;
    for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < rows; row += (long long)gridDim.x * LN_WARPS) {
        float f[LN_MAXV][8];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int v = lane + 32 * i;
            if (v < nv) {
                unpack8(ld_stream(x + row * C + v * 8), f[i]);
#pragma unroll
                for (int e = 0; e < 8; ++e) s += f[i][e];
            }
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int v = lane + 32 * i;
            if (v < nv) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { const float d = f[i][e] - mean; q = fmaf(d, d, q); }
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int v = lane + 32 * i;
            if (v < nv) {
                float gm[8], bt[8];
                unpack8(*reinterpret_cast<const uint4*>(gamma + v * 8), gm);
                unpack8(*reinterpret_cast<const uint4*>(beta + v * 8), bt);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[i][e] = (f[i][e] - mean) * rstd * gm[e] + bt[e];
                st_stream(y + row * C + v * 8, pack8(f[i]));
            }
        }
        if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
}

// backward: dx per row (one warp per row); dgamma/dbeta partials per block -> partial [gridDim.x][2][C].
// VPL = 16-byte vectors per lane (C <= 256*VPL).  x and dy stay packed (uint4) in registers between the two passes so
// the per-thread footprint is ~ 20*VPL registers of data + 16*VPL of dgamma/dbeta accumulators: no spills at VPL = 5.
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
              const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
              long long rows, int C, const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx,
              float* __restrict__ partial) {
    pdl_enter();
    extern __shared__ float sm[];      // [LN_WARPS][2][C] per-warp partials, reduced at the end
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = C / 8;
    uint4 gpk[VPL];
    float ag[VPL][8], ab[VPL][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
#pragma unroll
        for (int e = 0; e < 8; ++e) { ag[i][e] = 0.f; ab[i][e] = 0.f; }
        gpk[i] = v < nv ? *reinterpret_cast<const uint4*>(gamma + v * 8) : make_uint4(0, 0, 0, 0);
    }
    for (long long row = (long long)blockIdx.x * LN_WARPS + warp; row < rows; row += (long long)gridDim.x * LN_WARPS) {
        const float mu = mean[row], rs = rstd[row];
        uint4 xpk[VPL], dpk[VPL], rpk[VPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < nv) {
                xpk[i] = ld_stream(x + row * C + v * 8);
                dpk[i] = ld_stream(dy + row * C + v * 8);
                if (dres) rpk[i] = ld_stream(dres + row * C + v * 8);      // residual gradient: fetched with the rest, used last
            }
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < nv) {
                float fx[8], fd[8], gm[8];
                unpack8(xpk[i], fx); unpack8(dpk[i], fd); unpack8(gpk[i], gm);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float xh = (fx[e] - mu) * rs;
                    const float dg = fd[e] * gm[e];
                    s1 += dg;
                    s2 = fmaf(dg, xh, s2);
                    ag[i][e] = fmaf(fd[e], xh, ag[i][e]);
                    ab[i][e] += fd[e];
                }
            }
        }
        s1 = warp_sum(s1) / (float)C;
        s2 = warp_sum(s2) / (float)C;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int v = lane + 32 * i;
            if (v < nv) {
                float fx[8], fd[8], gm[8], o[8];
                unpack8(xpk[i], fx); unpack8(dpk[i], fd); unpack8(gpk[i], gm);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = rs * (fd[e] * gm[e] - s1 - (fx[e] - mu) * rs * s2);
                if (dres) {
                    float fr[8];
                    unpack8(rpk[i], fr);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = round_bf16(o[e]) + fr[e];
                }
                st_stream(dx + row * C + v * 8, pack8(o));
            }
        }
    }
    // per-warp partials -> smem (no atomics), then a column-parallel sum over the warps
    float* mine = sm + (size_t)warp * 2 * C;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { mine[v * 8 + e] = ag[i][e]; mine[C + v * 8 + e] = ab[i][e]; }
        }
    }
    __syncthreads();
    float* out = partial + (size_t)blockIdx.x * 2 * C;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) a += sm[(size_t)w * 2 * C + i];
        out[i] = a;
    }
}

// reduce [blocks][2][C] -> dgamma (first C), dbeta (second C)
// block = 32 columns x 8 row lanes: lane (cx, ry) sums partial rows ry, ry+8, ... of its column (a warp reads 128
// contiguous bytes per row), then the 8 row lanes are combined through shared memory
__global__ void __launch_bounds__(256)
ln_bwd_finalize_kernel(const float* __restrict__ partial, int blocks, int C, __nv_bfloat16* __restrict__ dgamma,
                       __nv_bfloat16* __restrict__ dbeta, int accumulate) {
    pdl_enter();
    __shared__ float sm[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (i < 2 * C)
        for (int b = ry; b < blocks; b += 8) s += partial[(size_t)b * 2 * C + i];
    sm[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && i < 2 * C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][cx];
        __nv_bfloat16* dst = i < C ? dgamma + i : dbeta + (i - C);
        if (accumulate) t = round_bf16(t) + __bfloat162float(*dst);
        *dst = __float2bfloat16_rn(t);
    }
}

}  // namespace aoz

using namespace aoz;

template <int VPL>
static int launch_ln_bwd(const void* dy, const void* x, const void* gamma, const void* mean, const void* rstd, long long rows, int C,
                         const void* dres, void* dx, void* workspace, int blocks, cudaStream_t s) {
    const size_t smem = (size_t)LN_WARPS * 2 * C * sizeof(float);
    static size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(ln_bwd_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    launch_k(ln_bwd_kernel<VPL>, dim3(blocks), dim3(LN_WARPS * 32), (size_t)(smem), s, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)gamma,
                                                          (const float*)mean, (const float*)rstd, rows, C, (const __nv_bfloat16*)dres,
                                                          (__nv_bfloat16*)dx, (float*)workspace);
    AOZ_CHECK_LAUNCH("ln_bwd_kernel");
    return AOZ_OK;
}


extern "C" {

static int gn_chunks(int NB, int HW) {
    int chunks = (sm_count() * 2 + NB - 1) / NB;
    if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
    if (chunks > HW) chunks = HW;
    if (chunks < 1) chunks = 1;
    return chunks;
}

// pixel chunks (grid.x) of the apply kernels: enough blocks to fill the GPU a few times over, at least 8 rows per thread
static int gn_apply_chunks(int NB, int HW, int C) {
    const int vec = C / 8;
    const int cols = vec < GN_THREADS ? vec : GN_THREADS;
    const int row_lanes = GN_THREADS / cols;
    int chunks = (HW + row_lanes * 8 - 1) / (row_lanes * 8);
    const int cap = (sm_count() * 8 + NB - 1) / NB;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    return chunks;
}

// workspace floats needed by the GroupNorm forward / backward (upper bound)
long long aoz_groupnorm_workspace_floats(int NB, int HW, int C) {
    const long long chunks = GN_MAX_CHUNKS;
    return (long long)NB * chunks * 2 * C + (long long)NB * GN_GROUPS * 2 + 64 + (long long)NB * 2 * C + 64;
}

// y = silu?(GroupNorm32(x)); x, y: [NB, HW, C] bf16 channels-last; mean/rstd out: [NB, 32] fp32
int aoz_groupnorm_fwd(const void* x, const void* gamma, const void* beta, int NB, int HW, int C, float eps, int silu,
                      void* y, void* mean, void* rstd, void* workspace, void* stream) {
    AOZ_CHECK_ARG(x && gamma && beta && y && mean && rstd && workspace, "aoz_groupnorm_fwd: null pointer");
    AOZ_CHECK_ARG(C % GN_GROUPS == 0 && C % 8 == 0, "aoz_groupnorm_fwd: C=%d must be a multiple of 32", C);
    AOZ_CHECK_ARG(NB > 0 && HW > 0, "aoz_groupnorm_fwd: empty input");
    AOZ_CHECK_ARG(2 * C * (int)sizeof(float) <= 48 * 1024, "aoz_groupnorm_fwd: C=%d too large", C);
    cudaStream_t s = (cudaStream_t)stream;
    const int chunks = gn_chunks(NB, HW);
    launch_k(gn_stats_kernel, dim3(chunks, NB), dim3(GN_THREADS), (size_t)(2 * C * sizeof(float)), s, (const __nv_bfloat16*)x, HW, C, (float*)workspace);
    AOZ_CHECK_LAUNCH("gn_stats_kernel");
    int gx = gn_apply_chunks(NB, HW, C);
    launch_k(gn_apply_kernel, dim3(gx, NB), dim3(GN_THREADS), (size_t)(0), s, (const __nv_bfloat16*)x, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta,
                                                       (const float*)workspace, chunks, HW, C, eps, silu, (__nv_bfloat16*)y,
                                                       (float*)mean, (float*)rstd);
    AOZ_CHECK_LAUNCH("gn_apply_kernel");
    return AOZ_OK;
}

int aoz_groupnorm_bwd(const void* dy, const void* x, const void* gamma, const void* beta, const void* mean, const void* rstd,
                      int NB, int HW, int C, int silu, const void* dres, void* dx, void* dgamma, void* dbeta, int accumulate,
                      void* workspace, void* stream) {
    AOZ_CHECK_ARG(dy && x && gamma && beta && mean && rstd && dx && workspace, "aoz_groupnorm_bwd: null pointer");
    AOZ_CHECK_ARG(C % GN_GROUPS == 0 && C % 8 == 0, "aoz_groupnorm_bwd: C=%d must be a multiple of 32", C);
    cudaStream_t s = (cudaStream_t)stream;
    const int chunks = gn_chunks(NB, HW);
    float* partial = (float*)workspace;
    float* group_terms = partial + (size_t)NB * GN_MAX_CHUNKS * 2 * C;
    launch_k(gn_bwd_stats_kernel, dim3(chunks, NB), dim3(GN_THREADS), (size_t)(2 * C * sizeof(float)), s, 
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta,
        (const float*)mean, (const float*)rstd, HW, C, silu, partial);
    AOZ_CHECK_LAUNCH("gn_bwd_stats_kernel");
    float* chansum = group_terms + (size_t)NB * GN_GROUPS * 2 + 64;
    launch_k(gn_bwd_group_kernel, dim3(GN_GROUPS, NB), dim3(128), (size_t)(0), s, partial, (const __nv_bfloat16*)gamma, chunks, C, chansum, group_terms);
    AOZ_CHECK_LAUNCH("gn_bwd_group_kernel");
    if (dgamma) {
        launch_k(gn_bwd_param_kernel, dim3((C + 255) / 256), dim3(256), (size_t)(0), s, chansum, NB, C, (__nv_bfloat16*)dgamma, (__nv_bfloat16*)dbeta, accumulate);
        AOZ_CHECK_LAUNCH("gn_bwd_param_kernel");
    }
    int gx = gn_apply_chunks(NB, HW, C);
    launch_k(gn_bwd_apply_kernel, dim3(gx, NB), dim3(GN_THREADS), (size_t)(0), s, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)gamma,
                                                           (const __nv_bfloat16*)beta, (const float*)mean, (const float*)rstd,
                                                           group_terms, HW, C, silu, (const __nv_bfloat16*)dres, (__nv_bfloat16*)dx);
    AOZ_CHECK_LAUNCH("gn_bwd_apply_kernel");
    return AOZ_OK;
}

int aoz_layernorm_fwd(const void* x, const void* gamma, const void* beta, long long rows, int C, float eps, void* y, void* mean,
                      void* rstd, void* stream) {
    AOZ_CHECK_ARG(x && gamma && beta && y && mean && rstd, "aoz_layernorm_fwd: null pointer");
    AOZ_CHECK_ARG(C % 8 == 0 && C <= LN_MAXV * 256, "aoz_layernorm_fwd: C=%d unsupported (multiple of 8, <= 2048)", C);
    if (rows <= 0) return AOZ_OK;
    long long blocks = (rows + LN_WARPS - 1) / LN_WARPS;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    launch_k(ln_fwd_kernel, dim3((int)blocks), dim3(LN_WARPS * 32), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)x, (const __nv_bfloat16*)gamma,
                                                                          (const __nv_bfloat16*)beta, rows, C, eps, (__nv_bfloat16*)y,
                                                                          (float*)mean, (float*)rstd);
    AOZ_CHECK_LAUNCH("ln_fwd_kernel");
    return AOZ_OK;
}

long long aoz_layernorm_bwd_workspace_floats(int C) { return (long long)sm_count() * 2 * 2 * C; }

int aoz_layernorm_bwd(const void* dy, const void* x, const void* gamma, const void* mean, const void* rstd, long long rows, int C,
                      const void* dres, void* dx, void* dgamma, void* dbeta, int accumulate, void* workspace, void* stream) {
    AOZ_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, "aoz_layernorm_bwd: null pointer");
    AOZ_CHECK_ARG(C % 8 == 0 && C <= LN_MAXV * 256, "aoz_layernorm_bwd: C=%d unsupported", C);
    if (rows <= 0) return AOZ_OK;
    long long blocks = (rows + LN_WARPS - 1) / LN_WARPS;
    if (blocks > sm_count() * 2) blocks = sm_count() * 2;
    cudaStream_t s = (cudaStream_t)stream;
    const int vpl = (C / 8 + 31) / 32;
    int rc;
    switch (vpl) {
        case 1: rc = launch_ln_bwd<1>(dy, x, gamma, mean, rstd, rows, C, dres, dx, workspace, (int)blocks, s); break;
        case 2: rc = launch_ln_bwd<2>(dy, x, gamma, mean, rstd, rows, C, dres, dx, workspace, (int)blocks, s); break;
        case 3: rc = launch_ln_bwd<3>(dy, x, gamma, mean, rstd, rows, C, dres, dx, workspace, (int)blocks, s); break;
        case 4: rc = launch_ln_bwd<4>(dy, x, gamma, mean, rstd, rows, C, dres, dx, workspace, (int)blocks, s); break;
        case 5: rc = launch_ln_bwd<5>(dy, x, gamma, mean, rstd, rows, C, dres, dx, workspace, (int)blocks, s); break;
        default: rc = launch_ln_bwd<8>(dy, x, gamma, mean, rstd, rows, C, dres, dx, workspace, (int)blocks, s); break;
    }
    if (rc != AOZ_OK) return rc;
    launch_k(ln_bwd_finalize_kernel, dim3((2 * C + 31) / 32), dim3(256), (size_t)(0), s, (const float*)workspace, (int)blocks, C, (__nv_bfloat16*)dgamma,
                                                               (__nv_bfloat16*)dbeta, accumulate);
    AOZ_CHECK_LAUNCH("ln_bwd_finalize_kernel");
    return AOZ_OK;
}

}  // extern "C"
