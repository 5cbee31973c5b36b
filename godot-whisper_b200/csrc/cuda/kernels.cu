// The SIMT kernels of the path: layout shuffles, LayerNorm, softmax, the skinny (decode-step) contraction, decoder
// attention against the f16 KV caches, and a plain tiled GEMM used only to cross-check the tcgen05 engine.
// Arithmetic follows the reference operator by operator — citations are on the launchers in kernels.cuh.
#include "kernels.cuh"

#include <cstdio>
#include <cstdlib>

namespace wb200 {

namespace {

// ---- mel window -> token-major f16 ----------------------------------------------------------------------------------------

__global__ void k_mel_to_tokens(const float * __restrict__ mel, __half * __restrict__ out, int n_mels, int n_frames) {
    __shared__ float tile[32][33];
    const int f0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int m = m0 + i, f = f0 + threadIdx.x;
        tile[i][threadIdx.x] = (m < n_mels && f < n_frames) ? mel[(int64_t) m * n_frames + f] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int f = f0 + i, m = m0 + threadIdx.x;
        if (f < n_frames && m < n_mels) out[(int64_t) (f + 1) * n_mels + m] = __float2half_rn(tile[threadIdx.x][i]);
    }
    if (blockIdx.x == 0 && blockIdx.y == 0) {
        for (int m = threadIdx.y * 32 + threadIdx.x; m < n_mels; m += 32 * blockDim.y) {
            out[m] = __float2half_rn(0.0f);
            out[(int64_t) (n_frames + 1) * n_mels + m] = __float2half_rn(0.0f);
        }
    }
}

// ---- LayerNorm --------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void layernorm_row(const float * __restrict__ x, const float * __restrict__ gamma,
                                              const float * __restrict__ beta, __half * out16, float * out32, int d,
                                              float eps, int lane) {
    double s = 0.0;
    for (int i = lane; i < d; i += 32) s += (double) x[i];
    s = warp_sum(s);
    const float mean = (float) (s / (double) d);
    double s2 = 0.0;
    for (int i = lane; i < d; i += 32) {
        const float v = __fsub_rn(x[i], mean);
        s2 += (double) __fmul_rn(v, v);
    }
    s2 = warp_sum(s2);
    const float var = (float) (s2 / (double) d);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
    for (int i = lane; i < d; i += 32) {
        float y = __fmul_rn(__fsub_rn(x[i], mean), scale);
        y = __fadd_rn(__fmul_rn(y, gamma[i]), beta[i]);
        if (out16) out16[i] = __float2half_rn(y);
        if (out32) out32[i] = y;
    }
}

__global__ void k_layernorm(const float * __restrict__ x, const float * __restrict__ gamma, const float * __restrict__ beta,
                            __half * __restrict__ out16, float * __restrict__ out32, int rows, int d, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    layernorm_row(x + (int64_t) row * d, gamma, beta, out16 ? out16 + (int64_t) row * d : nullptr,
                  out32 ? out32 + (int64_t) row * d : nullptr, d, eps, threadIdx.x & 31);
}

// The same arithmetic with the row held in registers: one 16-byte load per four features (lane l owns features 128 j + 4 l .. + 3),
// read once, 8-byte f16 / 16-byte f32 stores.  The f64 sums are exact to ~1e-16 in any order (d <= 1536 floats), so the order does
// not reach the f32 mean / variance; every other operation is per element.  d = 128 NV.
template <int NV>
__global__ void __launch_bounds__(256)
k_layernorm_vec(const float * __restrict__ x, const float * __restrict__ gamma, const float * __restrict__ beta,
                __half * __restrict__ out16, float * __restrict__ out32, int rows, int d, float eps) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float * xr = x + (int64_t) row * d;
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = *(const float4 *) (xr + 128 * j + 4 * lane);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += ((double) v[j].x + (double) v[j].y) + ((double) v[j].z + (double) v[j].w);
    s = warp_sum(s);
    const float mean = (float) (s / (double) d);
    double s2 = 0.0;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float a = __fsub_rn(v[j].x, mean), b = __fsub_rn(v[j].y, mean), c = __fsub_rn(v[j].z, mean), e = __fsub_rn(v[j].w, mean);
        s2 += ((double) __fmul_rn(a, a) + (double) __fmul_rn(b, b)) + ((double) __fmul_rn(c, c) + (double) __fmul_rn(e, e));
    }
    s2 = warp_sum(s2);
    const float var = (float) (s2 / (double) d);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, eps)));
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int e0 = 128 * j + 4 * lane;
        const float4 g = __ldg((const float4 *) (gamma + e0)), b = __ldg((const float4 *) (beta + e0));
        const float y0 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[j].x, mean), scale), g.x), b.x);
        const float y1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[j].y, mean), scale), g.y), b.y);
        const float y2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[j].z, mean), scale), g.z), b.z);
        const float y3 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(v[j].w, mean), scale), g.w), b.w);
        if (out16) {
            const __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
            uint2 pk;
            pk.x = *(const uint32_t *) &h0; pk.y = *(const uint32_t *) &h1;
            *(uint2 *) (out16 + (int64_t) row * d + e0) = pk;
        }
        if (out32) *(float4 *) (out32 + (int64_t) row * d + e0) = make_float4(y0, y1, y2, y3);
    }
}

// ---- softmax over rows --------------------------------------------------------------------------------------------------------

__device__ __forceinline__ float exp_table(const uint16_t * __restrict__ lut, float x) {
    const uint16_t h = __half_as_ushort(__float2half_rn(x));
    return __half2float(__ushort_as_half(__ldg(lut + h)));
}

// One warp per row, the row held in registers: one pass over HBM with 16-byte loads, one table look-up per score.  The sum of
// the table values is exact in any order: every f16 value is an integer multiple of 2^-24 and a row of <= 1536 of them sums to
// < 2^35 such units, so an integer sum equals the reference's f64 sum (ggml.c:11170-11192) bit for bit.
constexpr int kSmaxPer = 12;      // float4 loads per lane: rows of up to 32 * 4 * 12 = 1536 columns

__global__ void __launch_bounds__(256)
k_softmax_rows(const float * __restrict__ S, __half * __restrict__ P, int64_t rows, int n_cols, int ld_s,
               int ld_p, const uint16_t * __restrict__ lut) {
    const int64_t row = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int n4 = (n_cols + 3) >> 2;                       // float4 groups that hold at least one valid column
    const float * s = S + row * ld_s;
    __half * p = P + row * ld_p;
    float4 v[kSmaxPer];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < kSmaxPer; ++u) {
        const int g = lane + 32 * u;
        float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (g < n4) {
            t = __ldcs((const float4 *) s + g);
            const int c = 4 * g;
            if (c + 1 >= n_cols) t.y = -INFINITY;
            if (c + 2 >= n_cols) t.z = -INFINITY;
            if (c + 3 >= n_cols) t.w = -INFINITY;
        }
        v[u] = t;
        mx = fmaxf(mx, fmaxf(fmaxf(t.x, t.y), fmaxf(t.z, t.w)));
    }
    mx = warp_max(mx);
    unsigned int isum = 0;                                  // units of 2^-24; <= 48 terms of <= 2^24 each per lane
#pragma unroll
    for (int u = 0; u < kSmaxPer; ++u) {
        if (lane + 32 * u < n4) {
            float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float r = 0.0f;
                if (e[q] != -INFINITY) { r = exp_table(lut, __fsub_rn(e[q], mx)); isum += (unsigned int) (r * 16777216.0f); }
                e[q] = r;
            }
            v[u] = make_float4(e[0], e[1], e[2], e[3]);
        }
    }
    unsigned long long tot = isum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const float inv = (float) (1.0 / ((double) tot * (1.0 / 16777216.0)));
    const int p4 = ld_p >> 2;                               // P rows are written to their full padded width (zeros past n_cols)
#pragma unroll
    for (int u = 0; u < kSmaxPer; ++u) {
        const int g = lane + 32 * u;
        if (g < p4) {
            const float4 t = g < n4 ? v[u] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const __half2 h0 = __floats2half2_rn(__fmul_rn(t.x, inv), __fmul_rn(t.y, inv));
            const __half2 h1 = __floats2half2_rn(__fmul_rn(t.z, inv), __fmul_rn(t.w, inv));
            uint2 pk;
            pk.x = *(const uint32_t *) &h0; pk.y = *(const uint32_t *) &h1;
            *((uint2 *) p + g) = pk;
        }
    }
}

// ---- decoder embedding / gather ---------------------------------------------------------------------------------------------

__global__ void k_embed(const __half * __restrict__ te, const float * __restrict__ pe, const int * __restrict__ token,
                        const int * __restrict__ pos, float * __restrict__ x, int n, int d) {
    const int r = blockIdx.x;
    const int64_t t = token[r], p = pos[r];
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        x[(int64_t) r * d + i] = __fadd_rn(__half2float(te[t * d + i]), pe[p * d + i]);
    }
}

__global__ void k_gather_rows(const float * __restrict__ src, const int * __restrict__ idx, float * __restrict__ dst, int n, int d) {
    const int r = blockIdx.x;
    const int64_t s = idx[r];
    for (int i = threadIdx.x; i < d; i += blockDim.x) dst[(int64_t) r * d + i] = src[s * d + i];
}

// ---- skinny contraction (decode steps) -----------------------------------------------------------------------------------------

__device__ __forceinline__ void fma8(float & acc, const uint4 & w, const uint4 & x) {
    const __half2 * wh = (const __half2 *) &w;
    const __half2 * xh = (const __half2 *) &x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = __half22float2(wh[i]);
        const float2 b = __half22float2(xh[i]);
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
    }
}

template <int R, int RW>   // R = activation rows held per pass, RW = weight rows streamed concurrently per warp
__global__ void __launch_bounds__(128)
k_gemm_skinny(const SkinnyIn in, const __half * __restrict__ W, int n, int M, int K, const GemmEpi epi) {
    extern __shared__ __align__(16) uint8_t smem_sk[];
    __half * xs = (__half *) smem_sk;                 // [R][K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = blockDim.x >> 5;

    if (in.x16) {
        const int kc = K >> 3, total = n * kc;
        for (int i0 = threadIdx.x; i0 < total; i0 += blockDim.x * 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + blockDim.x * u, r = i / kc, k = (i - r * kc) << 3;
                v[u] = i < total ? __ldg((const uint4 *) (in.x16 + (int64_t) r * in.x16_ld + k)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + blockDim.x * u, r = i / kc, k = (i - r * kc) << 3;
                if (i < total) *(uint4 *) (xs + (int64_t) r * K + k) = v[u];
            }
        }
    } else {
        for (int r = warp; r < n; r += n_warps) {
            layernorm_row(in.x32 + (int64_t) r * in.x32_ld, in.gamma, in.beta, xs + (int64_t) r * K, nullptr, K, in.eps, lane);
        }
    }
    __syncthreads();

    const int m_base = (blockIdx.x * n_warps + warp) * RW;
    if (m_base >= M) return;

    float acc[RW][R];
#pragma unroll
    for (int j = 0; j < RW; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[j][r] = 0.0f;

    for (int k = lane * 8; k < K; k += 256) {
        uint4 w[RW];
#pragma unroll
        for (int j = 0; j < RW; ++j) {
            const int m = min(m_base + j, M - 1);
            w[j] = __ldg((const uint4 *) (W + (int64_t) m * K + k));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < n) {
                const uint4 x = *(const uint4 *) (xs + (int64_t) r * K + k);
#pragma unroll
                for (int j = 0; j < RW; ++j) fma8(acc[j][r], w[j], x);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < RW; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[j][r] = warp_sum(acc[j][r]);

    // lane r finishes row r for each of the RW features
#pragma unroll
    for (int j = 0; j < RW; ++j) {
        const int m = m_base + j;
        if (m >= M) break;
        float mine = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) if (lane == r) mine = acc[j][r];
        if (lane < n && lane < R) {
            const int seg_i = epi.nseg > 1 ? m / epi.seg_m : 0;
            const EpiSeg & sg = epi.seg[seg_i];
            const int ml = m - seg_i * epi.seg_m;
            float pre;
            const float v = epi_value(sg, epi.gelu_lut, mine, lane, ml, &pre);
            epi_store(sg, v, pre, lane, ml, 0, 0);
        }
    }
}


// ---- skinny contraction on mma.sync for 9..32 rows ---------------------------------------------------------------------------

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NT>      // NT tiles of 8 activation rows
__global__ void __launch_bounds__(128)
k_gemm_skinny_mma(const __half * __restrict__ x16, int64_t x_ld, const __half * __restrict__ W, int n, int M, int K, int kz,
                  const GemmEpi epi) {
    extern __shared__ __align__(16) uint8_t smem_mm[];
    const int Ks = K + 32;                                  // row stride = 16 mod 32 words: conflict-free 16-byte fragment loads
    __half * xs = (__half *) smem_mm;                       // [8 NT][Ks]
    float * red = (float *) (smem_mm + (size_t) 8 * NT * Ks * sizeof(__half));   // [4 warps][32 lanes][NT*4]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    {   // stage the activation rows: flat index over (row, 8-wide k chunk), four independent 16-byte loads in flight per thread
        const int kc = K >> 3, total = 8 * NT * kc;
        for (int i0 = threadIdx.x; i0 < total; i0 += 128 * 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 128 * u, r = i / kc, k = (i - r * kc) << 3;
                v[u] = (i < total && r < n) ? __ldg((const uint4 *) (x16 + (int64_t) r * x_ld + k)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 128 * u, r = i / kc, k = (i - r * kc) << 3;
                if (i < total) *(uint4 *) (xs + (int64_t) r * Ks + k) = v[u];
            }
        }
    }
    __syncthreads();

    const int tiles_per_cta = 4 / kz;
    const int m_tile = blockIdx.x * tiles_per_cta + warp / kz;
    const int ks = warp % kz;
    const int k_len = K / kz, k_beg = ks * k_len;
    const int m_tiles = (M + 15) >> 4;
    const bool live = m_tile < m_tiles;

    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    if (live) {
        const int r0 = min(m_tile * 16 + g, M - 1), r1 = min(m_tile * 16 + g + 8, M - 1);
        const __half * w0 = W + (int64_t) r0 * K + k_beg + 8 * t;
        const __half * w1 = W + (int64_t) r1 * K + k_beg + 8 * t;
        const __half * xb = xs + (int64_t) g * Ks + k_beg + 8 * t;
        constexpr int kU = 4;
        for (int k = 0; k < k_len; k += 32 * kU) {
            uint4 alo[kU], ahi[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int kk = k + 32 * u;
                if (kk < k_len) { alo[u] = __ldg((const uint4 *) (w0 + kk)); ahi[u] = __ldg((const uint4 *) (w1 + kk)); }
                else { alo[u] = make_uint4(0, 0, 0, 0); ahi[u] = make_uint4(0, 0, 0, 0); }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int kk = k + 32 * u;
                if (kk < k_len) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const uint4 b = *(const uint4 *) (xb + (int64_t) nt * 8 * Ks + kk);
                        // the 8 consecutive k of a lane are split 4 + 4 over two MMAs; A and B use the same permutation of k
                        mma_16816(acc[nt], alo[u].x, ahi[u].x, alo[u].y, ahi[u].y, b.x, b.y);
                        mma_16816(acc[nt], alo[u].z, ahi[u].z, alo[u].w, ahi[u].w, b.z, b.w);
                    }
                }
            }
        }
    }
    if (kz > 1) {
        float * mine = red + ((size_t) warp * 32 + lane) * (NT * 4);
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) mine[i * 4 + j] = acc[i][j];
        __syncthreads();
        if (ks != 0) return;
        for (int o = 1; o < kz; ++o) {
            const float * other = red + ((size_t) (warp + o) * 32 + lane) * (NT * 4);
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += other[i * 4 + j];
        }
    }
    if (!live) return;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int m = m_tile * 16 + g + ((c & 2) ? 8 : 0);
            const int row = nt * 8 + 2 * t + (c & 1);
            if (m < M && row < n) {
                const int seg_i = epi.nseg > 1 ? m / epi.seg_m : 0;
                const EpiSeg & sg = epi.seg[seg_i];
                const int ml = m - seg_i * epi.seg_m;
                float pre;
                const float v = epi_value(sg, epi.gelu_lut, acc[nt][c], row, ml, &pre);
                epi_store(sg, v, pre, row, ml, 0, 0);
            }
        }
    }
}

// ---- decoder attention: one thread-block CLUSTER per (head, row) -------------------------------------------------------------
//
// The keys of one (head, row) are split over the S CTAs of a cluster (S = 1, 2, 4 or 8) so that a single-token step still
// pulls the K / V^T stream with many SMs.  Softmax keeps the reference's arithmetic (global max first, table exp, f64 sum,
// p rounded to f16): the CTAs exchange their local max and local sum through distributed shared memory, each forms the
// partial P·V of its key range, and rank 0 adds the partials in rank order.

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// generic pointer into the shared memory of CTA `rank` of this cluster
template <class T> __device__ __forceinline__ const T * dsmem_ptr(const T * p, uint32_t rank) {
    uint64_t out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((uint64_t) p), "r"(rank));
    return (const T *) out;
}

__global__ void __launch_bounds__(256, 4)
k_decode_attention(const AttnArgs a) {
    extern __shared__ __align__(16) uint8_t smem_at[];
    const uint32_t S = cluster_size(), rank = cluster_rank();
    const int n_keys = a.n_keys_dev ? min(*a.n_keys_dev, a.n_keys) : a.n_keys;
    // key range of this CTA, multiples of 8 so that V^T is read with aligned 16-byte loads
    const int per = (((n_keys + (int) S - 1) / (int) S) + 7) & ~7;
    const int k0 = min((int) rank * per, n_keys), k1 = min(k0 + per, n_keys);
    const int n_own = k1 - k0, n_pad = (n_own + 7) & ~7;
    const int per_max = (((a.n_keys + (int) S - 1) / (int) S) + 7) & ~7;   // what the launcher sized shared memory for
    float *  sc  = (float *) smem_at;                               // [per_max] scores, then exp values
    __half * p16 = (__half *) (smem_at + sizeof(float) * per_max);  // [per_max]
    __shared__ float  red_f[8];
    __shared__ double red_d[8];
    __shared__ float  x_max;            // exchanged through DSMEM
    __shared__ double x_sum;
    __shared__ float  x_out[64];

    const int h = blockIdx.x / S, r = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane & 7;
    const int64_t koff = a.koff ? a.koff[r] : 0;
    const int64_t voff = a.voff ? a.voff[r] : 0;

    float q[8];
    {
        const uint4 qv = *(const uint4 *) (a.q + (int64_t) r * a.d + h * 64 + g * 8);
        const __half2 * qh = (const __half2 *) &qv;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(qh[i]); q[2 * i] = f.x; q[2 * i + 1] = f.y; }
    }
    const __half * Kb = a.K + koff + (int64_t) k0 * a.d + h * 64 + g * 8;
    const float * mrow = a.mask ? a.mask + (int64_t) r * a.ld_mask + k0 : nullptr;

    // scores of the own key range: 8 lanes per key, 4 keys per warp per pass, kU passes (= kU 16-byte loads per lane) in flight
    float mx = -INFINITY;
    constexpr int kU = 8;
    for (int j0 = warp * 4 + (lane >> 3); j0 < n_pad; j0 += 32 * kU) {
        uint4 kv[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int j = j0 + 32 * u;
            kv[u] = (j < n_own) ? __ldg((const uint4 *) (Kb + (int64_t) j * a.d)) : make_uint4(0, 0, 0, 0);
        }
        float dt[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const __half2 * hh = (const __half2 *) &kv[u];
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hh[i]);
                acc = fmaf(f.x, q[2 * i], acc); acc = fmaf(f.y, q[2 * i + 1], acc);
            }
            dt[u] = acc;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
            for (int u = 0; u < kU; ++u) dt[u] += __shfl_xor_sync(0xffffffffu, dt[u], o);
        }
        if (g == 0) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int j = j0 + 32 * u;
                if (j < n_pad) {
                    float v = -INFINITY;
                    if (j < n_own) v = mrow ? __fadd_rn(dt[u], mrow[j]) : dt[u];
                    sc[j] = v; mx = fmaxf(mx, v);
                }
            }
        }
    }
    mx = warp_max(mx);
    if (lane == 0) red_f[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red_f[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red_f[i]);
        x_max = m;
    }
    cluster_sync_all();
    mx = -INFINITY;
    for (uint32_t c = 0; c < S; ++c) mx = fmaxf(mx, *dsmem_ptr(&x_max, c));

    double sum = 0.0;
    for (int j = threadIdx.x; j < n_pad; j += 256) {
        const float v = sc[j];
        float e = 0.0f;
        if (v != -INFINITY) e = exp_table(a.exp_lut, __fsub_rn(v, mx));
        sc[j] = e;
        sum += (double) e;
    }
    sum = warp_sum(sum);
    if (lane == 0) red_d[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red_d[i];
        x_sum = t;
    }
    cluster_sync_all();
    sum = 0.0;
    for (uint32_t c = 0; c < S; ++c) sum += *dsmem_ptr(&x_sum, c);      // exact: every term is a multiple of 2^-24
    const float inv = (float) (1.0 / sum);
    for (int j = threadIdx.x; j < n_pad; j += 256) p16[j] = __float2half_rn(__fmul_rn(sc[j], inv));
    __syncthreads();

    // partial P·V over the own key range: each warp owns 8 of the 64 output features and streams their 8 V^T rows together
    {
        const __half * vbase = a.Vt + voff + (int64_t) (h * 64 + warp * 8) * a.ld_v + k0;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
        for (int j = lane * 8; j < n_pad; j += 256) {
            uint4 vv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] = __ldg((const uint4 *) (vbase + (int64_t) i * a.ld_v + j));
            const uint4 pv = *(const uint4 *) (p16 + j);
#pragma unroll
            for (int i = 0; i < 8; ++i) fma8(acc[i], vv[i], pv);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float t = warp_sum(acc[i]);
            if (lane == 0) x_out[warp * 8 + i] = t;
        }
    }
    cluster_sync_all();
    if (rank == 0 && threadIdx.x < 64) {
        float acc = 0.0f;
        for (uint32_t c = 0; c < S; ++c) acc += *dsmem_ptr(&x_out[threadIdx.x], c);
        a.out[(int64_t) r * a.d + h * 64 + threadIdx.x] = __float2half_rn(acc);
    }
    cluster_sync_all();       // keep every CTA's shared memory alive until rank 0 has read it
}


// ---- greedy sampler ----------------------------------------------------------------------------------------------------------------

struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax argmax_first(ArgMax a, ArgMax b) {      // larger value wins; on ties the smaller index
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

template <class T, class Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T * scratch /* [32] */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T other;
        if constexpr (sizeof(T) == 8) {
            long long t = __shfl_xor_sync(0xffffffffu, *(long long *) &v, o);
            other = *(T *) &t;
        } else {
            int t = __shfl_xor_sync(0xffffffffu, *(int *) &v, o);
            other = *(T *) &t;
        }
        v = op(v, other);
    }
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    T r = scratch[0];
    for (int i = 1; i < nw; ++i) r = op(r, scratch[i]);
    return r;
}

__device__ __forceinline__ bool token_masked(int i, int flags, int cls, int beg, int eot, int tid0_initial, int tid0_seek) {
    if (cls & 1) return true;                                                     // <|notimestamps|>, sot, nosp, translate, transcribe, prev, languages
    if ((cls & 2) && (flags & 8)) return true;                                    // non-speech symbols
    if ((cls & 4) && (flags & 1)) return true;                                    // blank / eot at the start of a sequence
    if ((cls & 8) && (flags & 4)) return true;                                    // solm unless tinydiarize
    if ((flags & 2) && i >= beg) return true;                                     // no_timestamps
    if (flags & 16) {                                                             // last token was a timestamp
        if (flags & 32) { if (i >= beg) return true; } else { if (i < eot) return true; }
    }
    if ((flags & 64) && i >= beg + tid0_initial + 1) return true;                 // max_initial_ts
    if ((flags & 128) && i >= beg && i < beg + tid0_seek) return true;            // timestamps do not decrease
    return false;
}

__global__ void __launch_bounds__(1024)
k_sample_greedy(const float * __restrict__ logits, int n_vocab, const int * __restrict__ rule, const uint8_t * __restrict__ cls,
                int beg, int eot, float * __restrict__ out) {
    extern __shared__ __align__(16) float row_s[];          // [n_vocab] the row with the rules applied (-inf = suppressed)
    __shared__ double sd[32];
    __shared__ float  sf[32];
    __shared__ ArgMax sa[32];
    const int row = blockIdx.x;
    const float * l = logits + (int64_t) row * n_vocab;
    const int flags = rule[4 * row], tid0_initial = rule[4 * row + 1], tid0_seek = rule[4 * row + 2];
    auto fmax_op = [](float a, float b) { return fmaxf(a, b); };
    auto dsum_op = [](double a, double b) { return a + b; };

    // pass 1, over HBM/L2: apply the suppression rules (whisper.cpp:4527-4635) while staging the row in shared memory; the largest
    // logit overall and per class (x -> x - lse is monotone, so the class maxima of the log-probabilities are taken on the logits)
    float mx = -INFINITY, ts_mx = -INFINITY, text_mx = -INFINITY;
    for (int i0 = threadIdx.x; i0 < n_vocab; i0 += blockDim.x * 4) {
        float v[4]; int c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + blockDim.x * u;
            v[u] = i < n_vocab ? __ldg(l + i) : -INFINITY;
            c[u] = i < n_vocab ? (int) __ldg(cls + i) : 1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + blockDim.x * u;
            if (i < n_vocab) {
                const float x = token_masked(i, flags, c[u], beg, eot, tid0_initial, tid0_seek) ? -INFINITY : v[u];
                row_s[i] = x;
                if (i >= beg) ts_mx = fmaxf(ts_mx, x); else text_mx = fmaxf(text_mx, x);
            }
        }
    }
    ts_mx = block_reduce(ts_mx, fmax_op, sf);             // (block_reduce synchronises: row_s is complete afterwards)
    text_mx = block_reduce(text_mx, fmax_op, sf);
    mx = fmaxf(ts_mx, text_mx);

    // pass 2: log-softmax normaliser over the unmasked entries (whisper.cpp:4637-4655)
    double sum = 0.0;
    for (int i = threadIdx.x; i < n_vocab; i += blockDim.x) {
        const float x = row_s[i];
        if (x > -INFINITY) sum += (double) expf(x - mx);
    }
    sum = block_reduce(sum, dsum_op, sd);
    const float lse = logf((float) sum) + mx;
    // class maxima of the log-probabilities (whisper.cpp:4659-4684): float subtraction is monotone, so max(x - lse) = max(x) - lse
    const float ts_max = ts_mx > -INFINITY ? ts_mx - lse : -INFINITY, text_max = text_mx > -INFINITY ? text_mx - lse : -INFINITY;

    // pass 3: timestamp mass, probabilities, the best text token and the best timestamp token, timestamp statistics (:4789-4819).
    // Whether text is switched off is only known after the timestamp mass, so both class winners are kept and merged afterwards:
    // first maximum wins, text tokens come first.
    double ts_sum = 0.0, p_ts_sum = 0.0;
    ArgMax best_text{0.0f, 0x7fffffff}, best_ts{0.0f, 0x7fffffff};
    for (int i = threadIdx.x; i < n_vocab; i += blockDim.x) {
        const float x = row_s[i];
        if (!(x > -INFINITY)) continue;
        const float lp = x - lse;
        const float p = expf(lp);
        if (i >= beg) {
            ts_sum += (double) expf(lp - ts_max);
            p_ts_sum += (double) p;
            if (p > best_ts.v) best_ts = ArgMax{p, i};
        } else {
            if (p > best_text.v) best_text = ArgMax{p, i};
        }
    }
    ts_sum = block_reduce(ts_sum, dsum_op, sd);
    p_ts_sum = block_reduce(p_ts_sum, dsum_op, sd);
    best_text = block_reduce(best_text, argmax_first, sa);
    best_ts = block_reduce(best_ts, argmax_first, sa);
    float ts_logprob = -INFINITY;
    if ((float) ts_sum > 0.0f) ts_logprob = logf((float) ts_sum) + ts_max;
    const bool text_off = ts_logprob > text_max;
    const ArgMax best = text_off ? best_ts : argmax_first(best_text, best_ts);
    if (threadIdx.x == 0) {
        int id = 0, tid = 0;
        float p = 0.0f, plog = 0.0f;
        if (best.v > 0.0f) { id = best.i; p = best.v; plog = row_s[id] - lse; }
        if (best_ts.v > 0.0f) tid = best_ts.i;
        float pt = (float) ((double) best_ts.v / (p_ts_sum + 1e-10));
        const float ptsum = (float) p_ts_sum;
        if (id >= beg) { tid = id; pt = p; }
        float * o = out + 6 * row;
        o[0] = __int_as_float(id); o[1] = __int_as_float(tid); o[2] = p; o[3] = plog; o[4] = pt; o[5] = ptsum;
    }
}

// ---- sampling from the distribution (t > 0 best-of decoders, beam search) -------------------------------------------------------------
//
// whisper_process_logits with a temperature (whisper.cpp:4493-4720) followed by whisper_sample_token(best = false) /
// whisper_sample_token_topk (whisper.cpp:4777-4909) for one row of logits per block.  The reference draws from
// std::discrete_distribution<>(probs): libstdc++ normalises the weights by their sum, forms the running sums cp_i, takes
// u = generate_canonical<double, 53>(rng) and returns the first i with cp_i >= u.  Here the host still owns the generators — it draws the
// uniform variates in the reference's order and sends them along (8 bytes per draw) — and the device inverts the same CDF: what comes
// back is 24 bytes per draw instead of a 207 KB logits row.  Differences to the host path: the normaliser of the log-softmax and the
// running sums are formed by parallel f64 sums instead of sequential ones, and expf is CUDA's; a draw can differ from the host path's only
// when u falls within that (1e-7-ish) distance of a bucket edge.
__global__ void __launch_bounds__(1024)
k_sample_dist(const float * __restrict__ logits, int n_vocab, const int * __restrict__ drule, const double * __restrict__ draws,
              const uint8_t * __restrict__ cls, int beg, int eot, float * __restrict__ out) {
    extern __shared__ __align__(16) float row_s[];          // [n_vocab] logits / T with the rules applied (-inf = suppressed)
    __shared__ double sd[32];
    __shared__ float  sf[32];
    __shared__ ArgMax sa[32];
    __shared__ double s_scan[32];
    const int row = blockIdx.x;
    const float * l = logits + (int64_t) row * n_vocab;
    const int * rl = drule + 8 * row;
    const int flags = rl[0], tid0_initial = rl[1], tid0_seek = rl[2], n_draws = rl[3], draw_off = rl[5];
    const float temperature = __int_as_float(rl[4]);
    auto fmax_op = [](float a, float b) { return fmaxf(a, b); };
    auto dsum_op = [](double a, double b) { return a + b; };
    const int tid = threadIdx.x, nthr = blockDim.x;

    float ts_mx = -INFINITY, text_mx = -INFINITY;
    for (int i = tid; i < n_vocab; i += nthr) {
        float x = __ldg(l + i);
        if (temperature > 0.0f) x = __fdiv_rn(x, temperature);                    // whisper.cpp:4518-4522
        if (token_masked(i, flags, (int) __ldg(cls + i), beg, eot, tid0_initial, tid0_seek)) x = -INFINITY;
        row_s[i] = x;
        if (i >= beg) ts_mx = fmaxf(ts_mx, x); else text_mx = fmaxf(text_mx, x);
    }
    ts_mx = block_reduce(ts_mx, fmax_op, sf);
    text_mx = block_reduce(text_mx, fmax_op, sf);
    const float mx = fmaxf(ts_mx, text_mx);
    double sum = 0.0;
    for (int i = tid; i < n_vocab; i += nthr) { const float x = row_s[i]; if (x > -INFINITY) sum += (double) expf(x - mx); }
    sum = block_reduce(sum, dsum_op, sd);
    const float lse = logf((float) sum) + mx;
    const float ts_max = ts_mx > -INFINITY ? ts_mx - lse : -INFINITY, text_max = text_mx > -INFINITY ? text_mx - lse : -INFINITY;
    double ts_sum = 0.0;
    for (int i = beg + tid; i < n_vocab; i += nthr) { const float x = row_s[i]; if (x > -INFINITY) ts_sum += (double) expf((x - lse) - ts_max); }
    ts_sum = block_reduce(ts_sum, dsum_op, sd);
    float ts_logprob = -INFINITY;
    if ((float) ts_sum > 0.0f) ts_logprob = logf((float) ts_sum) + ts_max;
    const bool text_off = ts_logprob > text_max;                                  // whisper.cpp:4659-4684

    // probabilities p_i = expf(logprob_i); their sum, the timestamp statistics (whisper.cpp:4851-4866), and the per-thread share of the
    // CDF: thread t owns the contiguous tokens [t * seg, (t + 1) * seg)
    const int seg = (n_vocab + nthr - 1) / nthr;
    const int i0 = min(n_vocab, tid * seg), i1 = min(n_vocab, i0 + seg);
    auto prob = [&](int i) -> float {
        const float x = row_s[i];
        if (!(x > -INFINITY) || (text_off && i < beg)) return 0.0f;
        return expf(x - lse);
    };
    double local = 0.0, p_ts_sum = 0.0;
    ArgMax best_ts{0.0f, 0x7fffffff};
    for (int i = i0; i < i1; ++i) {
        const float p = prob(i);
        local += (double) p;
        if (i >= beg) { p_ts_sum += (double) p; if (p > best_ts.v) best_ts = ArgMax{p, i}; }
    }
    const double total = block_reduce(local, dsum_op, sd);
    p_ts_sum = block_reduce(p_ts_sum, dsum_op, sd);
    best_ts = block_reduce(best_ts, argmax_first, sa);
    // exclusive prefix of the normalised shares over the threads
    const double share = total > 0.0 ? local / total : 0.0;
    double incl = share;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    __syncthreads();
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < warp; ++w) base += s_scan[w];
    const double lo = base + incl - share, hi = base + incl;                      // this thread covers cp in (lo, hi]
    const bool last_thread = i1 >= n_vocab && i0 < n_vocab;

    const int tid_ts = best_ts.v > 0.0f ? best_ts.i : rl[6];                      // no timestamp mass: 0 for whisper_sample_token (:4779), token_beg for _topk (:4847)
    const float pt_all = (float) ((double) best_ts.v / (p_ts_sum + 1e-10)), ptsum = (float) p_ts_sum;
    for (int d = 0; d < n_draws; ++d) {
        const double u = __ldg(draws + draw_off + d);
        // first i with cp_i >= u; the last cumulative probability counts as one (libstdc++ forces it), so every u < 1 finds its token
        const bool mine = (u > lo && u <= hi && share > 0.0) || (last_thread && u > hi) || (tid == 0 && u <= 0.0);
        if (mine && i0 < i1) {
            int pick = i1 - 1;
            double cp = lo;
            for (int i = i0; i < i1; ++i) {
                cp += (double) prob(i) / total;
                if (cp >= u) { pick = i; break; }
            }
            if (last_thread && u > hi) pick = n_vocab - 1;
            const float p = prob(pick);
            float * o = out + 6 * (size_t) (draw_off + d);
            int id = pick, tidv = tid_ts;
            float pt = pt_all;
            if (id >= beg) { tidv = id; pt = p; }                                 // whisper.cpp:4899-4902
            const float x = row_s[pick];
            o[0] = __int_as_float(id); o[1] = __int_as_float(tidv); o[2] = p; o[3] = (x > -INFINITY && !(text_off && pick < beg)) ? x - lse : -INFINITY;
            o[4] = pt; o[5] = ptsum;
        }
    }
}

// ---- SIMT tiled GEMM (debug engine) ----------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
k_gemm_simt(const Operand A, const Operand W, int N, int M, int K, int nb1, int w_batched, const GemmEpi epi) {
    __shared__ float As[16][65];
    __shared__ float Ws[16][65];
    const int b1 = blockIdx.z % nb1, b2 = blockIdx.z / nb1;
    const __half * Ap = A.p + (int64_t) b2 * A.bs2 + (int64_t) b1 * A.bs1;
    const __half * Wp = W.p + (w_batched ? (int64_t) b2 * W.bs2 + (int64_t) b1 * W.bs1 : 0);
    const int n0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx -> features, ty -> rows
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            const int rr = i >> 4, kk = i & 15;
            const int n = n0 + rr, m = m0 + rr, k = k0 + kk;
            As[kk][rr] = (n < N && n < A.rows && k < K) ? __half2float(Ap[(int64_t) n * A.ld + k]) : 0.0f;
            Ws[kk][rr] = (m < M && m < W.rows && k < K) ? __half2float(Wp[(int64_t) m * W.ld + k]) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], wv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; wv[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + tx * 4 + j;
            if (m >= M) continue;
            const int seg_i = epi.nseg > 1 ? m / epi.seg_m : 0;
            const EpiSeg & sg = epi.seg[seg_i];
            const int ml = m - seg_i * epi.seg_m;
            float pre;
            const float v = epi_value(sg, epi.gelu_lut, acc[i][j], n, ml, &pre);
            epi_store(sg, v, pre, n, ml, b1, b2);
        }
    }
}

}  // namespace

// ---- launchers ---------------------------------------------------------------------------------------------------------------------

void launch_mel_to_tokens(const float * mel, __half * out, int n_mels, int n_frames, cudaStream_t st) {
    dim3 grid((n_frames + 31) / 32, (n_mels + 31) / 32), block(32, 8);
    k_mel_to_tokens<<<grid, block, 0, st>>>(mel, out, n_mels, n_frames);
}

void launch_layernorm(const float * x, const float * gamma, const float * beta, __half * out16, float * out32, int rows,
                      int d, float eps, cudaStream_t st) {
    const int wpb = 8;
    const unsigned grid = (unsigned) ((rows + wpb - 1) / wpb);
    const bool aligned = (((uintptr_t) x | (uintptr_t) gamma | (uintptr_t) beta | (uintptr_t) out32) & 15) == 0 && ((uintptr_t) out16 & 7) == 0;
    if (aligned && d % 128 == 0) {
        switch (d / 128) {
            case 3:  k_layernorm_vec<3><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, out16, out32, rows, d, eps); return;
            case 4:  k_layernorm_vec<4><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, out16, out32, rows, d, eps); return;
            case 6:  k_layernorm_vec<6><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, out16, out32, rows, d, eps); return;
            case 8:  k_layernorm_vec<8><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, out16, out32, rows, d, eps); return;
            case 10: k_layernorm_vec<10><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, out16, out32, rows, d, eps); return;
            default: break;
        }
    }
    k_layernorm<<<grid, wpb * 32, 0, st>>>(x, gamma, beta, out16, out32, rows, d, eps);
}

void launch_softmax_rows(const float * S, __half * P, int64_t rows, int n_cols, int ld_s, int ld_p, const uint16_t * exp_lut,
                         cudaStream_t st) {
    if (n_cols > 32 * 4 * kSmaxPer || (ld_s & 3) || (ld_p & 3)) { fprintf(stderr, "whisper_b200: softmax row shape %d / %d / %d not supported\n", n_cols, ld_s, ld_p); return; }
    const int wpb = 8;
    k_softmax_rows<<<(unsigned) ((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(S, P, rows, n_cols, ld_s, ld_p, exp_lut);
}

void launch_embed(const __half * te, const float * pe, const int * token, const int * pos, float * x, int n, int d,
                  cudaStream_t st) {
    k_embed<<<n, 128, 0, st>>>(te, pe, token, pos, x, n, d);
}

void launch_gather_rows(const float * src, const int * idx, float * dst, int n, int d, cudaStream_t st) {
    k_gather_rows<<<n, 128, 0, st>>>(src, idx, dst, n, d);
}

template <int R, int RW>
static void launch_skinny_t(const SkinnyIn & in, const __half * W, int n, int M, int K, const GemmEpi & epi, cudaStream_t st) {
    const int warps = 4;
    const size_t smem = (size_t) R * K * sizeof(__half);
    static bool attr_done[16] = {};                // (function attributes are per device: a process may drive several)
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && !attr_done[dev & 15]) {
        cudaFuncSetAttribute(k_gemm_skinny<R, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_done[dev & 15] = true;
    }
    const int grid = (M + warps * RW - 1) / (warps * RW);
    k_gemm_skinny<R, RW><<<grid, warps * 32, smem, st>>>(in, W, n, M, K, epi);
}

template <int R>
static void launch_skinny_r(const SkinnyIn & in, const __half * W, int n, int M, int K, const GemmEpi & epi, cudaStream_t st) {
    if (M <= 768)       launch_skinny_t<R, 1>(in, W, n, M, K, epi, st);
    else if (M <= 4096) launch_skinny_t<R, 2>(in, W, n, M, K, epi, st);
    else                launch_skinny_t<R, 4>(in, W, n, M, K, epi, st);
}

void launch_gemm_skinny(const SkinnyIn & in, const __half * W, int n, int M, int K, const GemmEpi & epi, cudaStream_t st) {
    if (n <= 1)      launch_skinny_r<1>(in, W, n, M, K, epi, st);
    else if (n <= 2) launch_skinny_r<2>(in, W, n, M, K, epi, st);
    else if (n <= 4) launch_skinny_r<4>(in, W, n, M, K, epi, st);
    else             launch_skinny_r<8>(in, W, n, M, K, epi, st);
}

template <int NT>
static void launch_skinny_mma_t(const __half * x16, int64_t x_ld, const __half * W, int n, int M, int K, const GemmEpi & epi, cudaStream_t st) {
    const int m_tiles = (M + 15) / 16;
    int kz = 1;
    if (m_tiles <= 296 && K % 128 == 0) kz = 4;
    else if (m_tiles <= 592 && K % 64 == 0) kz = 2;
    const size_t smem = (size_t) 8 * NT * (K + 32) * sizeof(__half) + (size_t) 4 * 32 * NT * 4 * sizeof(float);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 15]) {
        cudaFuncSetAttribute(k_gemm_skinny_mma<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done[dev & 15] = true;
    }
    const int grid = (m_tiles + (4 / kz) - 1) / (4 / kz);
    k_gemm_skinny_mma<NT><<<grid, 128, smem, st>>>(x16, x_ld, W, n, M, K, kz, epi);
}

void launch_gemm_skinny_mma(const __half * x16, int64_t x_ld, const __half * W, int n, int M, int K, const GemmEpi & epi, cudaStream_t st) {
    if (n <= 16) launch_skinny_mma_t<2>(x16, x_ld, W, n, M, K, epi, st);
    else         launch_skinny_mma_t<4>(x16, x_ld, W, n, M, K, epi, st);
}

// Decoder SELF-attention of wide passes: one WARP per (row, head).  The live self-attention cache of a sequence holds a few dozen to a
// few hundred keys — 12 KB of K per (row, head) — so a 256-thread CTA (let alone a cluster) per item spends its time in block barriers:
// 27 us per launch at 512 rows, four launches per token step.  Here a warp walks the whole item alone: no block-wide barrier, eight
// items per CTA.  The arithmetic is k_decode_attention's for a one-CTA cluster, operation for operation — 8 lanes x 8 features per key
// with the same shuffle tree, the exact f64 sum of the table exponentials, p rounded to f16, per-lane partial P V over keys
// 8 lane .. 8 lane + 7 (+ 256 m) followed by the same warp reduction — so the output bits are the same.
__global__ void __launch_bounds__(256)
k_decode_self_attention_warp(const AttnArgs a) {
    extern __shared__ __align__(16) uint8_t smem_sw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + warp;
    if (item >= a.n * a.n_head) return;                                   // (whole warps leave; nothing below synchronises the block)
    const int r = item / a.n_head, h = item - r * a.n_head;
    const int n_keys = a.n_keys_dev ? min(*a.n_keys_dev, a.n_keys) : a.n_keys;
    const int n_pad = (n_keys + 7) & ~7;
    const int per_max = (a.n_keys + 7) & ~7;                              // what the launcher sized shared memory for
    float *  sc  = (float *) (smem_sw + (size_t) warp * per_max * (sizeof(float) + sizeof(__half)));
    __half * p16 = (__half *) (sc + per_max);
    const int g = lane & 7;
    const int64_t koff = a.koff ? a.koff[r] : 0;
    const int64_t voff = a.voff ? a.voff[r] : 0;

    float q[8];
    {
        const uint4 qv = *(const uint4 *) (a.q + (int64_t) r * a.d + h * 64 + g * 8);
        const __half2 * qh = (const __half2 *) &qv;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(qh[i]); q[2 * i] = f.x; q[2 * i + 1] = f.y; }
    }
    const __half * Kb = a.K + koff + h * 64 + g * 8;
    const float * mrow = a.mask ? a.mask + (int64_t) r * a.ld_mask : nullptr;

    // scores: 8 lanes per key, 4 keys per pass, kU passes in flight
    float mx = -INFINITY;
    constexpr int kU = 4;
    for (int j0 = lane >> 3; j0 < n_pad; j0 += 4 * kU) {
        uint4 kv[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int j = j0 + 4 * u;
            kv[u] = (j < n_keys) ? __ldg((const uint4 *) (Kb + (int64_t) j * a.d)) : make_uint4(0, 0, 0, 0);
        }
        float dt[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const __half2 * hh = (const __half2 *) &kv[u];
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hh[i]);
                acc = fmaf(f.x, q[2 * i], acc); acc = fmaf(f.y, q[2 * i + 1], acc);
            }
            dt[u] = acc;
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
            for (int u = 0; u < kU; ++u) dt[u] += __shfl_xor_sync(0xffffffffu, dt[u], o);
        }
        if (g == 0) {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int j = j0 + 4 * u;
                if (j < n_pad) {
                    float v = -INFINITY;
                    if (j < n_keys) v = mrow ? __fadd_rn(dt[u], mrow[j]) : dt[u];
                    sc[j] = v; mx = fmaxf(mx, v);
                }
            }
        }
    }
    mx = warp_max(mx);
    __syncwarp();

    double sum = 0.0;
    for (int j = lane; j < n_pad; j += 32) {
        const float v = sc[j];
        float e = 0.0f;
        if (v != -INFINITY) e = exp_table(a.exp_lut, __fsub_rn(v, mx));
        sc[j] = e;
        sum += (double) e;                                                // exact in any order: every term is a multiple of 2^-24
    }
    sum = warp_sum(sum);
    const float inv = (float) (1.0 / sum);
    for (int j = lane; j < n_pad; j += 32) p16[j] = __float2half_rn(__fmul_rn(sc[j], inv));
    __syncwarp();

    // P V: feature group fg = the eight features warp fg of k_decode_attention owns; lane <-> keys 8 lane .. 8 lane + 7 (+ 256 m)
    const __half * vhead = a.Vt + voff + (int64_t) (h * 64) * a.ld_v;
    __half * orow = a.out + (int64_t) r * a.d + h * 64;
#pragma unroll 1
    for (int fg = 0; fg < 8; ++fg) {
        const __half * vbase = vhead + (int64_t) (fg * 8) * a.ld_v;
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
        for (int j = lane * 8; j < n_pad; j += 256) {
            uint4 vv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] = __ldg((const uint4 *) (vbase + (int64_t) i * a.ld_v + j));
            const uint4 pv = *(const uint4 *) (p16 + j);
#pragma unroll
            for (int i = 0; i < 8; ++i) fma8(acc[i], vv[i], pv);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float t = warp_sum(acc[i]);
            if (lane == 0) orow[fg * 8 + i] = __float2half_rn(0.0f + t);
        }
    }
}

void launch_decode_attention(const AttnArgs & a, cudaStream_t st) {
    // short key ranges (self-attention of wide passes): a warp per (row, head)
    static const int warp_path = [] { const char * e = getenv("WHISPER_B200_SELF_ATTN_WARP"); return e ? atoi(e) : 1; }();
    if (warp_path && a.n_keys <= 512 && a.n * a.n_head >= 64) {
        const size_t smem = (size_t) 8 * ((a.n_keys + 7) & ~7) * (sizeof(float) + sizeof(__half));      // <= 24 KB
        k_decode_self_attention_warp<<<(a.n * a.n_head + 7) / 8, 256, smem, st>>>(a);
        return;
    }

    // split the keys of one (head, row) over S CTAs so that about two waves of CTAs are in flight
    int S = 1;
    const int pairs = a.n_head * a.n;
    while (S < 8 && pairs * S * 2 <= 1184 && a.n_keys / (S * 2) >= 96) S *= 2;
    const int per = ((((a.n_keys + S - 1) / S) + 7) & ~7);
    const size_t smem = (size_t) per * (sizeof(float) + sizeof(__half));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.n_head * S, a.n, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_decode_attention, a);
}

void launch_sample_greedy(const float * logits, int rows, int n_vocab, const int * rule, const uint8_t * cls, int token_beg,
                          int token_eot, float * out, cudaStream_t st) {
    const size_t smem = (size_t) n_vocab * sizeof(float);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 15]) {
        cudaFuncSetAttribute(k_sample_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done[dev & 15] = true;
    }
    k_sample_greedy<<<rows, 1024, smem, st>>>(logits, n_vocab, rule, cls, token_beg, token_eot, out);
}

void launch_sample_dist(const float * logits, int rows, int n_vocab, const int * drule, const double * draws, const uint8_t * cls, int token_beg,
                        int token_eot, float * out, cudaStream_t st) {
    const size_t smem = (size_t) n_vocab * sizeof(float);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 15]) {
        cudaFuncSetAttribute(k_sample_dist, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done[dev & 15] = true;
    }
    k_sample_dist<<<rows, 1024, smem, st>>>(logits, n_vocab, drule, draws, cls, token_beg, token_eot, out);
}

void launch_gemm_simt(const Operand & A, const Operand & W, const GemmShape & sh, const GemmEpi & epi, cudaStream_t st) {
    const int w_batched = (sh.nb1 * sh.nb2 > 1) && (W.bs1 != 0 || W.bs2 != 0);
    dim3 grid((sh.N + 63) / 64, (sh.M + 63) / 64, sh.nb1 * sh.nb2);
    k_gemm_simt<<<grid, 256, 0, st>>>(A, W, sh.N, sh.M, sh.K, sh.nb1, w_batched, epi);
}

}  // namespace wb200
