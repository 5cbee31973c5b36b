// Device-side vocabulary shared by every kernel of the B200 forward pass.
//
// Activations live token-major in HBM: a "row" is one audio frame / one text token, features are contiguous.
// Every linear map of the reference is the contraction  D[n][m] = sum_k A[n][k] * W[m][k]  with both operands
// K-major f16 and f32 accumulation — exactly ggml's mul_mat contract (ggml.c:9737-9948: src1 rounded to f16 row-wise,
// f32 accumulate).  What follows the contraction in the reference graph (bias add, scale, GELU table, residual add,
// f16 stores into K/V layouts) is folded into the epilogue described by EpiSeg.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wb200 {

// One operand of a (batched) contraction: element (row, k, b1, b2) = p[b2*bs2 + b1*bs1 + row*ld + k]   (f16)
struct Operand {
    const __half * p = nullptr;
    int64_t ld = 0;        // elements between consecutive rows (may be < K: overlapping rows = implicit im2col)
    int64_t bs1 = 0;       // inner batch stride (e.g. attention head)
    int64_t bs2 = 0;       // outer batch stride (e.g. audio chunk)
    int     rows = 0;      // rows per batch entry (bounds for zero fill)
};

// Epilogue of one feature segment.  Order of operations = order of the reference graph nodes:
//   v = acc; v += bias[m]; v *= scale; v = gelu_table(v); v += res[n % res_mod][m];   then the stores.
struct EpiSeg {
    const float *  bias     = nullptr;   // [seg_m] f32
    float          scale    = 1.0f;
    int            gelu     = 0;
    const float *  res      = nullptr;   // f32 [.. ][res_ld]
    int64_t        res_ld   = 0;
    int            res_mod  = 0;         // 0 => row n, else row n % res_mod (positional embedding)
    float *        out32    = nullptr;   // f32 [n][m]
    int64_t        out32_ld = 0, out32_bs1 = 0, out32_bs2 = 0;
    __half *       out16    = nullptr;   // f16 [n][m]  (row index optionally remapped: KV-cache cells)
    int64_t        out16_ld = 0, out16_bs1 = 0, out16_bs2 = 0;
    const int *    rowmap16 = nullptr;
    int            out16_pre = 0;        // 1 => out16 receives the value before the residual add
    __half *       out16t   = nullptr;   // f16 transposed [m][n] (V layouts); column optionally remapped
    int64_t        out16t_ld = 0, out16t_bs1 = 0, out16t_bs2 = 0;
    const int *    rowmap16t = nullptr;
    const int *    bmap2    = nullptr;   // optional: outer batch index b2 -> index used with the *_bs2 strides (device slots of a pass)
};

struct GemmEpi {
    EpiSeg seg[3];
    int    seg_m = 0;      // features per segment (M = nseg * seg_m); tiles never straddle a segment
    int    nseg  = 1;
    const uint16_t * gelu_lut = nullptr;   // 65536 x f16 bits (ggml.c:1416-1423 table semantics)
};

__device__ __forceinline__ float gelu_table(const uint16_t * __restrict__ lut, float x) {
    const uint16_t h = __half_as_ushort(__float2half_rn(x));
    return __half2float(__ushort_as_half(__ldg(lut + h)));
}

// scalar epilogue for element (n, m_local) of a segment; *pre receives the value before the residual add
__device__ __forceinline__ float epi_value(const EpiSeg & s, const uint16_t * lut, float acc, int n, int m, float * pre = nullptr) {
    float v = acc;
    if (s.bias)          v = __fadd_rn(v, __ldg(s.bias + m));
    if (s.scale != 1.0f) v = __fmul_rn(v, s.scale);
    if (s.gelu)          v = gelu_table(lut, v);
    if (pre) *pre = v;
    if (s.res) {
        const int rn = s.res_mod ? (n % s.res_mod) : n;
        v = __fadd_rn(v, __ldg(s.res + (int64_t) rn * s.res_ld + m));
    }
    return v;
}

__device__ __forceinline__ void epi_store(const EpiSeg & s, float v, float v_pre, int n, int m, int b1, int b2) {
    if (s.bmap2) b2 = __ldg(s.bmap2 + b2);
    if (s.out32)  s.out32[(int64_t) b2 * s.out32_bs2 + (int64_t) b1 * s.out32_bs1 + (int64_t) n * s.out32_ld + m] = v;
    if (s.out16) {
        const int64_t r = s.rowmap16 ? (int64_t) __ldg(s.rowmap16 + n) : (int64_t) n;
        s.out16[(int64_t) b2 * s.out16_bs2 + (int64_t) b1 * s.out16_bs1 + r * s.out16_ld + m] = __float2half_rn(s.out16_pre ? v_pre : v);
    }
    if (s.out16t) {
        const int64_t c = s.rowmap16t ? (int64_t) __ldg(s.rowmap16t + n) : (int64_t) n;
        s.out16t[(int64_t) b2 * s.out16t_bs2 + (int64_t) b1 * s.out16t_bs1 + (int64_t) m * s.out16t_ld + c] = __float2half_rn(v);
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- host-side launch API (implemented in gemm_tc.cu / kernels.cu) -----------------------------------------------------

struct GemmShape {
    int N = 0;     // rows of A per batch entry (tokens)
    int M = 0;     // rows of W (features)
    int K = 0;
    int nb1 = 1, nb2 = 1;
};

// engine 0: TMA + tcgen05 (UMMA 128 x BN x 16, f32 accumulators in TMEM).  Returns false if a tensor map cannot be
// encoded (misaligned operand) — callers treat that as a hard error.
bool launch_gemm_tc(const Operand & A, const Operand & W, const GemmShape & sh, const GemmEpi & epi, cudaStream_t st);
// engine 1: plain SIMT tiles, same operands / epilogue (debug cross-check of the tensor-core path)
void launch_gemm_simt(const Operand & A, const Operand & W, const GemmShape & sh, const GemmEpi & epi, cudaStream_t st);

}  // namespace wb200
