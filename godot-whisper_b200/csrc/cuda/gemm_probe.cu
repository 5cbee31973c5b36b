// whisper_b200_gemm_f16 — stand-alone contraction on host buffers (include/whisper_b200.h).  Used by the kernel unit
// tests (tcgen05 engine vs SIMT engine vs numpy) and by bench.py's tensor-core roofline probe.
#include "../common.h"
#include "dev.cuh"

#include <vector>

using namespace wb200;

extern "C" WHISPER_B200_API int whisper_b200_gemm_f16(const void * A_host_f16, const void * B_host_f16, float * C_host, int M, int N,
                                                      int K, int engine, int iters, float * ms_per_iter) {
    // C[n][m] = sum_k A[m][k] * B[n][k]   (A plays the weight role, B the activation role)
    if (M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0) return -1;
    __half * dA = nullptr, * dB = nullptr;
    float * dC = nullptr;
    cudaStream_t st = nullptr;
    int rc = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    do {
        if (cudaMalloc(&dA, (size_t) M * K * 2) != cudaSuccess || cudaMalloc(&dB, (size_t) N * K * 2) != cudaSuccess ||
            cudaMalloc(&dC, (size_t) N * M * 4) != cudaSuccess) { rc = -2; break; }
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaMemcpy(dA, A_host_f16, (size_t) M * K * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B_host_f16, (size_t) N * K * 2, cudaMemcpyHostToDevice);
        cudaMemset(dC, 0, (size_t) N * M * 4);
        cudaDeviceSynchronize();
        Operand act; act.p = dB; act.ld = K; act.rows = N;
        Operand wgt; wgt.p = dA; wgt.ld = K; wgt.rows = M;
        GemmShape sh; sh.N = N; sh.M = M; sh.K = K;
        GemmEpi epi; epi.seg[0].out32 = dC; epi.seg[0].out32_ld = M;
        auto run = [&]() -> bool {
            if (engine == 1) { launch_gemm_simt(act, wgt, sh, epi, st); return true; }
            return launch_gemm_tc(act, wgt, sh, epi, st);
        };
        if (!run()) { rc = -3; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = -4; break; }
        if (iters > 0) {
            cudaEventRecord(e0, st);
            for (int i = 0; i < iters; ++i) run();
            cudaEventRecord(e1, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) { rc = -4; break; }
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms_per_iter) *ms_per_iter = ms / iters;
        }
        if (cudaMemcpy(C_host, dC, (size_t) N * M * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = -5; break; }
    } while (0);
    if (rc != 0) {
        WB_LOG_ERROR("%s: failed (%d): %s\n", __func__, rc, cudaGetErrorString(cudaGetLastError()));
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
    return rc;
}

#include "gemm_enc.cuh"
#include "../tables.h"

// whisper_b200_gemm_enc_probe — the encoder GEMM with the TMA-store epilogue (gemm_enc.cu) on host buffers, one mode per call:
//   mode 0: out f16 [N][M] = acc + bias          1: ... then GELU table          2: out f16 [M][ldt] transposed (ldt = N rounded up to 8)
//   mode 3: out f32 [N][M] = acc + bias + res    (res f32 [N][M])                4: mode 0 with scale 0.25 and three feature segments
// act: f16 [N][K], wgt: f16 [M][K], bias: f32 [M] or NULL.  Returns 0, or a negative code.
extern "C" WHISPER_B200_API int whisper_b200_gemm_enc_probe(const void * act_f16, const void * wgt_f16, const float * bias, const float * res,
                                                            void * out, int N, int M, int K, int mode, int iters, float * ms_per_iter) {
    if (M <= 0 || N <= 0 || K <= 0 || (K % 8) != 0 || mode < 0 || mode > 4) return -1;
    const int ldt = (N + 7) & ~7;
    const size_t out_bytes = mode == 3 ? (size_t) N * M * 4 : mode == 2 ? (size_t) M * ldt * 2 : (size_t) N * M * 2;
    __half * dA = nullptr, * dW = nullptr;
    float * dBias = nullptr, * dRes = nullptr;
    void * dOut = nullptr;
    uint16_t * dLut = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 0;
    do {
        if (cudaMalloc(&dA, (size_t) N * K * 2) != cudaSuccess || cudaMalloc(&dW, (size_t) M * K * 2) != cudaSuccess ||
            cudaMalloc(&dOut, out_bytes) != cudaSuccess || cudaMalloc(&dLut, 65536 * 2) != cudaSuccess) { rc = -2; break; }
        if (bias && cudaMalloc(&dBias, (size_t) M * 4) != cudaSuccess) { rc = -2; break; }
        if (mode == 3 && (!res || cudaMalloc(&dRes, (size_t) N * M * 4) != cudaSuccess)) { rc = -2; break; }
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaMemcpy(dA, act_f16, (size_t) N * K * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dW, wgt_f16, (size_t) M * K * 2, cudaMemcpyHostToDevice);
        if (bias) cudaMemcpy(dBias, bias, (size_t) M * 4, cudaMemcpyHostToDevice);
        if (dRes) cudaMemcpy(dRes, res, (size_t) N * M * 4, cudaMemcpyHostToDevice);
        {
            std::vector<uint16_t> g(65536), e(65536);
            build_f16_tables(g.data(), e.data());
            cudaMemcpy(dLut, g.data(), 65536 * 2, cudaMemcpyHostToDevice);
        }
        cudaMemset(dOut, 0, out_bytes);
        cudaDeviceSynchronize();
        EncGemm g;
        g.A = dA; g.a_ld = K; g.a_rows = N; g.W = dW; g.w_ld = K; g.N = N; g.M = M; g.K = K; g.gelu_lut = dLut;
        g.out[0].p = dOut; g.out[0].ld = M; g.out[0].bias = dBias;
        if (mode == 1) g.out[0].gelu = 1;
        if (mode == 2) { g.out[0].transposed = 1; g.out[0].ld = ldt; }
        if (mode == 3) { g.res32 = true; g.res = dRes; g.res_ld = M; g.res_rows = N; }
        if (mode == 4) {
            if (M % 3 != 0) { rc = -1; break; }
            g.nseg = 3; g.seg_m = M / 3;
            for (int i = 0; i < 3; ++i) {
                g.out[i].p = (__half *) dOut + (size_t) i * g.seg_m; g.out[i].ld = M; g.out[i].bias = dBias ? dBias + (size_t) i * g.seg_m : nullptr;
                g.out[i].scale = 0.25f;
            }
        }
        if (!gemm_enc_usable(g)) { rc = -6; break; }
        if (!launch_gemm_enc(g, st)) { rc = -3; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = -4; break; }
        if (iters < 0) iters = -iters;
        if (iters > 0) {                    // (mode 3: res and out are distinct buffers here, so repeats compute the same thing)
            cudaEventRecord(e0, st);
            for (int i = 0; i < iters; ++i) launch_gemm_enc(g, st);
            cudaEventRecord(e1, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) { rc = -4; break; }
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms_per_iter) *ms_per_iter = ms / iters;
        }
        if (cudaMemcpy(out, dOut, out_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = -5; break; }
    } while (0);
    if (rc != 0) WB_LOG_ERROR("%s: failed (%d): %s\n", __func__, rc, cudaGetErrorString(cudaGetLastError()));
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    cudaFree(dA); cudaFree(dW); cudaFree(dBias); cudaFree(dRes); cudaFree(dOut); cudaFree(dLut);
    return rc;
}

#include "kernels.cuh"

// whisper_b200_attn_enc_probe — the fused encoder attention (attn_enc.cu) on host buffers: q, k f16 [B][T][d] (head h = columns
// 64 h .. 64 h + 63), vt f16 [B][d][Tp] (V transposed, Tp = T rounded up to 8), out f16 [B][T][d].  `variant` selects the kernel
// configuration (see launch_attention_enc; < 0 = the default).  Returns 0, or a negative code.
extern "C" WHISPER_B200_API int whisper_b200_attn_enc_probe(const void * q_f16, const void * k_f16, const void * vt_f16, void * out_f16, int B, int T,
                                                            int d, int n_head, int variant, int iters, float * ms_per_iter) {
    if (B <= 0 || T <= 0 || n_head <= 0 || d != 64 * n_head) return -1;
    const int Tp = (T + 7) & ~7;
    const size_t qk_bytes = (size_t) B * T * d * 2, v_bytes = (size_t) B * d * Tp * 2;
    __half * dQ = nullptr, * dK = nullptr, * dV = nullptr, * dO = nullptr;
    uint16_t * dLut = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 0;
    do {
        if (cudaMalloc(&dQ, qk_bytes) != cudaSuccess || cudaMalloc(&dK, qk_bytes) != cudaSuccess || cudaMalloc(&dV, v_bytes) != cudaSuccess ||
            cudaMalloc(&dO, qk_bytes) != cudaSuccess || cudaMalloc(&dLut, 65536 * 2) != cudaSuccess) { rc = -2; break; }
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaMemcpy(dQ, q_f16, qk_bytes, cudaMemcpyHostToDevice);
        cudaMemcpy(dK, k_f16, qk_bytes, cudaMemcpyHostToDevice);
        cudaMemcpy(dV, vt_f16, v_bytes, cudaMemcpyHostToDevice);
        {
            std::vector<uint16_t> g(65536), e(65536);
            build_f16_tables(g.data(), e.data());
            cudaMemcpy(dLut, e.data(), 65536 * 2, cudaMemcpyHostToDevice);
        }
        cudaMemset(dO, 0, qk_bytes);
        cudaDeviceSynchronize();
        if (!launch_attention_enc(dQ, dK, dV, dO, B, T, Tp, d, n_head, dLut, st, variant)) { rc = -3; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = -4; break; }
        if (iters > 0) {
            cudaEventRecord(e0, st);
            for (int i = 0; i < iters; ++i) launch_attention_enc(dQ, dK, dV, dO, B, T, Tp, d, n_head, dLut, st, variant);
            cudaEventRecord(e1, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) { rc = -4; break; }
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms_per_iter) *ms_per_iter = ms / iters;
        }
        if (cudaMemcpy(out_f16, dO, qk_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = -5; break; }
    } while (0);
    if (rc != 0) WB_LOG_ERROR("%s: failed (%d): %s\n", __func__, rc, cudaGetErrorString(cudaGetLastError()));
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (st) cudaStreamDestroy(st);
    cudaFree(dQ); cudaFree(dK); cudaFree(dV); cudaFree(dO); cudaFree(dLut);
    return rc;
}
