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
