// Host interface of the encoder GEMM with TMA-store epilogue (gemm_enc.cu).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wb200 {

// device-side description of one feature segment (tiles never straddle a segment)
struct EncSeg {
    const float * bias = nullptr;     // [seg_m] or nullptr
    float scale = 1.0f;               // applied after the bias
    int gelu = 0;                     // GELU table after bias / scale (f16 outputs only)
    int transposed = 0;               // f16 output stored as [feature][token] (V^T layouts)
    const int * bmap = nullptr;       // optional: batch entry -> index along the output's batch dimension (device slots)
};
struct EncGemmArgs {
    int N = 0, M = 0, K = 0, nseg = 1, seg_m = 0, tiles_n = 0, tiles_m = 0, n_tiles = 0, any_gelu = 0, res_batched = 0;
    EncSeg seg[3];
    const uint16_t * gelu_lut = nullptr;
    const float * res = nullptr;      // RES32 mode: f32 residual [res_batched ? nb : 1][res_rows][res_ld], read by the epilogue threads (a row segment each)
    int64_t res_ld = 0, res_bs = 0;
    int res_rows = 0;
    int dbg = 0;                      // timing experiments (WHISPER_B200_GEMM_DBG): 1 = no weight-tile traffic, 2 = no activation-tile traffic after a CTA's first tile, 4 = no epilogue work, 8 = no MMAs
};

// host-side description
struct EncOut {
    void * p = nullptr;               // f16 (HALF mode) or f32 (RES32 mode) output base of this segment
    int64_t ld = 0;                   // elements between rows (tokens; features when transposed)
    int64_t bs = 0;                   // elements between batch entries of the OUTPUT
    int n_batch_out = 0;              // size of the output's batch dimension (0: same as the input's)
    const float * bias = nullptr; float scale = 1.0f; int gelu = 0; int transposed = 0; const int * bmap = nullptr;
};
struct EncGemm {
    const __half * A = nullptr; int64_t a_ld = 0, a_bs = 0; int a_rows = 0;   // activations [nb][a_rows][K], rows a_ld apart
    const __half * W = nullptr; int64_t w_ld = 0;                             // weights [M][K]
    int N = 0, M = 0, K = 0, nb = 1;                                          // N = tokens per batch entry
    int nseg = 1, seg_m = 0;
    EncOut out[3];
    bool res32 = false;                                                       // RES32 mode: out[0] is f32 = acc + bias + res
    const float * res = nullptr; int64_t res_ld = 0, res_bs = 0; int res_rows = 0;
    const uint16_t * gelu_lut = nullptr;
};

bool gemm_enc_usable(const EncGemm & g);
bool launch_gemm_enc(const EncGemm & g, cudaStream_t st);
void gemm_enc_forget_maps();

}  // namespace wb200
