// Launchers of the non-tensor-core kernels of the path (kernels.cu).  All are HBM/latency bound; see DESIGN.md.
#pragma once

#include "dev.cuh"

namespace wb200 {

// mel window f32 [n_mels][n_frames] (one chunk) -> f16 token-major [n_frames + 2][n_mels] with zero rows at both ends
// (the zero padding of ggml_conv_1d_ph, ggml.c:5322-5366, realised once instead of per im2col element).
void launch_mel_to_tokens(const float * mel, __half * out, int n_mels, int n_frames, cudaStream_t st);

// LayerNorm with the reference's arithmetic (ggml.c:9301-9352 then separate mul / add, whisper.cpp:1821-1826):
// f64 sums, mean/variance rounded to f32, (x - mean) * (1/sqrtf(var + eps)) * gamma + beta with one rounding per op.
// Writes f16 (next GEMM operand) and / or f32 (embd_enc).
void launch_layernorm(const float * x, const float * gamma, const float * beta, __half * out16, float * out32,
                      int rows, int d, float eps, cudaStream_t st);

// Row softmax with the reference's arithmetic (ggml.c:11116-11201): f32 max, exp through the f16 table, f64 sum,
// multiply by (float)(1/sum); result rounded to f16 (what the following mul_mat does to it, ggml.c:9841-9857).
// S: f32 [rows][ld_s], P: f16 [rows][ld_p]; columns [n_cols, ld_p) of P are zeroed.
void launch_softmax_rows(const float * S, __half * P, int64_t rows, int n_cols, int ld_s, int ld_p,
                         const uint16_t * exp_lut, cudaStream_t st);

// Fused encoder self-attention (attn_enc.cu): O = softmax(Q K^T / sqrt(64)) V per (chunk, head) on tcgen05 with the scores kept in
// TMEM / shared memory.  q16, k16: f16 [B][T][d] (head h = columns 64h..64h+63); vt16: f16 [B][d][Tp] (V transposed);
// out16: f16 [B][T][d].  Same arithmetic as launch_softmax_rows between two mul_mats (see there).  exp_lut must be zero from entry
// 0x8000 + attention_enc_table_entries() - 1 on (the caller checks once): only that many entries are staged on chip.
bool launch_attention_enc(const __half * q16, const __half * k16, const __half * vt16, __half * out16, int B, int T, int Tp, int d,
                          int n_head, const uint16_t * exp_lut, cudaStream_t st, int variant = -1);
int attention_enc_table_entries();

// ---- decoder ------------------------------------------------------------------------------------------------------------

// x[r][:] = f32(te[token[r]][:]) + pe[pos[r]][:]      (whisper.cpp:2229-2233)
void launch_embed(const __half * te, const float * pe, const int * token, const int * pos, float * x, int n, int d,
                  cudaStream_t st);

// dst[i][:] = src[idx[i]][:]  (f32 rows) — picks the rows whose logits were requested
void launch_gather_rows(const float * src, const int * idx, float * dst, int n, int d, cudaStream_t st);

// Skinny contraction for n <= 8 rows per pass (decode steps): each warp streams weight rows with 16-byte loads.
// Input is either f16 rows (x16) or f32 rows with a fused LayerNorm prologue (x32 + gamma/beta).
struct SkinnyIn {
    const __half * x16 = nullptr; int64_t x16_ld = 0;
    const float *  x32 = nullptr; int64_t x32_ld = 0;
    const float *  gamma = nullptr; const float * beta = nullptr; float eps = 1e-5f;
};
void launch_gemm_skinny(const SkinnyIn & in, const __half * W, int n, int M, int K, const GemmEpi & epi, cudaStream_t st);

// Contraction for 9..32 rows (batched decode steps): weights stream straight from HBM into mma.sync.m16n8k16 A fragments
// (two 16-byte loads per lane per 32 k), activations sit in shared memory as B fragments, K is split over the warps of
// a CTA when there are few weight rows.  HBM-bound by design: the tensor-core instruction is only there so that 32
// activation rows cost no more shared-memory traffic than one.
void launch_gemm_skinny_mma(const __half * x16, int64_t x_ld, const __half * W, int n, int M, int K, const GemmEpi & epi, cudaStream_t st);

// One (row, head) of decoder attention against an f16 cache:  softmax(K q + mask) V   with table exp / f64 sum.
//   q    : f16 [n][d] (already scaled), head h uses columns [64h, 64h+64)
//   K    : f16 rows of d, row j of sequence slot s at  Kbase + koff[r] + j*d          (koff in elements, per row)
//   Vt   : f16 [d][ld_v] transposed,     column j at   Vbase + voff[r] + (64h+i)*ld_v + j
//   mask : f32 [n][ld_mask] additive (0 / -inf) or nullptr
//   out  : f16 [n][d]
struct AttnArgs {
    const __half * q = nullptr; const __half * K = nullptr; const __half * Vt = nullptr;
    const int64_t * koff = nullptr; const int64_t * voff = nullptr;   // per-row cache offsets (nullptr => 0)
    const float * mask = nullptr; int ld_mask = 0;
    __half * out = nullptr;
    int n = 0, d = 0, n_head = 0, n_keys = 0; int64_t ld_v = 0;
    const int * n_keys_dev = nullptr;   // if set: the live key count is read on the device (CUDA-graph replay); n_keys is its upper bound
    const uint16_t * exp_lut = nullptr;
};
void launch_decode_attention(const AttnArgs & a, cudaStream_t st);

// Greedy step on the device: whisper_process_logits' rules + log-softmax + timestamp-mass test + whisper_sample_token's
// argmax (whisper.cpp:4493-4834) for `rows` rows of logits [rows][n_vocab].  rule: 4 x int32 per row (forward.h
// SampleRule); cls: per-token class bits (1 always suppressed, 2 non-speech symbol, 4 blank/eot at sequence start,
// 8 solm); out: 6 x 32-bit per row = {id, tid, p, plog, pt, ptsum}.
void launch_sample_greedy(const float * logits, int rows, int n_vocab, const int * rule, const uint8_t * cls, int token_beg,
                          int token_eot, float * out, cudaStream_t st);

// Sampling from the distribution on the device (t > 0 best-of decoders, beam search): whisper_process_logits with a temperature +
// whisper_sample_token(best = false) / _topk (whisper.cpp:4493-4720, 4777-4909) for `rows` rows of logits.  drule: 8 x int32 per row =
// {flags, tid0_initial, tid0_seek (forward.h SampleRule), n_draws, temperature as float bits, offset of the row's first draw, tid when
// no timestamp has mass, 0};
// draws: the uniform variates in [0, 1) the host took from each decoder's generator (std::generate_canonical<double, 53>), all rows
// concatenated; out: 6 x 32-bit per DRAW = {id, tid, p, plog, pt, ptsum}.
void launch_sample_dist(const float * logits, int rows, int n_vocab, const int * drule, const double * draws, const uint8_t * cls, int token_beg,
                        int token_eot, float * out, cudaStream_t st);

}  // namespace wb200
