// One decoder token step as ONE persistent cooperative kernel (decode_step.cu).
//
// whisper_decode_internal (/root/reference/thirdparty/whisper.cpp/whisper.cpp:2517-2595) evaluates a graph of ~40 small
// operators per generated token (whisper_build_graph_decoder, :2148-2505), each reading a few hundred KB.  On a B200 such
// a step is bound by launch / dependency latency, not by bytes, so the whole step — embedding, every decoder layer, the
// logits contraction against the token embedding, the logits rules and the greedy pick (whisper_process_logits +
// whisper_sample_token, :4493-4834) — runs here as one launch of one CTA per SM with grid-wide barriers between the
// dependent phases, and the weights of a CTA's next pieces of work are already in flight (cp.async) while it waits.
#pragma once

#include "dev.cuh"

#include <cuda.h>

namespace wb200 {

struct StepLayerW {
    const float * ln1_g, * ln1_b, * lnc_g, * lnc_b, * ln2_g, * ln2_b;
    const __half * wqkv; const float * bqkv;     // [3d][d], bias of the key part is zero (whisper.cpp:2255-2258 has none)
    const __half * wo;   const float * bo;
    const __half * wcq;  const float * bcq;
    const __half * wco;  const float * bco;
    const __half * w1;   const float * b1;
    const __half * w2;   const float * b2;
};

constexpr int kStepMaxRows   = 16;    // decoder rows (= live sequences) one group of CTAs handles
constexpr int kStepMaxGroups = 4;     // independent row groups per launch
constexpr int kStepMaxPhases = 112;   // 3 + 8 * n_text_layer  (up to 13 layers)

// One phase of the step, built on the host (the launch geometry is a pure function of the model and n).
enum { STEP_EMBED = 0, STEP_GEMM = 1, STEP_SELF = 2, STEP_CROSS = 3, STEP_FINAL = 4 };
enum { EPI_QKV = 0, EPI_RESID = 1, EPI_Q = 2, EPI_FC1 = 3, EPI_LOGITS = 4 };
struct StepPhase {
    int type = STEP_EMBED;
    int epi = 0;
    int n_jobs = 0;             // GEMM: blocks of 16*tj weight rows; attention: (row, head) items
    int tj = 1;                 // 16-row tiles per GEMM job (1, 2, 4 or 8); the 8 warps split K 8/tj ways
    int M = 0, K = 0;
    int layer = 0;
    int src_ln = 0;             // 1: operand = LayerNorm(x32) with g / b; 0: operand = x16 rows
    int x16_ld = 0;
    int ksplit = 1;             // a 16*tj-row block wider than one ring slot is streamed as ksplit sub-jobs of kc columns
    int kc = 0;
    int tma = 0;                // 1: the weight block is fetched by TMA into the swizzled box layout (logits phase)
    const __half * W = nullptr;
    const float * g = nullptr, * b = nullptr;
    const __half * x16 = nullptr;
    const float * bias = nullptr;
};

struct StepArgs {
    // model
    int d = 0, n_head = 0, n_layer = 0, n_vocab = 0;
    const StepPhase * phases = nullptr; int n_phases = 0;       // device array (group 0 / the only group)
    // Independent row groups: with n_groups > 1 the grid is cut into n_groups equal parts, part g serves rows
    // [sum n_grp[<g], + n_grp[g]) with its own phase table (planned for grid / n_groups CTAs), barrier words and partials.
    int n_groups = 1; int n_grp[4] = {0, 0, 0, 0}; const StepPhase * phases_grp[4] = {nullptr, nullptr, nullptr, nullptr};
    const __half * te = nullptr; const float * pe = nullptr;
    const uint16_t * gelu_lut = nullptr, * exp_lut = nullptr; const uint8_t * cls = nullptr;
    int token_beg = 0, token_eot = 0; float eps = 1e-5f, qscale = 1.0f;
    // caches: per layer strides are kv_cells*d (self) and Tmax*d / d*Tpmax (cross)
    __half * self_k = nullptr; __half * self_v = nullptr; const __half * cross_k = nullptr; const __half * cross_v = nullptr;
    int kv_cells = 0, Tmax = 0, Tpmax = 0;
    // this step (arrays live in the staging block the host copies before the launch)
    int n = 0, n_full = 0, n_audio_ctx = 0, ld_mask = 0;
    const int * token = nullptr, * pos = nullptr, * wslot = nullptr, * rule = nullptr, * rowmap_k = nullptr, * rowmap_v = nullptr;
    const int64_t * koff_self = nullptr, * voff_self = nullptr, * koff_cross = nullptr, * voff_cross = nullptr;
    const float * mask = nullptr; const int * n_kv_dev = nullptr;
    // workspaces
    float * x32 = nullptr; __half * q16 = nullptr; __half * attn16 = nullptr; __half * h16 = nullptr;
    float * logits = nullptr;          // [n_full][n_vocab] rows that go back to the host
    float * sampled = nullptr;         // [n - n_full][6]   {id, tid, p, plog, pt, ptsum}
    double * records = nullptr;        // [grid][kStepMaxRows][6] per-CTA sampler partials
    unsigned long long * bar = nullptr;        // monotonically increasing arrival counter
    unsigned long long * trace = nullptr;      // optional [grid][kStepMaxPhases][8] globaltimer stamps (diagnostics)
    // shared-memory plan (bytes)
    int xs_bytes = 0, slot_bytes = 0, chunk_keys = 0;
    // cross-attention K [slots*Lt*Tmax rows][d] and V^T [slots*Lt*d rows][Tpmax] as TMA tensor maps (128-byte swizzle;
    // boxes of 64 columns x 128 rows / 64 columns x 64 rows); chunk_keys_cross keys per cross-attention chunk
    alignas(64) CUtensorMap tm_cross_k;
    alignas(64) CUtensorMap tm_cross_v;
    // token embedding [n_vocab rows][d] for the logits phase: boxes of 64 columns x (16 * tj of that phase) rows
    alignas(64) CUtensorMap tm_te;
    int chunk_keys_cross = 0;
};

// Dynamic shared memory the kernel needs for a model of width d, or 0 if the step kernel cannot serve it.
size_t decode_step_smem_bytes(int d, int * xs_bytes, int * slot_bytes, int * chunk_keys);
// Number of CTAs to launch on the current device (one per SM), 0 if the kernel cannot be co-resident.
int decode_step_grid(size_t smem_bytes);
// Fills `out` (capacity kStepMaxPhases) for a step of n rows; returns the number of phases.
int decode_step_plan(const StepLayerW * layers_host, int n_layer, int d, int n_head, int n_vocab, const __half * te, const float * ln_g,
                     const float * ln_b, const __half * attn16, const __half * h16, int n, int grid, int slot_bytes, StepPhase * out);
bool launch_decode_step(const StepArgs & a, int grid, size_t smem_bytes, cudaStream_t st);
// gemm_tc.cu
bool make_tensor_map_2d_f16(void * out_map, const void * base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes,
                            uint32_t box_cols, uint32_t box_rows);

}  // namespace wb200
