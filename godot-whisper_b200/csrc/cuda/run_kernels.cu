// Device side of the greedy runs (run_state.h): the two small kernels that bracket every decoder step of a run.
//
//   k_run_prep     fills the step's staging block — token, position, sampler rule, cache-cell maps, per-slot offsets, visibility
//                  mask, live key count — from the RunSeq of every listed sequence.  It writes exactly what CudaForward::decode_enqueue
//                  copies from the host for an ordinary pass, so the decoder kernels (k_decode_step, or the wide-pass graph of
//                  tcgen05 GEMMs + k_decode_attention + k_sample_greedy) do not know the difference.
//   k_run_advance  appends the token the step sampled to the sequence's token list and applies run_advance(): window slide,
//                  completion rules, repetition guard (whisper.cpp:5425-5507).
//
// Together they replace the host round trip per token of whisper_full_with_state's loop (whisper.cpp:5288-5606: sample on the
// host, build the next batch, call whisper_decode).
#include "run_kernels.cuh"

namespace wb200 {

namespace {

__global__ void __launch_bounds__(128)
k_run_prep(const RunPrepArgs a) {
    const int r = blockIdx.x;                       // one block per row: the mask row is written by the whole block
    __shared__ int s_len;
    if (threadIdx.x == 0) {
        const int slot = a.row_slot[r];
        bool live = false;
        RunSeq s;
        if (slot >= 0) { s = a.seqs[slot]; live = s.status == RUN_LIVE; }
        int32_t rule[4] = {0, 0, 0, 0};
        int token = 0, pos = 0, len = 1;
        int64_t self_slot = a.n_slots, cross_slot = 0;      // idle rows: K / V of the step go to the scratch slot behind the last real one
        int cell = r;
        if (live) {
            run_rule(s, a.token_beg, rule);
            token = s.token; pos = s.pos; len = s.pos + 1;
            self_slot = slot; cross_slot = slot; cell = s.pos;
        } else if (slot >= 0) {
            cross_slot = slot;
        }
        a.token[r] = token; a.pos[r] = pos; a.want[r] = r; a.wslot[r] = r;
        a.rule[4 * r + 0] = rule[0]; a.rule[4 * r + 1] = rule[1]; a.rule[4 * r + 2] = rule[2]; a.rule[4 * r + 3] = rule[3];
        a.rowmap_k[r] = (int) (self_slot * (int64_t) a.n_layer * a.kv_cells + cell);
        a.rowmap_v[r] = (int) (self_slot * a.self_v_slot + cell);
        a.koff_self[r]  = self_slot * a.self_k_slot;   a.voff_self[r]  = self_slot * a.self_v_slot;
        a.koff_cross[r] = cross_slot * a.cross_k_slot; a.voff_cross[r] = cross_slot * a.cross_v_slot;
        if (live) atomicMax(a.n_kv, len);
        else      atomicMax(a.n_kv, 1);
        s_len = len;                                    // (an idle row attends to cell 0 of the scratch slot: zeros or what idle row 0 wrote — finite)
    }
    __syncthreads();
    const int len = s_len;
    float * m = a.mask + (size_t) r * a.ld_mask;
    for (int c = threadIdx.x; c < a.ld_mask; c += blockDim.x) m[c] = c < len ? 0.0f : -INFINITY;
}

__global__ void __launch_bounds__(128)
k_run_advance(RunSeq * __restrict__ seqs, const int * __restrict__ row_slot, int n, const float * __restrict__ sampled,
              float * __restrict__ tokens_out, int * __restrict__ status_out, int token_beg, int token_eot) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int slot = row_slot[r];
    if (slot < 0) { status_out[r] = RUN_COMPLETED; return; }
    RunSeq s = seqs[slot];
    if (s.status == RUN_LIVE) {
        const float * o = sampled + 6 * (size_t) r;
        if (s.n_out < kRunTokenCap) {
            float * t = tokens_out + ((size_t) slot * kRunTokenCap + s.n_out) * 6;
#pragma unroll
            for (int k = 0; k < 6; ++k) t[k] = o[k];
        }
        run_advance(s, __float_as_int(o[0]), token_beg, token_eot);
        seqs[slot] = s;
    }
    status_out[r] = s.status;
}

}  // namespace

void launch_run_prep(const RunPrepArgs & a, int n_rows, cudaStream_t st) {
    k_run_prep<<<n_rows, 128, 0, st>>>(a);
}

void launch_run_advance(RunSeq * seqs, const int * row_slot, int n_rows, const float * sampled, float * tokens_out, int * status_out,
                        int token_beg, int token_eot, cudaStream_t st) {
    k_run_advance<<<(n_rows + 127) / 128, 128, 0, st>>>(seqs, row_slot, n_rows, sampled, tokens_out, status_out, token_beg, token_eot);
}

}  // namespace wb200
