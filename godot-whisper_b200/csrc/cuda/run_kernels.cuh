// Launchers of the run kernels (run_kernels.cu) — see run_state.h for the state machine they carry.
#pragma once

#include "../run_state.h"

#include <cuda_runtime.h>

namespace wb200 {

// Everything k_run_prep writes lives in the step's staging block (same arrays CudaForward::decode_enqueue fills from the host).
struct RunPrepArgs {
    const RunSeq * seqs = nullptr;        // [n_slots]
    const int *    row_slot = nullptr;    // [rows] device slot of every row, -1 = padding row
    int * token = nullptr, * pos = nullptr, * want = nullptr, * wslot = nullptr, * rule = nullptr, * rowmap_k = nullptr, * rowmap_v = nullptr;
    int64_t * koff_self = nullptr, * voff_self = nullptr, * koff_cross = nullptr, * voff_cross = nullptr;
    float * mask = nullptr; int ld_mask = 0;
    int *  n_kv = nullptr;                // live key count of the step (max over rows), zeroed before the launch
    int    n_slots = 0, n_layer = 0, kv_cells = 0, token_beg = 0;
    int64_t self_k_slot = 0, self_v_slot = 0, cross_k_slot = 0, cross_v_slot = 0;   // elements per slot
};

void launch_run_prep(const RunPrepArgs & a, int n_rows, cudaStream_t st);
void launch_run_advance(RunSeq * seqs, const int * row_slot, int n_rows, const float * sampled, float * tokens_out, int * status_out,
                        int token_beg, int token_eot, cudaStream_t st);

}  // namespace wb200
