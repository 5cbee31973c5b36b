// The greedy token loop as a per-sequence state machine that runs where the logits are: on the device.
//
// whisper_full_with_state's token loop (/root/reference/thirdparty/whisper.cpp/whisper.cpp:5288-5606) alternates one decoder pass
// with a few integer decisions per sequence: push the sampled token, slide the timestamp window (:5436-5450), test the completion
// rules (:5467-5490), guard against repetition loops (:5501-5506), derive the next step's logits rules from the tokens so far
// (whisper_process_logits, :4527-4635).  For greedy decoding at temperature 0 nothing in there needs the host: RunSeq holds the
// state, run_rule() derives the sampler rule of the next step, run_advance() applies one sampled token.  Both functions compile
// for the device (cuda/run_kernels.cu: one thread per sequence after every step) and for the host (the test-only checker forward
// replays them on the CPU, tests/hostlogic/forward_checker.cpp), so the state machine the GPU runs is the one the CPU tests pin
// against the reference's whisper_full.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define WB_HD __host__ __device__ __forceinline__
#else
#define WB_HD inline
#endif

namespace wb200 {

enum { RUN_LIVE = 0, RUN_COMPLETED = 1, RUN_FAILED = 2, RUN_EXHAUSTED = 3 };

constexpr int kRunTokenCap = 256;        // sampled tokens kept per run (the loop of whisper.cpp:5288 stops at n_text_ctx / 2 - 4 = 220)

struct RunSeq {
    // constants of the run (set by the host from whisper_full_params and the seek position)
    int32_t max_tokens     = 0;          // params.max_tokens
    int32_t seek           = 0;
    int32_t seek_end       = 0;
    int32_t single_segment = 0;
    int32_t n_max          = 0;          // n_text_ctx / 2 - 4
    int32_t rule_static    = 0;          // SampleRule flags that hold for every step: NO_TIMESTAMPS | SUPPRESS_SOLM | NON_SPEECH
    int32_t rule_initial   = 0;          // SampleRule flags of the first sampled token only: INITIAL_BLANK | INITIAL_MAX_TS
    int32_t tid0_initial   = 0;
    // state
    int32_t token      = 0;              // input of the next decoder step: the last prompt token, then the last sampled token
    int32_t pos        = 0;              // its position, which is also its self-attention cache cell (one sequence per slot, cells 0..pos-1 hold the past)
    int32_t i          = 0;              // index of the token the next step samples (the loop variable of whisper.cpp:5288)
    int32_t last_id    = -1;             // the two most recently sampled tokens (-1: none yet)
    int32_t penult_id  = -1;
    int32_t has_ts     = 0;
    int32_t seek_delta = 0;
    int32_t result_len = 0;
    int32_t status     = RUN_LIVE;
    int32_t n_out      = 0;              // tokens sampled so far
    int32_t reserved[2] = {0, 0};
};
static_assert(sizeof(RunSeq) == 80, "RunSeq is copied between host and device as raw bytes");

// SampleRule bits (forward.h) restated here so that this header stands alone on the device side
enum { RULE_INITIAL_BLANK = 1, RULE_NO_TIMESTAMPS = 2, RULE_SUPPRESS_SOLM = 4, RULE_NON_SPEECH = 8, RULE_LAST_TS = 16, RULE_PENULT_TS = 32,
       RULE_INITIAL_MAX_TS = 64, RULE_HAS_TS = 128 };

// The logits rule of the step that samples token s.i (make_sample_rule of decode_host.cpp, with the decoder state read from s).
WB_HD void run_rule(const RunSeq & s, int token_beg, int32_t * rule4) {
    int flags = s.rule_static;
    if (s.i == 0) flags |= s.rule_initial;                                        // whisper.cpp:4532-4537, 4618-4625
    if (s.last_id >= token_beg) flags |= RULE_LAST_TS;                            // whisper.cpp:4598-4614 (last_id = -1 when nothing was sampled yet)
    if (s.penult_id < 0 || s.penult_id >= token_beg) flags |= RULE_PENULT_TS;     // "tokens_cur.size() < 2 || tokens_cur[size - 2].id >= token_beg"
    int tid0_seek = 0;
    if (s.has_ts) { flags |= RULE_HAS_TS; tid0_seek = s.seek_delta / 2; }         // whisper.cpp:4629-4635
    rule4[0] = flags; rule4[1] = s.tid0_initial; rule4[2] = tid0_seek; rule4[3] = 0;
}

// Applies the token just sampled (the caller has stored it as token number s.n_out of the run): whisper.cpp:5425-5507 for one greedy decoder.
WB_HD void run_advance(RunSeq & s, int id, int token_beg, int token_eot) {
    const int i = s.i;
    s.penult_id = s.last_id;
    s.last_id   = id;
    s.n_out    += 1;
    if (id > token_beg) {                                                         // timestamp token: slide the window (:5436-5450)
        const int seek_delta_new = 2 * (id - token_beg);
        if (s.has_ts && s.seek_delta > seek_delta_new && s.result_len < i) { s.status = RUN_FAILED; return; }   // going back in time
        s.seek_delta = seek_delta_new;
        s.result_len = i + 1;
        s.has_ts     = 1;
    }
    if (id == token_eot || (s.max_tokens > 0 && i >= s.max_tokens) || (s.has_ts && s.seek + s.seek_delta + 100 >= s.seek_end)) {   // :5467-5490
        if (s.result_len == 0) {
            if (s.seek + s.seek_delta + 100 >= s.seek_end) s.result_len = i + 1;
            else { s.status = RUN_FAILED; return; }
        }
        if (s.single_segment) { s.result_len = i + 1; s.seek_delta = 100 * 30; }
        s.status = RUN_COMPLETED;
        return;
    }
    if (i == s.n_max - 1 && (s.result_len == 0 || s.seek_delta < 100 * 30 / 2)) { s.status = RUN_FAILED; return; }   // repetition guard (:5501-5506)
    s.i     = i + 1;
    s.token = id;
    s.pos  += 1;
    if (s.i >= s.n_max) s.status = RUN_EXHAUSTED;                                 // the for loop of :5288 ends; neither completed nor failed
}

}  // namespace wb200
