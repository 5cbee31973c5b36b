// The device boundary inside the library: everything numeric (conv stem, encoder, cross-KV, decoder) sits behind
// this interface.  The shipped library has exactly one implementation — CudaForward (cuda/forward_cuda.cu, sm_100a
// kernels).  There is no CPU implementation in the product; the host-logic tests under tests/ provide their own
// checker-backed implementation in a separate test-only library.
#pragma once

#include "decode_host.h"
#include "model.h"
#include "run_state.h"

#include <cstdint>
#include <vector>

namespace wb200 {

// What whisper_process_logits (whisper.cpp:4493-4775) needs to know about the decoder a row belongs to; lets the
// device apply the suppression / timestamp rules, the log-softmax and the greedy pick itself, so that a greedy step
// returns 24 bytes per sequence instead of n_vocab logits.
struct SampleRule {
    enum { INITIAL_BLANK = 1, NO_TIMESTAMPS = 2, SUPPRESS_SOLM = 4, NON_SPEECH = 8, LAST_TS = 16, PENULT_TS = 32,
           INITIAL_MAX_TS = 64, HAS_TS = 128 };
    int32_t flags = 0;
    int32_t tid0_initial = 0;    // round(max_initial_ts / precision)      (whisper.cpp:4618-4625)
    int32_t tid0_seek = 0;       // seek_delta / 2                         (whisper.cpp:4629-4635)
    int32_t reserved = 0;
};

struct DecodeInput {
    int n_tokens = 0;
    const int32_t * token = nullptr;
    const int32_t * pos   = nullptr;
    const int32_t * seq   = nullptr;
    const int8_t  * want_logits = nullptr;
    int kv_head = 0;                 // first cell written by this batch (find_slot result)
    int n_kv    = 0;                 // cells visible to attention (cell_max)
    const KvCell * cells = nullptr;  // the cell table (size >= n_kv) AFTER find_slot, for the visibility mask
    // non-null => every row flagged in want_logits is post-processed and greedily sampled ON THE DEVICE with rule
    // sample[row]; the job then receives whisper_token_data in sampled_out[row] instead of logits
    const SampleRule * sample = nullptr;
    // ... and rows with n_draws[row] > 0 are sampled from their distribution instead (Forward::can_sample_dist): logits / temperature,
    // the rules of sample[row], then one token per uniform variate of `draws` (all rows concatenated, the host's generators in the
    // reference's order).  The job receives one whisper_token_data per draw in dist_out, rows in order.
    const int32_t * n_draws = nullptr;
    const double * draws = nullptr;
    float temperature = 0.0f;
    int   tid_default = 0;           // token_data.tid when no timestamp has probability mass: 0 (whisper_sample_token), token_beg (_topk)
};

enum StageId {
    STAGE_MEL_WINDOW = 0, STAGE_EMBD_CONV = 1, STAGE_EMBD_ENC = 2, STAGE_CROSS_K = 3, STAGE_CROSS_V = 4,
    STAGE_SELF_K = 5, STAGE_SELF_V = 6, STAGE_HOST_MEL = 7,
};

// One audio chunk of an encoder batch / one decode request of a decoder batch.  `slot` selects the per-chunk device
// state (cross-attention K/V + self-attention cache) the job reads and writes.
struct EncodeJob {
    const float * mel_window = nullptr;   // host f32 [n_mels][2*n_ctx] (spectrogram computed on the host), or ...
    // ... mel_offset >= 0: the window starts at this frame of the SLOT's device-resident spectrogram.  With pcm set (a buffer from
    // pcm_stage_acquire holding the clip) that spectrogram is computed first; with pcm == nullptr the one of an earlier job is reused.
    const float * pcm = nullptr;
    int n_samples = 0;
    int mel_offset = -1;
    float * energy_out = nullptr;         // with pcm: the clip's energy envelope (whisper.cpp:6350-6366) is copied here (pinned host memory, n_samples floats)
    int slot = 0;
};
struct DecodeJob {
    DecodeInput in;
    int     slot = 0;
    float * logits_out = nullptr;         // host [n_tokens][n_vocab]; only rows flagged in want_logits are written
    whisper_token_data * sampled_out = nullptr;   // host [n_tokens]; written instead of logits when in.sample != nullptr
    whisper_token_data * dist_out = nullptr;      // host [sum of in.n_draws]; the tokens drawn for the rows with n_draws > 0
};

class Forward {
    bool sync_ok_ = false, enc_sync_ok_ = false;
public:
    virtual ~Forward() {}

    // Number of independent per-chunk device states; ensure_slots may grow it (only while no call is in flight).
    virtual int  n_slots() const { return 1; }
    virtual bool ensure_slots(int n) { return n <= 1; }

    // Batched forms: all jobs share n_ctx / n_audio_ctx and run as ONE set of kernel launches.  The defaults serve
    // single-state implementations (slot 0, one job at a time).
    virtual bool encode_batch(const EncodeJob * jobs, int n_jobs, int n_ctx) {
        for (int i = 0; i < n_jobs; ++i) {
            if (jobs[i].slot != 0 || !encode(jobs[i].mel_window, n_ctx)) return false;
        }
        return true;
    }
    virtual bool decode_batch(const DecodeJob * jobs, int n_jobs, int n_audio_ctx) {
        for (int i = 0; i < n_jobs; ++i) {
            if (jobs[i].slot != 0 || !decode(jobs[i].in, n_audio_ctx, jobs[i].logits_out)) return false;
        }
        return true;
    }
    // Log-mel spectrogram on the device: true if encode_batch takes PCM jobs.  pcm_stage_acquire hands out a pinned
    // staging buffer for a clip of n_samples (blocking while all are in use; nullptr if the clip is longer than the device path
    // takes), to be filled by the caller — any thread — and given back with pcm_stage_release once its encoder pass has returned.
    virtual bool mel_on_device() const { return false; }
    virtual int  pcm_stage_samples() const { return 0; }     // longest clip the device path takes
    virtual float * pcm_stage_acquire(int /*n_samples*/) { return nullptr; }
    virtual void pcm_stage_release(float * /*buf*/) {}
    virtual bool is_pinned_host(const void * /*p*/) const { return false; }   // page-locked host memory: uploaded from where it lies, no staging copy
    virtual float * energy_buffer(int /*slot*/) { return nullptr; }          // pinned per-slot destination of the energy envelope (pcm_stage_samples() floats)

    // Pipelined encoder passes: encode_sets() passes may be queued at once (each on its own set of staging buffers: the inputs of the
    // second are copied while the first computes); encode_collect waits for one.  The defaults run the pass inside encode_enqueue.
    virtual int  encode_sets() const { return 1; }
    virtual bool encode_enqueue(const EncodeJob * jobs, int n_jobs, int n_ctx, int set) {
        if (set != 0) return false;
        enc_sync_ok_ = encode_batch(jobs, n_jobs, n_ctx);
        return true;
    }
    virtual bool encode_collect(int set) { return set == 0 && enc_sync_ok_; }

    // Pipelined decoder passes: decode_sets() passes may be queued at once, each on its own staging set; decode_collect waits for
    // one and delivers its results.  The defaults run the pass synchronously inside decode_enqueue.
    virtual bool encoder_concurrent() const { return false; }   // encode_batch may run on its own host thread next to decoder passes
    virtual int  decode_sets() const { return 1; }
    virtual int  decode_rows_per_pass() const { return 16; }      // rows one decoder pass should carry (batching hint)
    virtual bool decode_enqueue(const DecodeJob * jobs, int n_jobs, int n_audio_ctx, int set) {
        if (set != 0) return false;
        sync_ok_ = decode_batch(jobs, n_jobs, n_audio_ctx);
        return true;
    }
    virtual bool decode_collect(int set) { return set == 0 && sync_ok_; }
    virtual long long read_stage_slot(int slot, int what, void * dst, long long cap_bytes) {
        return slot == 0 ? read_stage(what, dst, cap_bytes) : -1;
    }

    // conv stem + encoder blocks + ln_post + cross-attention K/V for one mel window.
    // mel_window: host f32 [n_mels][2*n_ctx] (already zero padded).  (whisper.cpp:2086-2146)
    virtual bool encode(const float * mel_window, int n_ctx) = 0;

    // One decoder pass over `in` with n_audio_ctx cross-attention keys; writes the rows flagged in want_logits
    // to logits_out[row * n_vocab ...] (host).  (whisper.cpp:2517-2595)
    virtual bool decode(const DecodeInput & in, int n_audio_ctx, float * logits_out) = 0;

    // Copies a stage tensor to the host (see whisper_b200_read_stage in include/whisper_b200.h).
    virtual long long read_stage(int what, void * dst, long long cap_bytes) = 0;

    // true if decode() can run the logits rules + greedy pick on the device (DecodeInput::sample)
    virtual bool can_sample() const { return false; }
    // true if decode() can draw from the distribution on the device (DecodeInput::n_draws / draws)
    virtual bool can_sample_dist() const { return false; }

    // ---- device-resident greedy runs (run_state.h) --------------------------------------------------------------------
    // A run is one greedy t = 0 sequence whose token loop stays on the device: run_start stores its RunSeq in the slot, every
    // run_step_enqueue queues ONE decoder step for a list of slots (the step reads each sequence's token / position / rule from its
    // RunSeq, samples, and advances the RunSeq — sequences that are not RUN_LIVE idle through it), run_step_wait returns the status
    // of every listed sequence after that step, run_fetch the final RunSeq and the sampled tokens.  Steps may be queued ahead of the
    // waits (run_depth() of them): the host learns about finished sequences a few steps late and never sits between two steps.
    virtual bool supports_runs() const { return false; }
    virtual int  run_rows_max() const { return 0; }          // sequences one step can carry
    virtual int  run_depth() const { return 1; }             // steps that may be queued before the oldest is waited for
    virtual bool run_start(int /*slot*/, const RunSeq & /*init*/) { return false; }
    virtual int  run_step_enqueue(const int * /*slots*/, int /*n*/, int /*n_audio_ctx*/) { return -1; }      // -> ticket (>= 0) or -1
    virtual bool run_step_wait(int /*ticket*/, int32_t * /*status*/) { return false; }
    virtual bool run_fetch(int /*slot*/, int /*ticket*/, RunSeq & /*out*/, std::vector<whisper_token_data> & /*tokens*/) { return false; }

    virtual int64_t kernel_launches() const = 0;

    // Device-side clocks (CUDA events on the launching stream).  out[0..5] = ms spent in encode calls, ms spent in
    // decode calls, number of encode calls, number of decode calls, host->device bytes, device->host bytes, launches of the
    // persistent decode-step kernel and their algorithmic bytes — accumulated since the context was created.
    virtual double busy_ms() const { return 0.0; }   // union of the device-busy intervals of all passes so far
    virtual double mel_ms() const { return 0.0; }    // ms inside the spectrogram stage of encoder passes (not part of gpu_times()[0])
    virtual void gpu_times(double * out8) const { for (int i = 0; i < 8; ++i) out8[i] = 0.0; }
    // Per-kernel-class profile (event pair around every launch while enabled; adds launch overhead, so bench.py turns
    // it on only for its profiled pass).  out[kind][0..3] = launches, total ms, algorithmic FLOP, algorithmic bytes.
    enum { PROF_GEMM_ENC = 0, PROF_GEMM_ATTN = 1, PROF_SOFTMAX = 2, PROF_LAYERNORM = 3, PROF_SKINNY = 4, PROF_DEC_ATTN = 5,
           PROF_MISC = 6, PROF_GEMM_DEC = 7, PROF_STEP = 8, PROF_KINDS = 9 };
    virtual void set_profiling(bool /*on*/) {}
    virtual void profile(double * out /*[PROF_KINDS][4]*/) const { for (int i = 0; i < PROF_KINDS * 4; ++i) out[i] = 0.0; }
    virtual void set_gemm_engine(int /*engine*/) {}
    virtual const char * name() const = 0;
};

// Factory, defined exactly once per link: in the product by cuda/forward_cuda.cu (the sm_100a implementation; returns
// nullptr after logging the reason when no usable device / kernel image exists — there is no fallback), and in the
// test-only host-logic library by tests/hostlogic/forward_checker.cpp.
Forward * create_forward(const ModelFile & model, int kv_self_cells, int device);

// Page-locked host memory for callers that want their PCM uploaded from where it lies (whisper_b200_host_alloc); nullptr on failure.
void * host_alloc_pinned(size_t bytes);
void   host_free_pinned(void * p);

}  // namespace wb200
