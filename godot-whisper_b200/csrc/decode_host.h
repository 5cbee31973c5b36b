// Host-side decode bookkeeping: unified KV-cell table, per-decoder state, logits post-processing, samplers,
// sequence scoring and token-level timestamps.  Everything here is integer / f32 control logic restated from the
// reference so that, given the same logits, the token stream is identical.
#pragma once

#include "../../include/whisper_b200.h"
#include "model.h"

#include <cstdint>
#include <random>
#include <set>
#include <string>
#include <vector>

namespace wb200 {

constexpr int kMaxDecoders = 8;   // WHISPER_MAX_DECODERS (whisper.cpp:148)

// ---- unified self-attention KV cells (whisper.cpp:639-664, 938-1054) --------------------------------------------------

struct KvCell {
    int32_t pos = -1;
    uint32_t seq_mask = 0;   // bit s set <=> cell belongs to sequence s (s < 2*kMaxDecoders; the reference uses a std::set)
    bool has_seq(int s) const { return (seq_mask >> s) & 1u; }
};

struct KvCells {
    uint32_t head = 0;
    uint32_t size = 0;
    uint32_t n    = 0;       // cells visible to the next decode (highest used cell + 1)
    std::vector<KvCell> cells;

    void init(uint32_t n_cells) { size = n_cells; head = 0; n = 0; cells.assign(n_cells, KvCell()); }
    void clear();
    // Finds `n_tokens` consecutive free cells starting the search at `head`; marks them with pos[i] / seq[i].
    bool find_slot(int n_tokens, const int32_t * pos, const int32_t * seq);
    int32_t cell_max() const;
    void seq_rm(int seq, int32_t p0, int32_t p1);
    void seq_cp(int seq_src, int seq_dst, int32_t p0, int32_t p1);
};

// ---- decode batch (whisper.cpp:407-458) -------------------------------------------------------------------------------

struct SampleRule;

struct Batch {
    int n_tokens = 0;
    std::vector<int32_t> token, pos, seq;
    std::vector<int8_t>  logits;   // 1 => the caller wants this row's logits
    std::vector<int32_t> rule;     // 4 x int32 per row (SampleRule) — used when sample_on_device is set
    bool sample_on_device = false; // rows flagged in `logits` are sampled greedily on the device
    // rows flagged in `logits` with n_draws[row] > 0 are sampled from their distribution on the device: `draws` holds the uniform
    // variates of all such rows in row order (taken from the decoders' generators the way std::discrete_distribution would)
    bool dist_on_device = false;
    std::vector<int32_t> n_draws;
    std::vector<double>  draws;
    float temperature = 0.0f;
    int   tid_default = 0;
    void reserve(int n) { token.resize(n); pos.resize(n); seq.resize(n); logits.resize(n); rule.resize(4 * (size_t) n); n_draws.resize(n); }
    // whisper_batch_prep_legacy: one sequence, positions n_past.., logits for the last row only
    void prep_legacy(const int32_t * tokens, int n, int n_past, int seq_id);
};

// ---- per-decoder state (whisper.cpp:730-768) --------------------------------------------------------------------------

struct Sequence {
    std::vector<whisper_token_data> tokens;
    int    result_len       = 0;
    double sum_logprobs_all = 0.0;
    double sum_logprobs     = 0.0;
    double avg_logprobs     = 0.0;
    double entropy          = 0.0;
    double score            = 0.0;
};

struct Decoder {
    Sequence sequence;
    whisper_token_data pending{};      // token picked on the device for the next sampling step
    bool has_pending = false;
    std::vector<whisper_token_data> cands;   // tokens drawn on the device for the next sampling step (t > 0 / beam search)
    bool has_cands = false;
    int  i_batch    = 0;
    int  seek_delta = 0;
    bool failed = false, completed = false, has_ts = false;
    std::vector<float> probs, logits, logprobs;
    struct LogitId { double first; int32_t second; };
    std::vector<LogitId> logits_id;
    mutable std::mt19937 rng;
};

struct Segment {                 // whisper.cpp:396-405
    int64_t t0 = 0, t1 = 0;
    std::string text;
    std::vector<whisper_token_data> tokens;
    bool speaker_turn_next = false;
};

// ---- logits -> probabilities -> token (whisper.cpp:4493-4909) ---------------------------------------------------------

struct LogitsRules {
    // token ids to suppress when suppress_non_speech_tokens is set (resolved once per model, whisper.cpp:4576-4593)
    std::vector<int32_t> non_speech;
    int32_t blank = -1;           // id of " " (whisper.cpp:4535)
    void build(const Vocab & vocab);
};

// Applies temperature, suppression rules and timestamp constraints to `raw` (one row of n_vocab logits), then
// fills decoder.logits / logprobs / probs.  n_audio_ctx_model is hparams.n_audio_ctx (for max_initial_ts).
void process_logits(const Vocab & vocab, const LogitsRules & rules, int n_audio_ctx_model,
                    const whisper_full_params & params, struct whisper_context * ctx, struct whisper_state * state,
                    const float * raw, Decoder & decoder, float temperature);

// The decoder-state part of process_logits as a rule for the device-side sampler (same conditions, same order).
void make_sample_rule(const Vocab & vocab, int n_audio_ctx_model, const whisper_full_params & params, const Decoder & decoder,
                      int32_t * rule4);

whisper_token_data sample_token(const Vocab & vocab, const Decoder & decoder, bool best);
std::vector<whisper_token_data> sample_token_topk(const Vocab & vocab, Decoder & decoder, int k);
void sequence_score(const whisper_full_params & params, Sequence & sequence);

// ---- token-level timestamps (whisper.cpp:6315-6599) -------------------------------------------------------------------

struct TimestampState {
    int64_t t_beg = 0, t_last = 0;
    int32_t tid_last = 0;
    std::vector<float> energy;
    // the energy envelope is computed when the first segment asks for it (about a millisecond of host time per 30 s: not paid before
    // the chunk's encoder request is on its way); pending_pcm is borrowed from the running whisper_full call
    const float * pending_pcm = nullptr;
    int pending_n = 0;
    void ensure_energy();
    // ... or it was computed on the device and lies in pinned host memory owned by the forward pass (valid while the chunk owns its slot)
    const float * energy_ext = nullptr;
    int energy_ext_n = 0;
    const float * energy_data() const { return energy_ext ? energy_ext : energy.data(); }
    int energy_size() const { return energy_ext ? energy_ext_n : (int) energy.size(); }
};

void compute_token_level_timestamps(const Vocab & vocab, TimestampState & ts, Segment & segment,
                                    float thold_pt, float thold_ptsum);

// Splits the last segment at max_len characters (whisper.cpp:4428-4480). Returns the number of resulting segments.
int wrap_segment(const Vocab & vocab, std::vector<Segment> & result_all, int max_len, bool split_on_word);

}  // namespace wb200
