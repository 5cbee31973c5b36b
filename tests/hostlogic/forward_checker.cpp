// TEST-ONLY implementation of wb200::Forward.  NOT part of libwhisper_b200.so.
//
// tests/conftest.py links this file with the product's HOST sources (model/mel/decode_host/full/api .cpp) into
// tests/_build/libwhisper_hostlogic.so so that the C++ driver — KV-cell bookkeeping, prompt construction, temperature
// fallback, beam management, logits rules, samplers, segment assembly, token timestamps — can be checked on a CPU-only
// box: the tensor math is delegated to the compiled reference (oracle/_ref, dlopen'ed here), so any difference in the
// resulting token stream is a host-logic bug.  The GPU tests (-m gpu) exercise the real CudaForward instead.
#include "../../godot-whisper_b200/csrc/common.h"
#include "../../godot-whisper_b200/csrc/forward.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <vector>

namespace wb200 {

namespace {

struct RefApi {
    void * h = nullptr;
    void * (*init)(void *, size_t, whisper_context_params) = nullptr;
    void   (*free_)(void *) = nullptr;
    int    (*set_mel)(void *, const float *, int, int) = nullptr;
    int    (*encode)(void *, int, int) = nullptr;
    float *(*get_logits)(void *) = nullptr;
    void   (*log_set)(ggml_log_callback, void *) = nullptr;
    void   (*set_audio_ctx)(void *, int) = nullptr;
    void   (*set_cells)(void *, const int *, const unsigned *, int, int) = nullptr;
    int    (*decode_batch)(void *, const int *, const int *, const int *, const signed char *, int, int) = nullptr;
    int    (*embd_enc)(void *, float *, int, int *, int *) = nullptr;
};

void quiet_log(ggml_log_level, const char *, void *) {}

class CheckerForward : public Forward {
public:
    RefApi api;
    void * rctx = nullptr;               // reference context of the slot being served
    std::vector<void *> slot_ctx;        // one reference context per device slot
    std::vector<uint8_t> model_copy;
    int n_vocab = 0, kv_cells = 0, n_audio_ctx_model = 0, n_threads = 4;
    std::atomic<int64_t> calls{0};
    // WHISPER_HOSTLOGIC_CONCURRENT_ENC=1: behave like the CUDA forward with its encoder stream — encode_batch may be called from the
    // batcher's encoder thread while the decoder driver runs decode passes (every job only touches its own slot's reference context)
    bool concurrent_enc = getenv("WHISPER_HOSTLOGIC_CONCURRENT_ENC") != nullptr;
    bool encoder_concurrent() const override { return concurrent_enc; }

    ~CheckerForward() override {
        for (void * c : slot_ctx) if (c) api.free_(c);
        if (api.h) dlclose(api.h);
    }

    int n_slots() const override { return (int) slot_ctx.size(); }
    bool ensure_slots(int n) override {
        while ((int) slot_ctx.size() < n) {
            whisper_context_params cp = { false };
            void * c = api.init(model_copy.data(), model_copy.size(), cp);
            if (!c) return false;
            slot_ctx.push_back(c);
        }
        return true;
    }
    // "batched" passes of the checker: the jobs run one after the other, each against its slot's reference context
    bool encode_batch(const EncodeJob * jobs, int n_jobs, int n_ctx) override {
        for (int i = 0; i < n_jobs; ++i) {
            if (jobs[i].slot < 0 || jobs[i].slot >= (int) slot_ctx.size()) return false;
            if (!encode_on(slot_ctx[jobs[i].slot], jobs[i].mel_window, n_ctx)) return false;
        }
        return true;
    }
    bool decode_batch(const DecodeJob * jobs, int n_jobs, int n_audio_ctx) override {
        for (int i = 0; i < n_jobs; ++i) {
            if (jobs[i].slot < 0 || jobs[i].slot >= (int) slot_ctx.size()) return false;
            const DecodeInput & in = jobs[i].in;
            if (!in.n_draws) {
                if (!decode_on(slot_ctx[jobs[i].slot], in, n_audio_ctx, jobs[i].logits_out)) return false;
                continue;
            }
            // rows sampled from their distribution "on the device": the reference's logits, the product's host rules with the row's
            // temperature, then libstdc++'s std::discrete_distribution inverted for the uniform variates the caller drew
            std::vector<float> rows((size_t) in.n_tokens * n_vocab);
            if (!decode_on(slot_ctx[jobs[i].slot], in, n_audio_ctx, rows.data())) return false;
            int at = 0, at_u = 0;
            for (int r = 0; r < in.n_tokens; ++r) {
                if (!in.want_logits[r] || in.n_draws[r] <= 0) continue;
                int32_t rule[4] = { in.sample[r].flags, in.sample[r].tid0_initial, in.sample[r].tid0_seek, 0 };
                Decoder d;
                rules_to_decoder(rule, in.temperature, rows.data() + (size_t) r * n_vocab, d);
                whisper_token_data stats = sample_token(vocab, d, true);            // (for tid / pt / ptsum: whisper.cpp:4789-4803 = :4851-4866)
                std::vector<double> prob(d.probs.begin(), d.probs.begin() + n_vocab), cp;
                double sum = 0.0;
                for (double p : prob) sum += p;
                for (double & p : prob) p /= sum;
                cp.reserve(prob.size());
                double acc = 0.0;
                for (double p : prob) { acc += p; cp.push_back(acc); }
                cp.back() = 1.0;
                for (int k = 0; k < in.n_draws[r]; ++k) {
                    const double u = in.draws[at_u++];
                    const int id = (int) (std::lower_bound(cp.begin(), cp.end(), u) - cp.begin());
                    whisper_token_data td = { id, stats.ptsum > 0.0f ? stats.tid : in.tid_default, d.probs[id], d.logprobs[id], stats.pt, stats.ptsum, -1, -1, 0.0f };
                    if (id >= vocab.token_beg) { td.tid = id; td.pt = td.p; }
                    if (jobs[i].dist_out) jobs[i].dist_out[at++] = td;
                }
            }
        }
        return true;
    }
    bool can_sample_dist() const override { return getenv("WHISPER_HOSTLOGIC_DIST") == nullptr || atoi(getenv("WHISPER_HOSTLOGIC_DIST")) != 0; }
    // fills decoder d (probs / logprobs) the way whisper_process_logits would for a decoder whose state produced `rule`
    void rules_to_decoder(const int32_t * rule, float temperature, const float * logits, Decoder & d) {
        whisper_full_params p;
        memset(&p, 0, sizeof(p));
        const int flags = rule[0];
        const bool initial = (flags & (SampleRule::INITIAL_BLANK | SampleRule::INITIAL_MAX_TS)) != 0;
        p.suppress_blank = (flags & SampleRule::INITIAL_BLANK) != 0;
        p.no_timestamps = (flags & SampleRule::NO_TIMESTAMPS) != 0;
        p.tdrz_enable = (flags & SampleRule::SUPPRESS_SOLM) == 0;
        p.suppress_non_speech_tokens = (flags & SampleRule::NON_SPEECH) != 0;
        p.max_initial_ts = (flags & SampleRule::INITIAL_MAX_TS) ? rule[1] * (float(WHISPER_CHUNK_SIZE) / n_audio_ctx_model) : 0.0f;
        if (!initial) {
            whisper_token_data t = { 0, 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, 0.0f };
            t.id = (flags & SampleRule::PENULT_TS) ? vocab.token_beg : 0; d.sequence.tokens.push_back(t);
            t.id = (flags & SampleRule::LAST_TS) ? vocab.token_beg : 0;   d.sequence.tokens.push_back(t);
        }
        d.has_ts = (flags & SampleRule::HAS_TS) != 0;
        d.seek_delta = 2 * rule[2];
        process_logits(vocab, lrules, n_audio_ctx_model, p, nullptr, nullptr, logits, d, temperature);
    }

    bool encode(const float * mel_window, int n_ctx) override { return encode_on(rctx, mel_window, n_ctx); }
    bool decode(const DecodeInput & in, int n_audio_ctx, float * logits_out) override { return decode_on(rctx, in, n_audio_ctx, logits_out); }

    bool encode_on(void * rctx, const float * mel_window, int n_ctx) {
        ++calls;
        const int n_mels = 80;
        if (api.set_mel(rctx, mel_window, 2 * n_ctx, n_mels) != 0) return false;
        api.set_audio_ctx(rctx, n_ctx == n_audio_ctx_model ? 0 : n_ctx);
        return api.encode(rctx, 0, n_threads) == 0;
    }

    bool decode_on(void * rctx, const DecodeInput & in, int n_audio_ctx, float * logits_out) {
        ++calls;
        api.set_audio_ctx(rctx, n_audio_ctx == n_audio_ctx_model ? 0 : n_audio_ctx);
        // replay the caller's cell table as it was BEFORE its find_slot, with the search head on the chosen slot
        std::vector<int> pos(kv_cells);
        std::vector<unsigned> mask(kv_cells);
        for (int i = 0; i < kv_cells; ++i) { pos[i] = in.cells[i].pos; mask[i] = in.cells[i].seq_mask; }
        for (int i = 0; i < in.n_tokens; ++i) { pos[in.kv_head + i] = -1; mask[in.kv_head + i] = 0; }
        api.set_cells(rctx, pos.data(), mask.data(), kv_cells, in.kv_head);
        if (api.decode_batch(rctx, in.token, in.pos, in.seq, (const signed char *) in.want_logits, in.n_tokens, n_threads) != 0) {
            return false;
        }
        const float * rows = api.get_logits(rctx);
        for (int i = 0; i < in.n_tokens; ++i) {
            if (in.want_logits[i]) memcpy(logits_out + (size_t) i * n_vocab, rows + (size_t) i * n_vocab, sizeof(float) * n_vocab);
        }
        return true;
    }

    // ---- greedy runs on the CPU: the reference's logits, the product's host rules for the pick (process_logits / sample_token, proven
    // against the reference by the per-token tests), and the SAME run_rule / run_advance the device executes (run_state.h) ----
    Vocab vocab; LogitsRules lrules;
    std::vector<RunSeq> run_seq;
    std::vector<std::vector<whisper_token_data>> run_tok;
    std::vector<std::vector<int32_t>> run_status;       // by ticket
    bool runs_on = getenv("WHISPER_HOSTLOGIC_RUNS") == nullptr || atoi(getenv("WHISPER_HOSTLOGIC_RUNS")) != 0;
    bool supports_runs() const override { return runs_on; }
    int  run_rows_max() const override { return 3; }    // small on purpose: more runs than rows must queue
    int  run_depth() const override { return 2; }
    bool run_start(int slot, const RunSeq & init) override {
        if (slot < 0 || slot >= (int) slot_ctx.size()) return false;
        if ((int) run_seq.size() < (int) slot_ctx.size()) { run_seq.resize(slot_ctx.size()); run_tok.resize(slot_ctx.size()); }
        run_seq[slot] = init; run_tok[slot].clear();
        return true;
    }
    whisper_token_data pick_by_rule(const float * logits, const int32_t * rule) {
        Decoder d;
        rules_to_decoder(rule, 0.0f, logits, d);
        return sample_token(vocab, d, true);
    }
    int run_step_enqueue(const int * slots, int n, int n_audio_ctx) override {
        std::vector<int32_t> status(n);
        std::vector<float> logits((size_t) n_vocab);
        for (int r = 0; r < n; ++r) {
            const int slot = slots[r];
            if (slot < 0 || slot >= (int) run_seq.size()) return -1;
            RunSeq & s = run_seq[slot];
            if (s.status == RUN_LIVE) {
                int32_t rule[4];
                run_rule(s, vocab.token_beg, rule);
                std::vector<KvCell> cells(kv_cells);
                for (int c = 0; c <= s.pos; ++c) { cells[c].pos = c; cells[c].seq_mask = 1u; }
                const int32_t tok = s.token, pos = s.pos, seq = 0;
                const int8_t want = 1;
                DecodeInput in;
                in.n_tokens = 1; in.token = &tok; in.pos = &pos; in.seq = &seq; in.want_logits = &want;
                in.kv_head = s.pos; in.n_kv = s.pos + 1; in.cells = cells.data();
                if (!decode_on(slot_ctx[slot], in, n_audio_ctx, logits.data())) return -1;
                const whisper_token_data td = pick_by_rule(logits.data(), rule);
                run_tok[slot].push_back(td);
                run_advance(s, td.id, vocab.token_beg, vocab.token_eot);
            }
            status[r] = s.status;
        }
        run_status.push_back(status);
        return (int) run_status.size() - 1;
    }
    bool run_step_wait(int ticket, int32_t * status) override {
        if (ticket < 0 || ticket >= (int) run_status.size()) return false;
        memcpy(status, run_status[ticket].data(), run_status[ticket].size() * sizeof(int32_t));
        return true;
    }
    bool run_fetch(int slot, int, RunSeq & out, std::vector<whisper_token_data> & tokens) override {
        if (slot < 0 || slot >= (int) run_seq.size()) return false;
        out = run_seq[slot]; tokens = run_tok[slot];
        return true;
    }

    long long read_stage(int, void *, long long) override { return -1; }
    int64_t kernel_launches() const override { return 0; }
    const char * name() const override { return "TEST-ONLY checker forward (compiled reference)"; }
};

}  // namespace

void * host_alloc_pinned(size_t bytes) { return malloc(bytes); }
void   host_free_pinned(void * p) { free(p); }

Forward * create_forward(const ModelFile & model, int kv_self_cells, int /*device*/) {
    const char * path = getenv("WHISPER_HOSTLOGIC_REF_LIB");
    if (!path) {
        WB_LOG_ERROR("%s: WHISPER_HOSTLOGIC_REF_LIB not set (test-only library)\n", __func__);
        return nullptr;
    }
    auto * f = new CheckerForward;
    f->api.h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!f->api.h) {
        WB_LOG_ERROR("%s: dlopen(%s) failed: %s\n", __func__, path, dlerror());
        delete f;
        return nullptr;
    }
#define LOAD(field, sym) *(void **) (&f->api.field) = dlsym(f->api.h, sym); if (!f->api.field) { WB_LOG_ERROR("missing %s\n", sym); delete f; return nullptr; }
    LOAD(init, "whisper_init_from_buffer_with_params")
    LOAD(free_, "whisper_free")
    LOAD(set_mel, "whisper_set_mel")
    LOAD(encode, "whisper_encode")
    LOAD(get_logits, "whisper_get_logits")
    LOAD(log_set, "whisper_log_set")
    LOAD(set_audio_ctx, "probe_set_audio_ctx")
    LOAD(set_cells, "probe_kv_self_set_cells")
    LOAD(decode_batch, "probe_decode_batch")
    LOAD(embd_enc, "probe_embd_enc")
#undef LOAD
    f->api.log_set(quiet_log, nullptr);
    f->model_copy.assign((const uint8_t *) model.raw, (const uint8_t *) model.raw + model.raw_size);
    if (!f->ensure_slots(1)) { delete f; return nullptr; }
    f->rctx = f->slot_ctx[0];
    f->n_vocab = model.hparams.n_vocab;
    f->vocab = model.vocab;
    f->lrules.build(f->vocab);
    f->kv_cells = kv_self_cells;
    f->n_audio_ctx_model = model.hparams.n_audio_ctx;
    if (const char * t = getenv("WHISPER_HOSTLOGIC_THREADS")) f->n_threads = atoi(t);
    return f;
}

}  // namespace wb200
