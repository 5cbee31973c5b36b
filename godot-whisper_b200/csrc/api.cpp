// extern "C" surface of libwhisper_b200.so — see include/whisper_b200.h for the contract and the reference
// interface (file:line) each entry point replaces.
#include "context.h"
#include "common.h"

#include <atomic>
#include <sys/resource.h>
#include <sched.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <mutex>
#include <string>
#include <thread>

// ---- ABI layout checks (the same numbers tests/test_abi.py reads from the compiled reference) -------
static_assert(sizeof(whisper_context_params) == 1, "whisper_context_params ABI");
static_assert(sizeof(whisper_token_data) == 48, "whisper_token_data ABI");
static_assert(offsetof(whisper_token_data, t0) == 24 && offsetof(whisper_token_data, vlen) == 40, "whisper_token_data ABI");
static_assert(sizeof(whisper_full_params) == 256, "whisper_full_params ABI");
static_assert(offsetof(whisper_full_params, n_threads) == 4, "ABI");
static_assert(offsetof(whisper_full_params, single_segment) == 23, "ABI");
static_assert(offsetof(whisper_full_params, token_timestamps) == 28, "ABI");
static_assert(offsetof(whisper_full_params, split_on_word) == 44, "ABI");
static_assert(offsetof(whisper_full_params, max_tokens) == 48, "ABI");
static_assert(offsetof(whisper_full_params, speed_up) == 52, "ABI");
static_assert(offsetof(whisper_full_params, audio_ctx) == 56, "ABI");
static_assert(offsetof(whisper_full_params, initial_prompt) == 64, "ABI");
static_assert(offsetof(whisper_full_params, language) == 88, "ABI");
static_assert(offsetof(whisper_full_params, suppress_non_speech_tokens) == 98, "ABI");
static_assert(offsetof(whisper_full_params, temperature) == 100, "ABI");
static_assert(offsetof(whisper_full_params, temperature_inc) == 112, "ABI");
static_assert(offsetof(whisper_full_params, entropy_thold) == 116, "ABI");
static_assert(offsetof(whisper_full_params, greedy) == 128, "ABI");
static_assert(offsetof(whisper_full_params, beam_search) == 132, "ABI");
static_assert(offsetof(whisper_full_params, new_segment_callback) == 144, "ABI");
static_assert(offsetof(whisper_full_params, grammar_rules) == 224, "ABI");
static_assert(offsetof(whisper_full_params, grammar_penalty) == 248, "ABI");

namespace wb200 {

namespace {
std::mutex        g_log_mutex;
ggml_log_callback g_log_cb = nullptr;
void *            g_log_ud = nullptr;
thread_local int  g_device = -1;
}  // namespace

void log_set(ggml_log_callback cb, void * user_data) {
    std::lock_guard<std::mutex> lk(g_log_mutex);
    g_log_cb = cb;
    g_log_ud = user_data;
}

void log_msg(ggml_log_level level, const char * fmt, ...) {
    char stack_buf[1024];
    va_list args;
    va_start(args, fmt);
    va_list args2;
    va_copy(args2, args);
    const int len = vsnprintf(stack_buf, sizeof(stack_buf), fmt, args);
    va_end(args);
    std::string heap;
    const char * text = stack_buf;
    if (len >= (int) sizeof(stack_buf)) {
        heap.resize(len + 1);
        vsnprintf(&heap[0], heap.size(), fmt, args2);
        text = heap.c_str();
    }
    va_end(args2);
    ggml_log_callback cb;
    void * ud;
    {
        std::lock_guard<std::mutex> lk(g_log_mutex);
        cb = g_log_cb;
        ud = g_log_ud;
    }
    if (cb) {
        cb(level, text, ud);
    } else {
        fputs(text, stderr);
        fflush(stderr);
    }
}

int selected_device() { return g_device; }

}  // namespace wb200

using namespace wb200;

extern "C" {

struct whisper_context_params whisper_context_default_params(void) {
    struct whisper_context_params p = { true };
    return p;
}

struct whisper_context * whisper_init_from_buffer_with_params(void * buffer, size_t buffer_size,
                                                              struct whisper_context_params params) {
    if (buffer == nullptr || buffer_size < 4) {
        WB_LOG_ERROR("%s: empty model buffer\n", __func__);
        return nullptr;
    }
    const int64_t t_start = time_us();
    WB_LOG_INFO("%s: loading model from buffer\n", __func__);

    std::unique_ptr<whisper_context> ctx(new whisper_context);
    ctx->params = params;
    ctx->t_start_us = t_start;
    if (!params.use_gpu) {
        // the reference would fall back to its CPU backend here (whisper.cpp:1056-1089); this library has none
        WB_LOG_WARN("%s: use_gpu = false requested, but this backend only runs on the GPU - ignoring\n", __func__);
    }

    ModelFile mf;
    if (!parse_model_file(buffer, buffer_size, mf)) {
        WB_LOG_ERROR("%s: failed to load model\n", __func__);
        return nullptr;
    }
    ctx->hparams  = mf.hparams;
    ctx->filters  = mf.filters;
    ctx->vocab    = mf.vocab;
    ctx->n_loaded = mf.n_loaded;
    ctx->rules.build(ctx->vocab);

    ctx->fwd.reset(create_forward(mf, 3 * mf.hparams.n_text_ctx, selected_device()));
    if (!ctx->fwd) {
        WB_LOG_ERROR("%s: no usable device backend - model not loaded\n", __func__);
        return nullptr;
    }
    ctx->batcher.reset(new Batcher(ctx->fwd.get()));
    ctx->state = new_state(*ctx);
    ctx->t_load_us = time_us() - t_start;
    WB_LOG_INFO("%s: backend = %s, load time = %.2f ms\n", __func__, ctx->fwd->name(), ctx->t_load_us / 1000.0);
    return ctx.release();
}

void whisper_free(struct whisper_context * ctx) {
    if (!ctx) return;
    for (whisper_context * p : ctx->peers) whisper_free(p);
    ctx->peers.clear();
    ctx->pool.reset();                      // (joins the chunk workers before anything they use goes away)
    delete ctx->state;
    ctx->state = nullptr;
    delete ctx;
}

const char * whisper_print_system_info(void) {
    static std::string s;
    static std::once_flag once;
    std::call_once(once, [] {
        s = "B200_NATIVE = 1 | CUDA = 1 | SM_100A = 1 | TCGEN05 = 1 | TMA = 1 | CPU_FALLBACK = 0 | BLAS = 0 | "
            "HOST_THREADS = " + std::to_string(std::thread::hardware_concurrency()) + " | ";
    });
    return s.c_str();
}

struct whisper_full_params whisper_full_default_params(enum whisper_sampling_strategy strategy) {
    struct whisper_full_params p;
    memset(&p, 0, sizeof(p));
    p.strategy         = strategy;
    p.n_threads        = std::min(4, (int32_t) std::thread::hardware_concurrency());
    p.n_max_text_ctx   = 16384;
    p.no_context       = true;
    p.print_progress   = true;
    p.print_timestamps = true;
    p.thold_pt         = 0.01f;
    p.thold_ptsum      = 0.01f;
    p.language         = "en";
    p.suppress_blank   = true;
    p.temperature      = 0.0f;
    p.max_initial_ts   = 1.0f;
    p.length_penalty   = -1.0f;
    p.temperature_inc  = 0.2f;
    p.entropy_thold    = 2.4f;
    p.logprob_thold    = -1.0f;
    p.no_speech_thold  = 0.6f;
    p.greedy.best_of        = -1;
    p.beam_search.beam_size = -1;
    p.beam_search.patience  = -1.0f;
    p.grammar_penalty  = 100.0f;
    switch (strategy) {
        case WHISPER_SAMPLING_GREEDY:      p.greedy.best_of = 5; break;
        case WHISPER_SAMPLING_BEAM_SEARCH: p.beam_search.beam_size = 5; p.beam_search.patience = -1.0f; break;
    }
    return p;
}

int whisper_full(struct whisper_context * ctx, struct whisper_full_params params, const float * samples, int n_samples) {
    if (!ctx || !ctx->state) return -1;
    return full_with_state(*ctx, *ctx->state, params, samples, n_samples);
}

int whisper_full_n_segments(struct whisper_context * ctx) { return (int) ctx->state->result_all.size(); }
int whisper_full_lang_id(struct whisper_context * ctx) { return ctx->state->lang_id; }
int64_t whisper_full_get_segment_t0(struct whisper_context * ctx, int i) { return ctx->state->result_all[i].t0; }
int64_t whisper_full_get_segment_t1(struct whisper_context * ctx, int i) { return ctx->state->result_all[i].t1; }
const char * whisper_full_get_segment_text(struct whisper_context * ctx, int i) { return ctx->state->result_all[i].text.c_str(); }
int whisper_full_n_tokens(struct whisper_context * ctx, int i) { return (int) ctx->state->result_all[i].tokens.size(); }
const char * whisper_full_get_token_text(struct whisper_context * ctx, int i, int j) {
    return ctx->vocab.id_to_token[ctx->state->result_all[i].tokens[j].id].c_str();
}
whisper_token whisper_full_get_token_id(struct whisper_context * ctx, int i, int j) { return ctx->state->result_all[i].tokens[j].id; }
whisper_token_data whisper_full_get_token_data(struct whisper_context * ctx, int i, int j) { return ctx->state->result_all[i].tokens[j]; }

void whisper_log_set(ggml_log_callback log_callback, void * user_data) { log_set(log_callback, user_data); }

// ---- stage API --------------------------------------------------------------------------------------------------------

int whisper_pcm_to_mel(struct whisper_context * ctx, const float * samples, int n_samples, int n_threads) {
    const int64_t t0 = time_us();
    if (!log_mel_spectrogram(samples, n_samples, n_threads, ctx->filters, ctx->state->mel)) {
        WB_LOG_ERROR("%s: failed to compute mel spectrogram\n", __func__);
        return -1;
    }
    ctx->state->t_mel_us += time_us() - t0;
    ctx->state->mel_pcm = nullptr;
    return 0;
}

int whisper_set_mel(struct whisper_context * ctx, const float * data, int n_len, int n_mel) {
    if (n_mel != ctx->filters.n_mel) {
        WB_LOG_ERROR("%s: invalid number of mel bands: %d (expected %d)\n", __func__, n_mel, ctx->filters.n_mel);
        return -1;
    }
    auto & mel = ctx->state->mel;
    mel.n_len = n_len;
    mel.n_len_org = n_len;
    mel.n_mel = n_mel;
    mel.data.assign(data, data + (size_t) n_len * n_mel);
    ctx->state->mel_pcm = nullptr;
    return 0;
}

int whisper_encode(struct whisper_context * ctx, int offset, int /*n_threads*/) {
    if (!encode_internal(*ctx, *ctx->state, offset, nullptr, nullptr)) {
        WB_LOG_ERROR("%s: failed to eval\n", __func__);
        return -1;
    }
    return 0;
}

int whisper_decode(struct whisper_context * ctx, const whisper_token * tokens, int n_tokens, int n_past, int /*n_threads*/) {
    auto & state = *ctx->state;
    state.batch.prep_legacy(tokens, n_tokens, n_past, 0);
    state.kv_self.seq_rm(0, n_past, -1);
    if (!decode_internal(*ctx, state, state.batch, nullptr, nullptr)) {
        WB_LOG_ERROR("%s: failed to eval\n", __func__);
        return 1;
    }
    return 0;
}

float * whisper_get_logits(struct whisper_context * ctx) { return ctx->state->logits.data(); }

int whisper_tokenize(struct whisper_context * ctx, const char * text, whisper_token * tokens, int n_max_tokens) {
    const auto res = tokenize(ctx->vocab, text);
    if (n_max_tokens < (int) res.size()) {
        WB_LOG_ERROR("%s: too many resulting tokens: %d (max %d)\n", __func__, (int) res.size(), n_max_tokens);
        return -1;
    }
    for (int i = 0; i < (int) res.size(); i++) tokens[i] = res[i];
    return (int) res.size();
}

int          whisper_lang_max_id(void) { return lang_max_id(); }
int          whisper_lang_id(const char * lang) { return lang_id(lang); }
const char * whisper_lang_str(int id) { return lang_str(id); }

int whisper_lang_auto_detect(struct whisper_context * ctx, int offset_ms, int /*n_threads*/, float * lang_probs) {
    return lang_auto_detect(*ctx, *ctx->state, offset_ms, lang_probs);
}

int whisper_n_len(struct whisper_context * ctx) { return ctx->state->mel.n_len_org; }
int whisper_n_vocab(struct whisper_context * ctx) { return ctx->vocab.n_vocab; }
int whisper_n_text_ctx(struct whisper_context * ctx) { return ctx->hparams.n_text_ctx; }
int whisper_n_audio_ctx(struct whisper_context * ctx) { return ctx->hparams.n_audio_ctx; }
int whisper_is_multilingual(struct whisper_context * ctx) { return ctx->vocab.is_multilingual() ? 1 : 0; }
int whisper_model_n_audio_state(struct whisper_context * ctx) { return ctx->hparams.n_audio_state; }
int whisper_model_n_audio_head(struct whisper_context * ctx) { return ctx->hparams.n_audio_head; }
int whisper_model_n_audio_layer(struct whisper_context * ctx) { return ctx->hparams.n_audio_layer; }
int whisper_model_n_text_layer(struct whisper_context * ctx) { return ctx->hparams.n_text_layer; }

const char * whisper_token_to_str(struct whisper_context * ctx, whisper_token token) {
    if (token < 0 || token >= (int) ctx->vocab.id_to_token.size()) return "";
    return ctx->vocab.id_to_token[token].c_str();
}
whisper_token whisper_token_eot(struct whisper_context * ctx) { return ctx->vocab.token_eot; }
whisper_token whisper_token_sot(struct whisper_context * ctx) { return ctx->vocab.token_sot; }
whisper_token whisper_token_solm(struct whisper_context * ctx) { return ctx->vocab.token_solm; }
whisper_token whisper_token_prev(struct whisper_context * ctx) { return ctx->vocab.token_prev; }
whisper_token whisper_token_nosp(struct whisper_context * ctx) { return ctx->vocab.token_nosp; }
whisper_token whisper_token_not(struct whisper_context * ctx) { return ctx->vocab.token_not; }
whisper_token whisper_token_beg(struct whisper_context * ctx) { return ctx->vocab.token_beg; }
whisper_token whisper_token_lang(struct whisper_context * ctx, int lang_id) { return ctx->vocab.token_lang(lang_id); }
whisper_token whisper_token_translate(struct whisper_context * ctx) { return ctx->vocab.token_translate; }
whisper_token whisper_token_transcribe(struct whisper_context * ctx) { return ctx->vocab.token_transcribe; }

void whisper_print_timings(struct whisper_context * ctx) {
    const auto & s = *ctx->state;
    const int32_t n_sample = std::max(1, s.n_sample), n_encode = std::max(1, s.n_encode), n_decode = std::max(1, s.n_decode);
    const int32_t n_batchd = std::max(1, s.n_batchd), n_prompt = std::max(1, s.n_prompt);
    WB_LOG_INFO("\n");
    WB_LOG_INFO("%s:     load time = %8.2f ms\n", __func__, ctx->t_load_us / 1000.0f);
    WB_LOG_INFO("%s:     fallbacks = %3d p / %3d h\n", __func__, s.n_fail_p, s.n_fail_h);
    WB_LOG_INFO("%s:      mel time = %8.2f ms\n", __func__, s.t_mel_us / 1000.0f);
    WB_LOG_INFO("%s:   sample time = %8.2f ms / %5d runs (%8.2f ms per run)\n", __func__, 1e-3f * s.t_sample_us, n_sample, 1e-3f * s.t_sample_us / n_sample);
    WB_LOG_INFO("%s:   encode time = %8.2f ms / %5d runs (%8.2f ms per run)\n", __func__, 1e-3f * s.t_encode_us, n_encode, 1e-3f * s.t_encode_us / n_encode);
    WB_LOG_INFO("%s:   decode time = %8.2f ms / %5d runs (%8.2f ms per run)\n", __func__, 1e-3f * s.t_decode_us, n_decode, 1e-3f * s.t_decode_us / n_decode);
    WB_LOG_INFO("%s:   batchd time = %8.2f ms / %5d runs (%8.2f ms per run)\n", __func__, 1e-3f * s.t_batchd_us, n_batchd, 1e-3f * s.t_batchd_us / n_batchd);
    WB_LOG_INFO("%s:   prompt time = %8.2f ms / %5d runs (%8.2f ms per run)\n", __func__, 1e-3f * s.t_prompt_us, n_prompt, 1e-3f * s.t_prompt_us / n_prompt);
    WB_LOG_INFO("%s:    total time = %8.2f ms\n", __func__, (time_us() - ctx->t_start_us) / 1000.0f);
}

void whisper_reset_timings(struct whisper_context * ctx) {
    ctx->t_start_us = time_us();
    auto & s = *ctx->state;
    s.t_mel_us = s.t_sample_us = s.t_encode_us = s.t_decode_us = s.t_batchd_us = s.t_prompt_us = 0;
    s.n_sample = s.n_encode = s.n_decode = s.n_batchd = s.n_prompt = 0;
}

// ---- additive entry points ----------------------------------------------------------------------------------------------

void whisper_b200_set_device(int device) { g_device = device; }

void whisper_b200_counters(struct whisper_context * ctx, int64_t * out) {
    const auto & s = *ctx->state;
    out[0] = s.n_sample; out[1] = s.n_encode; out[2] = s.n_decode; out[3] = s.n_batchd;
    out[4] = s.n_prompt; out[5] = s.n_fail_p; out[6] = s.n_fail_h;
    out[7] = ctx->fwd->kernel_launches();
}

void whisper_b200_timings_us(struct whisper_context * ctx, int64_t * out) {
    const auto & s = *ctx->state;
    out[0] = s.t_mel_us; out[1] = s.t_sample_us; out[2] = s.t_encode_us;
    out[3] = s.t_decode_us; out[4] = s.t_batchd_us; out[5] = s.t_prompt_us;
}

long long whisper_b200_read_stage(struct whisper_context * ctx, int what, void * dst, long long cap_bytes) {
    if (what == STAGE_HOST_MEL) {
        const auto & mel = ctx->state->mel;
        const long long nbytes = (long long) mel.data.size() * 4;
        if (dst) memcpy(dst, mel.data.data(), (size_t) std::min(nbytes, cap_bytes));
        return nbytes;
    }
    return ctx->fwd->read_stage(what, dst, cap_bytes);
}

void whisper_b200_gpu_times(struct whisper_context * ctx, double * out8) { ctx->fwd->gpu_times(out8); }
double whisper_b200_gpu_busy_ms(struct whisper_context * ctx) { return ctx->fwd->busy_ms(); }
double whisper_b200_gpu_mel_ms(struct whisper_context * ctx) { return ctx->fwd->mel_ms(); }
void whisper_b200_set_profiling(struct whisper_context * ctx, int on) { ctx->fwd->set_profiling(on != 0); }
void whisper_b200_profile(struct whisper_context * ctx, double * out36) { ctx->fwd->profile(out36); }

void whisper_b200_set_gemm_engine(struct whisper_context * ctx, int engine) { ctx->fwd->set_gemm_engine(engine); }

// Host cores this process may count on: its CPU affinity mask, shared evenly with the other ranks of a torchrun / mpirun launch
// on the same box (LOCAL_WORLD_SIZE); WHISPER_B200_HOST_THREADS overrides.
static int usable_cores() {
    if (const char * e = getenv("WHISPER_B200_HOST_THREADS")) return std::max(1, atoi(e));
    int n = (int) std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    int ranks = 1;
    if (const char * e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
    return std::max(1, n / ranks);
}

struct whisper_context * whisper_b200_init_multi(void * buffer, size_t buffer_size, struct whisper_context_params params,
                                                 const int * devices, int n_devices) {
    if (!devices || n_devices <= 0) return nullptr;
    // one replica per device, loaded side by side (each thread parses the caller's buffer and uploads to its own GPU)
    std::vector<whisper_context *> reps((size_t) n_devices, nullptr);
    std::vector<std::thread> threads;
    for (int i = 0; i < n_devices; ++i) {
        threads.emplace_back([&, i] {
            whisper_b200_set_device(devices[i]);          // (thread-local: selects the device of the init call below)
            reps[i] = whisper_init_from_buffer_with_params(buffer, buffer_size, params);
        });
    }
    for (auto & t : threads) t.join();
    bool ok = true;
    for (whisper_context * r : reps) ok = ok && r != nullptr;
    if (!ok) {
        WB_LOG_ERROR("%s: could not load the model on every device\n", __func__);
        for (whisper_context * r : reps) whisper_free(r);
        return nullptr;
    }
    reps[0]->peers.assign(reps.begin() + 1, reps.end());
    return reps[0];
}

int whisper_b200_dequantize(int ggml_type, const void * blocks, long long n, float * out) {
    return wb200::dequantize_blocks(ggml_type, blocks, n, out) ? 0 : -1;
}

void * whisper_b200_host_alloc(size_t bytes) { return wb200::host_alloc_pinned(bytes); }
void whisper_b200_host_free(void * p) { wb200::host_free_pinned(p); }

int whisper_b200_n_devices(struct whisper_context * ctx) { return ctx ? 1 + (int) ctx->peers.size() : 0; }

static int full_batch_single(struct whisper_context * ctx, struct whisper_full_params params,
                             const float * const * samples, const int * n_samples, int n_chunks);

int whisper_b200_full_batch(struct whisper_context * ctx, struct whisper_full_params params,
                            const float * const * samples, const int * n_samples, int n_chunks) {
    if (!ctx || n_chunks <= 0) return -1;
    if (ctx->peers.empty()) return full_batch_single(ctx, params, samples, n_samples, n_chunks);
    // chunk i -> replica i mod n; every replica runs its share as an ordinary single-device batch on its own host thread
    const int n_dev = 1 + (int) ctx->peers.size();
    std::vector<whisper_context *> reps{ctx};
    reps.insert(reps.end(), ctx->peers.begin(), ctx->peers.end());
    std::vector<std::vector<const float *>> ptrs(n_dev);
    std::vector<std::vector<int>> lens(n_dev), index(n_dev);
    for (int i = 0; i < n_chunks; ++i) { const int d = i % n_dev; ptrs[d].push_back(samples[i]); lens[d].push_back(n_samples[i]); index[d].push_back(i); }
    std::vector<int> rc(n_dev, 0);
    std::vector<std::thread> threads;
    for (int d = 0; d < n_dev; ++d) {
        if (index[d].empty()) continue;
        threads.emplace_back([&, d] { rc[d] = full_batch_single(reps[d], params, ptrs[d].data(), lens[d].data(), (int) index[d].size()); });
    }
    for (auto & t : threads) t.join();
    std::vector<std::unique_ptr<whisper_state>> all((size_t) n_chunks);
    int ret = 0;
    for (int d = 0; d < n_dev; ++d) {
        if (rc[d] != 0 && ret == 0) ret = rc[d];
        for (size_t k = 0; k < index[d].size() && k < reps[d]->chunk_states.size(); ++k) all[index[d][k]] = std::move(reps[d]->chunk_states[k]);
        reps[d]->chunk_states.clear();
        if (d > 0) {      // counters of the replicas are reported through the first context
            whisper_state * a = ctx->state; const whisper_state * b = reps[d]->state;
            a->n_encode += b->n_encode; a->n_decode += b->n_decode; a->n_fail_p += b->n_fail_p; a->n_fail_h += b->n_fail_h;
        }
    }
    for (auto & p : all) if (!p) p.reset(new_state(*ctx));      // (a replica that failed leaves empty results, never a null state)
    ctx->chunk_states = std::move(all);
    return ret;
}

static int full_batch_single(struct whisper_context * ctx, struct whisper_full_params params,
                             const float * const * samples, const int * n_samples, int n_chunks) {
    if (!ctx || n_chunks <= 0) return -1;
    // One decode state per chunk, shared read-only weights — the layout whisper_full_parallel uses (whisper.cpp:5840).
    // One host thread per in-flight chunk runs the ordinary whisper_full() state machine; their device passes are
    // merged by the Batcher so the encoder sees B chunks and every decoder step sees one row per live sequence.
    // More workers than cores: a worker is parked on a device request (an encoder pass, a whole greedy run) for most of its life,
    // so host and device work overlap once more chunks are given than there are cores.
    const int hw = usable_cores();
    const int64_t t_batch0 = time_us();
    struct rusage ru0; getrusage(RUSAGE_SELF, &ru0);
    int max_workers = std::max(16, hw + 3 * ctx->fwd->decode_rows_per_pass());
    if (const char * e = getenv("WHISPER_B200_MAX_WORKERS")) max_workers = std::max(1, atoi(e));
    const int n_workers = std::min(n_chunks, max_workers);
    ctx->batcher->set_max_host(hw);
    if (!ctx->fwd->ensure_slots(n_workers)) {
        WB_LOG_ERROR("%s: cannot allocate %d device slots\n", __func__, n_workers);
        return -1;
    }
    ctx->state->ts.energy_ext = nullptr;          // (growing the slots moves the forward pass's pinned buffers: nothing may point into the old ones)
    // the chunk states of the previous call are freed, and the new ones built (1 344 cache cells each), by the workers as they reach a
    // chunk: on the calling thread that was several milliseconds in front of the first encoder pass
    std::vector<std::unique_ptr<whisper_state>> old_states = std::move(ctx->chunk_states);
    ctx->chunk_states.clear();
    ctx->chunk_states.resize(n_chunks);
    // host log-mel threads: share the cores between the workers
    params.n_threads = std::max(1, std::min(params.n_threads, hw / n_workers));

    std::atomic<int> next{0};
    std::vector<int> rcs(n_chunks, 0);
    ctx->batcher->add_workers(n_workers);
    auto worker = [&](int w) {
        ctx->batcher->worker_attach();
        for (;;) {
            const int c = next.fetch_add(1);
            if (c >= n_chunks) break;
            if (c < (int) old_states.size()) old_states[c].reset();
            ctx->chunk_states[c].reset(new_state(*ctx));
            whisper_state & st = *ctx->chunk_states[c];
            st.slot = w;
            rcs[c] = full_with_state(*ctx, st, params, samples[c], n_samples[c]);
        }
        ctx->batcher->worker_end();
    };
    static const bool use_pool = [] { const char * e = getenv("WHISPER_B200_WORKER_POOL"); return !e || atoi(e) != 0; }();
    if (use_pool) {
        if (!ctx->pool) ctx->pool.reset(new wb200::WorkerPool);
        ctx->pool->run(n_workers, worker);
    } else {
        std::vector<std::thread> threads;
        for (int w = 0; w < n_workers; ++w) threads.emplace_back([&worker, w] { worker(w); });
        for (auto & t : threads) t.join();
    }
    if (getenv("WHISPER_B200_HOST_TRACE")) {
        Batcher & b = *ctx->batcher;
        int64_t mel_us = 0;
        for (int c = 0; c < n_chunks; ++c) mel_us += ctx->chunk_states[c]->t_mel_us;
        struct rusage ru1; getrusage(RUSAGE_SELF, &ru1);
        auto tv_ms = [](const timeval & a, const timeval & b) { return (b.tv_sec - a.tv_sec) * 1e3 + (b.tv_usec - a.tv_usec) / 1e3; };
        fprintf(stderr, "full_batch: process cpu user %.1f ms, sys %.1f ms, voluntary ctx switches %ld, involuntary %ld\n", tv_ms(ru0.ru_utime, ru1.ru_utime),
                tv_ms(ru0.ru_stime, ru1.ru_stime), ru1.ru_nvcsw - ru0.ru_nvcsw, ru1.ru_nivcsw - ru0.ru_nivcsw);
        fprintf(stderr, "full_batch: %d chunks, %d workers, %d cores | wall %.1f ms | log-mel %.1f ms per chunk (%.1f ms of core time per core) | driver: stage %.1f, device wait %.1f, "
                        "wake %.1f, encoder+other passes %.1f, idle %.1f ms | passes %lld, requests %lld\n",
                n_chunks, n_workers, hw, (time_us() - t_batch0) / 1e3, mel_us / 1e3 / n_chunks, mel_us / 1e3 / hw, b.t_stage_us / 1e3, b.t_device_wait_us / 1e3,
                b.t_complete_us / 1e3, b.t_run_us / 1e3, b.t_idle_us / 1e3, (long long) b.n_passes.load(), (long long) b.n_requests.load());
        fprintf(stderr, "full_batch: run steps %lld, rows per step %.1f\n", (long long) b.n_run_steps.load(),
                b.n_run_steps.load() ? (double) b.n_run_rows.load() / (double) b.n_run_steps.load() : 0.0);
        b.t_stage_us = b.t_device_wait_us = b.t_complete_us = b.t_run_us = b.t_idle_us = 0; b.n_passes = b.n_requests = 0; b.n_run_steps = b.n_run_rows = 0;
    }

    int ret = 0;
    whisper_state * agg = ctx->state;
    for (int c = 0; c < n_chunks; ++c) {
        const whisper_state & st = *ctx->chunk_states[c];
        if (rcs[c] != 0 && ret == 0) ret = rcs[c];
        // aggregate timers / counters into the default state, like whisper.cpp:5900-5912
        agg->t_mel_us += st.t_mel_us; agg->t_sample_us += st.t_sample_us; agg->t_encode_us += st.t_encode_us;
        agg->t_decode_us += st.t_decode_us; agg->t_batchd_us += st.t_batchd_us; agg->t_prompt_us += st.t_prompt_us;
        agg->n_sample += st.n_sample; agg->n_encode += st.n_encode; agg->n_decode += st.n_decode;
        agg->n_batchd += st.n_batchd; agg->n_prompt += st.n_prompt;
        agg->n_fail_p += st.n_fail_p; agg->n_fail_h += st.n_fail_h;
    }
    return ret;
}

int whisper_b200_chunk_n_segments(struct whisper_context * ctx, int c) { return (int) ctx->chunk_states[c]->result_all.size(); }
int whisper_b200_chunk_n_tokens(struct whisper_context * ctx, int c, int i) { return (int) ctx->chunk_states[c]->result_all[i].tokens.size(); }
const char * whisper_b200_chunk_segment_text(struct whisper_context * ctx, int c, int i) { return ctx->chunk_states[c]->result_all[i].text.c_str(); }
whisper_token_data whisper_b200_chunk_token_data(struct whisper_context * ctx, int c, int i, int j) { return ctx->chunk_states[c]->result_all[i].tokens[j]; }
int whisper_b200_chunk_token_ids(struct whisper_context * ctx, int c, whisper_token * out, int cap) {
    if (!ctx || c < 0 || c >= (int) ctx->chunk_states.size()) return -1;
    int n = 0;
    for (const auto & seg : ctx->chunk_states[c]->result_all)
        for (const auto & t : seg.tokens) { if (out && n < cap) out[n] = t.id; ++n; }
    return n;
}

}  // extern "C"
