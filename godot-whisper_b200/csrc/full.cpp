// whisper_full(): the C++ driver of the hot path.  PCM -> host log-mel -> GPU encoder -> GPU decoder token loop with
// temperature fallback, greedy / beam sampling, segment assembly and token-level timestamps.
//
// Control flow restated from /root/reference/thirdparty/whisper.cpp/whisper.cpp:4960-5807 (line numbers below refer to
// that file).  All tensor math happens behind wb200::Forward (CUDA); this file is integer / f32 / f64 bookkeeping and
// must make the same decisions as the reference given the same logits.
#include "context.h"
#include "common.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <random>
#include <thread>

namespace wb200 {

whisper_state * new_state(const whisper_context & ctx) {              // whisper_init_state, :3001-3120
    whisper_state * state = new whisper_state;
    // 3x n_text_ctx cells: the reference over-allocates for up to 8 concurrent decoders (:3006-3012)
    state->kv_self.init(3 * ctx.hparams.n_text_ctx);
    // (logits / probability buffers are sized where they are first written: a chunk whose token loop runs on the device never touches
    // them, and 512 chunk states per batch call should not each map megabytes they will not use)
    state->batch.reserve(ctx.hparams.n_text_ctx);
    state->decoders[0].sequence.tokens.reserve(ctx.hparams.n_text_ctx / 2);
    state->decoders[0].rng = std::mt19937(0);                         // seeded once, never reseeded (:3064)
    return state;
}

bool encode_internal(whisper_context & ctx, whisper_state & state, int mel_offset,
                     whisper_abort_callback abort_cb, void * abort_ud) {
    const int64_t t_start_us = time_us();
    const auto & hp  = ctx.hparams;
    const auto & mel = state.mel;
    const int n_ctx  = state.exp_n_audio_ctx > 0 ? state.exp_n_audio_ctx : hp.n_audio_ctx;
    const int n_mels = hp.n_mels;

    if (mel.n_mel != n_mels) {
        WB_LOG_ERROR("%s: mel spectrogram has %d bands, model expects %d\n", __func__, mel.n_mel, n_mels);
        return false;
    }

    if (!state.mel_pcm && mel.data.size() < (size_t) mel.n_mel * mel.n_len) {
        WB_LOG_ERROR("%s: no spectrogram (call whisper_pcm_to_mel / whisper_set_mel / whisper_full first)\n", __func__);
        return false;
    }
    if (state.mel_pcm) {
        // the spectrogram lives on the device: the first encode of the clip brings its PCM along (through a pinned staging buffer this
        // thread fills), later windows of the same clip only name their first frame
        float * stage = nullptr;
        const float * pcm = nullptr;
        float * energy_out = nullptr;
        if (!state.mel_dev_ready) {
            if (ctx.fwd->is_pinned_host(state.mel_pcm)) pcm = state.mel_pcm;       // page-locked by the caller: uploaded from where it lies
            else {
                stage = ctx.fwd->pcm_stage_acquire(state.mel_pcm_n);
                if (!stage) { WB_LOG_ERROR("%s: no staging buffer for %d samples\n", __func__, state.mel_pcm_n); return false; }
                memcpy(stage, state.mel_pcm, sizeof(float) * (size_t) state.mel_pcm_n);
                pcm = stage;
            }
            // the same pass computes the clip's energy envelope (token timestamps snap to it): into the slot's pinned buffer, where it is
            // read in place; without one it comes back through the staging buffer
            if (state.ts.pending_pcm == state.mel_pcm && state.ts.pending_n == state.mel_pcm_n) {
                energy_out = ctx.fwd->energy_buffer(state.slot);
                if (!energy_out) energy_out = stage;
            }
        }
        const bool ok = ctx.batcher->encode_pcm(state.slot, pcm, state.mel_pcm_n, mel_offset, n_ctx, energy_out);
        if (ok && energy_out) {
            if (energy_out == stage) { state.ts.energy.assign(stage, stage + state.mel_pcm_n); state.ts.energy_ext = nullptr; }
            else { state.ts.energy_ext = energy_out; state.ts.energy_ext_n = state.mel_pcm_n; }
            state.ts.pending_pcm = nullptr;
        }
        if (stage) ctx.fwd->pcm_stage_release(stage);
        if (!ok) return false;
        state.mel_dev_ready = true;
        state.t_encode_us += time_us() - t_start_us;
        state.n_encode++;
        return !(abort_cb && abort_cb(abort_ud));
    }

    // window copy with zero padding to 2*n_ctx frames (:1692-1706)
    state.mel_window.assign((size_t) n_mels * 2 * n_ctx, 0.0f);
    const int i0 = std::min(mel_offset, mel.n_len);
    const int i1 = std::min(mel_offset + 2 * n_ctx, mel.n_len);
    for (int j = 0; j < n_mels; ++j) {
        if (i1 > i0) {
            memcpy(state.mel_window.data() + (size_t) j * 2 * n_ctx, mel.data.data() + (size_t) j * mel.n_len + i0,
                   sizeof(float) * (size_t) (i1 - i0));
        }
    }

    if (!ctx.batcher->encode(state.slot, state.mel_window.data(), n_ctx)) return false;

    state.t_encode_us += time_us() - t_start_us;
    state.n_encode++;
    return !(abort_cb && abort_cb(abort_ud));
}

bool decode_internal(whisper_context & ctx, whisper_state & state, const Batch & batch,
                     whisper_abort_callback abort_cb, void * abort_ud) {
    const int64_t t_start_us = time_us();
    const int n_vocab  = ctx.hparams.n_vocab;
    const int n_tokens = batch.n_tokens;
    auto & kv = state.kv_self;

    if (!kv.find_slot(n_tokens, batch.pos.data(), batch.seq.data())) return false;
    kv.n = kv.cell_max();

    DecodeInput in;
    in.n_tokens = n_tokens;
    in.token = batch.token.data();
    in.pos   = batch.pos.data();
    in.seq   = batch.seq.data();
    in.want_logits = batch.logits.data();
    in.kv_head = (int) kv.head;
    in.n_kv    = (int) kv.n;
    in.cells   = kv.cells.data();

    const int n_audio_ctx = state.exp_n_audio_ctx > 0 ? state.exp_n_audio_ctx : ctx.hparams.n_audio_ctx;

    if (batch.dist_on_device) {
        size_t total = 0;
        for (int i = 0; i < n_tokens; ++i) if (batch.logits[i]) total += (size_t) batch.n_draws[i];
        in.sample = (const SampleRule *) batch.rule.data();
        in.n_draws = batch.n_draws.data(); in.draws = batch.draws.data(); in.temperature = batch.temperature; in.tid_default = batch.tid_default;
        state.drawn.resize(total);
    } else if (batch.sample_on_device) {
        in.sample = (const SampleRule *) batch.rule.data();
        state.sampled.resize(n_tokens);
    } else {
        bool any = false;
        for (int i = 0; i < n_tokens; ++i) any = any || batch.logits[i];
        if (any) state.logits.resize((size_t) n_tokens * n_vocab);     // (a prefill pass that wants no logits moves none)
    }
    if (!ctx.batcher->decode(state.slot, in, n_audio_ctx, state.logits.data(), state.sampled.data(), state.drawn.data())) return false;

    if (n_tokens == 1) {
        state.t_decode_us += time_us() - t_start_us;
        state.n_decode++;
    } else if (n_tokens < 16) {
        state.t_batchd_us += time_us() - t_start_us;
        state.n_batchd += n_tokens;
    } else {
        state.t_prompt_us += time_us() - t_start_us;
        state.n_prompt += n_tokens;
    }
    return !(abort_cb && abort_cb(abort_ud));
}

int lang_auto_detect(whisper_context & ctx, whisper_state & state, int offset_ms, float * lang_probs) {   // :3569-3642
    const int seek = offset_ms / 10;
    if (seek < 0) {
        WB_LOG_ERROR("%s: offset %dms is before the start of the audio\n", __func__, offset_ms);
        return -1;
    }
    if (seek >= state.mel.n_len_org) {
        WB_LOG_ERROR("%s: offset %dms is past the end of the audio (%dms)\n", __func__, offset_ms, state.mel.n_len_org * 10);
        return -2;
    }
    if (!encode_internal(ctx, state, seek, nullptr, nullptr)) {
        WB_LOG_ERROR("%s: failed to encode\n", __func__);
        return -6;
    }
    const int32_t prompt[1] = { ctx.vocab.token_sot };
    state.batch.prep_legacy(prompt, 1, 0, 0);
    state.kv_self.seq_rm(0, 0, -1);
    if (!decode_internal(ctx, state, state.batch, nullptr, nullptr)) {
        WB_LOG_ERROR("%s: failed to decode\n", __func__);
        return -7;
    }

    // the reference walks its language map in key (code) order before an unstable sort (:3603-3616)
    std::vector<std::pair<std::string, int>> by_code;
    for (int i = 0; i < lang_count(); ++i) by_code.emplace_back(lang_str(i), i);
    std::sort(by_code.begin(), by_code.end());

    auto & logits_id = state.decoders[0].logits_id;
    logits_id.clear();
    for (const auto & kv : by_code) {
        const int tok = ctx.vocab.token_lang(kv.second);
        const float l = (tok >= 0 && tok < ctx.vocab.n_vocab) ? state.logits[tok] : -INFINITY;
        logits_id.push_back({ (double) l, kv.second });
    }
    std::sort(logits_id.begin(), logits_id.end(),
              [](const Decoder::LogitId & a, const Decoder::LogitId & b) { return a.first > b.first; });
    {
        const auto max = logits_id[0].first;
        double sum = 0.0f;
        for (auto & kv : logits_id) {
            kv.first = exp(kv.first - max);
            sum += kv.first;
        }
        for (auto & kv : logits_id) kv.first /= sum;
    }
    if (lang_probs) {
        for (const auto & prob : logits_id) lang_probs[prob.second] = (float) prob.first;
    }
    return logits_id[0].second;
}

namespace {

struct BeamCandidate {       // :5136-5144
    int  decoder_idx;
    int  seek_delta;
    bool has_ts;
    Sequence sequence;
};

template <typename F>
void for_each_decoder_parallel(int n_threads_req, int n_decoders_cur, F && fn) {
    // same fan-out as the reference (:5300-5357): an atomic work index shared by min(n_threads, n_decoders) threads
    std::atomic<int> j_cur(0);
    auto process = [&]() {
        while (true) {
            const int j = j_cur.fetch_add(1);
            if (j >= n_decoders_cur) break;
            fn(j);
        }
    };
    const int n_threads = std::min(n_threads_req, n_decoders_cur);
    if (n_threads <= 1) {
        process();
    } else {
        std::vector<std::thread> threads(n_threads - 1);
        for (auto & t : threads) t = std::thread(process);
        process();
        for (auto & t : threads) t.join();
    }
}

}  // namespace

int full_with_state(whisper_context & ctx, whisper_state & state, whisper_full_params params,
                    const float * samples, int n_samples) {
    auto & result_all = state.result_all;
    result_all.clear();
    struct PcmGuard {          // `samples` is borrowed for this call only: once it returns, the state may at most refer to the device's copy
        whisper_state & st;
        ~PcmGuard() { if (!st.mel_dev_ready) st.mel_pcm = nullptr; st.ts.ensure_energy(); }
    } pcm_guard{state};

    const Vocab & vocab = ctx.vocab;
    const int n_text_ctx = ctx.hparams.n_text_ctx;

    // Greedy decoding at temperature 0 lets the device apply whisper_process_logits' rules and pick the token
    // (SURVEY.md §8f.1).  Anything that needs the full distribution on the host — beam search, best-of sampling at
    // t > 0, a logits filter callback — keeps the logits download + host path.  WHISPER_B200_DEVICE_SAMPLING=0 forces it.
    bool greedy_on_device = params.strategy == WHISPER_SAMPLING_GREEDY && params.logits_filter_callback == nullptr;
    if (const char * e = getenv("WHISPER_B200_DEVICE_SAMPLING")) greedy_on_device = greedy_on_device && atoi(e) != 0;
    const bool device_sampling = greedy_on_device && ctx.fwd->can_sample();
    // ... and lets the whole token loop of such a pass stay on the device (Forward::run_*): no host round trip per token.  A weight-less
    // test model (n_loaded == 0) completes after one token on the host (:5492-5497) and keeps that path.
    const bool use_runs = greedy_on_device && ctx.fwd->supports_runs() && ctx.n_loaded > 0;
    // Passes that draw from the distribution — best-of decoders at t > 0, beam search — send the uniform variates of the decoders'
    // generators along and get the drawn tokens back (Forward::can_sample_dist) instead of the logits.
    bool dist_sampling = ctx.fwd->can_sample_dist() && params.logits_filter_callback == nullptr;
    if (const char * e = getenv("WHISPER_B200_DEVICE_SAMPLING")) dist_sampling = dist_sampling && atoi(e) != 0;
    auto draw_uniform = [](std::mt19937 & rng) { return std::generate_canonical<double, std::numeric_limits<double>::digits>(rng); };   // what std::discrete_distribution takes per sample
    const int n_cand = params.strategy == WHISPER_SAMPLING_BEAM_SEARCH ? params.beam_search.beam_size : 1;

    if (params.grammar_rules != nullptr && params.n_grammar_rules > 0) {
        WB_LOG_ERROR("%s: grammar-constrained sampling is not supported by this backend\n", __func__);
        return -9;
    }

    if (n_samples > 0) {
        if (params.speed_up) {                                           // :4973-4976
            WB_LOG_ERROR("%s: failed to compute log mel spectrogram\n", __func__);
            return -1;
        }
        state.mel_pcm = nullptr; state.mel_dev_ready = false;
        bool dev_mel = ctx.fwd->mel_on_device();
        if (const char * e = getenv("WHISPER_B200_HOST_MEL")) dev_mel = dev_mel && atoi(e) == 0;
        dev_mel = dev_mel && n_samples <= ctx.fwd->pcm_stage_samples();      // (longer clips keep the host transform)
        if (dev_mel) {
            // one window or less: the device computes the spectrogram (cuda/mel_kernels.cu, same arithmetic); only its shape is needed here
            int n_calc = 0;
            mel_shape(n_samples, state.mel.n_len, state.mel.n_len_org, n_calc);
            state.mel.n_mel = ctx.filters.n_mel;
            state.mel.data.clear();
            state.mel_pcm = samples; state.mel_pcm_n = n_samples;
        }
        if (!dev_mel) {
            ctx.batcher->host_phase_begin();          // long host-only phase: batches of the other chunk workers do not wait for it
            const int64_t t0 = time_us();             // (the time spent queueing for a core is not log-mel time)
            const bool mel_ok = log_mel_spectrogram(samples, n_samples, params.n_threads, ctx.filters, state.mel);
            ctx.batcher->host_phase_end();
            if (!mel_ok) {
                WB_LOG_ERROR("%s: failed to compute log mel spectrogram\n", __func__);
                return -2;
            }
            state.t_mel_us += time_us() - t0;
        }
    }

    // language auto-detection (:4986-5001)
    if (params.language == nullptr || strlen(params.language) == 0 || strcmp(params.language, "auto") == 0 || params.detect_language) {
        std::vector<float> probs(lang_max_id() + 1, 0.0f);
        const int lid = lang_auto_detect(ctx, state, 0, probs.data());
        if (lid < 0) {
            WB_LOG_ERROR("%s: failed to auto-detect language\n", __func__);
            return -3;
        }
        state.lang_id = lid;
        params.language = lang_str(lid);
        WB_LOG_INFO("%s: auto-detected language: %s (p = %f)\n", __func__, params.language, probs[lang_id(params.language)]);
        if (params.detect_language) return 0;
    }

    if (params.token_timestamps) {                                       // :5003-5010
        state.ts.t_beg = 0;
        state.ts.t_last = 0;
        state.ts.tid_last = 0;
        if (n_samples > 0) { state.ts.pending_pcm = samples; state.ts.pending_n = n_samples; state.ts.energy_ext = nullptr; }     // computed at the first segment (ensure_energy) or by the encoder pass
    }

    const int seek_start = params.offset_ms / 10;
    const int seek_end   = params.duration_ms == 0 ? state.mel.n_len_org : seek_start + params.duration_ms / 10;

    // less than 1.0 s of audio: nothing to do (:5015-5021)
    if (seek_end < seek_start + (params.speed_up ? 50 : 100)) return 0;

    std::vector<float> temperatures;                                     // :5025-5032
    if (params.temperature_inc > 0.0f) {
        for (float t = params.temperature; t < 1.0f + 1e-6f; t += params.temperature_inc) temperatures.push_back(t);
    } else {
        temperatures.push_back(params.temperature);
    }

    int n_decoders = 1;                                                  // :5035-5052
    switch (params.strategy) {
        case WHISPER_SAMPLING_GREEDY:      n_decoders = params.greedy.best_of; break;
        case WHISPER_SAMPLING_BEAM_SEARCH: n_decoders = std::max(params.greedy.best_of, params.beam_search.beam_size); break;
    }
    n_decoders = std::max(1, n_decoders);
    if (n_decoders > kMaxDecoders) {
        WB_LOG_ERROR("%s: too many decoders requested (%d), max = %d\n", __func__, n_decoders, kMaxDecoders);
        return -4;
    }

    for (int j = 1; j < n_decoders; j++) {                               // :5055-5067
        auto & decoder = state.decoders[j];
        decoder.sequence.tokens.reserve(state.decoders[0].sequence.tokens.capacity());
        // (probs / logits / logprobs are sized where they are first written: process_logits, or the copy from decoder 0 after the
        // prompt pass — a greedy t = 0 chunk that is sampled on the device never touches them)
        decoder.rng = std::mt19937(0);
    }

    auto & prompt_past = state.prompt_past;
    if (params.no_context) prompt_past.clear();

    std::vector<int32_t> prompt_tokens_buf;                              // :5076-5094
    if (!params.prompt_tokens && params.initial_prompt) {
        prompt_tokens_buf = tokenize(vocab, params.initial_prompt);
        if (prompt_tokens_buf.size() > 1024) {
            WB_LOG_ERROR("%s: too many resulting tokens: %d (max %d)\n", __func__, (int) prompt_tokens_buf.size(), 1024);
            prompt_tokens_buf.clear();   // the reference's whisper_tokenize returns -1 here; treat as no prompt
        }
        params.prompt_tokens   = prompt_tokens_buf.data();
        params.prompt_n_tokens = (int) prompt_tokens_buf.size();
    }
    if (params.prompt_tokens && params.prompt_n_tokens > 0) {
        for (int i = 0; i < params.prompt_n_tokens; i++) prompt_past.push_back(params.prompt_tokens[i]);
        std::rotate(prompt_past.begin(), prompt_past.end() - params.prompt_n_tokens, prompt_past.end());
    }

    if (params.audio_ctx > ctx.hparams.n_audio_ctx) {                    // :5098-5102
        WB_LOG_ERROR("%s: audio_ctx is larger than the maximum allowed (%d > %d)\n", __func__, params.audio_ctx, ctx.hparams.n_audio_ctx);
        return -5;
    }
    state.exp_n_audio_ctx = params.audio_ctx;

    std::vector<int32_t> prompt_init = { vocab.token_sot };              // :5105-5126
    if (vocab.is_multilingual()) {
        const int lid = lang_id(params.language);
        state.lang_id = lid;
        prompt_init.push_back(vocab.token_lang(lid));
        prompt_init.push_back(params.translate ? vocab.token_translate : vocab.token_transcribe);
    }
    {
        const bool is_distil = ctx.hparams.n_text_layer == 2;
        if (is_distil && !params.no_timestamps) {
            WB_LOG_WARN("%s: using distilled model - forcing no_timestamps\n", __func__);
            params.no_timestamps = true;
        }
    }
    if (params.no_timestamps) prompt_init.push_back(vocab.token_not);

    int seek = seek_start;

    std::vector<int32_t> prompt;
    prompt.reserve(n_text_ctx);

    std::vector<std::vector<BeamCandidate>> bc_per_dec(n_decoders);
    std::vector<BeamCandidate> beam_candidates;

    struct DecodePhase {          // released on every return path
        Batcher * b; bool held = false;
        explicit DecodePhase(Batcher * b_) : b(b_) {}
        void acquire() { if (!held) { b->decode_phase_begin(); held = true; } }
        ~DecodePhase() { if (held) b->decode_phase_end(); }
    } decode_phase(ctx.batcher.get());

    // main loop over 30 s windows (:5150)
    while (true) {
        if (params.progress_callback) {
            const int progress_cur = (100 * (seek - seek_start)) / (seek_end - seek_start);
            params.progress_callback(&ctx, &state, progress_cur, params.progress_callback_user_data);
        }
        if (seek + 100 >= seek_end) break;

        if (params.encoder_begin_callback) {
            if (params.encoder_begin_callback(&ctx, &state, params.encoder_begin_callback_user_data) == false) {
                WB_LOG_ERROR("%s: encoder_begin_callback returned false - aborting\n", __func__);
                break;
            }
        }

        if (!encode_internal(ctx, state, seek, params.abort_callback, params.abort_callback_user_data)) {
            WB_LOG_ERROR("%s: failed to encode\n", __func__);
            return -6;
        }
        {
            // measurement hook (bench.py's encoder-only leg): stop behind the encoder, so that batched encoder passes can be timed with
            // nothing else on the device.  Never set by the product.
            const char * e = getenv("WHISPER_B200_ENCODE_ONLY");
            if (e && atoi(e) != 0) break;
        }
        decode_phase.acquire();       // chunk workers: at most one decoder pass worth of sequences decode at a time (Batcher)

        if (seek > seek_start && seek + 500 >= seek_end) prompt_past.clear();   // :5177-5179

        int best_decoder_id = 0;

        for (int it = 0; it < (int) temperatures.size(); ++it) {
            const float t_cur = temperatures[it];

            int n_decoders_cur = 1;                                      // :5187-5207
            switch (params.strategy) {
                case WHISPER_SAMPLING_GREEDY:
                    if (t_cur > 0.0f) n_decoders_cur = params.greedy.best_of;
                    break;
                case WHISPER_SAMPLING_BEAM_SEARCH:
                    n_decoders_cur = t_cur > 0.0f ? params.greedy.best_of : params.beam_search.beam_size;
                    break;
            }
            n_decoders_cur = std::max(1, n_decoders_cur);

            for (int j = 0; j < n_decoders_cur; ++j) {                   // :5212-5234
                auto & decoder = state.decoders[j];
                decoder.sequence.tokens.clear();
                decoder.sequence.result_len       = 0;
                decoder.sequence.sum_logprobs_all = 0.0;
                decoder.sequence.sum_logprobs     = -INFINITY;
                decoder.sequence.avg_logprobs     = -INFINITY;
                decoder.sequence.entropy          = 0.0;
                decoder.sequence.score            = -INFINITY;
                decoder.seek_delta = 100 * WHISPER_CHUNK_SIZE;
                decoder.failed    = false;
                decoder.completed = false;
                decoder.has_ts    = false;
            }

            // Greedy at temperature 0 with one decoder: prompt + token loop as ONE device-resident run (run_state.h).  The
            // prompt's last token is the first input of the run; whatever precedes it is prefilled in one pass without logits.
            bool ran_on_device = false;
            if (use_runs && t_cur < 1e-6f && n_decoders_cur == 1) {
                prompt.clear();
                if (!prompt_past.empty() && t_cur < 0.5f && params.n_max_text_ctx > 0) {
                    const int n_take = std::min(std::min(params.n_max_text_ctx, n_text_ctx / 2), int(prompt_past.size()));
                    prompt = { vocab.token_prev };
                    prompt.insert(prompt.begin() + 1, prompt_past.end() - n_take, prompt_past.end());
                }
                prompt.insert(prompt.end(), prompt_init.begin(), prompt_init.end());
                const int P = (int) prompt.size();
                state.kv_self.clear();
                state.decoders[0].has_pending = false;
                if (P > 1) {
                    state.batch.prep_legacy(prompt.data(), P - 1, 0, 0);
                    state.batch.logits[P - 2] = 0;
                    if (!decode_internal(ctx, state, state.batch, nullptr, nullptr)) {
                        WB_LOG_ERROR("%s: failed to decode\n", __func__);
                        return -7;
                    }
                }
                auto & decoder = state.decoders[0];
                RunSeq rs;
                rs.max_tokens = params.max_tokens; rs.seek = seek; rs.seek_end = seek_end; rs.single_segment = params.single_segment ? 1 : 0;
                rs.n_max = n_text_ctx / 2 - 4;
                {
                    int32_t rule4[4];
                    make_sample_rule(vocab, ctx.hparams.n_audio_ctx, params, decoder, rule4);     // decoder.sequence.tokens is empty: the initial rule
                    rs.rule_static  = rule4[0] & (SampleRule::NO_TIMESTAMPS | SampleRule::SUPPRESS_SOLM | SampleRule::NON_SPEECH);
                    rs.rule_initial = rule4[0] & (SampleRule::INITIAL_BLANK | SampleRule::INITIAL_MAX_TS);
                    rs.tid0_initial = rule4[1];
                }
                rs.token = prompt[P - 1]; rs.pos = P - 1; rs.i = 0;
                rs.seek_delta = decoder.seek_delta; rs.has_ts = 0; rs.result_len = 0;
                const int n_audio_ctx = state.exp_n_audio_ctx > 0 ? state.exp_n_audio_ctx : ctx.hparams.n_audio_ctx;
                const int64_t t_run0 = time_us();
                RunSeq fin;
                if (!ctx.batcher->run(state.slot, rs, n_audio_ctx, fin, decoder.sequence.tokens)) {
                    WB_LOG_ERROR("%s: failed to decode\n", __func__);
                    return -8;
                }
                // the cells the run wrote, mirrored on the host (whisper_kv_cache_find_slot would have handed out the same ones)
                for (int c = P - 1; c < std::min((int) state.kv_self.size, P - 1 + fin.n_out); ++c) { state.kv_self.cells[c].pos = c; state.kv_self.cells[c].seq_mask |= 1u; }
                state.kv_self.n = state.kv_self.cell_max();
                for (const auto & tok : decoder.sequence.tokens) decoder.sequence.sum_logprobs_all += tok.plog;
                decoder.seek_delta = fin.seek_delta;
                decoder.has_ts     = fin.has_ts != 0;
                decoder.sequence.result_len = fin.result_len;
                decoder.failed     = fin.status == RUN_FAILED;
                decoder.completed  = fin.status == RUN_COMPLETED;
                state.t_decode_us += time_us() - t_run0;
                state.n_decode    += fin.n_out;
                if (params.abort_callback && params.abort_callback(params.abort_callback_user_data)) {
                    WB_LOG_ERROR("%s: failed to decode\n", __func__);
                    return -8;
                }
                ran_on_device = true;
            }

            // prompt pass (:5238-5286)
            if (!ran_on_device) {
                prompt.clear();
                if (!prompt_past.empty() && t_cur < 0.5f && params.n_max_text_ctx > 0) {
                    const int n_take = std::min(std::min(params.n_max_text_ctx, n_text_ctx / 2), int(prompt_past.size()));
                    prompt = { vocab.token_prev };
                    prompt.insert(prompt.begin() + 1, prompt_past.end() - n_take, prompt_past.end());
                }
                prompt.insert(prompt.end(), prompt_init.begin(), prompt_init.end());

                state.kv_self.clear();
                state.batch.prep_legacy(prompt.data(), (int) prompt.size(), 0, 0);
                // greedy at temperature 0: the logits rules and the pick run on the device (24 bytes come back per sequence)
                const bool dev_sample = device_sampling && t_cur < 1e-6f && n_decoders_cur == 1;
                const bool dist_pass = dist_sampling && !dev_sample && n_cand >= 1 && n_cand * n_decoders_cur <= 64 &&
                                       (params.strategy == WHISPER_SAMPLING_BEAM_SEARCH || t_cur >= 1e-6f);
                state.batch.sample_on_device = dev_sample;
                state.batch.dist_on_device = dist_pass;
                for (int j = 0; j < n_decoders_cur; ++j) { state.decoders[j].has_pending = false; state.decoders[j].has_cands = false; }
                if (dev_sample || dist_pass) {
                    make_sample_rule(vocab, ctx.hparams.n_audio_ctx, params, state.decoders[0], state.batch.rule.data() + 4 * (prompt.size() - 1));
                }
                if (dist_pass) {
                    // every decoder of this pass samples its first token(s) from the prompt's distribution with its own generator (:5276-5284)
                    state.batch.temperature = t_cur;
                    state.batch.tid_default = params.strategy == WHISPER_SAMPLING_BEAM_SEARCH ? vocab.token_beg : 0;
                    state.batch.n_draws[prompt.size() - 1] = n_cand * n_decoders_cur;
                    state.batch.draws.clear();
                    for (int j = 0; j < n_decoders_cur; ++j)
                        for (int c = 0; c < n_cand; ++c) state.batch.draws.push_back(draw_uniform(state.decoders[j].rng));
                }

                if (!decode_internal(ctx, state, state.batch, params.abort_callback, params.abort_callback_user_data)) {
                    WB_LOG_ERROR("%s: failed to decode\n", __func__);
                    return -7;
                }
                {
                    const int64_t t_start_sample_us = time_us();
                    state.decoders[0].i_batch = (int) prompt.size() - 1;
                    if (dist_pass) {
                        for (int j = 0; j < n_decoders_cur; ++j) {
                            state.decoders[j].cands.assign(state.drawn.begin() + (size_t) j * n_cand, state.drawn.begin() + (size_t) (j + 1) * n_cand);
                            state.decoders[j].has_cands = true;
                        }
                    } else if (dev_sample) {
                        state.decoders[0].pending = state.sampled[state.decoders[0].i_batch];
                        state.decoders[0].has_pending = true;
                    } else
                    process_logits(vocab, ctx.rules, ctx.hparams.n_audio_ctx, params, &ctx, &state,
                                   state.logits.data() + (size_t) state.decoders[0].i_batch * vocab.n_vocab,
                                   state.decoders[0], t_cur);
                    for (int j = 1; j < n_decoders_cur; ++j) {
                        auto & decoder = state.decoders[j];
                        state.kv_self.seq_cp(0, j, -1, -1);
                        if (dist_pass) continue;
                        decoder.probs    = state.decoders[0].probs;
                        decoder.logits   = state.decoders[0].logits;
                        decoder.logprobs = state.decoders[0].logprobs;
                    }
                    state.t_sample_us += time_us() - t_start_sample_us;
                }
            }

            // token loop (:5288)
            for (int i = 0, n_max = ran_on_device ? 0 : n_text_ctx / 2 - 4; i < n_max; ++i) {
                const int64_t t_start_sample_us = time_us();

                if (params.strategy == WHISPER_SAMPLING_BEAM_SEARCH) {
                    for (auto & bc : bc_per_dec) bc.clear();
                }

                // sampling (:5297-5357)
                for_each_decoder_parallel(params.n_threads, n_decoders_cur, [&](int j) {
                    auto & decoder = state.decoders[j];
                    if (decoder.completed || decoder.failed) return;
                    switch (params.strategy) {
                        case WHISPER_SAMPLING_GREEDY: {
                            decoder.sequence.tokens.push_back(decoder.has_pending ? decoder.pending : decoder.has_cands ? decoder.cands[0]
                                                                                                   : sample_token(vocab, decoder, t_cur < 1e-6f));
                            decoder.sequence.sum_logprobs_all += decoder.sequence.tokens.back().plog;
                        } break;
                        case WHISPER_SAMPLING_BEAM_SEARCH: {
                            const auto tokens_new = decoder.has_cands ? decoder.cands : sample_token_topk(vocab, decoder, params.beam_search.beam_size);
                            for (const auto & token : tokens_new) {
                                bc_per_dec[j].push_back({ j, decoder.seek_delta, decoder.has_ts, decoder.sequence });
                                bc_per_dec[j].back().sequence.tokens.push_back(token);
                                bc_per_dec[j].back().sequence.sum_logprobs_all += token.plog;
                            }
                        } break;
                    }
                });

                beam_candidates.clear();
                for (const auto & bc : bc_per_dec) {
                    beam_candidates.insert(beam_candidates.end(), bc.begin(), bc.end());
                    if (!bc.empty()) state.n_sample += 1;
                }

                // beam search: keep the best candidates, re-point the KV cells (:5370-5419)
                if (params.strategy == WHISPER_SAMPLING_BEAM_SEARCH) {
                    std::sort(beam_candidates.begin(), beam_candidates.end(),
                              [](const BeamCandidate & a, const BeamCandidate & b) {
                                  return a.sequence.sum_logprobs_all > b.sequence.sum_logprobs_all;
                              });
                    uint32_t cur_c = 0;
                    for (int j = 0; j < n_decoders_cur; ++j) {
                        auto & decoder = state.decoders[j];
                        if (decoder.completed || decoder.failed) continue;
                        if (cur_c >= beam_candidates.size()) cur_c = 0;
                        auto & cur = beam_candidates[cur_c++];
                        while (beam_candidates.size() > cur_c &&
                               beam_candidates[cur_c].sequence.sum_logprobs_all == cur.sequence.sum_logprobs_all && i > 0) {
                            ++cur_c;
                        }
                        decoder.seek_delta = cur.seek_delta;
                        decoder.has_ts     = cur.has_ts;
                        decoder.sequence   = cur.sequence;
                        state.kv_self.seq_cp(cur.decoder_idx, kMaxDecoders + j, -1, -1);
                    }
                    for (int j = 0; j < n_decoders_cur; ++j) {
                        auto & decoder = state.decoders[j];
                        if (decoder.completed || decoder.failed) continue;
                        state.kv_self.seq_rm(j, -1, -1);
                        state.kv_self.seq_cp(kMaxDecoders + j, j, -1, -1);
                        state.kv_self.seq_rm(kMaxDecoders + j, -1, -1);
                    }
                }

                // per-decoder state update (:5425-5507)
                for (int j = 0; j < n_decoders_cur; ++j) {
                    auto & decoder = state.decoders[j];
                    if (decoder.completed || decoder.failed) continue;

                    auto & has_ts     = decoder.has_ts;
                    auto & failed     = decoder.failed;
                    auto & completed  = decoder.completed;
                    auto & seek_delta = decoder.seek_delta;
                    auto & result_len = decoder.sequence.result_len;
                    {
                        const auto & token = decoder.sequence.tokens.back();

                        if (token.id > vocab.token_beg) {               // timestamp token: slide the window
                            const int seek_delta_new = 2 * (token.id - vocab.token_beg);
                            if (has_ts && seek_delta > seek_delta_new && result_len < i) {
                                failed = true;                         // going back in time
                                continue;
                            }
                            seek_delta = seek_delta_new;
                            result_len = i + 1;
                            has_ts = true;
                        }

                        if (token.id == vocab.token_eot ||                                  // end of text
                            (params.max_tokens > 0 && i >= params.max_tokens) ||            // per-segment token budget
                            (has_ts && seek + seek_delta + 100 >= seek_end)) {              // end of audio
                            if (result_len == 0) {
                                if (seek + seek_delta + 100 >= seek_end) {
                                    result_len = i + 1;
                                } else {
                                    failed = true;
                                    continue;
                                }
                            }
                            if (params.single_segment) {
                                result_len = i + 1;
                                seek_delta = 100 * WHISPER_CHUNK_SIZE;
                            }
                            completed = true;
                            continue;
                        }

                        if (ctx.n_loaded == 0) {                        // weight-less test model (:5492-5497)
                            seek_delta = 100 * WHISPER_CHUNK_SIZE;
                            completed = true;
                            continue;
                        }
                    }
                    // repetition-loop guard (:5501-5506)
                    if (i == n_max - 1 && (result_len == 0 || seek_delta < 100 * WHISPER_CHUNK_SIZE / 2)) {
                        failed = true;
                        continue;
                    }
                }

                {
                    bool completed_all = true;
                    for (int j = 0; j < n_decoders_cur; ++j) {
                        auto & decoder = state.decoders[j];
                        if (decoder.completed || decoder.failed) continue;
                        completed_all = false;
                    }
                    if (completed_all) break;
                }

                state.t_sample_us += time_us() - t_start_sample_us;

                // next-token logits for every live decoder (:5531-5605)
                {
                    auto & batch = state.batch;
                    batch.n_tokens = 0;
                    const int n_past = (int) prompt.size() + i;
                    for (int j = 0; j < n_decoders_cur; ++j) {
                        auto & decoder = state.decoders[j];
                        if (decoder.failed || decoder.completed) continue;
                        decoder.i_batch = batch.n_tokens;
                        batch.token [batch.n_tokens] = decoder.sequence.tokens.back().id;
                        batch.pos   [batch.n_tokens] = n_past;
                        batch.seq   [batch.n_tokens] = j;
                        batch.logits[batch.n_tokens] = 1;
                        if (batch.sample_on_device || batch.dist_on_device) {
                            make_sample_rule(vocab, ctx.hparams.n_audio_ctx, params, decoder, batch.rule.data() + 4 * (size_t) batch.n_tokens);
                        }
                        if (batch.dist_on_device) {
                            if (batch.n_tokens == 0) batch.draws.clear();
                            // the variates this decoder's NEXT sampling step would take from its generator; the last turn of the loop
                            // samples nothing any more (the reference decodes once more and drops the logits, :5288)
                            const int k = i + 1 < n_max ? n_cand : 0;
                            batch.n_draws[batch.n_tokens] = k;
                            if (k == 0) batch.logits[batch.n_tokens] = 0;
                            for (int c = 0; c < k; ++c) batch.draws.push_back(draw_uniform(decoder.rng));
                        }
                        batch.n_tokens++;
                    }

                    if (!decode_internal(ctx, state, batch, params.abort_callback, params.abort_callback_user_data)) {
                        WB_LOG_ERROR("%s: failed to decode\n", __func__);
                        return -8;
                    }

                    const int64_t t_start_sample_us2 = time_us();
                    for_each_decoder_parallel(params.n_threads, n_decoders_cur, [&](int j) {
                        auto & decoder = state.decoders[j];
                        if (decoder.failed || decoder.completed) return;
                        if (state.batch.dist_on_device) return;        // (handed out below: the drawn tokens are in row order)
                        if (state.batch.sample_on_device) {
                            decoder.pending = state.sampled[decoder.i_batch];
                            decoder.has_pending = true;
                            return;
                        }
                        process_logits(vocab, ctx.rules, ctx.hparams.n_audio_ctx, params, &ctx, &state,
                                       state.logits.data() + (size_t) decoder.i_batch * vocab.n_vocab, decoder, t_cur);
                    });
                    if (state.batch.dist_on_device) {
                        size_t at = 0;
                        for (int j = 0; j < n_decoders_cur; ++j) {
                            auto & decoder = state.decoders[j];
                            if (decoder.failed || decoder.completed) continue;
                            const int k = state.batch.n_draws[decoder.i_batch];
                            decoder.cands.assign(state.drawn.begin() + at, state.drawn.begin() + at + k);
                            decoder.has_cands = k > 0;
                            at += (size_t) k;
                        }
                    }
                    state.t_sample_us += time_us() - t_start_sample_us2;
                }
            }

            // rank the sequences (:5609-5645)
            {
                double best_score = -INFINITY;
                for (int j = 0; j < n_decoders_cur; ++j) {
                    auto & decoder = state.decoders[j];
                    if (decoder.failed) continue;
                    decoder.sequence.tokens.resize(decoder.sequence.result_len);
                    sequence_score(params, decoder.sequence);
                    if (decoder.sequence.result_len > 32 && decoder.sequence.entropy < params.entropy_thold) {
                        decoder.failed = true;
                        state.n_fail_h++;
                        continue;
                    }
                    if (best_score < decoder.sequence.score) {
                        best_score = decoder.sequence.score;
                        best_decoder_id = j;
                    }
                }
            }

            // fallback decision (:5647-5668)
            bool success = true;
            if (it != (int) temperatures.size() - 1) {
                const auto & decoder = state.decoders[best_decoder_id];
                if (decoder.failed || decoder.sequence.avg_logprobs < params.logprob_thold) {
                    success = false;
                    state.n_fail_p++;
                }
            }
            if (success) break;
        }

        // emit results (:5674-5800)
        {
            const auto & best_decoder = state.decoders[best_decoder_id];
            const auto seek_delta = best_decoder.seek_delta;
            const auto result_len = best_decoder.sequence.result_len;
            const auto & tokens_cur = best_decoder.sequence.tokens;

            prompt_past.clear();
            if (prompt.front() == vocab.token_prev) {
                prompt_past.insert(prompt_past.end(), prompt.begin() + 1, prompt.end() - prompt_init.size());
            }
            for (int i = 0; i < result_len; ++i) prompt_past.push_back(tokens_cur[i].id);

            auto finish_segment = [&]() {
                int n_new = 1;
                if (params.token_timestamps) {
                    compute_token_level_timestamps(vocab, state.ts, result_all.back(), params.thold_pt, params.thold_ptsum);
                    if (params.max_len > 0) n_new = wrap_segment(vocab, result_all, params.max_len, params.split_on_word);
                }
                if (params.new_segment_callback) {
                    params.new_segment_callback(&ctx, &state, n_new, params.new_segment_callback_user_data);
                }
            };

            if (!tokens_cur.empty() && ctx.n_loaded > 0) {
                int  i0 = 0;
                auto t0 = seek + 2 * (tokens_cur.front().tid - vocab.token_beg);
                std::string text;
                bool speaker_turn_next = false;

                for (int i = 0; i < (int) tokens_cur.size(); i++) {
                    if (params.print_special || tokens_cur[i].id < vocab.token_eot) {
                        text += vocab.id_to_token[tokens_cur[i].id].c_str();
                    }
                    if (params.tdrz_enable && tokens_cur[i].id == vocab.token_solm) speaker_turn_next = true;

                    if (tokens_cur[i].id > vocab.token_beg && !params.single_segment) {
                        const auto t1 = seek + 2 * (tokens_cur[i].tid - vocab.token_beg);
                        if (!text.empty()) {
                            const auto tt0 = params.speed_up ? 2 * t0 : t0;
                            const auto tt1 = params.speed_up ? 2 * t1 : t1;
                            result_all.push_back({ tt0, tt1, text, {}, speaker_turn_next });
                            for (int j = i0; j <= i; j++) result_all.back().tokens.push_back(tokens_cur[j]);
                            finish_segment();
                        }
                        text = "";
                        while (i < (int) tokens_cur.size() && tokens_cur[i].id > vocab.token_beg) i++;
                        i--;
                        t0 = t1;
                        i0 = i + 1;
                        speaker_turn_next = false;
                    }
                }

                if (!text.empty()) {
                    const auto t1 = seek + seek_delta;
                    const auto tt0 = params.speed_up ? 2 * t0 : t0;
                    const auto tt1 = params.speed_up ? 2 * t1 : t1;
                    result_all.push_back({ tt0, tt1, text, {}, speaker_turn_next });
                    for (int j = i0; j < (int) tokens_cur.size(); j++) result_all.back().tokens.push_back(tokens_cur[j]);
                    finish_segment();
                }
            }

            seek += seek_delta;
        }
    }

    return 0;
}

}  // namespace wb200
