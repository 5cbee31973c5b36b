// TEST INFRASTRUCTURE — not part of the product.
//
// This translation unit is compiled by oracle/Makefile into oracle/_ref/libwhisper_ref*.so.
// It textually includes the UNMODIFIED reference implementation from where it lies
// (/root/reference/thirdparty/whisper.cpp/whisper.cpp, via -I) so that the file-static internals of the
// reference (stage tensors, logits post-processing, samplers, KV-cache bookkeeping) become reachable from tests
// through a handful of extern "C" probe_* entry points.  Nothing of the reference is copied into this repo.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the result.

#include "whisper.cpp"  // reference source, resolved through -I$(W) in oracle/Makefile

#include <cstring>

extern "C" {

// ---- stage tensors -------------------------------------------------------------------------------------------------

// log-mel produced by whisper_pcm_to_mel (whisper.cpp:2793). Returns n_len; writes n_len_org.
__attribute__((visibility("default")))
int probe_mel(struct whisper_context * ctx, float * dst, int cap, int * n_len_org) {
    const auto & mel = ctx->state->mel;
    if (n_len_org) *n_len_org = mel.n_len_org;
    const int n = mel.n_mel * mel.n_len;
    if (dst && cap >= n) memcpy(dst, mel.data.data(), sizeof(float) * n);
    return mel.n_len;
}

static int probe_tensor_f32(struct ggml_tensor * t, float * dst, int cap, int * ne0, int * ne1) {
    if (!t) return -1;
    if (ne0) *ne0 = (int) t->ne[0];
    if (ne1) *ne1 = (int) t->ne[1];
    const int n = (int) ggml_nelements(t);
    if (dst && cap >= n) ggml_backend_tensor_get(t, dst, 0, sizeof(float) * n);
    return n;
}

// output of the conv stem (whisper.cpp:1709-1723): ggml [n_ctx (fast), n_state]
__attribute__((visibility("default")))
int probe_embd_conv(struct whisper_context * ctx, float * dst, int cap, int * ne0, int * ne1) {
    return probe_tensor_f32(ctx->state->embd_conv, dst, cap, ne0, ne1);
}

// output of the encoder after ln_post (whisper.cpp:1975-1985): ggml [n_state (fast), n_ctx]
__attribute__((visibility("default")))
int probe_embd_enc(struct whisper_context * ctx, float * dst, int cap, int * ne0, int * ne1) {
    return probe_tensor_f32(ctx->state->embd_enc, dst, cap, ne0, ne1);
}

// raw f16 words of the cross-attention cache (whisper.cpp:2057-2066). Returns element count per tensor.
__attribute__((visibility("default")))
int probe_kv_cross(struct whisper_context * ctx, uint16_t * k, uint16_t * v, int cap) {
    auto & kv = ctx->state->kv_cross;
    const int n = (int) ggml_nelements(kv.k);
    if (k && cap >= n) ggml_backend_tensor_get(kv.k, k, 0, sizeof(uint16_t) * n);
    if (v && cap >= n) ggml_backend_tensor_get(kv.v, v, 0, sizeof(uint16_t) * n);
    return n;
}

// raw f16 words of the unified self-attention cache (whisper.cpp:2282-2288)
__attribute__((visibility("default")))
int probe_kv_self(struct whisper_context * ctx, uint16_t * k, uint16_t * v, int cap) {
    auto & kv = ctx->state->kv_self;
    const int n = (int) ggml_nelements(kv.k);
    if (k && cap >= n) ggml_backend_tensor_get(kv.k, k, 0, sizeof(uint16_t) * n);
    if (v && cap >= n) ggml_backend_tensor_get(kv.v, v, 0, sizeof(uint16_t) * n);
    return n;
}

__attribute__((visibility("default")))
void probe_set_audio_ctx(struct whisper_context * ctx, int n) { ctx->state->exp_n_audio_ctx = n; }

// ---- decoder with an arbitrary batch (whisper.cpp:2517) ---------------------------------------------------------------

__attribute__((visibility("default")))
void probe_kv_self_clear(struct whisper_context * ctx) { whisper_kv_cache_clear(ctx->state->kv_self); }

__attribute__((visibility("default")))
void probe_kv_self_seq_cp(struct whisper_context * ctx, int src, int dst, int p0, int p1) {
    whisper_kv_cache_seq_cp(ctx->state->kv_self, src, dst, p0, p1);
}

__attribute__((visibility("default")))
void probe_kv_self_seq_rm(struct whisper_context * ctx, int seq, int p0, int p1) {
    whisper_kv_cache_seq_rm(ctx->state->kv_self, seq, p0, p1);
}

// cells: writes pos[i] and a bitmask of seq ids (bit s set if cell i holds seq s, s < 32); returns head
__attribute__((visibility("default")))
int probe_kv_self_cells(struct whisper_context * ctx, int * pos, unsigned * seq_mask, int cap, int * n_out) {
    auto & kv = ctx->state->kv_self;
    const int n = (int) kv.size;
    for (int i = 0; i < n && i < cap; ++i) {
        pos[i] = kv.cells[i].pos;
        unsigned m = 0;
        for (auto s : kv.cells[i].seq_id) if (s >= 0 && s < 32) m |= 1u << s;
        seq_mask[i] = m;
    }
    if (n_out) *n_out = (int) kv.n;
    return (int) kv.head;
}

// Overwrites the cell table (pos + seq-id bitmask per cell) and the search head — lets a test replay a KV state.
__attribute__((visibility("default")))
void probe_kv_self_set_cells(struct whisper_context * ctx, const int * pos, const unsigned * seq_mask, int n, int head) {
    auto & kv = ctx->state->kv_self;
    for (int i = 0; i < n && i < (int) kv.size; ++i) {
        kv.cells[i].pos = pos[i];
        kv.cells[i].seq_id.clear();
        for (int s = 0; s < 32; ++s) if ((seq_mask[i] >> s) & 1u) kv.cells[i].seq_id.insert(s);
    }
    kv.head = head;
}

// Returns 0 on success. Logits rows for which want_logits[i] != 0 are then readable through whisper_get_logits().
__attribute__((visibility("default")))
int probe_decode_batch(struct whisper_context * ctx, const int * tokens, const int * pos, const int * seq,
                       const signed char * want_logits, int n_tokens, int n_threads) {
    auto & batch = ctx->state->batch;
    batch.n_tokens = n_tokens;
    for (int i = 0; i < n_tokens; ++i) {
        batch.token[i]     = tokens[i];
        batch.pos[i]       = pos[i];
        batch.n_seq_id[i]  = 1;
        batch.seq_id[i][0] = seq[i];
        batch.logits[i]    = want_logits[i];
    }
    return whisper_decode_internal(*ctx, *ctx->state, batch, n_threads, nullptr, nullptr) ? 0 : 1;
}

// ---- logits post-processing and sampling (whisper.cpp:4493, 4777, 4836) ------------------------------------------------

// Runs whisper_process_logits on decoder 0 with a caller-provided raw logits row and decoding history.
__attribute__((visibility("default")))
void probe_process_logits(struct whisper_context * ctx, struct whisper_full_params params, const float * logits_in,
                          const int * tokens_cur, int n_tokens_cur, int has_ts, int seek_delta, float temperature,
                          float * logits_out, float * logprobs_out, float * probs_out) {
    auto & state   = *ctx->state;
    auto & decoder = state.decoders[0];
    const int n_vocab = ctx->vocab.n_vocab;

    state.logits.resize(n_vocab);
    memcpy(state.logits.data(), logits_in, sizeof(float) * n_vocab);

    decoder.i_batch = 0;
    decoder.sequence.tokens.clear();
    for (int i = 0; i < n_tokens_cur; ++i) {
        whisper_token_data td = { tokens_cur[i], 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, 0.0f };
        decoder.sequence.tokens.push_back(td);
    }
    decoder.has_ts     = has_ts != 0;
    decoder.seek_delta = seek_delta;

    whisper_process_logits(*ctx, state, decoder, params, temperature);

    if (logits_out)   memcpy(logits_out,   decoder.logits.data(),   sizeof(float) * n_vocab);
    if (logprobs_out) memcpy(logprobs_out, decoder.logprobs.data(), sizeof(float) * n_vocab);
    if (probs_out)    memcpy(probs_out,    decoder.probs.data(),    sizeof(float) * n_vocab);
}

// Samples from decoder 0's current probs/logprobs (as left by probe_process_logits).
__attribute__((visibility("default")))
whisper_token_data probe_sample_token(struct whisper_context * ctx, int best) {
    return whisper_sample_token(*ctx, ctx->state->decoders[0], best != 0);
}

__attribute__((visibility("default")))
int probe_sample_token_topk(struct whisper_context * ctx, int k, whisper_token_data * out) {
    auto res = whisper_sample_token_topk(*ctx, ctx->state->decoders[0], k);
    for (size_t i = 0; i < res.size(); ++i) out[i] = res[i];
    return (int) res.size();
}

__attribute__((visibility("default")))
void probe_seed_rng(struct whisper_context * ctx, int decoder, unsigned seed) {
    ctx->state->decoders[decoder].rng = std::mt19937(seed);
}

// ---- counters (whisper.cpp:770-783) ---------------------------------------------------------------------------------

// out[0..6] = n_sample, n_encode, n_decode, n_batchd, n_prompt, n_fail_p, n_fail_h
__attribute__((visibility("default")))
void probe_counters(struct whisper_context * ctx, int * out) {
    const auto & s = *ctx->state;
    out[0] = s.n_sample; out[1] = s.n_encode; out[2] = s.n_decode; out[3] = s.n_batchd;
    out[4] = s.n_prompt; out[5] = s.n_fail_p; out[6] = s.n_fail_h;
}

// out[0..5] = t_mel_us, t_sample_us, t_encode_us, t_decode_us, t_batchd_us, t_prompt_us
__attribute__((visibility("default")))
void probe_timings_us(struct whisper_context * ctx, long long * out) {
    const auto & s = *ctx->state;
    out[0] = s.t_mel_us; out[1] = s.t_sample_us; out[2] = s.t_encode_us;
    out[3] = s.t_decode_us; out[4] = s.t_batchd_us; out[5] = s.t_prompt_us;
}

// ---- mel front-end internals (whisper.cpp:2634-2725), for pinning the restated FFT / window ------------------------------

__attribute__((visibility("default")))
void probe_fft(const float * in, int n, float * out /* 2n */) {
    fill_sin_cos_table();
    std::vector<float> vin(in, in + n), vout;
    fft(vin, vout);
    memcpy(out, vout.data(), sizeof(float) * 2 * n);
}

__attribute__((visibility("default")))
void probe_hann(int length, int periodic, float * out) {
    std::vector<float> h;
    hann_window(length, periodic != 0, h);
    memcpy(out, h.data(), sizeof(float) * length);
}

// ---- ABI facts -------------------------------------------------------------------------------------------------------

__attribute__((visibility("default")))
int probe_sizeof_full_params(void) { return (int) sizeof(whisper_full_params); }

__attribute__((visibility("default")))
int probe_sizeof_token_data(void) { return (int) sizeof(whisper_token_data); }

// ggml's 65 536-entry f16 tables (ggml.c:2218-2236), for pinning the restated GELU / exp numerics
__attribute__((visibility("default")))
void probe_f16_tables(uint16_t * gelu, uint16_t * exp_) {
    // make sure tables are initialised
    struct ggml_init_params p = { 1024, nullptr, true };
    struct ggml_context * c = ggml_init(p);
    ggml_free(c);
    for (int i = 0; i < 65536; ++i) {
        ggml_fp16_t h; uint16_t u = (uint16_t) i; memcpy(&h, &u, 2);
        float f = ggml_fp16_to_fp32(h);
        // GELU through the public f32 path is not exported; recompute like ggml.c:1404 / 2229-2233
        float g = 0.5f*f*(1.0f + tanhf(0.79788456080286535587989211986876f*f*(1.0f + 0.044715f*f*f)));
        ggml_fp16_t gh = ggml_fp32_to_fp16(g);
        ggml_fp16_t eh = ggml_fp32_to_fp16(expf(f));
        memcpy(&gelu[i], &gh, 2);
        memcpy(&exp_[i], &eh, 2);
    }
}

} // extern "C"
