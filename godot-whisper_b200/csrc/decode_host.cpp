// Host-side decode bookkeeping — see decode_host.h.  Reference line numbers are for
// /root/reference/thirdparty/whisper.cpp/whisper.cpp.
#include "decode_host.h"
#include "forward.h"
#include "common.h"
#include "mel.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>

namespace wb200 {

// ---- KV cells ---------------------------------------------------------------------------------------------------------

void KvCells::clear() {                                   // :1000-1006
    for (auto & c : cells) { c.pos = -1; c.seq_mask = 0; }
    head = 0;
}

bool KvCells::find_slot(int n_tokens_i, const int32_t * pos, const int32_t * seq) {   // :938-986
    const uint32_t n_ctx = size;
    const uint32_t n_tokens = (uint32_t) n_tokens_i;
    if (n_tokens > n_ctx) {
        WB_LOG_ERROR("%s: n_tokens=%d > n_ctx=%d\n", __func__, n_tokens, n_ctx);
        return false;
    }
    uint32_t n_tested = 0;
    while (true) {
        if (head + n_tokens > n_ctx) {
            n_tested += n_ctx - head;
            head = 0;
            continue;
        }
        bool found = true;
        for (uint32_t i = 0; i < n_tokens; i++) {
            if (cells[head + i].pos >= 0) {
                found = false;
                head     += i + 1;
                n_tested += i + 1;
                break;
            }
        }
        if (found) break;
        if (n_tested >= n_ctx) return false;
    }
    for (uint32_t i = 0; i < n_tokens; i++) {
        cells[head + i].pos = pos[i];
        cells[head + i].seq_mask |= 1u << seq[i];
    }
    return true;
}

int32_t KvCells::cell_max() const {                       // :989-997
    for (uint32_t i = size - 1; i > 0; --i) {
        if (cells[i].pos >= 0 && cells[i].seq_mask != 0) return (int32_t) i + 1;
    }
    return 1;
}

void KvCells::seq_rm(int seq, int32_t p0, int32_t p1) {    // :1008-1036
    uint32_t new_head = size;
    if (p0 < 0) p0 = 0;
    if (p1 < 0) p1 = std::numeric_limits<int32_t>::max();
    for (uint32_t i = 0; i < size; ++i) {
        if (cells[i].pos >= p0 && cells[i].pos < p1) {
            if (seq < 0) {
                cells[i].seq_mask = 0;
            } else if (cells[i].has_seq(seq)) {
                cells[i].seq_mask &= ~(1u << seq);
            } else {
                continue;
            }
            if (cells[i].seq_mask == 0) {
                cells[i].pos = -1;
                if (new_head == size) new_head = i;
            }
        }
    }
    if (new_head != size) head = new_head;
}

void KvCells::seq_cp(int seq_src, int seq_dst, int32_t p0, int32_t p1) {   // :1038-1054
    if (p0 < 0) p0 = 0;
    if (p1 < 0) p1 = std::numeric_limits<int32_t>::max();
    head = 0;
    for (uint32_t i = 0; i < size; ++i) {
        if (cells[i].has_seq(seq_src) && cells[i].pos >= p0 && cells[i].pos < p1) {
            cells[i].seq_mask |= 1u << seq_dst;
        }
    }
}

void Batch::prep_legacy(const int32_t * tokens, int n, int n_past, int seq_id) {   // :446-458
    sample_on_device = false;
    dist_on_device = false;
    if ((int) token.size() < n) reserve(n);
    if ((int) token.size() < n) reserve(n);
    n_tokens = n;
    for (int i = 0; i < n; ++i) {
        if (tokens) token[i] = tokens[i];
        pos[i]    = n_past + i;
        seq[i]    = seq_id;
        logits[i] = 0;
        n_draws[i] = 0;
    }
    draws.clear();
    if (n > 0) logits[n - 1] = 1;
}

// ---- logits rules -----------------------------------------------------------------------------------------------------

void LogitsRules::build(const Vocab & vocab) {
    // the symbol list of whisper.cpp:4482-4487 (OpenAI tokenizer.py non_speech_tokens)
    static const char * const kSymbols[] = {
        "\"", "#", "(", ")", "*", "+", "/", ":", ";", "<", "=", ">", "@", "[", "\\", "]", "^",
        "_", "`", "{", "|", "}", "~", "「", "」", "『", "』", "<<", ">>", "<<<", ">>>", "--",
        "---", "-(", "-[", "('", "(\"", "((", "))", "(((", ")))", "[[", "]]", "{{", "}}", "♪♪",
        "♪♪♪", "♩", "♪", "♫", "♬", "♭", "♮", "♯",
    };
    non_speech.clear();
    auto add = [&](const std::string & s) {
        auto it = vocab.token_to_id.find(s);
        if (it != vocab.token_to_id.end()) non_speech.push_back(it->second);
    };
    for (const char * sym : kSymbols) {
        add(sym);
        add(std::string(" ") + sym);
    }
    add(" -");   // hyphens / quotes allowed inside words but not at word start (:4586-4592)
    add(" '");
    auto it = vocab.token_to_id.find(" ");
    blank = it != vocab.token_to_id.end() ? it->second : -1;
}

namespace {

const float NEG_INF = -INFINITY;

// log_softmax over the finite entries (:4637-4655)
void log_softmax(const std::vector<float> & logits, std::vector<float> & logprobs, int n) {
    const float logit_max = *std::max_element(logits.begin(), logits.begin() + n);
    float logsumexp = 0.0f;
    for (int i = 0; i < n; ++i) {
        if (logits[i] > NEG_INF) logsumexp += expf(logits[i] - logit_max);
    }
    logsumexp = logf(logsumexp) + logit_max;
    for (int i = 0; i < n; ++i) {
        logprobs[i] = logits[i] > NEG_INF ? logits[i] - logsumexp : NEG_INF;
    }
}

}  // namespace

void process_logits(const Vocab & vocab, const LogitsRules & rules, int n_audio_ctx_model,
                    const whisper_full_params & params, struct whisper_context * ctx, struct whisper_state * state,
                    const float * raw, Decoder & decoder, float temperature) {
    const auto & tokens_cur = decoder.sequence.tokens;
    const bool is_initial = tokens_cur.empty();
    const int  n_logits   = vocab.n_vocab;

    auto & probs    = decoder.probs;
    auto & logits   = decoder.logits;
    auto & logprobs = decoder.logprobs;

    logits.resize(n_logits);
    memcpy(logits.data(), raw, sizeof(float) * n_logits);
    if (temperature > 0.0f) {
        for (int i = 0; i < n_logits; i++) logits[i] /= temperature;
    }
    probs.resize(n_logits);
    logprobs.resize(n_logits);

    // suppression rules (:4527-4594)
    if (params.suppress_blank && is_initial) {
        logits[vocab.token_eot] = NEG_INF;
        if (rules.blank >= 0) logits[rules.blank] = NEG_INF;
    }
    logits[vocab.token_not] = NEG_INF;
    if (params.no_timestamps) {
        for (int i = vocab.token_beg; i < n_logits; ++i) logits[i] = NEG_INF;
    }
    logits[vocab.token_sot]  = NEG_INF;
    logits[vocab.token_nosp] = NEG_INF;
    if (!params.tdrz_enable) logits[vocab.token_solm] = NEG_INF;
    logits[vocab.token_translate]  = NEG_INF;
    logits[vocab.token_transcribe] = NEG_INF;
    logits[vocab.token_prev]       = NEG_INF;
    for (int i = 0; i < lang_count(); ++i) {
        const int id = vocab.token_lang(i);
        if (id >= 0 && id < n_logits) logits[id] = NEG_INF;
    }

    if (params.logits_filter_callback) {
        params.logits_filter_callback(ctx, state, tokens_cur.data(), (int) tokens_cur.size(), logits.data(),
                                      params.logits_filter_callback_user_data);
    }

    if (params.suppress_non_speech_tokens) {
        for (int32_t id : rules.non_speech) logits[id] = NEG_INF;
    }

    // timestamps come in pairs, except directly before EOT (:4598-4614)
    {
        const bool last_was_timestamp        = !tokens_cur.empty() && tokens_cur.back().id >= vocab.token_beg;
        const bool penultimate_was_timestamp = tokens_cur.size() < 2 || tokens_cur[tokens_cur.size() - 2].id >= vocab.token_beg;
        if (last_was_timestamp) {
            if (penultimate_was_timestamp) {
                for (int i = vocab.token_beg; i < n_logits; ++i) logits[i] = NEG_INF;
            } else {
                for (int i = 0; i < vocab.token_eot; ++i) logits[i] = NEG_INF;
            }
        }
    }

    // the first timestamp may not exceed max_initial_ts (:4618-4625)
    if (is_initial && params.max_initial_ts > 0.0f) {
        const float precision = float(WHISPER_CHUNK_SIZE) / n_audio_ctx_model;
        const int   tid0      = std::round(params.max_initial_ts / precision);
        for (int i = vocab.token_beg + tid0 + 1; i < n_logits; ++i) logits[i] = NEG_INF;
    }

    // timestamps must not decrease (:4629-4635)
    if (decoder.has_ts) {
        const int tid0 = decoder.seek_delta / 2;
        for (int i = vocab.token_beg; i < vocab.token_beg + tid0 && i < n_logits; ++i) logits[i] = NEG_INF;
    }

    log_softmax(logits, logprobs, n_logits);

    // if the timestamp mass beats every text token, force a timestamp (:4659-4684)
    {
        float timestamp_logprob = NEG_INF;
        {
            float logsumexp = 0.0f;
            const float logprob_max = *std::max_element(logprobs.begin() + vocab.token_beg, logprobs.begin() + n_logits);
            for (int i = vocab.token_beg; i < n_logits; ++i) {
                if (logprobs[i] > NEG_INF) logsumexp += expf(logprobs[i] - logprob_max);
            }
            if (logsumexp > 0.0f) timestamp_logprob = logf(logsumexp) + logprob_max;
        }
        const float max_text_token_logprob = *std::max_element(logprobs.begin(), logprobs.begin() + vocab.token_beg);
        if (timestamp_logprob > max_text_token_logprob) {
            for (int i = 0; i < vocab.token_beg; ++i) {
                logits[i]   = NEG_INF;
                logprobs[i] = NEG_INF;
            }
        }
    }

    for (int i = 0; i < n_logits; ++i) {
        probs[i] = logits[i] == NEG_INF ? 0.0f : expf(logprobs[i]);
    }
}

void make_sample_rule(const Vocab & vocab, int n_audio_ctx_model, const whisper_full_params & params, const Decoder & decoder,
                      int32_t * rule4) {
    const auto & tokens_cur = decoder.sequence.tokens;
    const bool is_initial = tokens_cur.empty();
    SampleRule r;
    if (params.suppress_blank && is_initial) r.flags |= SampleRule::INITIAL_BLANK;                  // :4532-4537
    if (params.no_timestamps)                r.flags |= SampleRule::NO_TIMESTAMPS;                  // :4543-4547
    if (!params.tdrz_enable)                 r.flags |= SampleRule::SUPPRESS_SOLM;                  // :4553-4555
    if (params.suppress_non_speech_tokens)   r.flags |= SampleRule::NON_SPEECH;                     // :4576-4593
    const bool last_was_timestamp        = !tokens_cur.empty() && tokens_cur.back().id >= vocab.token_beg;
    const bool penultimate_was_timestamp = tokens_cur.size() < 2 || tokens_cur[tokens_cur.size() - 2].id >= vocab.token_beg;
    if (last_was_timestamp)        r.flags |= SampleRule::LAST_TS;                                  // :4598-4614
    if (penultimate_was_timestamp) r.flags |= SampleRule::PENULT_TS;
    if (is_initial && params.max_initial_ts > 0.0f) {                                               // :4618-4625
        const float precision = float(WHISPER_CHUNK_SIZE) / n_audio_ctx_model;
        r.flags |= SampleRule::INITIAL_MAX_TS;
        r.tid0_initial = (int32_t) std::round(params.max_initial_ts / precision);
    }
    if (decoder.has_ts) {                                                                           // :4629-4635
        r.flags |= SampleRule::HAS_TS;
        r.tid0_seek = decoder.seek_delta / 2;
    }
    rule4[0] = r.flags; rule4[1] = r.tid0_initial; rule4[2] = r.tid0_seek; rule4[3] = 0;
}

whisper_token_data sample_token(const Vocab & vocab, const Decoder & decoder, bool best) {   // :4777-4834
    whisper_token_data result = { 0, 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, 0.0f };
    const auto & probs    = decoder.probs;
    const auto & logprobs = decoder.logprobs;
    const int n_logits = vocab.n_vocab;
    {
        double sum_ts = 0.0;
        double max_ts = 0.0;
        for (int i = vocab.token_beg; i < n_logits; i++) {
            if (probs[i] == NEG_INF) continue;
            sum_ts += probs[i];
            if (max_ts < probs[i]) {
                max_ts = probs[i];
                result.tid = i;
            }
        }
        result.pt    = max_ts / (sum_ts + 1e-10);
        result.ptsum = sum_ts;
    }
    if (best) {
        for (int i = 0; i < n_logits; ++i) {
            if (result.p < probs[i]) {
                result.id   = i;
                result.p    = probs[i];
                result.plog = logprobs[i];
            }
        }
    } else {
        std::discrete_distribution<> dist(probs.begin(), probs.begin() + n_logits);
        result.id   = dist(decoder.rng);
        result.p    = probs[result.id];
        result.plog = logprobs[result.id];
    }
    if (result.id >= vocab.token_beg) {
        result.tid = result.id;
        result.pt  = result.p;
    }
    return result;
}

std::vector<whisper_token_data> sample_token_topk(const Vocab & vocab, Decoder & decoder, int k) {   // :4836-4909
    const auto & probs    = decoder.probs;
    const auto & logprobs = decoder.logprobs;
    const int n_logits = vocab.n_vocab;
    // NOTE: the reference partial_sorts (logit, id) pairs here but never uses the result for selection; the k
    // candidates are k independent draws from the full distribution, which is what is restated.
    std::vector<whisper_token_data> result;
    result.reserve(k);
    whisper_token tid = vocab.token_beg;
    float pt = 0.0f, ptsum = 0.0f;
    {
        double sum_ts = 0.0;
        double max_ts = 0.0;
        for (int i = vocab.token_beg; i < n_logits; i++) {
            if (probs[i] == NEG_INF) continue;
            sum_ts += probs[i];
            if (max_ts < probs[i]) {
                max_ts = probs[i];
                tid = i;
            }
        }
        pt    = max_ts / (sum_ts + 1e-10);
        ptsum = sum_ts;
    }
    std::discrete_distribution<> dist(probs.begin(), probs.begin() + n_logits);
    for (int i = 0; i < k; ++i) {
        const auto id = dist(decoder.rng);
        result.push_back({ (whisper_token) id, tid, probs[id], logprobs[id], pt, ptsum, -1, -1, 0.0f });
        if (result[i].id >= vocab.token_beg) {
            result[i].tid = result[i].id;
            result[i].pt  = result[i].p;
        }
    }
    return result;
}

void sequence_score(const whisper_full_params & params, Sequence & sequence) {   // :4912-4958
    if (sequence.result_len == 0) return;
    double result = 0.0f;
    for (int i = 0; i < sequence.result_len; ++i) result += sequence.tokens[i].plog;
    sequence.sum_logprobs = result;
    sequence.avg_logprobs = result / sequence.result_len;
    double penalty = sequence.result_len;
    if (params.length_penalty > 0.0f) penalty = pow((5.0 + penalty) / 6.0, params.length_penalty);
    sequence.score = result / penalty;
    {
        const int n = 32;
        int cnt = 0;
        double entropy = 0.0f;
        std::map<whisper_token, int> token_counts;
        for (int i = std::max(0, sequence.result_len - n); i < sequence.result_len; ++i) {
            token_counts[sequence.tokens[i].id]++;
            cnt++;
        }
        for (const auto & kv : token_counts) {
            const auto p = kv.second / (double) cnt;
            entropy -= p * log(p);
        }
        sequence.entropy = entropy;
    }
}

// ---- token-level timestamps -------------------------------------------------------------------------------------------

namespace {

int timestamp_to_sample(int64_t t, int n_samples) {                       // :6315-6317
    return std::max(0, std::min((int) n_samples - 1, (int) ((t * WHISPER_SAMPLE_RATE) / 100)));
}
int64_t sample_to_timestamp(int i_sample) { return (100ll * i_sample) / WHISPER_SAMPLE_RATE; }   // :6319-6321

float voice_length(const std::string & text) {                            // :6325-6347
    float res = 0.0f;
    for (char c : text) {
        if (c == ' ')                  res += 0.01f;
        else if (c == ',')             res += 2.00f;
        else if (c == '.' || c == '!' || c == '?') res += 3.00f;
        else if (c >= '0' && c <= '9') res += 3.00f;
        else                           res += 1.00f;
    }
    return res;
}

}  // namespace

void TimestampState::ensure_energy() {
    if (!pending_pcm) return;
    signal_energy(pending_pcm, pending_n, 32, energy);                    // :5003-5010 -> :6350-6366
    pending_pcm = nullptr;
    energy_ext = nullptr;
}

void compute_token_level_timestamps(const Vocab & vocab, TimestampState & ts, Segment & segment,
                                    float thold_pt, float thold_ptsum) {   // :6368-6578
    ts.ensure_energy();
    auto & tokens = segment.tokens;
    const int n_samples = ts.energy_size();
    if (n_samples == 0) {
        WB_LOG_ERROR("%s: no signal data available\n", __func__);
        return;
    }
    const int64_t t0 = segment.t0;
    const int64_t t1 = segment.t1;
    const int n = (int) tokens.size();
    if (n == 0) return;
    if (n == 1) {
        tokens[0].t0 = t0;
        tokens[0].t1 = t1;
        return;
    }
    auto & t_beg    = ts.t_beg;
    auto & t_last   = ts.t_last;
    auto & tid_last = ts.tid_last;

    for (int j = 0; j < n; ++j) {
        auto & token = tokens[j];
        if (j == 0) {
            if (token.id == vocab.token_beg) {
                tokens[j    ].t0 = t0;
                tokens[j    ].t1 = t0;
                tokens[j + 1].t0 = t0;
                t_beg    = t0;
                t_last   = t0;
                tid_last = vocab.token_beg;
            } else {
                tokens[j].t0 = t_last;
            }
        }
        const int64_t tt = t_beg + 2 * (token.tid - vocab.token_beg);
        token.vlen = voice_length(std::string(vocab.id_to_token[token.id].c_str()));
        if (token.pt > thold_pt && token.ptsum > thold_ptsum && token.tid > tid_last && tt <= t1) {
            if (j > 0) tokens[j - 1].t1 = tt;
            tokens[j].t0 = tt;
            tid_last = token.tid;
        }
    }
    tokens[n - 2].t1 = t1;
    tokens[n - 1].t0 = t1;
    tokens[n - 1].t1 = t1;
    t_last = t1;

    // spread unknown timestamps proportionally to the voice length (:6448-6490)
    {
        int p0 = 0, p1 = 0;
        while (true) {
            while (p1 < n && tokens[p1].t1 < 0) p1++;
            if (p1 >= n) p1--;
            if (p1 > p0) {
                double psum = 0.0;
                for (int j = p0; j <= p1; j++) psum += tokens[j].vlen;
                const double dt = tokens[p1].t1 - tokens[p0].t0;
                for (int j = p0 + 1; j <= p1; j++) {
                    const double ct = tokens[j - 1].t0 + dt * tokens[j - 1].vlen / psum;
                    tokens[j - 1].t1 = ct;
                    tokens[j    ].t0 = ct;
                }
            }
            p1++;
            p0 = p1;
            if (p1 >= n) break;
        }
    }

    // fix-up pass (:6493-6504)
    for (int j = 0; j < n - 1; j++) {
        if (tokens[j].t1 < 0) tokens[j + 1].t0 = tokens[j].t1;
        if (j > 0) {
            if (tokens[j - 1].t1 > tokens[j].t0) {
                tokens[j].t0 = tokens[j - 1].t1;
                tokens[j].t1 = std::max(tokens[j].t0, tokens[j].t1);
            }
        }
    }

    // snap token boundaries to voice activity (:6508-6567)
    {
        const float * energy = ts.energy_data();
        const int hw = WHISPER_SAMPLE_RATE / 8;
        for (int j = 0; j < n; j++) {
            if (tokens[j].id >= vocab.token_eot) continue;
            int s0 = timestamp_to_sample(tokens[j].t0, n_samples);
            int s1 = timestamp_to_sample(tokens[j].t1, n_samples);
            const int ss0 = std::max(s0 - hw, 0);
            const int ss1 = std::min(s1 + hw, n_samples);
            const int ns = ss1 - ss0;
            float sum = 0.0f;
            for (int k = ss0; k < ss1; k++) sum += energy[k];
            const float thold = 0.5 * sum / ns;
            {
                int k = s0;
                if (energy[k] > thold && j > 0) {
                    while (k > 0 && energy[k] > thold) k--;
                    tokens[j].t0 = sample_to_timestamp(k);
                    if (tokens[j].t0 < tokens[j - 1].t1) tokens[j].t0 = tokens[j - 1].t1;
                    else s0 = k;
                } else {
                    while (energy[k] < thold && k < s1) k++;
                    s0 = k;
                    tokens[j].t0 = sample_to_timestamp(k);
                }
            }
            {
                int k = s1;
                if (energy[k] > thold) {
                    while (k < n_samples - 1 && energy[k] > thold) k++;
                    tokens[j].t1 = sample_to_timestamp(k);
                    // (the reference tests "j < ns - 1" here — ns is the SAMPLE count of the window — and so reads tokens[j + 1] one past
                    // the end for the last token, whisper.cpp:6547; that value is heap garbage there, so the last token is simply not clamped here)
                    if (j < ns - 1 && j + 1 < n && tokens[j].t1 > tokens[j + 1].t0) tokens[j].t1 = tokens[j + 1].t0;
                    else s1 = k;
                } else {
                    while (energy[k] < thold && k > s0) k--;
                    s1 = k;
                    tokens[j].t1 = sample_to_timestamp(k);
                }
            }
        }
    }
}

int wrap_segment(const Vocab & vocab, std::vector<Segment> & result_all, int max_len, bool split_on_word) {   // :4428-4480
    Segment segment = result_all.back();
    int res = 1;
    int acc = 0;
    std::string text;
    for (int i = 0; i < (int) segment.tokens.size(); i++) {
        const auto & token = segment.tokens[i];
        if (token.id >= vocab.token_eot) continue;
        const std::string & txt = vocab.id_to_token[token.id];
        const int cur = (int) strlen(txt.c_str());
        const bool may_split = !split_on_word || txt.c_str()[0] == ' ';
        if (acc + cur > max_len && i > 0 && may_split) {
            result_all.back().text = std::move(text);
            result_all.back().t1 = token.t0;
            result_all.back().tokens.resize(i);
            result_all.back().speaker_turn_next = false;

            result_all.push_back({});
            result_all.back().t0 = token.t0;
            result_all.back().t1 = segment.t1;
            result_all.back().tokens.insert(result_all.back().tokens.end(), segment.tokens.begin() + i, segment.tokens.end());
            result_all.back().speaker_turn_next = segment.speaker_turn_next;

            acc = 0;
            text = "";
            segment = result_all.back();
            i = -1;
            res++;
        } else {
            acc += cur;
            text += txt;
        }
    }
    result_all.back().text = std::move(text);
    return res;
}

}  // namespace wb200
