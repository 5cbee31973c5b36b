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
            if (!decode_on(slot_ctx[jobs[i].slot], jobs[i].in, n_audio_ctx, jobs[i].logits_out)) return false;
        }
        return true;
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

    long long read_stage(int, void *, long long) override { return -1; }
    int64_t kernel_launches() const override { return 0; }
    const char * name() const override { return "TEST-ONLY checker forward (compiled reference)"; }
};

}  // namespace

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
    f->kv_cells = kv_self_cells;
    f->n_audio_ctx_model = model.hparams.n_audio_ctx;
    if (const char * t = getenv("WHISPER_HOSTLOGIC_THREADS")) f->n_threads = atoi(t);
    return f;
}

}  // namespace wb200
