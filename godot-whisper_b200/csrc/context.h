// whisper_context / whisper_state as this library defines them (opaque to the host).
#pragma once

#include "batcher.h"
#include "decode_host.h"
#include "forward.h"
#include "mel.h"
#include "model.h"

#include <memory>
#include <string>
#include <vector>
#include "worker_pool.h"

struct whisper_state {
    // phase timers and counters, same meaning as whisper.cpp:770-783
    int64_t t_sample_us = 0, t_encode_us = 0, t_decode_us = 0, t_batchd_us = 0, t_prompt_us = 0, t_mel_us = 0;
    int32_t n_sample = 0, n_encode = 0, n_decode = 0, n_batchd = 0, n_prompt = 0, n_fail_p = 0, n_fail_h = 0;

    wb200::KvCells kv_self;            // host mirror of the unified self-attention cache cells
    wb200::Mel     mel;
    wb200::Batch   batch;
    wb200::Decoder decoders[wb200::kMaxDecoders];

    std::vector<float> mel_window;     // [n_mels][2*n_ctx] staging for the conv stem
    std::vector<float> logits;         // [n_tokens][n_vocab] of the last decode
    std::vector<whisper_token_data> sampled;   // [n_tokens] device-sampled tokens of the last decode (greedy fast path)
    std::vector<whisper_token_data> drawn;     // tokens drawn on the device from the rows' distributions, in row order (t > 0 / beam search)

    std::vector<wb200::Segment> result_all;
    std::vector<int32_t>        prompt_past;

    int lang_id = 0;
    wb200::TimestampState ts;
    int32_t exp_n_audio_ctx = 0;
    // spectrogram on the device (Forward::mel_on_device): `mel` then only carries its shape; the PCM is staged with the first encode
    const float * mel_pcm = nullptr;
    int     mel_pcm_n = 0;
    bool    mel_dev_ready = false;
    int     slot = 0;                   // device slot (cross-KV + self-KV cache) this state decodes against
};

struct whisper_context {
    int64_t t_load_us = 0, t_start_us = 0;
    whisper_context_params params{};

    wb200::HParams     hparams;
    wb200::MelFilters  filters;
    wb200::Vocab       vocab;
    wb200::LogitsRules rules;
    int                n_loaded = 0;

    std::unique_ptr<wb200::Forward> fwd;
    std::unique_ptr<wb200::Batcher> batcher;
    whisper_state * state = nullptr;

    // results of whisper_b200_full_batch, one state per chunk
    std::vector<std::unique_ptr<whisper_state>> chunk_states;
    // ... and the threads that run them, kept between calls (worker_pool.h)
    std::unique_ptr<wb200::WorkerPool> pool;
    // whisper_b200_init_multi: replicas of the model on further devices (this context is the one on devices[0]); whisper_b200_full_batch
    // deals chunk i to replica i mod n — independent chunks, no exchange step (whisper_full_parallel, whisper.cpp:5817-5930, across GPUs)
    std::vector<whisper_context *> peers;
};

namespace wb200 {

whisper_state * new_state(const whisper_context & ctx);

// whisper_encode_internal / whisper_decode_internal equivalents (whisper.cpp:2086, 2517)
bool encode_internal(whisper_context & ctx, whisper_state & state, int mel_offset,
                     whisper_abort_callback abort_cb, void * abort_ud);
bool decode_internal(whisper_context & ctx, whisper_state & state, const Batch & batch,
                     whisper_abort_callback abort_cb, void * abort_ud);

int lang_auto_detect(whisper_context & ctx, whisper_state & state, int offset_ms, float * lang_probs);

// whisper_full_with_state (whisper.cpp:4960-5807)
int full_with_state(whisper_context & ctx, whisper_state & state, whisper_full_params params,
                    const float * samples, int n_samples);

}  // namespace wb200
