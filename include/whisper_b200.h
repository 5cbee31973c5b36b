/*
 * whisper_b200.h — C ABI of the B200-native Whisper backend (libwhisper_b200.so).
 *
 * DROP-IN BOUNDARY.  godot-whisper's GDExtension host (src/speech_to_text.cpp, src/register_types.cpp) calls eleven
 * functions of thirdparty/whisper.cpp/whisper.h and nothing else.  This header declares those eleven entry points with
 * the SAME names, argument order, by-value struct layouts and return conventions, so that the unmodified host — compiled
 * against its own <whisper.cpp/whisper.h> — links against this library instead of the vendored whisper.cpp objects.
 * Each declaration cites the reference interface it replaces (paths relative to /root/reference).
 *
 * Layout contract (checked by static_asserts in csrc/api.cpp and by tests/test_abi.py against the compiled reference):
 *   sizeof(whisper_context_params) == 1, sizeof(whisper_token_data) == 48, sizeof(whisper_full_params) == 256.
 *
 * Everything behind these calls runs on the GPU (sm_100a); if no CUDA device / kernel image is usable,
 * whisper_init_from_buffer_with_params() logs an error and returns NULL.  There is no CPU fallback.
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#  define WHISPER_B200_API __attribute__((visibility("default")))
#else
#  define WHISPER_B200_API
#endif

/* thirdparty/whisper.cpp/whisper.h:32-35 */
#define WHISPER_SAMPLE_RATE 16000
#define WHISPER_N_FFT       400
#define WHISPER_HOP_LENGTH  160
#define WHISPER_CHUNK_SIZE  30

#ifdef __cplusplus
extern "C" {
#endif

/* ---- types shared with the host -------------------------------------------------------------------------------- */

struct whisper_context;   /* opaque; owns weights in HBM, device arenas, streams, the decode state and the results */
struct whisper_state;     /* opaque; only ever seen by callbacks */

typedef int32_t whisper_pos;
typedef int32_t whisper_token;
typedef int32_t whisper_seq_id;

/* thirdparty/whisper.cpp/ggml.h:486-491 — log levels handed to the callback */
enum ggml_log_level {
    GGML_LOG_LEVEL_ERROR = 2,
    GGML_LOG_LEVEL_WARN  = 3,
    GGML_LOG_LEVEL_INFO  = 4,
    GGML_LOG_LEVEL_DEBUG = 5
};

/* thirdparty/whisper.cpp/ggml.h:1914 */
typedef void (*ggml_log_callback)(enum ggml_log_level level, const char * text, void * user_data);

/* thirdparty/whisper.cpp/whisper.h:87-89.  use_gpu == false is refused (this backend IS the GPU path). */
struct whisper_context_params {
    bool use_gpu;
};

/* thirdparty/whisper.cpp/whisper.h:91-106 (48 bytes, returned by value) */
typedef struct whisper_token_data {
    whisper_token id;     /* token id */
    whisper_token tid;    /* most probable timestamp token */
    float   p;            /* probability of id */
    float   plog;         /* log-probability of id */
    float   pt;           /* probability of tid */
    float   ptsum;        /* probability mass of all timestamp tokens */
    int64_t t0;           /* token start, 10 ms units (token_timestamps) */
    int64_t t1;           /* token end */
    float   vlen;         /* voice-length heuristic of the token text */
} whisper_token_data;

/* thirdparty/whisper.cpp/whisper.h:111-133 — only the shape is needed: the grammar sampler is out of scope
 * (SURVEY.md §2.1 row 19); a non-NULL grammar_rules makes whisper_full() fail with -9 instead of being ignored. */
enum whisper_gretype {
    WHISPER_GRETYPE_END = 0, WHISPER_GRETYPE_ALT = 1, WHISPER_GRETYPE_RULE_REF = 2, WHISPER_GRETYPE_CHAR = 3,
    WHISPER_GRETYPE_CHAR_NOT = 4, WHISPER_GRETYPE_CHAR_RNG_UPPER = 5, WHISPER_GRETYPE_CHAR_ALT = 6
};
typedef struct whisper_grammar_element {
    enum whisper_gretype type;
    uint32_t             value;
} whisper_grammar_element;

/* thirdparty/whisper.cpp/whisper.h:395-399 */
enum whisper_sampling_strategy {
    WHISPER_SAMPLING_GREEDY      = 0,
    WHISPER_SAMPLING_BEAM_SEARCH = 1
};

/* thirdparty/whisper.cpp/whisper.h:401-428 — callback signatures (the godot host leaves all of them NULL) */
typedef void (*whisper_new_segment_callback)  (struct whisper_context *, struct whisper_state *, int n_new, void * user_data);
typedef void (*whisper_progress_callback)     (struct whisper_context *, struct whisper_state *, int progress, void * user_data);
typedef bool (*whisper_encoder_begin_callback)(struct whisper_context *, struct whisper_state *, void * user_data);
typedef bool (*whisper_abort_callback)        (void * user_data);
typedef void (*whisper_logits_filter_callback)(struct whisper_context *, struct whisper_state *,
                                               const whisper_token_data * tokens, int n_tokens,
                                               float * logits, void * user_data);

/* thirdparty/whisper.cpp/whisper.h:433-526 — passed BY VALUE (256 bytes); field order and padding are the ABI. */
struct whisper_full_params {
    enum whisper_sampling_strategy strategy;

    int  n_threads;          /* host threads for the log-mel front end */
    int  n_max_text_ctx;
    int  offset_ms;
    int  duration_ms;

    bool translate;
    bool no_context;
    bool no_timestamps;
    bool single_segment;
    bool print_special;
    bool print_progress;
    bool print_realtime;
    bool print_timestamps;

    bool  token_timestamps;
    float thold_pt;
    float thold_ptsum;
    int   max_len;
    bool  split_on_word;
    int   max_tokens;

    bool speed_up;           /* true => whisper_full returns -1 (whisper.cpp:4973-4976) */
    bool debug_mode;
    int  audio_ctx;          /* 0 = model default (1500); > n_audio_ctx => -5 */

    bool tdrz_enable;

    const char *          initial_prompt;
    const whisper_token * prompt_tokens;
    int                   prompt_n_tokens;

    const char * language;   /* "en", ..., or NULL / "" / "auto" for detection */
    bool         detect_language;

    bool suppress_blank;
    bool suppress_non_speech_tokens;

    float temperature;
    float max_initial_ts;
    float length_penalty;

    float temperature_inc;
    float entropy_thold;
    float logprob_thold;
    float no_speech_thold;

    struct { int best_of; } greedy;
    struct { int beam_size; float patience; } beam_search;

    whisper_new_segment_callback   new_segment_callback;
    void *                         new_segment_callback_user_data;
    whisper_progress_callback      progress_callback;
    void *                         progress_callback_user_data;
    whisper_encoder_begin_callback encoder_begin_callback;
    void *                         encoder_begin_callback_user_data;
    whisper_abort_callback         abort_callback;
    void *                         abort_callback_user_data;
    whisper_logits_filter_callback logits_filter_callback;
    void *                         logits_filter_callback_user_data;

    const whisper_grammar_element ** grammar_rules;
    size_t                           n_grammar_rules;
    size_t                           i_start_rule;
    float                            grammar_penalty;
};

/* ---- the eleven entry points the GDExtension host binds --------------------------------------------------------- */

/* replaces whisper.h:151 / whisper.cpp:3286; caller: src/speech_to_text.cpp:342-345.
 * Parses a ggml-format Whisper file held in `buffer` (caller-owned, may be freed after return), uploads the
 * tensors to HBM, allocates the device arenas / KV caches.  NULL on any failure. */
WHISPER_B200_API struct whisper_context * whisper_init_from_buffer_with_params(
        void * buffer, size_t buffer_size, struct whisper_context_params params);

/* replaces whisper.h:205 / whisper.cpp:3373; callers: src/speech_to_text.cpp:332,349.  NULL-safe. */
WHISPER_B200_API void whisper_free(struct whisper_context * ctx);

/* replaces whisper.h:391 / whisper.cpp:3851; caller: src/speech_to_text.cpp:334.  Static storage. */
WHISPER_B200_API const char * whisper_print_system_info(void);

/* replaces whisper.h:532 / whisper.cpp:4311-4410; caller: src/speech_to_text.cpp:403. */
WHISPER_B200_API struct whisper_full_params whisper_full_default_params(enum whisper_sampling_strategy strategy);

/* replaces whisper.h:537-541 / whisper.cpp:5809,4960; caller: src/speech_to_text.cpp:419.
 * PCM (16 kHz mono f32) -> host log-mel -> GPU encoder -> GPU decoder token loop -> segments.
 * Returns 0, or the reference's negative codes (-1 speed_up, -2 mel, -3 language, -4 decoders, -5 audio_ctx,
 * -6 encode, -7 prompt decode, -8 decode) plus -9 for the unsupported grammar sampler. */
WHISPER_B200_API int whisper_full(struct whisper_context * ctx, struct whisper_full_params params,
                                  const float * samples, int n_samples);

/* replaces whisper.h:564 / whisper.cpp:5940; caller: src/speech_to_text.cpp:424 */
WHISPER_B200_API int whisper_full_n_segments(struct whisper_context * ctx);
/* replaces whisper.h:589 / whisper.cpp:5988; caller: src/speech_to_text.cpp:427 */
WHISPER_B200_API int whisper_full_n_tokens(struct whisper_context * ctx, int i_segment);
/* replaces whisper.h:585 / whisper.cpp:5980; caller: src/speech_to_text.cpp:428.  Valid until the next whisper_full/free. */
WHISPER_B200_API const char * whisper_full_get_segment_text(struct whisper_context * ctx, int i_segment);
/* replaces whisper.h:593 / whisper.cpp:5996; caller: src/speech_to_text.cpp:432 */
WHISPER_B200_API const char * whisper_full_get_token_text(struct whisper_context * ctx, int i_segment, int i_token);
/* replaces whisper.h:601 / whisper.cpp:6012; caller: src/speech_to_text.cpp:431 */
WHISPER_B200_API whisper_token_data whisper_full_get_token_data(struct whisper_context * ctx, int i_segment, int i_token);

/* replaces whisper.h:619 / whisper.cpp:6601; caller: src/register_types.cpp:58.  Global; NULL restores stderr. */
WHISPER_B200_API void whisper_log_set(ggml_log_callback log_callback, void * user_data);

/* ---- per-stage entry points with the reference's semantics (parity hooks; also what examples/bench drives) ------ */

WHISPER_B200_API struct whisper_context_params whisper_context_default_params(void);              /* whisper.h:529 */
WHISPER_B200_API int64_t whisper_full_get_segment_t0(struct whisper_context * ctx, int i_segment); /* whisper.h:575 */
WHISPER_B200_API int64_t whisper_full_get_segment_t1(struct whisper_context * ctx, int i_segment); /* whisper.h:578 */
WHISPER_B200_API whisper_token whisper_full_get_token_id(struct whisper_context * ctx, int i_segment, int i_token); /* whisper.h:596 */
WHISPER_B200_API int whisper_full_lang_id(struct whisper_context * ctx);                           /* whisper.h:568 */

/* whisper.h:223 / whisper.cpp:3412 — host log-mel of `samples` into the context */
WHISPER_B200_API int whisper_pcm_to_mel(struct whisper_context * ctx, const float * samples, int n_samples, int n_threads);
/* whisper.h:246 / whisper.cpp:3463 — install a caller-computed log-mel [n_mel][n_len] */
WHISPER_B200_API int whisper_set_mel(struct whisper_context * ctx, const float * data, int n_len, int n_mel);
/* whisper.h:263 / whisper.cpp:3481 — conv stem + encoder + cross-KV on the GPU for the window starting at frame `offset` */
WHISPER_B200_API int whisper_encode(struct whisper_context * ctx, int offset, int n_threads);
/* whisper.h:280 / whisper.cpp:3503 — decoder pass for sequence 0; logits of the LAST row via whisper_get_logits */
WHISPER_B200_API int whisper_decode(struct whisper_context * ctx, const whisper_token * tokens, int n_tokens, int n_past, int n_threads);
/* whisper.h:364 / whisper.cpp:3741 — [n_tokens][n_vocab]; only rows that were requested are defined */
WHISPER_B200_API float * whisper_get_logits(struct whisper_context * ctx);
/* whisper.h:295 / whisper.cpp:3507 */
WHISPER_B200_API int whisper_tokenize(struct whisper_context * ctx, const char * text, whisper_token * tokens, int n_max_tokens);
/* whisper.h:308-323 / whisper.cpp:3523-3558 */
WHISPER_B200_API int          whisper_lang_max_id(void);
WHISPER_B200_API int          whisper_lang_id(const char * lang);
WHISPER_B200_API const char * whisper_lang_str(int id);
/* whisper.h:330-336 / whisper.cpp:3569,3644 */
WHISPER_B200_API int whisper_lang_auto_detect(struct whisper_context * ctx, int offset_ms, int n_threads, float * lang_probs);

/* whisper.h:338-353 / whisper.cpp:3717-3739 */
WHISPER_B200_API int whisper_n_len(struct whisper_context * ctx);
WHISPER_B200_API int whisper_n_vocab(struct whisper_context * ctx);
WHISPER_B200_API int whisper_n_text_ctx(struct whisper_context * ctx);
WHISPER_B200_API int whisper_n_audio_ctx(struct whisper_context * ctx);
WHISPER_B200_API int whisper_is_multilingual(struct whisper_context * ctx);
WHISPER_B200_API int whisper_model_n_audio_state(struct whisper_context * ctx);
WHISPER_B200_API int whisper_model_n_audio_head(struct whisper_context * ctx);
WHISPER_B200_API int whisper_model_n_audio_layer(struct whisper_context * ctx);
WHISPER_B200_API int whisper_model_n_text_layer(struct whisper_context * ctx);

/* whisper.h:367-384 / whisper.cpp:3749-3791 */
WHISPER_B200_API const char *  whisper_token_to_str(struct whisper_context * ctx, whisper_token token);
WHISPER_B200_API whisper_token whisper_token_eot(struct whisper_context * ctx);
WHISPER_B200_API whisper_token whisper_token_sot(struct whisper_context * ctx);
WHISPER_B200_API whisper_token whisper_token_solm(struct whisper_context * ctx);
WHISPER_B200_API whisper_token whisper_token_prev(struct whisper_context * ctx);
WHISPER_B200_API whisper_token whisper_token_nosp(struct whisper_context * ctx);
WHISPER_B200_API whisper_token whisper_token_not(struct whisper_context * ctx);
WHISPER_B200_API whisper_token whisper_token_beg(struct whisper_context * ctx);
WHISPER_B200_API whisper_token whisper_token_lang(struct whisper_context * ctx, int lang_id);
WHISPER_B200_API whisper_token whisper_token_translate(struct whisper_context * ctx);
WHISPER_B200_API whisper_token whisper_token_transcribe(struct whisper_context * ctx);

/* whisper.h:387-388 / whisper.cpp:3793-3832 — phase timers and fallback counters */
WHISPER_B200_API void whisper_print_timings(struct whisper_context * ctx);
WHISPER_B200_API void whisper_reset_timings(struct whisper_context * ctx);

/* ---- additive B200 entry points (no reference counterpart) -------------------------------------------------------- */

/* Independent chunks as one data-parallel batch (the CPU analogue is whisper_full_parallel, whisper.cpp:5817-5930:
 * one state per chunk, shared read-only weights).  Runs whisper_full() semantics on n_chunks PCM buffers; the
 * encoder is batched across chunks on the device.  Results are read per chunk with the whisper_b200_chunk_*
 * accessors.  Returns 0 or the first non-zero per-chunk code. */
WHISPER_B200_API int whisper_b200_full_batch(struct whisper_context * ctx, struct whisper_full_params params,
                                             const float * const * samples, const int * n_samples, int n_chunks);
/* Several GPUs of one box behind ONE context, for a host that is a single process (the GDExtension is): loads a replica of the model on
 * each of devices[0..n_devices) and returns the context of devices[0]; whisper_b200_full_batch on it deals chunk i to replica
 * i mod n_devices (independent chunks, no exchange step — whisper_full_parallel, whisper.cpp:5817-5930, across GPUs instead of CPU
 * threads) and the chunk accessors below see all results in the caller's order.  NULL if any device fails.  whisper_free frees all. */
WHISPER_B200_API struct whisper_context * whisper_b200_init_multi(void * buffer, size_t buffer_size, struct whisper_context_params params,
                                                                  const int * devices, int n_devices);
WHISPER_B200_API int          whisper_b200_n_devices(struct whisper_context * ctx);
/* Loader unit-test hook: n elements (a multiple of 32) of block-quantised ggml data (type ids of ggml.h:327-333: Q4_0 2, Q4_1 3, Q5_0 6,
 * Q5_1 7, Q8_0 8) -> f32, the values ggml-quants.c dequantize_row_* yields.  The model loader expands such matrices to f16 with it. */
WHISPER_B200_API int whisper_b200_dequantize(int ggml_type, const void * blocks, long long n, float * out);

/* Page-locked host memory for PCM: whisper_full / whisper_b200_full_batch upload samples that lie in such memory straight from there
 * (any page-locked memory is recognised, e.g. cudaHostRegister'ed by the host); pageable samples are first copied to an internal
 * staging buffer.  The analogue on the reference side is none: its whisper_full reads the samples on the CPU. */
WHISPER_B200_API void * whisper_b200_host_alloc(size_t bytes);
WHISPER_B200_API void   whisper_b200_host_free(void * p);
WHISPER_B200_API int          whisper_b200_chunk_n_segments(struct whisper_context * ctx, int i_chunk);
WHISPER_B200_API int          whisper_b200_chunk_n_tokens(struct whisper_context * ctx, int i_chunk, int i_segment);
WHISPER_B200_API const char * whisper_b200_chunk_segment_text(struct whisper_context * ctx, int i_chunk, int i_segment);
WHISPER_B200_API whisper_token_data whisper_b200_chunk_token_data(struct whisper_context * ctx, int i_chunk, int i_segment, int i_token);
/* All token ids of a chunk (segments concatenated) in one call: writes min(count, cap) ids to out, returns the count.  The bulk form of
 * the per-token loop of src/speech_to_text.cpp:424-445 for hosts that marshal many chunks. */
WHISPER_B200_API int          whisper_b200_chunk_token_ids(struct whisper_context * ctx, int i_chunk, whisper_token * out, int cap);

/* Device selection for the NEXT whisper_init_* call on this thread (default: current CUDA device, else 0). */
WHISPER_B200_API void whisper_b200_set_device(int device);

/* Counters: out[0..6] = n_sample, n_encode, n_decode, n_batchd, n_prompt, n_fail_p, n_fail_h (whisper.cpp:770-783);
 * out[7] = kernels launched by this context so far (graph nodes counted individually). */
WHISPER_B200_API void whisper_b200_counters(struct whisper_context * ctx, int64_t * out8);
/* Phase times accumulated like the reference's t_*_us: out[0..5] = mel, sample, encode, decode, batchd, prompt (us). */
WHISPER_B200_API void whisper_b200_timings_us(struct whisper_context * ctx, int64_t * out6);

/* Device clocks (CUDA events on the launching stream), accumulated since init: out[0] = ms inside encoder passes (conv stem to cross K / V; the spectrogram stage in front is whisper_b200_gpu_mel_ms),
 * out[1] = ms inside decoder passes, out[2] / out[3] = number of encoder / decoder passes, out[4] / out[5] = bytes copied
 * host->device / device->host by those passes, out[6] = launches of the persistent decode-step kernel, out[7] = their
 * algorithmic bytes (decoder weights once per launch + cross-attention K/V of every row). */
WHISPER_B200_API void whisper_b200_gpu_times(struct whisper_context * ctx, double * out8);
/* Device-busy milliseconds since the context was created: the union of the intervals of all encoder / decoder passes (they overlap
 * on two streams).  Call between whisper_full / whisper_b200_full_batch calls. */
WHISPER_B200_API double whisper_b200_gpu_busy_ms(struct whisper_context * ctx);
/* Milliseconds the encoder passes spent in their spectrogram stage (device log-mel, energy envelope, window staging: the device side of
 * log_mel_spectrogram, whisper.cpp:2727-2887, and of the energy pass of whisper.cpp:6350-6366).  whisper_b200_gpu_times()[0] starts where
 * this ends — the same split as the reference's t_mel_us / t_encode_us timers (whisper.cpp:3793-3815). */
WHISPER_B200_API double whisper_b200_gpu_mel_ms(struct whisper_context * ctx);
/* Per-kernel-class profile: while enabled every launch is bracketed by an event pair.  whisper_b200_profile fills
 * out[9][4] = {launches, total ms, algorithmic FLOP, algorithmic bytes} for the classes
 * 0 encoder GEMM (tcgen05), 1 encoder attention GEMMs (tcgen05), 2 softmax, 3 LayerNorm, 4 skinny GEMM (multi-kernel decode),
 * 5 decoder attention, 6 misc (embed / gather / mel transpose), 7 decoder GEMM on tcgen05 (rows > 32),
 * 8 persistent decode-step kernel (one launch = one whole token step). */
WHISPER_B200_API void whisper_b200_set_profiling(struct whisper_context * ctx, int on);
WHISPER_B200_API void whisper_b200_profile(struct whisper_context * ctx, double * out36);

/* Stage tensors for parity tests (device -> host copies; same layouts as the reference's ggml tensors):
 *   what = 0: mel window fed to the conv stem  f32 [n_mels][2*n_ctx]
 *          1: conv stem output                 f32 [n_ctx][n_state]  (token-major; the reference holds its transpose)
 *          2: encoder output after ln_post     f32 [n_ctx][n_state]
 *          3: cross-attention K                f16 [n_text_layer][n_ctx][n_state]
 *          4: cross-attention V (transposed)   f16 [n_text_layer][n_state][n_ctx]
 *          5: self-attention  K                f16 [n_text_layer][kv_size][n_state]
 *          6: self-attention  V (transposed)   f16 [n_text_layer][n_state][kv_size]
 *          7: host log-mel                     f32 [n_mels][n_len]   (whisper_pcm_to_mel / whisper_set_mel, csrc/mel.cpp)
 *          9: device log-mel of the last whisper_full on a clip of <= 30 s, normalised, f32 [n_mels][n_len]   (cuda/mel_kernels.cu)
 * Returns the number of BYTES of the tensor; copies min(cap_bytes, that) bytes when dst != NULL. */
WHISPER_B200_API long long whisper_b200_read_stage(struct whisper_context * ctx, int what, void * dst, long long cap_bytes);

/* Engine selection: 0 = tcgen05/TMA GEMMs + persistent decode-step kernel (default), 1 = SIMT reference GEMMs (debug only),
 * 2 = tcgen05/TMA GEMMs with every decode step on the multi-kernel path (cross-check of the step kernel). */
WHISPER_B200_API void whisper_b200_set_gemm_engine(struct whisper_context * ctx, int engine);

/* Stand-alone f16 GEMM  C[n][m] (f32) = sum_k A[m][k] * B[n][k]  on device buffers of this context's device, used by
 * the kernel unit tests and by bench.py's roofline probe.  engine as above.  Returns 0 on success. */
WHISPER_B200_API int whisper_b200_gemm_f16(const void * A_host_f16, const void * B_host_f16, float * C_host,
                                           int M, int N, int K, int engine, int iters, float * ms_per_iter);

/* Kernel unit-test hook for the encoder GEMM with the TMA-store epilogue (csrc/cuda/gemm_enc.cu), on host buffers:
 * act f16 [N][K], wgt f16 [M][K], bias f32 [M] or NULL.  mode 0: out f16 [N][M] = acc + bias; 1: ... through the GELU table
 * (ggml.c:1416-1423); 2: out f16 [M][ldt] transposed, ldt = N rounded up to 8 (V^T layouts, whisper.cpp:1903-1909); 3: out f32 [N][M] =
 * acc + bias + res (res f32 [N][M]; whisper.cpp:1922-1930); 4: three feature segments, scale 0.25.  Returns 0 or a negative code. */
WHISPER_B200_API int whisper_b200_gemm_enc_probe(const void * act_f16, const void * wgt_f16, const float * bias, const float * res, void * out,
                                                 int N, int M, int K, int mode, int iters, float * ms_per_iter);

/* Host front of the realtime path: the GDExtension's end-of-speech test on the most recent audio window
 * (SpeechToText::voice_activity_detection, src/speech_to_text.cpp:378-399).  whisper_b200_high_pass_filter replaces _high_pass_filter
 * (src/speech_to_text.cpp:53-64): its recursion in place, including the reference's read of the already-overwritten previous element.  whisper_b200_vad_simple replaces _vad_simple (:67-104):
 * filters `pcmf32` in place when freq_thold > 0 (like the reference), then compares the mean |x| of the last `last_ms` with that of the
 * whole window; returns what the reference's bool returns (1 / 0).  Bit-identical with the reference on the same input. */
WHISPER_B200_API void whisper_b200_high_pass_filter(float * data, int n_samples, float cutoff, float sample_rate);
WHISPER_B200_API int whisper_b200_vad_simple(float * pcmf32, int n_samples, int sample_rate, int last_ms, float vad_thold, float freq_thold);

/* Test / bench hook: the fused encoder attention (csrc/cuda/attn_enc.cu; replaces the KQ mul_mat -> scale -> soft_max -> V mul_mat chain of
 * whisper.cpp:1880-1917) on host buffers.  q, k f16 [B][T][d] (head h = columns 64 h .. 64 h + 63), vt f16 [B][d][Tp] (V transposed,
 * Tp = T rounded up to 8), out f16 [B][T][d].  variant: kernel configuration (< 0 = the one the encoder uses).  Returns 0 or a negative code. */
WHISPER_B200_API int whisper_b200_attn_enc_probe(const void * q_f16, const void * k_f16, const void * vt_f16, void * out_f16, int B, int T, int d,
                                                 int n_head, int variant, int iters, float * ms_per_iter);

/* The two f16 activation tables (GELU, exp) the kernels index, as built on the host — ggml.c:2218-2236 semantics.
 * 65536 entries each.  Test hook: lets a CPU-only test compare them with the reference's tables. */
WHISPER_B200_API void whisper_b200_f16_tables(uint16_t * gelu_f16, uint16_t * exp_f16);

#ifdef __cplusplus
}
#endif

#endif /* WHISPER_B200_H */
