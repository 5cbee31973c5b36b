// ggml-format Whisper model file parser (host side of whisper_init_from_buffer_with_params).
//
// Restates the on-disk format read by the reference at thirdparty/whisper.cpp/whisper.cpp:1102-1640 and written by
// thirdparty/whisper.cpp/models/convert-pt-to-ggml.py:268-339.  The parser works IN PLACE on the
// caller's buffer: tensors are described by (pointer, shape, type) and uploaded to HBM straight from that buffer.
#pragma once

#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace wb200 {

struct HParams {            // whisper.cpp:537-550
    int32_t n_vocab       = 51864;
    int32_t n_audio_ctx   = 1500;
    int32_t n_audio_state = 384;
    int32_t n_audio_head  = 6;
    int32_t n_audio_layer = 4;
    int32_t n_text_ctx    = 448;
    int32_t n_text_state  = 384;
    int32_t n_text_head   = 6;
    int32_t n_text_layer  = 4;
    int32_t n_mels        = 80;
    int32_t ftype         = 1;
    float   eps           = 1e-5f;
};

struct MelFilters {         // whisper.cpp:357-362
    int32_t n_mel = 0;
    int32_t n_fft = 0;
    std::vector<float> data;  // [n_mel][n_fft]
};

struct Vocab {              // whisper.cpp:364-394
    int n_vocab = 51864;
    std::map<std::string, int32_t> token_to_id;
    std::vector<std::string>       id_to_token;   // dense, size n_vocab (the reference uses a map with the same keys)

    int32_t token_eot        = 50256;
    int32_t token_sot        = 50257;
    int32_t token_translate  = 50357;
    int32_t token_transcribe = 50358;
    int32_t token_solm       = 50359;
    int32_t token_prev       = 50360;
    int32_t token_nosp       = 50361;
    int32_t token_not        = 50362;
    int32_t token_beg        = 50363;

    bool is_multilingual() const { return n_vocab >= 51865; }
    int  num_languages()   const { return n_vocab - 51765 - (is_multilingual() ? 1 : 0); }
    int32_t token_lang(int lang_id) const { return token_sot + 1 + lang_id; }   // whisper.cpp:3781
};

enum TensorType { TT_F32 = 0, TT_F16 = 1 };

struct TensorView {
    const uint8_t * data = nullptr;  // into the caller's buffer (valid only during init)
    int       n_dims = 0;
    int32_t   ne[4]  = {1, 1, 1, 1}; // ggml order: ne[0] innermost
    TensorType type  = TT_F32;
    size_t    nbytes = 0;
    int64_t   nelements() const { return (int64_t) ne[0] * ne[1] * ne[2] * ne[3]; }
};

// Block-quantised ggml tensor types the loader accepts (ggml.h:327-333) and turns into f16 at load time.
enum { GGML_T_F32 = 0, GGML_T_F16 = 1, GGML_T_Q4_0 = 2, GGML_T_Q4_1 = 3, GGML_T_Q5_0 = 6, GGML_T_Q5_1 = 7, GGML_T_Q8_0 = 8 };
// bytes of one 32-element block of `ggml_type`, 0 if the type is not block-quantised
size_t quant_block_bytes(int ggml_type);
// n elements (a multiple of 32) of block-quantised data -> f32, the arithmetic of ggml-quants.c dequantize_row_q{4_0,4_1,5_0,5_1,8_0}
bool dequantize_blocks(int ggml_type, const void * blocks, int64_t n, float * out);

struct ModelFile {
    HParams    hparams;
    MelFilters filters;
    Vocab      vocab;
    std::map<std::string, TensorView> tensors;  // by OpenAI state-dict name
    std::vector<std::vector<uint16_t>> owned;   // f16 images of tensors that were block-quantised in the file (TensorView::data points here)
    int        n_loaded   = 0;                  // 0 => weight-less "test model" (whisper.cpp:1627-1628)
    int        n_expected = 0;
    size_t     total_bytes = 0;
    const void * raw = nullptr;                 // the caller's buffer (valid only during init)
    size_t     raw_size = 0;
};

// Returns false (after logging why) on bad magic, unsupported ftype, unknown / mis-shaped / missing tensors.
bool parse_model_file(const void * buffer, size_t size, ModelFile & out);

// language table (whisper.cpp:246-347)
int          lang_max_id();
int          lang_id(const char * code_or_name);   // -1 if unknown
const char * lang_str(int id);                     // nullptr if unknown
const char * lang_str_full(int id);
int          lang_count();                         // 100

// Greedy longest-match tokenizer over GPT-2 style pre-split words (whisper.cpp:2899-2947).
std::vector<int32_t> tokenize(const Vocab & vocab, const std::string & text);

}  // namespace wb200
