// ggml Whisper file parser, vocabulary, language table, tokenizer.  See model.h for the reference citations.
#include "model.h"
#include "common.h"

#include <cstring>
#include <exception>
#include <regex>

namespace wb200 {

namespace {

constexpr uint32_t kGgmlMagic = 0x67676d6c;  // "ggml" (ggml.h:209)

struct Cursor {
    const uint8_t * p;
    size_t size;
    size_t off = 0;
    // same truncating semantics as the buffer loader at whisper.cpp:3232-3240
    size_t read(void * dst, size_t n) {
        const size_t m = off + n < size ? n : size - off;
        memcpy(dst, p + off, m);
        off += m;
        return m;
    }
    template <typename T> void get(T & v) { v = T(); read(&v, sizeof(T)); }
    bool eof() const { return off >= size; }
};

struct Lang { const char * code; const char * name; };
// ids are the array indices (whisper.cpp:246-347)
const Lang kLangs[] = {
    {"en","english"},{"zh","chinese"},{"de","german"},{"es","spanish"},{"ru","russian"},{"ko","korean"},
    {"fr","french"},{"ja","japanese"},{"pt","portuguese"},{"tr","turkish"},{"pl","polish"},{"ca","catalan"},
    {"nl","dutch"},{"ar","arabic"},{"sv","swedish"},{"it","italian"},{"id","indonesian"},{"hi","hindi"},
    {"fi","finnish"},{"vi","vietnamese"},{"he","hebrew"},{"uk","ukrainian"},{"el","greek"},{"ms","malay"},
    {"cs","czech"},{"ro","romanian"},{"da","danish"},{"hu","hungarian"},{"ta","tamil"},{"no","norwegian"},
    {"th","thai"},{"ur","urdu"},{"hr","croatian"},{"bg","bulgarian"},{"lt","lithuanian"},{"la","latin"},
    {"mi","maori"},{"ml","malayalam"},{"cy","welsh"},{"sk","slovak"},{"te","telugu"},{"fa","persian"},
    {"lv","latvian"},{"bn","bengali"},{"sr","serbian"},{"az","azerbaijani"},{"sl","slovenian"},{"kn","kannada"},
    {"et","estonian"},{"mk","macedonian"},{"br","breton"},{"eu","basque"},{"is","icelandic"},{"hy","armenian"},
    {"ne","nepali"},{"mn","mongolian"},{"bs","bosnian"},{"kk","kazakh"},{"sq","albanian"},{"sw","swahili"},
    {"gl","galician"},{"mr","marathi"},{"pa","punjabi"},{"si","sinhala"},{"km","khmer"},{"sn","shona"},
    {"yo","yoruba"},{"so","somali"},{"af","afrikaans"},{"oc","occitan"},{"ka","georgian"},{"be","belarusian"},
    {"tg","tajik"},{"sd","sindhi"},{"gu","gujarati"},{"am","amharic"},{"yi","yiddish"},{"lo","lao"},
    {"uz","uzbek"},{"fo","faroese"},{"ht","haitian creole"},{"ps","pashto"},{"tk","turkmen"},{"nn","nynorsk"},
    {"mt","maltese"},{"sa","sanskrit"},{"lb","luxembourgish"},{"my","myanmar"},{"bo","tibetan"},{"tl","tagalog"},
    {"mg","malagasy"},{"as","assamese"},{"tt","tatar"},{"haw","hawaiian"},{"ln","lingala"},{"ha","hausa"},
    {"ba","bashkir"},{"jw","javanese"},{"su","sundanese"},{"yue","cantonese"},
};
constexpr int kNumLangs = (int)(sizeof(kLangs) / sizeof(kLangs[0]));
static_assert(kNumLangs == 100, "language table");

void expect(ModelFile & mf, const std::string & name, TensorType type, int n_dims, int e0, int e1 = 1, int e2 = 1) {
    TensorView tv;
    tv.type = type;
    tv.n_dims = n_dims;
    tv.ne[0] = e0; tv.ne[1] = e1; tv.ne[2] = e2; tv.ne[3] = 1;
    tv.nbytes = (size_t) tv.nelements() * (type == TT_F16 ? 2 : 4);
    mf.tensors[name] = tv;
}

// The tensor set and shapes the loader insists on (whisper.cpp:1320-1510); ggml dim order (ne[0] innermost).
void declare_expected_tensors(ModelFile & mf) {
    const HParams & hp = mf.hparams;
    const int d_a = hp.n_audio_state, d_t = hp.n_text_state;
    const TensorType W = TT_F16;  // wtype / vtype for ftype == 1

    expect(mf, "encoder.positional_embedding", TT_F32, 2, d_a, hp.n_audio_ctx);
    expect(mf, "encoder.conv1.weight", W, 3, 3, hp.n_mels, d_a);
    expect(mf, "encoder.conv1.bias", TT_F32, 2, 1, d_a);
    expect(mf, "encoder.conv2.weight", W, 3, 3, d_a, d_a);
    expect(mf, "encoder.conv2.bias", TT_F32, 2, 1, d_a);
    expect(mf, "encoder.ln_post.weight", TT_F32, 1, d_a);
    expect(mf, "encoder.ln_post.bias", TT_F32, 1, d_a);
    for (int i = 0; i < hp.n_audio_layer; ++i) {
        const std::string p = "encoder.blocks." + std::to_string(i) + ".";
        expect(mf, p + "mlp_ln.weight", TT_F32, 1, d_a);
        expect(mf, p + "mlp_ln.bias", TT_F32, 1, d_a);
        expect(mf, p + "mlp.0.weight", W, 2, d_a, 4 * d_a);
        expect(mf, p + "mlp.0.bias", TT_F32, 1, 4 * d_a);
        expect(mf, p + "mlp.2.weight", W, 2, 4 * d_a, d_a);
        expect(mf, p + "mlp.2.bias", TT_F32, 1, d_a);
        expect(mf, p + "attn_ln.weight", TT_F32, 1, d_a);
        expect(mf, p + "attn_ln.bias", TT_F32, 1, d_a);
        expect(mf, p + "attn.query.weight", W, 2, d_a, d_a);
        expect(mf, p + "attn.query.bias", TT_F32, 1, d_a);
        expect(mf, p + "attn.key.weight", W, 2, d_a, d_a);
        expect(mf, p + "attn.value.weight", W, 2, d_a, d_a);
        expect(mf, p + "attn.value.bias", TT_F32, 1, d_a);
        expect(mf, p + "attn.out.weight", W, 2, d_a, d_a);
        expect(mf, p + "attn.out.bias", TT_F32, 1, d_a);
    }
    expect(mf, "decoder.positional_embedding", TT_F32, 2, d_t, hp.n_text_ctx);
    expect(mf, "decoder.token_embedding.weight", W, 2, d_t, hp.n_vocab);
    expect(mf, "decoder.ln.weight", TT_F32, 1, d_t);
    expect(mf, "decoder.ln.bias", TT_F32, 1, d_t);
    for (int i = 0; i < hp.n_text_layer; ++i) {
        const std::string p = "decoder.blocks." + std::to_string(i) + ".";
        expect(mf, p + "mlp_ln.weight", TT_F32, 1, d_t);
        expect(mf, p + "mlp_ln.bias", TT_F32, 1, d_t);
        expect(mf, p + "mlp.0.weight", W, 2, d_t, 4 * d_t);
        expect(mf, p + "mlp.0.bias", TT_F32, 1, 4 * d_t);
        expect(mf, p + "mlp.2.weight", W, 2, 4 * d_t, d_t);
        expect(mf, p + "mlp.2.bias", TT_F32, 1, d_t);
        for (const char * a : {"attn", "cross_attn"}) {
            const std::string q = p + a;
            expect(mf, q + "_ln.weight", TT_F32, 1, d_t);
            expect(mf, q + "_ln.bias", TT_F32, 1, d_t);
            expect(mf, q + ".query.weight", W, 2, d_t, d_t);
            expect(mf, q + ".query.bias", TT_F32, 1, d_t);
            expect(mf, q + ".key.weight", W, 2, d_t, d_t);
            expect(mf, q + ".value.weight", W, 2, d_t, d_t);
            expect(mf, q + ".value.bias", TT_F32, 1, d_t);
            expect(mf, q + ".out.weight", W, 2, d_t, d_t);
            expect(mf, q + ".out.bias", TT_F32, 1, d_t);
        }
    }
    mf.n_expected = (int) mf.tensors.size();
}

}  // namespace

size_t quant_block_bytes(int t) {
    switch (t) {
        case GGML_T_Q4_0: return 2 + 16;           // f16 d, 32 x 4 bit
        case GGML_T_Q4_1: return 2 + 2 + 16;       // f16 d, f16 m, 32 x 4 bit
        case GGML_T_Q5_0: return 2 + 4 + 16;       // f16 d, 32 high bits, 32 x 4 bit
        case GGML_T_Q5_1: return 2 + 2 + 4 + 16;   // f16 d, f16 m, 32 high bits, 32 x 4 bit
        case GGML_T_Q8_0: return 2 + 32;           // f16 d, 32 x int8
        default: return 0;
    }
}

// ggml-quants.c: dequantize_row_q4_0 / q4_1 / q5_0 / q5_1 / q8_0 (block layouts ggml-quants.h: block_q4_0 ... block_q8_0)
bool dequantize_blocks(int t, const void * blocks, int64_t n, float * y) {
    const size_t bb = quant_block_bytes(t);
    if (bb == 0 || n % 32 != 0) return false;
    const uint8_t * b = (const uint8_t *) blocks;
    for (int64_t i = 0; i < n / 32; ++i, b += bb, y += 32) {
        uint16_t dh, mh = 0;
        memcpy(&dh, b, 2);
        const float d = f16_to_f32(dh);
        const uint8_t * p = b + 2;
        float m = 0.0f;
        if (t == GGML_T_Q4_1 || t == GGML_T_Q5_1) { memcpy(&mh, p, 2); m = f16_to_f32(mh); p += 2; }
        if (t == GGML_T_Q8_0) {
            const int8_t * q = (const int8_t *) p;
            for (int j = 0; j < 32; ++j) y[j] = q[j] * d;
            continue;
        }
        uint32_t qh = 0;
        if (t == GGML_T_Q5_0 || t == GGML_T_Q5_1) { memcpy(&qh, p, 4); p += 4; }
        for (int j = 0; j < 16; ++j) {
            int x0 = p[j] & 0x0F, x1 = p[j] >> 4;
            if (t == GGML_T_Q5_0 || t == GGML_T_Q5_1) {
                x0 |= (int) (((qh >> (j + 0)) << 4) & 0x10);
                x1 |= (int) ((qh >> (j + 12)) & 0x10);
            }
            if (t == GGML_T_Q4_0)      { y[j] = (x0 - 8) * d;  y[j + 16] = (x1 - 8) * d; }
            else if (t == GGML_T_Q5_0) { y[j] = (x0 - 16) * d; y[j + 16] = (x1 - 16) * d; }
            else                       { y[j] = x0 * d + m;    y[j + 16] = x1 * d + m; }
        }
    }
    return true;
}

int lang_max_id() { return kNumLangs - 1; }
int lang_count() { return kNumLangs; }

int lang_id(const char * s) {
    if (!s) return -1;
    for (int i = 0; i < kNumLangs; ++i) if (strcmp(kLangs[i].code, s) == 0) return i;
    for (int i = 0; i < kNumLangs; ++i) if (strcmp(kLangs[i].name, s) == 0) return i;
    WB_LOG_ERROR("%s: unknown language '%s'\n", __func__, s);
    return -1;
}

const char * lang_str(int id) {
    if (id >= 0 && id < kNumLangs) return kLangs[id].code;
    WB_LOG_ERROR("%s: unknown language id %d\n", __func__, id);
    return nullptr;
}

const char * lang_str_full(int id) {
    if (id >= 0 && id < kNumLangs) return kLangs[id].name;
    WB_LOG_ERROR("%s: unknown language id %d\n", __func__, id);
    return nullptr;
}

static bool parse_model_file_impl(const void * buffer, size_t size, ModelFile & mf) {
    Cursor c{(const uint8_t *) buffer, size};
    mf.raw = buffer;
    mf.raw_size = size;

    uint32_t magic = 0;
    c.get(magic);
    if (magic != kGgmlMagic) {
        WB_LOG_ERROR("%s: invalid model data (bad magic)\n", __func__);
        return false;
    }

    HParams & hp = mf.hparams;
    c.get(hp.n_vocab); c.get(hp.n_audio_ctx); c.get(hp.n_audio_state); c.get(hp.n_audio_head); c.get(hp.n_audio_layer);
    c.get(hp.n_text_ctx); c.get(hp.n_text_state); c.get(hp.n_text_head); c.get(hp.n_text_layer); c.get(hp.n_mels);
    c.get(hp.ftype);

    const int qntvr = hp.ftype / 1000;  // GGML_QNT_VERSION_FACTOR (ggml.h:213)
    hp.ftype %= 1000;
    if (hp.ftype != 1 && hp.ftype != 2 && hp.ftype != 3 && hp.ftype != 7 && hp.ftype != 8 && hp.ftype != 9) {
        // f16, and the block-quantised files of whisper.cpp's quantize tool (Q4_0 / Q4_1 / Q8_0 / Q5_0 / Q5_1, ggml.h:364-370), whose
        // matrices are expanded to f16 while loading; f32 files and the k-quants are refused cleanly instead of being mis-read
        WB_LOG_ERROR("%s: unsupported model ftype %d (qntvr %d): f16 and Q4_0 / Q4_1 / Q5_0 / Q5_1 / Q8_0 ggml models are supported\n",
                     __func__, hp.ftype, qntvr);
        return false;
    }
    if (hp.ftype != 1 && qntvr != 1 && qntvr != 2) {                       // GGML_QNT_VERSION 2 (ggml.h:210); 1 shares these block layouts
        WB_LOG_ERROR("%s: unsupported quantisation format version %d\n", __func__, qntvr);
        return false;
    }
    // every divisor is tested before it is used, and every count that sizes a loop or an allocation below is bounded (the largest
    // released model has 32 layers, 1280 features, 51 866 tokens, 128 mel bands): a crafted header must end in `return false`, never in
    // SIGFPE, a 2^31-iteration loop or bad_alloc across the C ABI
    if (hp.n_audio_state <= 0 || hp.n_text_state <= 0 || hp.n_audio_head <= 0 || hp.n_text_head <= 0 ||
        hp.n_audio_state != hp.n_text_state || hp.n_audio_state > 8192 ||
        hp.n_audio_state % hp.n_audio_head != 0 || hp.n_text_state % hp.n_text_head != 0 ||
        hp.n_audio_state / hp.n_audio_head != 64 || hp.n_text_state / hp.n_text_head != 64 ||
        hp.n_vocab <= 0 || hp.n_vocab > (1 << 20) || hp.n_audio_ctx <= 0 || hp.n_audio_ctx > 1 << 16 ||
        hp.n_text_ctx <= 0 || hp.n_text_ctx > 1 << 16 || hp.n_mels <= 0 || hp.n_mels > 1024 ||
        hp.n_audio_layer <= 0 || hp.n_audio_layer > 64 || hp.n_text_layer <= 0 || hp.n_text_layer > 64) {
        WB_LOG_ERROR("%s: invalid model hyper-parameters (state %d/%d, heads %d/%d)\n", __func__,
                     hp.n_audio_state, hp.n_text_state, hp.n_audio_head, hp.n_text_head);
        return false;
    }

    WB_LOG_INFO("%s: n_vocab = %d, n_audio_ctx = %d, n_audio_state = %d, n_audio_head = %d, n_audio_layer = %d\n",
                __func__, hp.n_vocab, hp.n_audio_ctx, hp.n_audio_state, hp.n_audio_head, hp.n_audio_layer);
    WB_LOG_INFO("%s: n_text_ctx = %d, n_text_state = %d, n_text_head = %d, n_text_layer = %d, n_mels = %d, ftype = %d\n",
                __func__, hp.n_text_ctx, hp.n_text_state, hp.n_text_head, hp.n_text_layer, hp.n_mels, hp.ftype);

    // mel filter bank
    {
        MelFilters & f = mf.filters;
        c.get(f.n_mel);
        c.get(f.n_fft);
        if (f.n_mel <= 0 || f.n_fft <= 0 || (size_t) f.n_mel * f.n_fft * 4 > size) {
            WB_LOG_ERROR("%s: invalid mel filter header (%d x %d)\n", __func__, f.n_mel, f.n_fft);
            return false;
        }
        f.data.resize((size_t) f.n_mel * f.n_fft);
        c.read(f.data.data(), f.data.size() * sizeof(float));
    }

    // vocabulary (whisper.cpp:1205-1291)
    {
        Vocab & v = mf.vocab;
        int32_t n_file = 0;
        c.get(n_file);
        if (n_file < 0 || n_file > (1 << 20) || (size_t) n_file * 4 > size) {
            WB_LOG_ERROR("%s: invalid vocabulary size %d\n", __func__, n_file);
            return false;
        }
        v.n_vocab = hp.n_vocab;
        v.id_to_token.assign(std::max(n_file, hp.n_vocab), std::string());
        std::string word;
        for (int i = 0; i < n_file; ++i) {
            uint32_t len = 0;
            c.get(len);
            if (len > size - std::min(size, c.off)) {
                WB_LOG_ERROR("%s: truncated vocabulary\n", __func__);
                return false;
            }
            word.assign((const char *) c.p + c.off, len);
            c.off += len;
            v.token_to_id[word] = i;
            v.id_to_token[i] = word;
        }
        if (v.is_multilingual()) {
            v.token_eot++;
            v.token_sot++;
            const int dt = v.num_languages() - 98;
            v.token_translate += dt; v.token_transcribe += dt; v.token_solm += dt; v.token_prev += dt;
            v.token_nosp += dt; v.token_not += dt; v.token_beg += dt;
        }
        if (n_file < hp.n_vocab) {
            WB_LOG_INFO("%s: adding %d extra tokens\n", __func__, hp.n_vocab - n_file);
            for (int i = n_file; i < hp.n_vocab; ++i) {
                if (i > v.token_beg)                word = "[_TT_" + std::to_string(i - v.token_beg) + "]";
                else if (i == v.token_eot)          word = "[_EOT_]";
                else if (i == v.token_sot)          word = "[_SOT_]";
                else if (i == v.token_translate)    word = "[_TRANSLATE_]";
                else if (i == v.token_transcribe)   word = "[_TRANSCRIBE_]";
                else if (i == v.token_solm)         word = "[_SOLM_]";
                else if (i == v.token_prev)         word = "[_PREV_]";
                else if (i == v.token_nosp)         word = "[_NOSP_]";
                else if (i == v.token_not)          word = "[_NOT_]";
                else if (i == v.token_beg)          word = "[_BEG_]";
                else if (i > v.token_sot && i <= v.token_sot + v.num_languages()) {
                    const char * l = lang_str(i - v.token_sot - 1);
                    word = "[_LANG_" + std::string(l ? l : "?") + "]";
                } else                              word = "[_extra_token_" + std::to_string(i) + "]";
                v.token_to_id[word] = i;
                v.id_to_token[i] = word;
            }
        }
    }

    declare_expected_tensors(mf);

    // tensor records until EOF (whisper.cpp:1531-1633)
    std::map<std::string, bool> seen;
    while (true) {
        int32_t n_dims = 0, name_len = 0, ttype = 0;
        c.get(n_dims);
        c.get(name_len);
        c.get(ttype);
        if (c.eof()) break;

        // the dims (4 bytes each) are read before the name: both must lie inside what is left of the buffer
        if (n_dims < 0 || n_dims > 4 || name_len < 0 || name_len > 4096 || c.off > size ||
            (size_t) name_len + 4u * (size_t) n_dims > size - c.off) {
            WB_LOG_ERROR("%s: corrupt tensor record (n_dims %d, name_len %d)\n", __func__, n_dims, name_len);
            return false;
        }
        int32_t ne[4] = {1, 1, 1, 1};
        int64_t nelements = 1;
        for (int i = 0; i < n_dims; ++i) {
            c.get(ne[i]);
            if (ne[i] <= 0 || ne[i] > (1 << 24)) {
                WB_LOG_ERROR("%s: corrupt tensor record (dimension %d = %d)\n", __func__, i, ne[i]);
                return false;
            }
            nelements *= ne[i];
        }
        std::string name((const char *) c.p + c.off, (size_t) name_len);
        c.off += name_len;

        auto it = mf.tensors.find(name);
        if (it == mf.tensors.end()) {
            WB_LOG_ERROR("%s: unknown tensor '%s' in model file\n", __func__, name.c_str());
            return false;
        }
        TensorView & tv = it->second;
        if (tv.nelements() != nelements) {
            WB_LOG_ERROR("%s: tensor '%s' has wrong size in model file\n", __func__, name.c_str());
            return false;
        }
        if (tv.ne[0] != ne[0] || tv.ne[1] != ne[1] || tv.ne[2] != ne[2]) {
            WB_LOG_ERROR("%s: tensor '%s' has wrong shape in model file: got [%d, %d, %d], expected [%d, %d, %d]\n",
                         __func__, name.c_str(), ne[0], ne[1], ne[2], tv.ne[0], tv.ne[1], tv.ne[2]);
            return false;
        }
        if (quant_block_bytes(ttype) != 0) {
            // a block-quantised matrix: expanded to the f16 image the kernels consume (the value ggml's dequantize_row_* yields, rounded
            // to f16).  The reference multiplies such matrices by activations it first quantises to 8 bits per block (ggml.c vec_dot_q*_q8_*);
            // that part is NOT restated — this backend's products use the f16 activations of the f16 path
            if (tv.type != TT_F16 || ne[0] % 32 != 0) {
                WB_LOG_ERROR("%s: tensor '%s' cannot be block-quantised (type %d)\n", __func__, name.c_str(), ttype);
                return false;
            }
            const size_t qbytes = (size_t) (nelements / 32) * quant_block_bytes(ttype);
            if (qbytes > size - std::min(size, c.off)) {
                WB_LOG_ERROR("%s: tensor '%s' is truncated\n", __func__, name.c_str());
                return false;
            }
            std::vector<float> tmp((size_t) nelements);
            dequantize_blocks(ttype, c.p + c.off, nelements, tmp.data());
            mf.owned.emplace_back((size_t) nelements);
            std::vector<uint16_t> & img = mf.owned.back();
            for (int64_t i = 0; i < nelements; ++i) img[(size_t) i] = f32_to_f16(tmp[(size_t) i]);
            tv.data = (const uint8_t *) img.data();
            c.off += qbytes;
            mf.total_bytes += qbytes;
            if (!seen[name]) { seen[name] = true; mf.n_loaded++; }
            continue;
        }
        if (ttype != 0 && ttype != 1) {
            WB_LOG_ERROR("%s: tensor '%s' has unsupported type %d\n", __func__, name.c_str(), ttype);
            return false;
        }
        const size_t bpe = ttype == 1 ? 2 : 4;
        if ((size_t) nelements * bpe != tv.nbytes) {
            WB_LOG_ERROR("%s: tensor '%s' has wrong size in model file: got %zu, expected %zu\n", __func__,
                         name.c_str(), (size_t) nelements * bpe, tv.nbytes);
            return false;
        }
        if (tv.nbytes > size - std::min(size, c.off)) {
            WB_LOG_ERROR("%s: tensor '%s' is truncated\n", __func__, name.c_str());
            return false;
        }
        tv.data = c.p + c.off;
        c.off += tv.nbytes;
        mf.total_bytes += tv.nbytes;
        if (!seen[name]) {
            seen[name] = true;
            mf.n_loaded++;
        }
    }

    WB_LOG_INFO("%s: model size = %7.2f MB (%d tensors)\n", __func__, mf.total_bytes / 1e6, mf.n_loaded);
    if (mf.n_loaded == 0) {
        WB_LOG_WARN("%s: WARN no tensors loaded from model file - assuming empty model for testing\n", __func__);
    } else if (mf.n_loaded != mf.n_expected) {
        WB_LOG_ERROR("%s: ERROR not all tensors loaded from model file - expected %d, got %d\n", __func__,
                     mf.n_expected, mf.n_loaded);
        return false;
    }
    return true;
}

bool parse_model_file(const void * buffer, size_t size, ModelFile & mf) {
    // allocation failures (a header that passes the bounds above can still ask for more than the host has) must not cross the C ABI
    try {
        return parse_model_file_impl(buffer, size, mf);
    } catch (const std::exception & e) {
        WB_LOG_ERROR("%s: exception while parsing the model file: %s\n", __func__, e.what());
        return false;
    }
}

std::vector<int32_t> tokenize(const Vocab & vocab, const std::string & text) {
    std::vector<std::string> words;
    {
        // GPT-2 pre-tokenisation pattern in its ECMAScript form (whisper.cpp:2899-2920)
        static const std::regex re(
            R"('s|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+)");
        std::string str = text;
        std::smatch m;
        while (std::regex_search(str, m, re)) {
            for (auto x : m) words.push_back(x);
            str = m.suffix();
        }
    }
    std::vector<int32_t> tokens;
    for (const auto & word : words) {
        if (word.empty()) continue;
        int i = 0;
        const int n = (int) word.size();
        while (i < n) {
            int j = n;
            bool found = false;
            while (j > i) {
                auto it = vocab.token_to_id.find(word.substr(i, j - i));
                if (it != vocab.token_to_id.end()) {
                    tokens.push_back(it->second);
                    i = j;
                    found = true;
                    break;
                }
                --j;
            }
            if (!found) {
                WB_LOG_ERROR("unknown token\n");
                ++i;
            }
        }
    }
    return tokens;
}

}  // namespace wb200
