// Host log-mel spectrogram.
//
// Arithmetic restated from thirdparty/whisper.cpp/whisper.cpp:2614-2887:
//   * periodic Hann window built with cosf                                  (whisper.cpp:2711-2725)
//   * 400-entry f32 sin/cos table                                           (whisper.cpp:2614-2629)
//   * radix-2 decimation in time 400 -> 200 -> 100 -> 50 -> 25, naive DFT at 25, all in f32 (whisper.cpp:2634-2709)
//   * power spectrum, 80x201 mel filter bank with f32 4-term partial sums accumulated in f64, log10 in f64
//                                                                           (whisper.cpp:2753-2779)
//   * clamp to (global max - 8), (x + 4) / 4                                (whisper.cpp:2856-2871)
// The recursion is unrolled into an iterative, allocation-free form: the sixteen length-25 sub-DFTs share one
// twiddle matrix and run as 16 SIMD lanes; the four butterfly levels vectorise over k.  Operation order per output
// value — including which products gcc 13 -O3 contracts into FMAs when it compiles the reference for an FMA target
// (checked against the disassembly of oracle/_ref) — is kept, so the result matches the compiled reference bit for bit
// on the same libm.  This file must be compiled with -ffp-contract=off: every fusion here is an explicit fmaf().
#include "mel.h"
#include "common.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#endif

namespace wb200 {

namespace {

constexpr int kN      = WHISPER_N_FFT;       // 400
constexpr int kHop    = WHISPER_HOP_LENGTH;  // 160
constexpr int kBins   = 1 + kN / 2;          // 201
constexpr int kSub    = 16;                  // 400 / 25 sub-sequences
constexpr int kLeaf   = 25;

struct Tables {
    float sin_t[kN], cos_t[kN];
    alignas(64) float hann[kN];
    // leaf DFT twiddles: [k][n] for the 25-point DFT (table step 16)
    float leaf_cos[kLeaf][kLeaf], leaf_sin[kLeaf][kLeaf];
    // butterfly twiddles per level: N = 50, 100, 200, 400 -> k < N/2, re = cos, im = -sin
    float tw_re[4][200], tw_im[4][200];
    Tables() {
        for (int i = 0; i < kN; i++) {
            double theta = (2 * M_PI * i) / kN;
            sin_t[i] = sinf(theta);
            cos_t[i] = cosf(theta);
        }
        for (int i = 0; i < kN; i++) {
            hann[i] = 0.5 * (1.0 - cosf((2.0 * M_PI * i) / (kN)));
        }
        for (int k = 0; k < kLeaf; ++k)
            for (int n = 0; n < kLeaf; ++n) {
                const int idx = (k * n * kSub) % kN;
                leaf_cos[k][n] = cos_t[idx];
                leaf_sin[k][n] = sin_t[idx];
            }
        int N = 50;
        for (int l = 0; l < 4; ++l, N *= 2) {
            const int step = kN / N;
            for (int k = 0; k < N / 2; ++k) {
                tw_re[l][k] = cos_t[k * step];
                tw_im[l][k] = -sin_t[k * step];
            }
        }
    }
};

const Tables & tables() {
    static const Tables t;
    return t;
}

struct Scratch {
    alignas(64) float in[kN];
    // leaf output, lane = sub-sequence r: [k][r]
    alignas(64) float lre[kLeaf][kSub], lim[kLeaf][kSub];
    // ping-pong buffers for the butterfly levels: [sequence][k]
    alignas(64) float are[kSub * kLeaf], aim[kSub * kLeaf];
    alignas(64) float bre[kSub * kLeaf], bim[kSub * kLeaf];
    alignas(64) float power[kBins + 3];
};

#if defined(__AVX2__) && defined(__FMA__)
__attribute__((target("avx512f"))) void leaf_dft_avx512(const Tables & T, Scratch & S) {
    // five output bins per pass: ten independent accumulation chains hide the add latency, every load of x feeds all five
    static_assert(kLeaf % 5 == 0, "bins are processed five at a time");
    for (int k0 = 0; k0 < kLeaf; k0 += 5) {
        __m512 re[5], im[5];
        for (int u = 0; u < 5; ++u) { re[u] = _mm512_setzero_ps(); im[u] = _mm512_setzero_ps(); }
        for (int n = 0; n < kLeaf - 1; ++n) {
            const __m512 x = _mm512_load_ps(S.in + n * kSub);
#pragma GCC unroll 5
            for (int u = 0; u < 5; ++u) {
                re[u] = _mm512_add_ps(re[u], _mm512_mul_ps(x, _mm512_set1_ps(T.leaf_cos[k0 + u][n])));
                im[u] = _mm512_sub_ps(im[u], _mm512_mul_ps(x, _mm512_set1_ps(T.leaf_sin[k0 + u][n])));
            }
        }
        const __m512 x = _mm512_load_ps(S.in + (kLeaf - 1) * kSub);
#pragma GCC unroll 5
        for (int u = 0; u < 5; ++u) {
            re[u] = _mm512_fmadd_ps(x, _mm512_set1_ps(T.leaf_cos[k0 + u][kLeaf - 1]), re[u]);
            im[u] = _mm512_fnmadd_ps(x, _mm512_set1_ps(T.leaf_sin[k0 + u][kLeaf - 1]), im[u]);
            _mm512_store_ps(S.lre[k0 + u], re[u]);
            _mm512_store_ps(S.lim[k0 + u], im[u]);
        }
    }
}

void leaf_dft_avx2(const Tables & T, Scratch & S) {
    auto one = [&](int k) {
        __m256 re0 = _mm256_setzero_ps(), re1 = re0, im0 = re0, im1 = re0;
        for (int n = 0; n < kLeaf - 1; ++n) {
            const __m256 x0 = _mm256_load_ps(S.in + n * kSub), x1 = _mm256_load_ps(S.in + n * kSub + 8);
            const __m256 c = _mm256_broadcast_ss(&T.leaf_cos[k][n]), s = _mm256_broadcast_ss(&T.leaf_sin[k][n]);
            re0 = _mm256_add_ps(re0, _mm256_mul_ps(x0, c)); re1 = _mm256_add_ps(re1, _mm256_mul_ps(x1, c));
            im0 = _mm256_sub_ps(im0, _mm256_mul_ps(x0, s)); im1 = _mm256_sub_ps(im1, _mm256_mul_ps(x1, s));
        }
        const __m256 x0 = _mm256_load_ps(S.in + (kLeaf - 1) * kSub), x1 = _mm256_load_ps(S.in + (kLeaf - 1) * kSub + 8);
        const __m256 c = _mm256_broadcast_ss(&T.leaf_cos[k][kLeaf - 1]), s = _mm256_broadcast_ss(&T.leaf_sin[k][kLeaf - 1]);
        re0 = _mm256_fmadd_ps(x0, c, re0); re1 = _mm256_fmadd_ps(x1, c, re1);
        im0 = _mm256_fnmadd_ps(x0, s, im0); im1 = _mm256_fnmadd_ps(x1, s, im1);
        _mm256_store_ps(S.lre[k], re0); _mm256_store_ps(S.lre[k] + 8, re1);
        _mm256_store_ps(S.lim[k], im0); _mm256_store_ps(S.lim[k] + 8, im1);
    };
    for (int k = 0; k < kLeaf; ++k) one(k);
}
#endif

void leaf_dft_scalar(const Tables & T, Scratch & S) {
    for (int k = 0; k < kLeaf; ++k) {
        float re[kSub], im[kSub];
        for (int r = 0; r < kSub; ++r) { re[r] = 0.0f; im[r] = 0.0f; }
        for (int n = 0; n < kLeaf; ++n) {
            const float c = T.leaf_cos[k][n], s = T.leaf_sin[k][n];
            const float * x = S.in + n * kSub;
            if (n < kLeaf - 1) {
                for (int r = 0; r < kSub; ++r) {
                    const float pc = x[r] * c, ps = x[r] * s;
                    re[r] = re[r] + pc;
                    im[r] = im[r] - ps;
                }
            } else {
                for (int r = 0; r < kSub; ++r) {
                    re[r] = fmaf(x[r], c, re[r]);
                    im[r] = fmaf(-x[r], s, im[r]);
                }
            }
        }
        for (int r = 0; r < kSub; ++r) { S.lre[k][r] = re[r]; S.lim[k][r] = im[r]; }
    }
}

void leaf_dft(const Tables & T, Scratch & S) {
#if defined(__AVX2__) && defined(__FMA__)
    static const bool has512 = __builtin_cpu_supports("avx512f") && !getenv("WHISPER_B200_NO_AVX512");
    if (has512) leaf_dft_avx512(T, S); else leaf_dft_avx2(T, S);
#else
    leaf_dft_scalar(T, S);
#endif
}

#if defined(__AVX2__) && defined(__FMA__)
// The four butterfly levels with 16 butterflies per pass (masked tail), then the power spectrum; per element the same two chained
// FMAs as the scalar form in frame_power.  Returns nothing: S.power is filled.
__attribute__((target("avx512f"))) void butterflies_power_avx512(const Tables & T, Scratch & S) {
    float * sre = S.are, * sim = S.aim, * dre = S.bre, * dim = S.bim;
    int nseq = kSub, len = kLeaf;
    for (int l = 0; l < 4; ++l) {
        const int half = nseq / 2;
        const float * wr = T.tw_re[l], * wi = T.tw_im[l];
        for (int q = 0; q < half; ++q) {
            const float * er = sre + q * len,           * ei = sim + q * len;
            const float * orr = sre + (q + half) * len, * oi = sim + (q + half) * len;
            float * o_r = dre + q * 2 * len, * o_i = dim + q * 2 * len;
            for (int k = 0; k < len; k += 16) {
                const __mmask16 m = (len - k >= 16) ? (__mmask16) 0xFFFF : (__mmask16) ((1u << (len - k)) - 1u);
                const __m512 re = _mm512_maskz_loadu_ps(m, wr + k), im = _mm512_maskz_loadu_ps(m, wi + k);
                const __m512 ro = _mm512_maskz_loadu_ps(m, orr + k), io = _mm512_maskz_loadu_ps(m, oi + k);
                const __m512 ere = _mm512_maskz_loadu_ps(m, er + k), eim = _mm512_maskz_loadu_ps(m, ei + k);
                _mm512_mask_storeu_ps(o_r + k,       m, _mm512_fnmadd_ps(im, io, _mm512_fmadd_ps(re, ro, ere)));
                _mm512_mask_storeu_ps(o_i + k,       m, _mm512_fmadd_ps(im, ro, _mm512_fmadd_ps(re, io, eim)));
                _mm512_mask_storeu_ps(o_r + k + len, m, _mm512_fmadd_ps(im, io, _mm512_fnmadd_ps(re, ro, ere)));
                _mm512_mask_storeu_ps(o_i + k + len, m, _mm512_fnmadd_ps(im, ro, _mm512_fnmadd_ps(re, io, eim)));
            }
        }
        std::swap(sre, dre);
        std::swap(sim, dim);
        nseq = half;
        len *= 2;
    }
    for (int j = 0; j < kBins; j += 16) {
        const __mmask16 m = (kBins - j >= 16) ? (__mmask16) 0xFFFF : (__mmask16) ((1u << (kBins - j)) - 1u);
        const __m512 re = _mm512_maskz_loadu_ps(m, sre + j), im = _mm512_maskz_loadu_ps(m, sim + j);
        _mm512_mask_storeu_ps(S.power + j, m, _mm512_fmadd_ps(re, re, _mm512_mul_ps(im, im)));
    }
}
#endif

#if defined(__AVX2__) && defined(__FMA__)
// dst[r][k] = src[k][r] for a 16 x 16 block held in sixteen registers (rows k) -> sixteen registers (rows r)
__attribute__((target("avx512f"))) inline void transpose16(__m512 (&v)[16]) {
    __m512 t[16];
#pragma GCC unroll 8
    for (int i = 0; i < 8; ++i) { t[2 * i] = _mm512_unpacklo_ps(v[2 * i], v[2 * i + 1]); t[2 * i + 1] = _mm512_unpackhi_ps(v[2 * i], v[2 * i + 1]); }
#pragma GCC unroll 4
    for (int i = 0; i < 4; ++i) {
        v[4 * i + 0] = _mm512_shuffle_ps(t[4 * i + 0], t[4 * i + 2], 0x44); v[4 * i + 1] = _mm512_shuffle_ps(t[4 * i + 0], t[4 * i + 2], 0xEE);
        v[4 * i + 2] = _mm512_shuffle_ps(t[4 * i + 1], t[4 * i + 3], 0x44); v[4 * i + 3] = _mm512_shuffle_ps(t[4 * i + 1], t[4 * i + 3], 0xEE);
    }
#pragma GCC unroll 2
    for (int i = 0; i < 2; ++i)
#pragma GCC unroll 4
        for (int j = 0; j < 4; ++j) {
            t[8 * i + j]     = _mm512_shuffle_f32x4(v[8 * i + j], v[8 * i + 4 + j], 0x88);
            t[8 * i + 4 + j] = _mm512_shuffle_f32x4(v[8 * i + j], v[8 * i + 4 + j], 0xDD);
        }
#pragma GCC unroll 8
    for (int j = 0; j < 8; ++j) {
        v[j]     = _mm512_shuffle_f32x4(t[j], t[8 + j], 0x88);
        v[8 + j] = _mm512_shuffle_f32x4(t[j], t[8 + j], 0xDD);
    }
}

// leaf output [k][r] (25 x 16) -> [r][k] (16 rows of 25) for both planes: two 16 x 16 register transposes per plane (k = 0..15 and
// k = 16..24 padded with don't-care rows), row r written as 16 + 9 floats
__attribute__((target("avx512f"))) void transpose_leaf_avx512(Scratch & S) {
    const float * src[2] = { &S.lre[0][0], &S.lim[0][0] };
    float * dst[2] = { S.are, S.aim };
    for (int p = 0; p < 2; ++p) {
        __m512 a[16], b[16];
        for (int k = 0; k < 16; ++k) a[k] = _mm512_load_ps(src[p] + k * kSub);
        for (int k = 0; k < 16; ++k) b[k] = k < kLeaf - 16 ? _mm512_load_ps(src[p] + (16 + k) * kSub) : _mm512_setzero_ps();
        transpose16(a);
        transpose16(b);
        for (int i = 0; i < 16; ++i) {
            _mm512_storeu_ps(dst[p] + i * kLeaf, a[i]);
            _mm512_mask_storeu_ps(dst[p] + i * kLeaf + 16, (__mmask16) ((1u << (kLeaf - 16)) - 1u), b[i]);
        }
    }
}
#endif

// One windowed frame (400 f32) -> power spectrum bins 0..200.
void frame_power(const Tables & T, Scratch & S) {
    // 16 interleaved 25-point DFTs: sub-sequence r holds in[16 n + r]; the 16 sub-sequences are the SIMD lanes (2 x 8 with AVX2,
    // 1 x 16 with AVX-512), two output bins k per pass so every load of x feeds both.  gcc -O3 vectorises the reference's dft()
    // loop with an in-order reduction — products rounded separately from the adds for n = 0..23 — and contracts only the scalar
    // remainder iteration (n = 24) into FMAs; the same here, explicitly.
    leaf_dft(T, S);
#if defined(__AVX2__) && defined(__FMA__)
    static const bool has512 = __builtin_cpu_supports("avx512f") && !getenv("WHISPER_B200_NO_AVX512");
    if (has512) { transpose_leaf_avx512(S); butterflies_power_avx512(T, S); return; }
#endif
    // transpose to [r][k]
    for (int r = 0; r < kSub; ++r)
        for (int k = 0; k < kLeaf; ++k) { S.are[r * kLeaf + k] = S.lre[k][r]; S.aim[r * kLeaf + k] = S.lim[k][r]; }

    // butterflies: at a level with `nseq` input sequences of length `len`, output sequence q (< nseq/2) combines
    // even = input q and odd = input q + nseq/2  (x[stride*n + q] split into even/odd n)
    float * sre = S.are, * sim = S.aim, * dre = S.bre, * dim = S.bim;
    int nseq = kSub, len = kLeaf;
    for (int l = 0; l < 4; ++l) {
        const int half = nseq / 2;
        const float * wr = T.tw_re[l], * wi = T.tw_im[l];
        for (int q = 0; q < half; ++q) {
            const float * er = sre + q * len,          * ei = sim + q * len;
            const float * orr = sre + (q + half) * len, * oi = sim + (q + half) * len;
            float * o_r = dre + q * 2 * len, * o_i = dim + q * 2 * len;
            int k = 0;
#if defined(__AVX2__) && defined(__FMA__)
            // eight butterflies per pass; per element the same two chained FMAs as the scalar form below
            for (; k + 8 <= len; k += 8) {
                const __m256 re = _mm256_loadu_ps(wr + k), im = _mm256_loadu_ps(wi + k);
                const __m256 ro = _mm256_loadu_ps(orr + k), io = _mm256_loadu_ps(oi + k);
                const __m256 ere = _mm256_loadu_ps(er + k), eim = _mm256_loadu_ps(ei + k);
                _mm256_storeu_ps(o_r + k,       _mm256_fnmadd_ps(im, io, _mm256_fmadd_ps(re, ro, ere)));
                _mm256_storeu_ps(o_i + k,       _mm256_fmadd_ps(im, ro, _mm256_fmadd_ps(re, io, eim)));
                _mm256_storeu_ps(o_r + k + len, _mm256_fmadd_ps(im, io, _mm256_fnmadd_ps(re, ro, ere)));
                _mm256_storeu_ps(o_i + k + len, _mm256_fnmadd_ps(im, ro, _mm256_fnmadd_ps(re, io, eim)));
            }
#endif
            for (; k < len; ++k) {
                const float re = wr[k], im = wi[k];
                const float ro = orr[k], io = oi[k];
                // even + re*re_odd - im*im_odd ; even + re*im_odd + im*re_odd   (whisper.cpp:2700-2704)
                o_r[k]       = fmaf(-im, io, fmaf(re, ro, er[k]));
                o_i[k]       = fmaf(im, ro, fmaf(re, io, ei[k]));
                o_r[k + len] = fmaf(im, io, fmaf(-re, ro, er[k]));
                o_i[k + len] = fmaf(-im, ro, fmaf(-re, io, ei[k]));
            }
        }
        std::swap(sre, dre);
        std::swap(sim, dim);
        nseq = half;
        len *= 2;
    }
    // sre/sim now hold the single length-400 spectrum
    for (int j = 0; j < kBins; ++j) {
        const float re = sre[j], im = sim[j];
        S.power[j] = fmaf(re, re, im * im);
    }
}

constexpr int kMaxMel = 128;      // mel bands (80; 128 for large-v3 style filter banks)

// out[i] = (float) log10(x[i]) for x[i] >= 1e-10, bit-identical to calling libm's log10 and rounding to f32.
// A SIMD evaluation of log10 (range reduction to [sqrt(1/2), sqrt(2)), atanh series; |error| < 1e-13) gives y; if y - 2e-11 and
// y + 2e-11 round to the same f32, every value in between does — in particular libm's result (its error is a few 1e-16) — so
// that f32 is the answer.  Otherwise (a rounding boundary within 2e-11 of y: a few dozen values per 30 s of audio) libm decides.
#if defined(__AVX2__) && defined(__FMA__)
// eight values per pass; same evaluation and the same rounding-safety check as the AVX2 loop in log10_to_f32
__attribute__((target("avx512f"))) int log10_to_f32_avx512(const double * x, float * out, int n) {
    const __m512d one = _mm512_set1_pd(1.0), half = _mm512_set1_pd(0.5), sqrt2 = _mm512_set1_pd(1.4142135623730951);
    const __m512d ln2 = _mm512_set1_pd(0.6931471805599453), inv_ln10 = _mm512_set1_pd(0.4342944819032518), eps = _mm512_set1_pd(2e-11);
    const __m512i mant_mask = _mm512_set1_epi64(0x000FFFFFFFFFFFFFLL), exp_one = _mm512_set1_epi64(0x3FF0000000000000LL);
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        const __m512d v = _mm512_loadu_pd(x + i);
        const __m512i bits = _mm512_castpd_si512(v);
        __m512d m = _mm512_castsi512_pd(_mm512_or_si512(_mm512_and_si512(bits, mant_mask), exp_one));       // [1, 2)
        __m512d e = _mm512_sub_pd(_mm512_cvtepi32_pd(_mm512_cvtepi64_epi32(_mm512_srli_epi64(bits, 52))), _mm512_set1_pd(1023.0));
        const __mmask8 big = _mm512_cmp_pd_mask(m, sqrt2, _CMP_GT_OQ);
        m = _mm512_mask_mul_pd(m, big, m, half);
        e = _mm512_mask_add_pd(e, big, e, one);
        const __m512d s = _mm512_div_pd(_mm512_sub_pd(m, one), _mm512_add_pd(m, one));
        const __m512d z = _mm512_mul_pd(s, s);
        __m512d p = _mm512_set1_pd(1.0 / 21.0);
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 19.0));
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 17.0));
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 15.0));
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 13.0));
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 11.0));
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 9.0));
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 7.0));
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 5.0));
        p = _mm512_fmadd_pd(p, z, _mm512_set1_pd(1.0 / 3.0));
        p = _mm512_fmadd_pd(p, z, one);
        const __m512d lnm = _mm512_mul_pd(_mm512_add_pd(s, s), p);
        const __m512d y = _mm512_mul_pd(_mm512_fmadd_pd(e, ln2, lnm), inv_ln10);
        const __m256 lo = _mm512_cvtpd_ps(_mm512_sub_pd(y, eps)), hi = _mm512_cvtpd_ps(_mm512_add_pd(y, eps));
        const int same = _mm256_movemask_ps(_mm256_cmp_ps(lo, hi, _CMP_EQ_OQ));
        _mm256_storeu_ps(out + i, lo);
        if (same != 0xFF) {
            for (int l = 0; l < 8; ++l) if (!((same >> l) & 1)) out[i + l] = (float) log10(x[i + l]);
        }
    }
    return i;
}
#endif

void log10_to_f32(const double * x, float * out, int n) {
    int i = 0;
#if defined(__AVX2__) && defined(__FMA__)
    static const bool has512 = __builtin_cpu_supports("avx512f") && !getenv("WHISPER_B200_NO_AVX512");
    if (has512) i = log10_to_f32_avx512(x, out, n);
    const __m256d one = _mm256_set1_pd(1.0), half = _mm256_set1_pd(0.5), sqrt2 = _mm256_set1_pd(1.4142135623730951);
    const __m256d ln2 = _mm256_set1_pd(0.6931471805599453), inv_ln10 = _mm256_set1_pd(0.4342944819032518), eps = _mm256_set1_pd(2e-11);
    const __m256i mant_mask = _mm256_set1_epi64x(0x000FFFFFFFFFFFFFLL), exp_one = _mm256_set1_epi64x(0x3FF0000000000000LL);
    const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 0, 0, 0);
    for (; i + 4 <= n; i += 4) {
        const __m256d v = _mm256_loadu_pd(x + i);
        const __m256i bits = _mm256_castpd_si256(v);
        __m256d m = _mm256_castsi256_pd(_mm256_or_si256(_mm256_and_si256(bits, mant_mask), exp_one));       // [1, 2)
        const __m128i e32 = _mm256_castsi256_si128(_mm256_permutevar8x32_epi32(_mm256_srli_epi64(bits, 52), pick));
        __m256d e = _mm256_sub_pd(_mm256_cvtepi32_pd(e32), _mm256_set1_pd(1023.0));
        const __m256d big = _mm256_cmp_pd(m, sqrt2, _CMP_GT_OQ);
        m = _mm256_blendv_pd(m, _mm256_mul_pd(m, half), big);
        e = _mm256_add_pd(e, _mm256_and_pd(big, one));
        const __m256d s = _mm256_div_pd(_mm256_sub_pd(m, one), _mm256_add_pd(m, one));
        const __m256d z = _mm256_mul_pd(s, s);
        __m256d p = _mm256_set1_pd(1.0 / 21.0);
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 19.0));
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 17.0));
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 15.0));
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 13.0));
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 11.0));
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 9.0));
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 7.0));
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 5.0));
        p = _mm256_fmadd_pd(p, z, _mm256_set1_pd(1.0 / 3.0));
        p = _mm256_fmadd_pd(p, z, one);
        const __m256d lnm = _mm256_mul_pd(_mm256_add_pd(s, s), p);
        const __m256d y = _mm256_mul_pd(_mm256_fmadd_pd(e, ln2, lnm), inv_ln10);
        const __m128 lo = _mm256_cvtpd_ps(_mm256_sub_pd(y, eps)), hi = _mm256_cvtpd_ps(_mm256_add_pd(y, eps));
        const int same = _mm_movemask_ps(_mm_cmpeq_ps(lo, hi));
        _mm_storeu_ps(out + i, lo);
        if (same != 0xF) {
            for (int l = 0; l < 4; ++l) if (!((same >> l) & 1)) out[i + l] = (float) log10(x[i + l]);
        }
    }
#endif
    for (; i < n; ++i) out[i] = (float) log10(x[i]);
}

// in[j] = hann[j] * src[j] for j < n_take, 0 behind.  With AVX-512 the stores are as wide as the loads of the leaf DFT that
// follows (a 64-byte load that spans two narrower stores still in flight cannot be forwarded and stalls).
#if defined(__AVX2__) && defined(__FMA__)
__attribute__((target("avx512f"))) void window_frame_avx512(const Tables & T, const float * src, float * in) {
    for (int j = 0; j < kN; j += 16) _mm512_store_ps(in + j, _mm512_mul_ps(_mm512_load_ps(T.hann + j), _mm512_loadu_ps(src + j)));
}
#endif
void window_frame(const Tables & T, const float * src, int n_take, float * in) {
#if defined(__AVX2__) && defined(__FMA__)
    static const bool has512 = __builtin_cpu_supports("avx512f") && !getenv("WHISPER_B200_NO_AVX512");
    if (has512 && n_take == kN) { window_frame_avx512(T, src, in); return; }
#endif
    for (int j = 0; j < n_take; ++j) in[j] = T.hann[j] * src[j];
    for (int j = n_take; j < kN; ++j) in[j] = 0.0f;
}

struct FilterSpan { int g0, g1; };  // 4-wide groups [g0, g1) that contain non-zero weights

// The filter bank re-laid for SIMD over groups: for filter j, lane u of vector v holds the weights of group g0 + 8 v + u in four
// planes (weight of bin 4 g + q in plane q); lanes behind the filter's last group hold zeros.
struct FilterPlan {
    std::vector<FilterSpan> spans;
    std::vector<int> first;                 // first vector of filter j in w[] (w holds 4 planes x 8 lanes per vector)
    std::vector<float> w;
    std::vector<float> last;                // weight of bin 200 (the remainder term)
    // the same groups as one flat list (AVX-512 path): entry e = group g of filter j, entries of a filter consecutive; e_first[j] is
    // the first entry of filter j; e_bin[e] = 4 g; e_w[q][e] = weight of bin 4 g + q; padded to a multiple of 16 with zero weights
    std::vector<int> e_first, e_bin;
    std::vector<float> e_w[4];
    explicit FilterPlan(const MelFilters & f) : spans(f.n_mel), first(f.n_mel + 1, 0), last(f.n_mel), e_first(f.n_mel + 1, 0) {
        for (int j = 0; j < f.n_mel; ++j) {
            int lo = 50, hi = 0;
            for (int g = 0; g < 50; ++g) {
                bool nz = false;
                for (int t = 0; t < 4; ++t) nz |= f.data[(size_t) j * kBins + 4 * g + t] != 0.0f;
                if (nz) { lo = std::min(lo, g); hi = std::max(hi, g + 1); }
            }
            spans[j] = { std::min(lo, hi), hi };
            first[j + 1] = first[j] + (spans[j].g1 - spans[j].g0 + 7) / 8;
            last[j] = f.data[(size_t) j * kBins + 200];
        }
        for (int j = 0; j < f.n_mel; ++j) e_first[j + 1] = e_first[j] + (spans[j].g1 - spans[j].g0);
        const size_t n_e = ((size_t) e_first[f.n_mel] + 15) & ~(size_t) 15;
        e_bin.assign(n_e, 0);
        for (int q = 0; q < 4; ++q) e_w[q].assign(n_e, 0.0f);
        for (int j = 0; j < f.n_mel; ++j)
            for (int g = spans[j].g0; g < spans[j].g1; ++g) {
                const int e = e_first[j] + g - spans[j].g0;
                e_bin[e] = 4 * g;
                for (int q = 0; q < 4; ++q) e_w[q][e] = f.data[(size_t) j * kBins + 4 * g + q];
            }
        w.assign((size_t) first[f.n_mel] * 32, 0.0f);
        for (int j = 0; j < f.n_mel; ++j)
            for (int g = spans[j].g0; g < spans[j].g1; ++g) {
                const int e = g - spans[j].g0;
                float * dst = w.data() + (size_t) (first[j] + e / 8) * 32 + (e & 7);
                for (int q = 0; q < 4; ++q) dst[8 * q] = f.data[(size_t) j * kBins + 4 * g + q];
            }
    }
};

// sums[j] = max(1e-10, sum over the groups of filter j, in order, of (f64) part(g)) with
//   part(g) = fma(P3, F3, fma(P2, F2, fma(P0, F0, P1 * F1)))     — gcc contracts p0 + p1 as fma(P0, F0, P1 * F1): the SECOND product
// is the rounded one — plus the remainder term P[200] * F[200] (whisper.cpp:2761-2773).  The parts of eight consecutive groups are
// computed in SIMD lanes from the de-interleaved power spectrum; the f64 accumulation stays sequential per filter.
#if defined(__AVX2__) && defined(__FMA__)
// AVX-512: the parts of all groups of all filters in one flat pass (16 groups per vector, the four bins of a group gathered from the
// power spectrum), then the in-order f64 sums per filter.
__attribute__((target("avx512f"))) void filter_bank_avx512(const float * P, const FilterPlan & FP, int n_mel, double * sums) {
    alignas(64) float parts[1024];
    const int n_e = (int) FP.e_bin.size();
    for (int e = 0; e < n_e; e += 16) {
        const __m512i b = _mm512_loadu_si512((const void *) (FP.e_bin.data() + e));
        __m512 part = _mm512_mul_ps(_mm512_i32gather_ps(b, P + 1, 4), _mm512_loadu_ps(FP.e_w[1].data() + e));
        part = _mm512_fmadd_ps(_mm512_i32gather_ps(b, P, 4), _mm512_loadu_ps(FP.e_w[0].data() + e), part);
        part = _mm512_fmadd_ps(_mm512_i32gather_ps(b, P + 2, 4), _mm512_loadu_ps(FP.e_w[2].data() + e), part);
        part = _mm512_fmadd_ps(_mm512_i32gather_ps(b, P + 3, 4), _mm512_loadu_ps(FP.e_w[3].data() + e), part);
        _mm512_store_ps(parts + e, part);
    }
    for (int j = 0; j < n_mel; ++j) {
        double sum = 0.0;
        for (int e = FP.e_first[j]; e < FP.e_first[j + 1]; ++e) sum += parts[e];
        sum += P[200] * FP.last[j];
        sums[j] = std::max(sum, 1e-10);
    }
}
#endif

void filter_bank(const float * P, const MelFilters & filters, const FilterPlan & FP, int n_mel, double * sums) {
#if defined(__AVX2__) && defined(__FMA__)
    static const bool has512 = __builtin_cpu_supports("avx512f") && !getenv("WHISPER_B200_NO_AVX512");
    if (has512 && FP.e_bin.size() <= 1024) { filter_bank_avx512(P, FP, n_mel, sums); return; }
    alignas(32) float pl[4][64];            // plane q: P[4 g + q] for g = 0..49, zeros behind
    for (int q = 0; q < 4; ++q) { for (int g = 0; g < 50; ++g) pl[q][g] = P[4 * g + q]; for (int g = 50; g < 64; ++g) pl[q][g] = 0.0f; }
    for (int j = 0; j < n_mel; ++j) {
        const int g0 = FP.spans[j].g0, ng = FP.spans[j].g1 - g0;
        double sum = 0.0;
        for (int v = 0; v * 8 < ng; ++v) {
            const float * w = FP.w.data() + (size_t) (FP.first[j] + v) * 32;
            const int g = g0 + 8 * v;
            __m256 part = _mm256_mul_ps(_mm256_loadu_ps(pl[1] + g), _mm256_loadu_ps(w + 8));
            part = _mm256_fmadd_ps(_mm256_loadu_ps(pl[0] + g), _mm256_loadu_ps(w), part);
            part = _mm256_fmadd_ps(_mm256_loadu_ps(pl[2] + g), _mm256_loadu_ps(w + 16), part);
            part = _mm256_fmadd_ps(_mm256_loadu_ps(pl[3] + g), _mm256_loadu_ps(w + 24), part);
            alignas(32) float parts[8];
            _mm256_store_ps(parts, part);
            const int n = std::min(8, ng - 8 * v);
            for (int u = 0; u < n; ++u) sum += parts[u];
        }
        sum += P[200] * FP.last[j];
        sums[j] = std::max(sum, 1e-10);
    }
    (void) filters;
#else
    for (int j = 0; j < n_mel; ++j) {
        const float * F = filters.data.data() + (size_t) j * kBins;
        double sum = 0.0;
        for (int g = FP.spans[j].g0; g < FP.spans[j].g1; ++g) {
            const int k = 4 * g;
            float part = P[k + 1] * F[k + 1];
            part = fmaf(P[k + 0], F[k + 0], part);
            part = fmaf(P[k + 2], F[k + 2], part);
            part = fmaf(P[k + 3], F[k + 3], part);
            sum += part;
        }
        sum += P[200] * F[200];
        sums[j] = std::max(sum, 1e-10);
    }
#endif
}

// Frames [i0, i1) (all of them real frames: i1 <= n_calc) -> log10 mel energies, written frame-block-wise so that each mel row
// receives 16 consecutive floats at a time; *out_max receives the largest value written.
// The padded signal of the reference is never materialised: `head` holds its first kHeadLen values (the reflected front pad and the
// first samples), everything behind is read straight from `samples` (padded[x] = samples[x - 200]).
constexpr int kHeadLen = kN / 2 + 2 * kN;
void mel_worker(int i0, int i1, const Tables & T, const float * head, const float * samples, int n_valid,
                const MelFilters & filters, const FilterPlan & FP, Mel & mel, float * out_max) {
    Scratch S;
    constexpr int kBlk = 16;
    alignas(64) float blk[kMaxMel][kBlk];
    float vmax = -1e20f;

    for (int ib = i0; ib < i1; ib += kBlk) {
        const int nb = std::min(kBlk, i1 - ib);
        for (int f = 0; f < nb; ++f) {
            const int offset = (ib + f) * kHop;
            const int n_take = std::max(0, std::min(kN, n_valid - offset));
            const float * src = offset + kN <= kHeadLen ? head + offset : samples + (offset - kN / 2);
            window_frame(T, src, n_take, S.in);

            frame_power(T, S);

            // mel filter bank: 4-term f32 partial sums accumulated into f64, k = 0,4,..,196 (whisper.cpp:2761-2768); groups whose
            // four weights are all zero add +0.0 and are skipped
            const float * P = S.power;
            alignas(32) double sums[kMaxMel];
            alignas(32) float logs[kMaxMel];
            filter_bank(P, filters, FP, mel.n_mel, sums);
            // (float) log10(sum) for the whole frame (whisper.cpp:2775-2777)
            log10_to_f32(sums, logs, mel.n_mel);
            for (int j = 0; j < mel.n_mel; ++j) { blk[j][f] = logs[j]; vmax = std::max(vmax, logs[j]); }
        }
        for (int j = 0; j < mel.n_mel; ++j) memcpy(mel.data.data() + (size_t) j * mel.n_len + ib, blk[j], (size_t) nb * sizeof(float));
    }
    *out_max = vmax;
}

}  // namespace

bool log_mel_spectrogram(const float * samples, int n_samples, int n_threads, const MelFilters & filters, Mel & mel) {
    if (filters.n_fft != kBins || filters.n_mel <= 0 || filters.n_mel > kMaxMel) {
        WB_LOG_ERROR("%s: unsupported mel filter bank %d x %d\n", __func__, filters.n_mel, filters.n_fft);
        return false;
    }
    const Tables & T = tables();
    n_threads = std::max(1, n_threads);

    const int64_t pad30 = (int64_t) WHISPER_SAMPLE_RATE * WHISPER_CHUNK_SIZE;  // 480 000 zeros
    const int     pad2  = kN / 2;                                              // 200 reflected / trailing

    // The reference pads with 30 s of zeros + 200 more (whisper.cpp:2815-2827); frames are only computed while they overlap
    // samples (i < n_calc below) and read nothing behind n_valid, so the zeros are implied instead of materialised.
    const int64_t padded_size = (int64_t) n_samples + pad30 + 2 * pad2;
    const int n_valid = n_samples + pad2;
    // head of the padded signal: samples[1..200] reflected in front (whisper.cpp:2827; guarded for clips shorter than 201 samples),
    // then the first samples; zeros where the clip has ended
    float head[kHeadLen];
    for (int i = 0; i < pad2; ++i) {
        const int s = pad2 - i;
        head[i] = s < n_samples ? samples[s] : 0.0f;
    }
    for (int i = pad2; i < kHeadLen; ++i) head[i] = i - pad2 < n_samples ? samples[i - pad2] : 0.0f;

    mel.n_mel     = filters.n_mel;
    mel.n_len     = (int) ((padded_size - kN) / kHop);
    mel.n_len_org = 1 + (n_samples + pad2 - kN) / kHop;
    mel.data.resize((size_t) mel.n_mel * mel.n_len);

    const FilterPlan FP(filters);           // (a few microseconds: 80 short filters)

    // frames [0, n_calc) overlap samples and are computed (split evenly over the threads: per-frame results are independent, the
    // reference interleaves them); every later frame is log10(1e-10) (whisper.cpp:2748-2752)
    const int n_calc = std::min(n_valid / kHop + 1, mel.n_len);
    const float low = (float) log10(1e-10);
    n_threads = std::min(n_threads, std::max(1, n_calc / 64));
    std::vector<float> tmax((size_t) n_threads, -1e20f);
    {
        const int per = ((n_calc + n_threads - 1) / n_threads + 15) & ~15;
        std::vector<std::thread> workers;
        workers.reserve(n_threads - 1);
        for (int iw = 1; iw < n_threads; ++iw) {
            const int i0 = std::min(n_calc, iw * per), i1 = std::min(n_calc, i0 + per);
            workers.emplace_back(mel_worker, i0, i1, std::cref(T), (const float *) head, samples, n_valid, std::cref(filters), std::cref(FP), std::ref(mel), &tmax[iw]);
        }
        mel_worker(0, std::min(n_calc, per), T, head, samples, n_valid, filters, FP, mel, &tmax[0]);
        for (auto & w : workers) w.join();
    }

    // clamping and normalisation (whisper.cpp:2856-2871): the maximum runs over all frames, the uncomputed ones included
    float fmax = -1e20f;
    for (float v : tmax) fmax = std::max(fmax, v);
    if (n_calc < mel.n_len) fmax = std::max(fmax, low);
    double mmax = fmax;
    mmax -= 8.0;
    const float fclamp = (float) mmax;
    auto norm = [&](float v) { if (v < mmax) v = fclamp; return (float) ((v + 4.0) / 4.0); };
    const float tail = norm(low);
    for (int j = 0; j < mel.n_mel; ++j) {
        float * row = mel.data.data() + (size_t) j * mel.n_len;
        for (int i = 0; i < n_calc; ++i) row[i] = norm(row[i]);
        std::fill(row + n_calc, row + mel.n_len, tail);
    }
    return true;
}

MelTablesView mel_tables_view() {
    const Tables & T = tables();
    return MelTablesView{T.hann, &T.leaf_cos[0][0], &T.leaf_sin[0][0], &T.tw_re[0][0], &T.tw_im[0][0]};
}

void mel_filter_spans(const MelFilters & filters, std::vector<int> & g0, std::vector<int> & g1) {
    const FilterPlan FP(filters);
    g0.resize(filters.n_mel); g1.resize(filters.n_mel);
    for (int j = 0; j < filters.n_mel; ++j) { g0[j] = FP.spans[j].g0; g1[j] = FP.spans[j].g1; }
}

void mel_shape(int n_samples, int & n_len, int & n_len_org, int & n_calc) {
    const int64_t padded_size = (int64_t) n_samples + (int64_t) WHISPER_SAMPLE_RATE * WHISPER_CHUNK_SIZE + kN;
    n_len     = (int) ((padded_size - kN) / kHop);
    n_len_org = 1 + (n_samples + kN / 2 - kN) / kHop;
    n_calc    = std::min((n_samples + kN / 2) / kHop + 1, n_len);
}

void signal_energy(const float * signal, int n_samples, int hw, std::vector<float> & out) {
    // result[i] = (sum_{j=-hw..hw, in range} |signal[i+j]|) / (2 hw + 1), each sum taken in increasing j in f32 (whisper.cpp:6350-6366).
    // Interior outputs keep their running sums in registers (32 outputs = 4 vectors per pass, one unaligned load + add per term);
    // the 2*hw border outputs take the scalar form.  Same order of additions per output either way.
    out.assign(n_samples, 0.0f);
    std::vector<float> mag(n_samples);
    for (int i = 0; i < n_samples; ++i) mag[i] = fabsf(signal[i]);
    const float denom = (float) (2 * hw + 1);
    auto scalar = [&](int i) {
        float sum = 0.0f;
        for (int j = -hw; j <= hw; ++j) if (i + j >= 0 && i + j < n_samples) sum += mag[i + j];
        out[i] = sum / denom;
    };
    int i = 0;
    for (; i < std::min(hw, n_samples); ++i) scalar(i);
#if defined(__AVX2__)
    const __m256 vden = _mm256_set1_ps(denom);
    for (; i + 32 + hw <= n_samples; i += 32) {
        __m256 a0 = _mm256_setzero_ps(), a1 = a0, a2 = a0, a3 = a0;
        const float * m = mag.data() + i - hw;
        for (int j = 0; j <= 2 * hw; ++j) {
            a0 = _mm256_add_ps(a0, _mm256_loadu_ps(m + j));
            a1 = _mm256_add_ps(a1, _mm256_loadu_ps(m + j + 8));
            a2 = _mm256_add_ps(a2, _mm256_loadu_ps(m + j + 16));
            a3 = _mm256_add_ps(a3, _mm256_loadu_ps(m + j + 24));
        }
        _mm256_storeu_ps(out.data() + i,      _mm256_div_ps(a0, vden));
        _mm256_storeu_ps(out.data() + i + 8,  _mm256_div_ps(a1, vden));
        _mm256_storeu_ps(out.data() + i + 16, _mm256_div_ps(a2, vden));
        _mm256_storeu_ps(out.data() + i + 24, _mm256_div_ps(a3, vden));
    }
#endif
    for (; i < n_samples; ++i) scalar(i);
}

}  // namespace wb200
