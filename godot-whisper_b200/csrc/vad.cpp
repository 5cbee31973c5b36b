// The host front of the realtime path: the energy-based end-of-speech test the GDExtension runs on the last three
// seconds of captured audio before it calls whisper_full (SpeechToText::voice_activity_detection,
// /root/reference/src/speech_to_text.cpp:378-399 -> _vad_simple :67-104 -> _high_pass_filter :53-64).  Host code: a first-order
// recursive filter over 48 000 samples is a serial chain of a few hundred microseconds and has nothing to gain from the device.
// Restated statement for statement — every sum in f32 in sample order, the filter constant through double like the reference's
// `1.0f / (2.0f * Math_PI * cutoff)` — so the filtered signal and the decision are bit-identical (tests/test_vad.py checks both
// against the reference's own function text, oracle/ref_vad.cpp).
#include "common.h"

#include <cmath>

extern "C" WHISPER_B200_API void whisper_b200_high_pass_filter(float * x, int n, float cutoff_hz, float rate_hz) {
    if (!x || n <= 0) return;
    // a = dt / (RC + dt)  with RC formed in double and rounded once, like the reference's expression
    const float time_const = 1.0f / (2.0f * 3.1415926535897932384626433833 * cutoff_hz);
    const float step = 1.0f / rate_hz;
    const float a = step / (time_const + step);
    // The reference subtracts data[i - 1] AFTER it has stored the previous output there, so the term is the previous OUTPUT, not the
    // previous input: y[i] = a ((y[i-1] + x[i]) - y[i-1]).  Restated as it is — the decision downstream depends on these very values.
    float out = x[0];
    for (float * p = x + 1; p != x + n; ++p) {
        out = a * ((out + *p) - out);
        *p = out;
    }
}

extern "C" WHISPER_B200_API int whisper_b200_vad_simple(float * window, int n, int rate_hz, int tail_ms, float ratio_thold, float cutoff_hz) {
    const int n_tail = (rate_hz * tail_ms) / 1000;
    if (!window || n_tail >= n) return 0;                                 // shorter than its own tail: "no end of speech"
    if (cutoff_hz > 0.0f) whisper_b200_high_pass_filter(window, n, cutoff_hz, (float) rate_hz);
    // mean magnitude of the whole window and of its tail: two sequential f32 sums in sample order
    float sum_all = 0.0f, sum_tail = 0.0f;
    const int tail_from = n - n_tail;
    for (int i = 0; i < n; ++i) {
        const float m = fabsf(window[i]);
        sum_all += m;
        if (i >= tail_from) sum_tail += m;
    }
    const float mean_all = sum_all / n;
    const float mean_tail = n_tail != 0 ? sum_tail / n_tail : sum_tail;
    // the host's variant of the test (src/speech_to_text.cpp:100-103): speech "has ended" only in a faint window whose tail is not louder
    const bool faint = mean_all < 0.0001f && mean_tail < 0.0001f;
    return (faint && !(mean_tail > ratio_thold * mean_all)) ? 1 : 0;
}
