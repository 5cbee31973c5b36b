// The host front of the realtime path: the energy-based end-of-speech test the GDExtension runs on the last three
// seconds of captured audio before it calls whisper_full (SpeechToText::voice_activity_detection,
// /root/reference/src/speech_to_text.cpp:378-399 -> _vad_simple :67-104 -> _high_pass_filter :53-64).  Host code: a first-order
// recursive filter over 48 000 samples is a serial chain of a few hundred microseconds and has nothing to gain from the device.
// Restated statement for statement — every sum in f32 in sample order, the filter constant through double like the reference's
// `1.0f / (2.0f * Math_PI * cutoff)` — so the filtered signal and the decision are bit-identical (tests/test_vad.py checks both
// against the reference's own function text, oracle/ref_vad.cpp).
#include "common.h"

#include <cmath>

extern "C" WHISPER_B200_API void whisper_b200_high_pass_filter(float * data, int n_samples, float cutoff, float sample_rate) {
    if (!data || n_samples <= 0) return;
    const float rc = 1.0f / (2.0f * 3.1415926535897932384626433833 * cutoff);
    const float dt = 1.0f / sample_rate;
    const float alpha = dt / (rc + dt);
    float y = data[0];
    for (int i = 1; i < n_samples; i++) {
        y = alpha * (y + data[i] - data[i - 1]);
        data[i] = y;
    }
}

extern "C" WHISPER_B200_API int whisper_b200_vad_simple(float * pcmf32, int n_samples, int sample_rate, int last_ms, float vad_thold, float freq_thold) {
    const int n_samples_last = (sample_rate * last_ms) / 1000;
    if (!pcmf32 || n_samples_last >= n_samples) return 0;                 // not enough samples - assume no speech
    if (freq_thold > 0.0f) whisper_b200_high_pass_filter(pcmf32, n_samples, freq_thold, (float) sample_rate);
    float energy_all = 0.0f, energy_last = 0.0f;
    for (int i = 0; i < n_samples; i++) {
        energy_all += fabsf(pcmf32[i]);
        if (i >= n_samples - n_samples_last) energy_last += fabsf(pcmf32[i]);
    }
    energy_all /= n_samples;
    if (n_samples_last != 0) energy_last /= n_samples_last;
    if (!(energy_all < 0.0001f && energy_last < 0.0001f) || energy_last > vad_thold * energy_all) return 0;
    return 1;
}
