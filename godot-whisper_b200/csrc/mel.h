// Host log-mel front end (stays on the host by design: BASELINE.json north_star).
#pragma once

#include "model.h"

#include <vector>

namespace wb200 {

struct Mel {                 // whisper.cpp:349-355
    int n_len     = 0;       // frames incl. the 30 s of zero padding
    int n_len_org = 0;       // frames covering the real audio
    int n_mel     = 0;
    std::vector<float> data; // [n_mel][n_len]
};

// PCM (16 kHz mono f32) -> log-mel, same arithmetic as whisper.cpp:2793-2887 (frame 400 / hop 160 only).
bool log_mel_spectrogram(const float * samples, int n_samples, int n_threads, const MelFilters & filters, Mel & mel);

// |signal| averaged over a (2*hw+1)-sample window, same summation order as whisper.cpp:6350-6366.
void signal_energy(const float * signal, int n_samples, int hw, std::vector<float> & out);

}  // namespace wb200
