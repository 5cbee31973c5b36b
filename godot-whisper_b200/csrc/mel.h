// Log-mel front end.  The host form (AVX2 / AVX-512, bit-exact with the compiled reference) serves the stage API and audio longer than
// one window; whisper_full on a chunk of up to 30 s computes the same arithmetic on the device (cuda/mel_kernels.cu).
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

// The constant tables of the transform exactly as the host code builds them (libm sinf / cosf of the reference's arguments): Hann window
// [400], leaf DFT twiddles cos / sin [25][25] (table step 16), butterfly twiddles re / im [4][200] for N = 50, 100, 200, 400.  The device
// kernel uploads these instead of recomputing them, so both forms multiply by the same bits.
struct MelTablesView { const float * hann, * leaf_cos, * leaf_sin, * tw_re, * tw_im; };
MelTablesView mel_tables_view();
// Filter j only has non-zero weights in the 4-bin groups [g0[j], g1[j]) (bins 4 g .. 4 g + 3); group sums outside add +0.0 and are skipped.
void mel_filter_spans(const MelFilters & filters, std::vector<int> & g0, std::vector<int> & g1);
// Shape of the spectrogram of an n_samples clip (whisper.cpp:2815-2842): frames in total, frames over real audio, frames that overlap
// samples (all later ones are log10(1e-10)).
void mel_shape(int n_samples, int & n_len, int & n_len_org, int & n_calc);

// |signal| averaged over a (2*hw+1)-sample window, same summation order as whisper.cpp:6350-6366.
void signal_energy(const float * signal, int n_samples, int hw, std::vector<float> & out);

}  // namespace wb200
