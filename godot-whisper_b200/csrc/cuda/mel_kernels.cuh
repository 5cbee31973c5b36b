// Log-mel spectrogram on the device (mel_kernels.cu): the arithmetic of csrc/mel.cpp — itself the reference's log_mel_spectrogram
// (/root/reference/thirdparty/whisper.cpp/whisper.cpp:2614-2887) operation for operation — one warp per frame.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wb200 {

struct MelDevTables {
    const float * hann = nullptr;       // [400]
    const float * leaf_cos = nullptr;   // [25][25]
    const float * leaf_sin = nullptr;
    const float * tw_re = nullptr;      // [4][200]
    const float * tw_im = nullptr;
    const float * filt = nullptr;       // [n_mel][201]
    const int *   g0 = nullptr;         // [n_mel] first / one-past-last 4-bin group with non-zero weights
    const int *   g1 = nullptr;
    int n_mel = 0;
};

// One clip of a batch: samples at pcm (device), spectrogram of the slot at raw (f32 [frames][n_mel], raw log10 values), its maximum
// at max_bits (order-preserving integer image of the float, reset by the launcher).
struct MelClip {
    const float * pcm;
    float *       raw;
    int *         max_bits;
    int           n_samples;
    int           n_calc;               // frames that overlap samples
};

// raw[i][m] = (float) log10(max(1e-10, sum_k filt[m][k] * |DFT_400(hann * frame_i)|^2[k]))  for i < n_calc, and the maximum over them.
void launch_logmel_frames(const MelDevTables & T, const MelClip * clips_dev, int n_clips, int max_calc, cudaStream_t st);

// The encoder's view of a window of the spectrogram: frames [offset, offset + n_frames) clamped and normalised like whisper.cpp:2856-2871
// (global maximum incl. the constant tail, max - 8 floor, (x + 4) / 4), zero beyond n_len (the window copy of whisper.cpp:1692-1706), as
// f16 token-major rows [n_frames + 2][n_mel] with a zero row at either end (the conv's padding).
struct MelWindow {
    const float * raw;
    const int *   max_bits;
    __half *      out;
    int           n_calc, n_len, offset;
};
void launch_mel_window(const MelWindow * wins_dev, int n_wins, int n_mel, int n_frames, float low, cudaStream_t st);

// The energy envelope the token-level timestamps snap to (whisper.cpp:6350-6366, get_signal_energy with hw = 32): same additions in the
// same order as csrc/mel.cpp's signal_energy, one thread per output sample.
struct EnergyClip { const float * pcm; float * out; int n_samples; };
void launch_signal_energy(const EnergyClip * clips_dev, int n_clips, int max_samples, int hw, cudaStream_t st);

// order-preserving float <-> int image used for the atomic maximum
__host__ __device__ inline int   mel_float_to_ordered(float f) { int i; memcpy(&i, &f, 4); return i >= 0 ? i : i ^ 0x7fffffff; }
__host__ __device__ inline float mel_ordered_to_float(int i) { i = i >= 0 ? i : i ^ 0x7fffffff; float f; memcpy(&f, &i, 4); return f; }

}  // namespace wb200
