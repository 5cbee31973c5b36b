// Log-mel spectrogram on the device — SURVEY.md §8f.2.
//
// Same operations in the same order as csrc/mel.cpp (which matches the compiled reference bit for bit, whisper.cpp:2614-2887):
//   window      in[j] = hann[j] * x[j]                                                    one f32 multiply
//   leaf DFTs   sixteen 25-point DFTs over x[16 n + r]: acc += x * tw with the product rounded on its own for n = 0..23 and
//               fused for n = 24 (what gcc makes of the reference's dft() loop)           (whisper.cpp:2634-2658)
//   butterflies four radix-2 levels 25 -> 50 -> 100 -> 200 -> 400, two chained FMAs per output   (whisper.cpp:2660-2709)
//   power       re * re + im * im with the second product rounded, the first fused       (whisper.cpp:2753-2757)
//   filter bank 4-term f32 partial sums (one rounded product, three FMAs) accumulated in f64 in bin order, plus the term of bin 200
//   log10       in f64, rounded to f32                                                    (whisper.cpp:2761-2777)
// Every multiply-add is written with __fmul_rn / __fadd_rn / __fmaf_rn so that nvcc neither fuses nor splits anything.  The one
// operation that is not the same code as on the host is the f64 log10 (CUDA's libdevice vs glibc): both are within 1 ulp of the true
// value, so the f32 results can differ only when the true value lies within ~1e-15 of an f32 rounding boundary (tests/test_gpu_parity.py
// asserts equality on the golden clips and states the bound).
//
// One warp per frame, frames of a clip side by side in a CTA; all intermediate vectors live in shared memory.
#include "mel_kernels.cuh"

#include <algorithm>
#include <cstring>

namespace wb200 {

namespace {

constexpr int kN = 400, kHop = 160, kBins = 201, kLeaf = 25, kSub = 16;
constexpr int kWarpsPerCta = 8;

struct FrameScratch {
    float in[kN];
    float are[kN], aim[kN], bre[kN], bim[kN];
    float power[kBins + 3];
};
// twiddles and window of the CTA in shared memory: the leaf DFT reads 26 twiddles per input sample and lane (two distinct addresses per
// warp-wide load: a broadcast), which as global loads was most of the kernel's time
struct MelSmemTables {
    float hann[kN];
    float leaf_cos[kLeaf * kLeaf], leaf_sin[kLeaf * kLeaf];
    float tw_re[4 * 200], tw_im[4 * 200];
};
constexpr int kMelSmem = (int) (sizeof(MelSmemTables) + kWarpsPerCta * sizeof(FrameScratch));

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_logmel_frames(const MelDevTables Tg, const MelClip * __restrict__ clips) {
    extern __shared__ __align__(16) uint8_t mel_smem[];
    MelSmemTables & T = *(MelSmemTables *) mel_smem;
    FrameScratch * scratch = (FrameScratch *) (mel_smem + sizeof(MelSmemTables));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = threadIdx.x; j < kN; j += blockDim.x) T.hann[j] = __ldg(Tg.hann + j);
    for (int j = threadIdx.x; j < kLeaf * kLeaf; j += blockDim.x) { T.leaf_cos[j] = __ldg(Tg.leaf_cos + j); T.leaf_sin[j] = __ldg(Tg.leaf_sin + j); }
    for (int j = threadIdx.x; j < 4 * 200; j += blockDim.x) { T.tw_re[j] = __ldg(Tg.tw_re + j); T.tw_im[j] = __ldg(Tg.tw_im + j); }
    __syncthreads();
    const MelClip c = clips[blockIdx.y];
    const int i = blockIdx.x * kWarpsPerCta + warp;            // frame
    if (i >= c.n_calc) return;                                  // (whole warps leave; no block-wide barrier below)
    FrameScratch & S = scratch[warp];
    const int n = c.n_samples;

    // the padded signal of the reference — 200 reflected samples in front, zeros behind — is never materialised
    const int offset = i * kHop;
    for (int j = lane; j < kN; j += 32) {
        const int x = offset + j;
        float v = 0.0f;
        if (x < kN / 2) { const int s = kN / 2 - x; if (s < n) v = __ldg(c.pcm + s); }
        else            { const int s = x - kN / 2; if (s < n) v = __ldg(c.pcm + s); }
        S.in[j] = __fmul_rn(T.hann[j], v);
    }
    __syncwarp();

    // leaf DFTs: output (k, r), k < 25, r < 16, stored as sequence r, index k.  Lane l owns r = l & 15 and the bins k = (l >> 4) + 2 t:
    // every x[16 n + r] it loads feeds all of its (up to 13) bins, whose accumulation chains are independent of each other.
    {
        const int r = lane & 15, k0 = lane >> 4;
        constexpr int kMine = 13;
        float re[kMine], im[kMine];
#pragma unroll
        for (int t = 0; t < kMine; ++t) { re[t] = 0.0f; im[t] = 0.0f; }
#pragma unroll 2
        for (int nn = 0; nn < kLeaf - 1; ++nn) {
            const float x = S.in[nn * kSub + r];
#pragma unroll
            for (int t = 0; t < kMine; ++t) {
                const int k = k0 + 2 * t;
                if (k < kLeaf) {
                    re[t] = __fadd_rn(re[t], __fmul_rn(x, T.leaf_cos[k * kLeaf + nn]));
                    im[t] = __fsub_rn(im[t], __fmul_rn(x, T.leaf_sin[k * kLeaf + nn]));
                }
            }
        }
        const float x = S.in[(kLeaf - 1) * kSub + r];
#pragma unroll
        for (int t = 0; t < kMine; ++t) {
            const int k = k0 + 2 * t;
            if (k < kLeaf) {
                S.are[r * kLeaf + k] = __fmaf_rn(x, T.leaf_cos[k * kLeaf + kLeaf - 1], re[t]);
                S.aim[r * kLeaf + k] = __fmaf_rn(-x, T.leaf_sin[k * kLeaf + kLeaf - 1], im[t]);
            }
        }
    }
    __syncwarp();

    // butterflies: at a level with nseq input sequences of length len, output sequence q (< nseq / 2) combines even = input q and
    // odd = input q + nseq / 2
    float * sre = S.are, * sim = S.aim, * dre = S.bre, * dim = S.bim;
    int nseq = kSub, len = kLeaf;
#pragma unroll 1
    for (int l = 0; l < 4; ++l) {
        const int half = nseq >> 1;
        const float * wr = T.tw_re + l * 200, * wi = T.tw_im + l * 200;
        for (int e = lane; e < half * len; e += 32) {
            const int q = e / len, k = e - q * len;
            const float re = wr[k], im = wi[k];
            const float er = sre[q * len + k], ei = sim[q * len + k];
            const float ro = sre[(q + half) * len + k], io = sim[(q + half) * len + k];
            float * o_r = dre + q * 2 * len, * o_i = dim + q * 2 * len;
            o_r[k]       = __fmaf_rn(-im, io, __fmaf_rn(re, ro, er));
            o_i[k]       = __fmaf_rn(im, ro, __fmaf_rn(re, io, ei));
            o_r[k + len] = __fmaf_rn(im, io, __fmaf_rn(-re, ro, er));
            o_i[k + len] = __fmaf_rn(-im, ro, __fmaf_rn(-re, io, ei));
        }
        __syncwarp();
        float * t = sre; sre = dre; dre = t;
        t = sim; sim = dim; dim = t;
        nseq = half;
        len <<= 1;
    }
    for (int j = lane; j < kBins; j += 32) {
        const float re = sre[j], im = sim[j];
        S.power[j] = __fmaf_rn(re, re, __fmul_rn(im, im));
    }
    __syncwarp();

    // mel filter bank + log10
    float vmax = -1e20f;
    const float * P = S.power;
    for (int j = lane; j < Tg.n_mel; j += 32) {
        const float * F = Tg.filt + (size_t) j * kBins;
        double sum = 0.0;
        const int g1 = __ldg(Tg.g1 + j);
        for (int g = __ldg(Tg.g0 + j); g < g1; ++g) {
            const int k = 4 * g;
            float part = __fmul_rn(P[k + 1], __ldg(F + k + 1));
            part = __fmaf_rn(P[k + 0], __ldg(F + k + 0), part);
            part = __fmaf_rn(P[k + 2], __ldg(F + k + 2), part);
            part = __fmaf_rn(P[k + 3], __ldg(F + k + 3), part);
            sum += (double) part;
        }
        sum += (double) __fmul_rn(P[200], __ldg(F + 200));
        sum = fmax(sum, 1e-10);
        const float lg = (float) log10(sum);
        c.raw[(size_t) i * Tg.n_mel + j] = lg;
        vmax = fmaxf(vmax, lg);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0) atomicMax(c.max_bits, mel_float_to_ordered(vmax));
}

__global__ void k_mel_reset_max(const MelClip * __restrict__ clips, int n) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n) *clips[b].max_bits = mel_float_to_ordered(-1e20f);
}

__global__ void __launch_bounds__(256)
k_mel_window(const MelWindow * __restrict__ wins, int n_mel, int n_frames, float low) {
    const MelWindow w = wins[blockIdx.y];
    // clamping and normalisation (whisper.cpp:2856-2871): the maximum runs over all frames, the constant tail included
    float fmax = mel_ordered_to_float(*w.max_bits);
    if (w.n_calc < w.n_len) fmax = fmaxf(fmax, low);
    const double mmax = (double) fmax - 8.0;
    const float fclamp = (float) mmax;
    const int total = (n_frames + 2) * n_mel;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int row = e / n_mel, m = e - row * n_mel;
        float out = 0.0f;                                       // rows 0 and n_frames + 1: the conv's zero padding
        if (row >= 1 && row <= n_frames) {
            const int i = w.offset + row - 1;
            if (i < w.n_len) {                                  // beyond the spectrogram the window stays zero (whisper.cpp:1692-1706)
                float v = i < w.n_calc ? w.raw[(size_t) i * n_mel + m] : low;
                if ((double) v < mmax) v = fclamp;
                out = (float) (((double) v + 4.0) / 4.0);
            }
        }
        w.out[e] = __float2half_rn(out);
    }
}

// out[i] = (sum over j = -hw .. hw, inside the clip, of |x[i + j]|, added in that order in f32) / (2 hw + 1)   — whisper.cpp:6350-6366
__global__ void __launch_bounds__(256)
k_signal_energy(const EnergyClip * __restrict__ clips, int hw) {
    const EnergyClip c = clips[blockIdx.y];
    const float denom = (float) (2 * hw + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c.n_samples; i += gridDim.x * blockDim.x) {
        float sum = 0.0f;
        const int j0 = max(0, i - hw), j1 = min(c.n_samples - 1, i + hw);
        for (int j = j0; j <= j1; ++j) sum = __fadd_rn(sum, fabsf(__ldg(c.pcm + j)));
        c.out[i] = __fdiv_rn(sum, denom);
    }
}

}  // namespace

void launch_signal_energy(const EnergyClip * clips_dev, int n_clips, int max_samples, int hw, cudaStream_t st) {
    if (n_clips <= 0 || max_samples <= 0) return;
    dim3 grid(std::min(512, (max_samples + 255) / 256), n_clips);
    k_signal_energy<<<grid, 256, 0, st>>>(clips_dev, hw);
}

void launch_logmel_frames(const MelDevTables & T, const MelClip * clips_dev, int n_clips, int max_calc, cudaStream_t st) {
    if (n_clips <= 0 || max_calc <= 0) return;
    dim3 grid((max_calc + kWarpsPerCta - 1) / kWarpsPerCta, n_clips);
    static bool attr_done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 15]) { cudaFuncSetAttribute(k_logmel_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, kMelSmem); attr_done[dev & 15] = true; }
    k_mel_reset_max<<<(n_clips + 127) / 128, 128, 0, st>>>(clips_dev, n_clips);
    k_logmel_frames<<<grid, kWarpsPerCta * 32, kMelSmem, st>>>(T, clips_dev);
}

void launch_mel_window(const MelWindow * wins_dev, int n_wins, int n_mel, int n_frames, float low, cudaStream_t st) {
    if (n_wins <= 0) return;
    const int total = (n_frames + 2) * n_mel;
    dim3 grid(std::min(64, (total + 255) / 256), n_wins);
    k_mel_window<<<grid, 256, 0, st>>>(wins_dev, n_mel, n_frames, low);
}

}  // namespace wb200
