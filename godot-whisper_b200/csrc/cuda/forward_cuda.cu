// CudaForward — the one and only implementation of wb200::Forward shipped in libwhisper_b200.so.
//
// Device-side restatement of whisper_encode_internal / whisper_decode_internal
// (/root/reference/thirdparty/whisper.cpp/whisper.cpp:2086-2146, 2517-2595 and the four graph builders :1660-2505).
// Layouts in HBM (all token-major, features contiguous; T = n_ctx, Tp = T rounded up to 8 so rows stay 16-byte aligned):
//   weights      f16 [out][in] exactly as in the ggml file; Q/K/V stacked to [3d][d]; cross K/V stacked to [2d][d];
//                conv kernels re-ordered to [out][tap][in] so the conv is an implicit GEMM over overlapping rows
//   residual x   f32 [B][T][d]          operands of the next GEMM  f16 [B][T][d] / [B][T][4d]
//   S, P         f32 / f16 [B][h][T][Tp]     V^T  f16 [B][d][Tp]
//   per slot     cross K f16 [Lt][Tmax][d], cross V^T f16 [Lt][d][Tpmax], self K f16 [Lt][cells][d], self V^T f16 [Lt][d][cells]
#include "../common.h"
#include "../forward.h"
#include "../tables.h"
#include "dev.cuh"
#include "kernels.cuh"
#include "gemm_enc.cuh"
#include "decode_step.cuh"
#include "run_kernels.cuh"
#include "mel_kernels.cuh"
#include "../mel.h"

#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <sched.h>
#include <condition_variable>
#include <map>
#include <tuple>
#include <string>
#include <vector>

namespace wb200 {

void gemm_tc_forget_maps();

namespace {

#define CUDA_OK(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            WB_LOG_ERROR("%s:%d: %s failed: %s\n", __FILE__, __LINE__, #expr, cudaGetErrorString(e_));      \
            return false;                                                                                   \
        }                                                                                                   \
    } while (0)

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct DevBuf {
    void * p = nullptr;
    size_t bytes = 0;
    bool ensure(size_t need) {
        if (need <= bytes) return true;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        if (cudaMalloc(&p, need) != cudaSuccess) { WB_LOG_ERROR("cudaMalloc(%zu) failed\n", need); return false; }
        bytes = need;
        cudaMemset(p, 0, need);
        cudaDeviceSynchronize();      // the arena must be zero before any stream of this context touches it
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T * as() const { return (T *) p; }
};

struct PinnedBuf {
    void * p = nullptr;
    size_t bytes = 0;
    bool ensure(size_t need) {
        if (need <= bytes) return true;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        if (cudaMallocHost(&p, need) != cudaSuccess) { WB_LOG_ERROR("cudaMallocHost(%zu) failed\n", need); return false; }
        bytes = need;
        return true;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    template <class T> T * as() const { return (T *) p; }
};

struct EncLayerW {
    const float * ln1_g, * ln1_b, * ln2_g, * ln2_b;
    const __half * wqkv; const float * bqkv;      // [3d][d], [3d] (K part of the bias is zero: the reference has none)
    const __half * wo;   const float * bo;
    const __half * w1;   const float * b1;
    const __half * w2;   const float * b2;
};
struct DecLayerW {
    const float * ln1_g, * ln1_b, * lnc_g, * lnc_b, * ln2_g, * ln2_b;
    const __half * wqkv; const float * bqkv;
    const __half * wo;   const float * bo;
    const __half * wcq;  const float * bcq;
    const __half * wckv; const float * bckv;      // [2d][d] cross K (no bias) + cross V
    const __half * wco;  const float * bco;
    const __half * w1;   const float * b1;
    const __half * w2;   const float * b2;
};

class CudaForward : public Forward {
public:
    int device = 0;
    cudaStream_t st = nullptr;
    HParams hp;
    int kv_cells = 0;
    int engine = 0;                 // 0 = tcgen05, 1 = SIMT cross-check
    std::atomic<int64_t> launches{0};
    std::string name_;

    // weights
    DevBuf wbuf;
    std::vector<EncLayerW> enc;
    std::vector<DecLayerW> dec;
    const __half * conv1_w = nullptr, * conv2_w = nullptr, * d_te = nullptr;
    const float * conv1_b = nullptr, * conv2_b = nullptr, * e_pe = nullptr, * e_ln_g = nullptr, * e_ln_b = nullptr;
    const float * d_pe = nullptr, * d_ln_g = nullptr, * d_ln_b = nullptr;
    const uint16_t * gelu_lut = nullptr, * exp_lut = nullptr;
    const uint8_t * cls_tab = nullptr;
    int token_beg = 0, token_eot = 0;

    // per-slot state
    int slots = 0;
    DevBuf cross_k, cross_v, self_k, self_v;
    int64_t cross_k_slot = 0, cross_v_slot = 0, self_k_slot = 0, self_v_slot = 0;   // elements per slot
    int Tmax = 0, Tpmax = 0;
    std::vector<int> slot_n_ctx;

    // encoder workspace (capacity enc_cap chunks)
    int enc_cap = 0;
    DevBuf mel_d, melT, act1, conv16, x32, xn16, q16, k16, vt16, S32, P16, attn16, h16, enc32;
    PinnedBuf mel_h;

    // log-mel spectrogram on the device (mel_kernels.cu): tables, per-slot raw spectrogram + maximum, PCM of the clips of an encoder pass,
    // and a small pool of pinned staging buffers the chunk workers copy their PCM into (pcm_stage_acquire / release)
    static constexpr int kPcmCap = 30 * 16000;                 // samples of the longest clip the device path takes (one 30 s window)
    static constexpr int kMelFramesCap = kPcmCap / 160 + 8;    // frames that can overlap samples
    static constexpr int kPcmStages = 64;
    bool mel_dev_on = true;
    MelDevTables meltab;
    int filt_n_mel = 0;
    float mel_low = -10.0f;
    // (two sets of everything an encoder pass copies in or out: the PCM of pass k + 1 is uploaded on st_h2d while pass k computes)
    DevBuf raw_mel, mel_max, pcm_d2[2], clips_d2[2], wins_d2[2], energy_d2[2], eclips_d2[2], slotmap_d2[2];
    PinnedBuf clips_h2[2], wins_h2[2], pcm_pool, eclips_h2[2], energy_pool, slotmap_h2[2];
    cudaStream_t st_h2d = nullptr;
    cudaEvent_t ev_h2d[2] = {}, ev_pcm_read[2] = {}, ev_energy2[2] = {}, ev_e2h2[2] = {}, ev_enc0s[2] = {}, ev_enc1s[2] = {}, ev_mel1s[2] = {};
    bool e2h_pending2[2] = {false, false}, enc_pending[2] = {false, false};
    std::vector<int> mel_n_calc, mel_n_len;
    std::mutex pool_mu;
    std::condition_variable pool_cv;
    std::vector<float *> pool_free;
    bool pool_ready = false;

    // decoder workspace (capacity dec_cap rows)
    int dec_cap = 0;
    DevBuf dx32, dxn16, dq16, dattn16, dh16, dxw32, dlogits, dstage, dsampled;
    // wide passes of staging set 1 run on their own stream with their own activations: the two alternating passes of the batcher
    // belong to different sequences, so their (latency-bound, few-CTA) kernels overlap on the device
    struct DecAct { DevBuf x32, xn16, q16, attn16, h16, xw32, logits; } act_b;
    cudaStream_t st_dec2 = nullptr;
    PinnedBuf hstage, hlogits, hsampled, hdist;
    DevBuf ddist;
    // second staging set: lets the next decode-step launch be staged and queued while the previous one still runs
    DevBuf dstage2, dsampled2;
    PinnedBuf hstage2, hsampled2;
    cudaEvent_t ev2_call0 = nullptr, ev2_call1 = nullptr;
    struct PendingPass { bool active = false; std::vector<DecodeJob> jobs; int n_full = 0, n_samp = 0, n_dist = 0; };
    PendingPass pend[2];

    // device-resident greedy runs (run_state.h, run_kernels.cu): per-slot sequence state and sampled tokens, the staging block the
    // prep kernel fills (staging set 2), and a ring of kRunRing queued steps (row lists in, statuses out, events)
    static constexpr int kRunRing = 8;
    DevBuf run_seqs, run_tokens, run_rows_d, run_status_d, dstage_run, dsampled_run;
    PinnedBuf run_init_h, run_fetch_h, run_rows_h[kRunRing], run_status_h[kRunRing];
    cudaEvent_t run_ev0[kRunRing] = {}, run_ev1[kRunRing] = {};
    int run_rows_of[kRunRing] = {};
    int64_t run_ticket = 0;
    std::vector<int> run_pos_ub;          // per slot: upper bound of RunSeq::pos on the device (start + steps queued)
    cudaStream_t st_copy = nullptr;       // result fetches of finished runs, next to the steps still queued on st
    int run_rows = 512, run_depth_ = 3;
    bool runs_on = true;
    bool blocking_sync = false;
    unsigned ev_flags = cudaEventDefault;
    cudaEvent_t ev_copy = nullptr, ev_energy = nullptr, ev_e2h = nullptr;
    cudaStream_t st_e2h = nullptr;        // energy envelopes back to the host, under the encoder kernels
    bool e2h_pending = false;

    // persistent decode-step kernel (decode_step.cu): device copy of the layer table, sampler partials, grid barrier words
    DevBuf step_plans, step_records, step_bar, step_trace;
    int    step_trace_groups = 1;
    int    step_groups_max = 2;       // independent row groups per launch (WHISPER_B200_STEP_GROUPS): 16 rows each
    int    step_n_phases = 0, step_slot = 0, step_chunk_keys = 0, step_chunk_keys_cross = 0;
    alignas(64) CUtensorMap step_tm_ck, step_tm_cv, step_tm_te;         // cross-attention K / V^T of all slots (rebuilt when the slots move)
    bool   fused_attn = true;         // encoder attention as one tcgen05 kernel (WHISPER_B200_FUSED_ATTN=0: three launches, scores through HBM)
    bool   use_step = true;
    int    wide_rows = 0;             // rows per decoder pass when many chunks decode at once (WHISPER_B200_DECODE_ROWS): the
                                      // multi-kernel path reads the weights once for all of them; 0 = decode-step passes only
    int    step_rows_max = 0;         // passes of up to this many rows take the decode-step kernel (WHISPER_B200_STEP_MAX_ROWS)
    int    step_grid = 0, step_xs = 0;
    size_t step_smem = 0;
    int64_t n_step_launches = 0;
    double  step_bytes_total = 0.0;    // algorithmic bytes of all decode-step launches (weights once per launch + cross-KV per row)

    // ---- device clocks ---------------------------------------------------------------------------------------------
    cudaEvent_t ev_call0 = nullptr, ev_call1 = nullptr, ev_enc0 = nullptr, ev_enc1 = nullptr;
    cudaStream_t st_enc = nullptr;     // encoder passes (own host thread in the batcher)
    double h2d_bytes_enc = 0.0;
    bool encoder_concurrent() const override { return !prof_on && st_enc != nullptr && !serial_enc; }
    bool serial_enc = false;           // WHISPER_B200_ENC_STREAM=0: encoder passes share the decoder's stream and driver thread
    double t_enc_ms = 0.0, t_mel_ms = 0.0, t_dec_ms = 0.0, h2d_bytes = 0.0, d2h_bytes = 0.0;
    int64_t n_enc_calls = 0, n_dec_calls = 0;
    bool prof_on = false;
    int  prof_kind = PROF_MISC;
    struct ProfRec { cudaEvent_t a, b; int kind; double flop, bytes; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_acc[PROF_KINDS][4] = {};

    cudaEvent_t prof_event() {
        if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    // brackets the launches issued between begin/end with an event pair (only while profiling)
    void prof_begin(int kind, double flop, double bytes) {
        if (!prof_on) return;
        prof_kind = kind;
        ProfRec r{prof_event(), prof_event(), kind, flop, bytes};
        cudaEventRecord(r.a, st);
        prof_pending.push_back(r);
    }
    void prof_end() {
        if (!prof_on || prof_pending.empty()) return;
        cudaEventRecord(prof_pending.back().b, st);
    }
    void prof_collect() {      // call after the stream has been synchronised
        for (auto & r : prof_pending) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
                prof_acc[r.kind][0] += 1.0; prof_acc[r.kind][1] += ms; prof_acc[r.kind][2] += r.flop; prof_acc[r.kind][3] += r.bytes;
            }
            prof_pool.push_back(r.a); prof_pool.push_back(r.b);
        }
        prof_pending.clear();
    }
    // Device-busy time: the union of the [start, end] intervals of all passes (encoder and decoder passes overlap on two streams,
    // so the sum of their durations counts the overlap twice).  Interval ends are event timestamps relative to ev_base.
    cudaEvent_t ev_base = nullptr;
    mutable std::mutex busy_mu;
    mutable std::vector<std::pair<float, float>> busy_iv;
    mutable double busy_done_ms = 0.0;     // union length of the intervals already folded away (all of them end before busy_iv starts)
    void note_busy(cudaEvent_t a, cudaEvent_t b) {
        float t0 = 0.0f, t1 = 0.0f;
        if (!ev_base || cudaEventElapsedTime(&t0, ev_base, a) != cudaSuccess || cudaEventElapsedTime(&t1, ev_base, b) != cudaSuccess) return;
        std::lock_guard<std::mutex> g(busy_mu);
        busy_iv.emplace_back(t0, t1);
    }
    double busy_ms() const override {
        std::lock_guard<std::mutex> g(busy_mu);
        std::sort(busy_iv.begin(), busy_iv.end());
        if (getenv("WHISPER_B200_HOST_TRACE") && !busy_iv.empty()) {
            // where the device idles inside the span of these passes: ten equal slices of the span, idle milliseconds in each
            const float t_lo = busy_iv.front().first;
            float t_hi = t_lo;
            for (const auto & iv : busy_iv) t_hi = std::max(t_hi, iv.second);
            double idle[10] = {0};
            float cur = t_lo;
            auto add_idle = [&](float a, float b) {
                for (int k = 0; k < 10; ++k) {
                    const float s0 = t_lo + (t_hi - t_lo) * k / 10.0f, s1 = t_lo + (t_hi - t_lo) * (k + 1) / 10.0f;
                    idle[k] += std::max(0.0f, std::min(b, s1) - std::max(a, s0));
                }
            };
            for (const auto & iv : busy_iv) { if (iv.first > cur) add_idle(cur, iv.first); cur = std::max(cur, iv.second); }
            fprintf(stderr, "device: %zu passes over %.1f ms; idle ms per tenth of the span:", busy_iv.size(), t_hi - t_lo);
            for (int k = 0; k < 10; ++k) fprintf(stderr, " %.1f", idle[k]);
            fprintf(stderr, "\n");
        }
        double total = busy_done_ms;
        std::vector<std::pair<float, float>> merged;
        for (const auto & iv : busy_iv) {
            if (!merged.empty() && iv.first <= merged.back().second) merged.back().second = std::max(merged.back().second, iv.second);
            else merged.push_back(iv);
        }
        for (const auto & iv : merged) total += (double) iv.second - iv.first;
        // fold: called between batches (nothing in flight), so no later interval can start before the last merged one ends
        if (!merged.empty()) { busy_done_ms = total; busy_iv.clear(); }
        return total;
    }
    double mel_ms() const override { return t_mel_ms; }
    void gpu_times(double * out) const override { out[0] = t_enc_ms; out[1] = t_dec_ms; out[2] = (double) n_enc_calls; out[3] = (double) n_dec_calls; out[4] = h2d_bytes + h2d_bytes_enc; out[5] = d2h_bytes; out[6] = (double) n_step_launches; out[7] = step_bytes_total; }
    void set_profiling(bool on) override { prof_on = on; if (on) memset(prof_acc, 0, sizeof(prof_acc)); }
    void profile(double * out) const override { memcpy(out, prof_acc, sizeof(prof_acc)); }

    ~CudaForward() override {
        cudaSetDevice(device);
        if (st) cudaStreamSynchronize(st);
        if (st_enc) { cudaStreamSynchronize(st_enc); cudaStreamDestroy(st_enc); }
        if (st_dec2) { cudaStreamSynchronize(st_dec2); cudaStreamDestroy(st_dec2); }
        for (DevBuf * b : {&act_b.x32, &act_b.xn16, &act_b.q16, &act_b.attn16, &act_b.h16, &act_b.xw32, &act_b.logits}) b->release();
        if (st_e2h) { cudaStreamSynchronize(st_e2h); cudaStreamDestroy(st_e2h); }
        if (ev_energy) cudaEventDestroy(ev_energy);
        if (ev_e2h) cudaEventDestroy(ev_e2h);
        if (ev_copy) cudaEventDestroy(ev_copy);
        if (ev_enc0) cudaEventDestroy(ev_enc0);
        if (ev_enc1) cudaEventDestroy(ev_enc1);
        if (ev_base) cudaEventDestroy(ev_base);
        drop_graphs();
        for (cudaEvent_t e : prof_pool) cudaEventDestroy(e);
        if (ev_call0) cudaEventDestroy(ev_call0);
        if (ev_call1) cudaEventDestroy(ev_call1);
        for (DevBuf * b : {&wbuf, &cross_k, &cross_v, &self_k, &self_v, &mel_d, &melT, &act1, &conv16, &x32, &xn16, &q16, &k16,
                           &vt16, &S32, &P16, &attn16, &h16, &enc32, &dx32, &dxn16, &dq16, &dattn16, &dh16, &dxw32, &dlogits,
                           &dstage, &dsampled, &dstage2, &dsampled2, &step_plans, &step_records, &step_bar, &step_trace}) b->release();
        for (DevBuf * b : {&raw_mel, &mel_max}) b->release();
        for (int i = 0; i < 2; ++i) {
            for (DevBuf * b : {&pcm_d2[i], &clips_d2[i], &wins_d2[i], &energy_d2[i], &eclips_d2[i], &slotmap_d2[i]}) b->release();
            clips_h2[i].release(); wins_h2[i].release(); eclips_h2[i].release(); slotmap_h2[i].release();
            for (cudaEvent_t e : {ev_h2d[i], ev_pcm_read[i], ev_energy2[i], ev_e2h2[i], ev_enc0s[i], ev_enc1s[i], ev_mel1s[i]}) if (e) cudaEventDestroy(e);
        }
        if (st_h2d) { cudaStreamSynchronize(st_h2d); cudaStreamDestroy(st_h2d); }
        pcm_pool.release(); energy_pool.release();
        if (st_copy) { cudaStreamSynchronize(st_copy); cudaStreamDestroy(st_copy); }
        for (DevBuf * b : {&run_seqs, &run_tokens, &run_rows_d, &run_status_d, &dstage_run, &dsampled_run}) b->release();
        run_init_h.release(); run_fetch_h.release();
        for (int i = 0; i < kRunRing; ++i) {
            run_rows_h[i].release(); run_status_h[i].release();
            if (run_ev0[i]) cudaEventDestroy(run_ev0[i]);
            if (run_ev1[i]) cudaEventDestroy(run_ev1[i]);
        }
        hdist.release(); ddist.release();
        mel_h.release(); hstage.release(); hlogits.release(); hsampled.release(); hstage2.release(); hsampled2.release();
        if (ev2_call0) cudaEventDestroy(ev2_call0);
        if (ev2_call1) cudaEventDestroy(ev2_call1);
        gemm_tc_forget_maps(); gemm_enc_forget_maps();
        if (st) cudaStreamDestroy(st);
    }

    const char * name() const override { return name_.c_str(); }
    int64_t kernel_launches() const override { return launches.load(); }
    void set_gemm_engine(int e) override { engine = e == 1 ? 1 : 0; force_multi = e == 2; }
    int n_slots() const override { return slots; }
    bool can_sample() const override { return true; }
    bool can_sample_dist() const override { return true; }

    // ---- init ------------------------------------------------------------------------------------------------------

    bool init(const ModelFile & mf, int kv_self_cells, int dev) {
        int n_dev = 0;
        if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
            WB_LOG_ERROR("%s: no CUDA device available - this backend has no CPU fallback\n", __func__);
            return false;
        }
        if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
        if (dev >= n_dev) { WB_LOG_ERROR("%s: device %d out of range (%d devices)\n", __func__, dev, n_dev); return false; }
        device = dev;
        CUDA_OK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) {
            WB_LOG_ERROR("%s: device %d (%s) is sm_%d%d; this library carries sm_100a kernels only\n", __func__, device, prop.name,
                         prop.major, prop.minor);
            return false;
        }
        name_ = std::string("CUDA sm_100a tcgen05/TMA on ") + prop.name;
        {
            int pr_lo = 0, pr_hi = 0;
            CUDA_OK(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
            CUDA_OK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, pr_hi));          // decoder passes: latency-bound, scheduled first
            CUDA_OK(cudaStreamCreateWithPriority(&st_enc, cudaStreamNonBlocking, pr_lo));
            CUDA_OK(cudaStreamCreateWithPriority(&st_dec2, cudaStreamNonBlocking, pr_hi));
            if (const char * e = getenv("WHISPER_B200_DEC_STREAMS")) { if (atoi(e) < 2) { cudaStreamDestroy(st_dec2); st_dec2 = nullptr; } }
            // Host threads that wait for the device sleep instead of spinning when cores are scarce (several ranks on one box: two
            // driver threads per rank spinning would take the cores the chunk workers need); with cores to spare they spin, which
            // wakes a few microseconds sooner.  WHISPER_B200_BLOCKING_SYNC overrides.
            {
                int n_cpu = (int) std::thread::hardware_concurrency();
                cpu_set_t set;
                if (sched_getaffinity(0, sizeof(set), &set) == 0) n_cpu = CPU_COUNT(&set);
                int ranks = 1;
                if (const char * e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
                blocking_sync = n_cpu / ranks < 8;
                if (const char * e = getenv("WHISPER_B200_BLOCKING_SYNC")) blocking_sync = atoi(e) != 0;
            }
            ev_flags = blocking_sync ? cudaEventBlockingSync : cudaEventDefault;
            CUDA_OK(cudaEventCreateWithFlags(&ev_enc0, ev_flags));
            CUDA_OK(cudaEventCreateWithFlags(&ev_enc1, ev_flags));
            CUDA_OK(cudaEventCreateWithFlags(&ev_copy, ev_flags));
            CUDA_OK(cudaEventCreateWithFlags(&ev_energy, cudaEventDisableTiming));
            CUDA_OK(cudaEventCreateWithFlags(&ev_e2h, ev_flags | cudaEventDisableTiming));
            CUDA_OK(cudaStreamCreateWithFlags(&st_e2h, cudaStreamNonBlocking));
            CUDA_OK(cudaStreamCreateWithFlags(&st_h2d, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                CUDA_OK(cudaEventCreateWithFlags(&ev_h2d[i], cudaEventDisableTiming));
                CUDA_OK(cudaEventCreateWithFlags(&ev_pcm_read[i], cudaEventDisableTiming));
                CUDA_OK(cudaEventCreateWithFlags(&ev_energy2[i], cudaEventDisableTiming));
                CUDA_OK(cudaEventCreateWithFlags(&ev_e2h2[i], ev_flags | cudaEventDisableTiming));
                CUDA_OK(cudaEventCreateWithFlags(&ev_enc0s[i], ev_flags));
                CUDA_OK(cudaEventCreateWithFlags(&ev_enc1s[i], ev_flags));
                CUDA_OK(cudaEventCreateWithFlags(&ev_mel1s[i], ev_flags));
            }
            CUDA_OK(cudaEventCreate(&ev_base));
            CUDA_OK(cudaEventRecord(ev_base, st));
            if (const char * e = getenv("WHISPER_B200_ENC_STREAM")) serial_enc = atoi(e) == 0;
        }
        CUDA_OK(cudaStreamCreateWithFlags(&st_copy, cudaStreamNonBlocking));
        for (int i = 0; i < kRunRing; ++i) { CUDA_OK(cudaEventCreateWithFlags(&run_ev0[i], ev_flags)); CUDA_OK(cudaEventCreateWithFlags(&run_ev1[i], ev_flags)); }
        if (const char * e = getenv("WHISPER_B200_RUNS")) runs_on = atoi(e) != 0;
        if (const char * e = getenv("WHISPER_B200_DEVICE_MEL")) mel_dev_on = atoi(e) != 0;
        if (const char * e = getenv("WHISPER_B200_RUN_ROWS")) run_rows = std::min(1024, std::max(1, atoi(e)));
        if (const char * e = getenv("WHISPER_B200_RUN_DEPTH")) run_depth_ = std::min(kRunRing - 1, std::max(1, atoi(e)));
        CUDA_OK(cudaEventCreateWithFlags(&ev_call0, ev_flags));
        CUDA_OK(cudaEventCreateWithFlags(&ev_call1, ev_flags));
        CUDA_OK(cudaEventCreateWithFlags(&ev2_call0, ev_flags));
        CUDA_OK(cudaEventCreateWithFlags(&ev2_call1, ev_flags));
        if (const char * e = getenv("WHISPER_B200_GEMM_ENGINE")) set_gemm_engine(atoi(e));
        if (const char * e = getenv("WHISPER_B200_GRAPHS")) use_graphs = atoi(e) != 0;
        if (const char * e = getenv("WHISPER_B200_STEP_KERNEL")) use_step = atoi(e) != 0;
        if (const char * e = getenv("WHISPER_B200_STEP_GROUPS")) step_groups_max = std::min(2, std::max(1, atoi(e)));    // 3 and 4 groups are not validated yet
        if (const char * e = getenv("WHISPER_B200_FUSED_ATTN")) fused_attn = atoi(e) != 0;
        wide_rows = 256;
        if (const char * e = getenv("WHISPER_B200_DECODE_ROWS")) wide_rows = std::min(1024, std::max(0, atoi(e)));
        step_rows_max = kStepMaxRows * step_groups_max;
        if (const char * e = getenv("WHISPER_B200_STEP_MAX_ROWS")) step_rows_max = std::min(step_rows_max, std::max(0, atoi(e)));

        hp = mf.hparams;
        kv_cells = kv_self_cells;
        if (hp.n_audio_state / hp.n_audio_head != 64 || hp.n_text_state / hp.n_text_head != 64) {
            WB_LOG_ERROR("%s: head size must be 64\n", __func__);
            return false;
        }
        if (hp.n_audio_state != hp.n_text_state) {
            WB_LOG_ERROR("%s: n_audio_state != n_text_state is not supported\n", __func__);
            return false;
        }
        Tmax = hp.n_audio_ctx;
        Tpmax = (int) align_up(Tmax, 8);
        if (!upload_weights(mf)) return false;
        if (!ensure_slots(1)) return false;
        if (!init_step_kernel()) return false;
        CUDA_OK(cudaStreamSynchronize(st));
        return true;
    }

    // One-time setup of the persistent decode-step kernel; leaves step_grid == 0 (multi-kernel path only) when the model is
    // too wide for its shared-memory plan or the device cannot co-schedule one CTA per SM.
    bool init_step_kernel() {
        step_grid = 0;
        // the plans point into the decoder workspace; it is sized once for the widest pass so that it never moves while a
        // pass is queued on the other staging set
        if (dec_cap == 0 && !ensure_dec(std::max(std::max(kStepMaxRows * step_groups_max, wide_rows), runs_on ? run_rows : 0))) return false;
        step_smem = decode_step_smem_bytes(hp.n_text_state, &step_xs, &step_slot, &step_chunk_keys);
        if (step_smem == 0 || hp.n_audio_ctx > 1536 || kv_cells > 1536 || 3 + 8 * hp.n_text_layer > kStepMaxPhases) {
            WB_LOG_INFO("%s: decode-step kernel not used for this model (n_text_state %d, %d layers)\n", __func__, hp.n_text_state, hp.n_text_layer);
            return true;
        }
        step_grid = decode_step_grid(step_smem);
        if (step_grid <= 0) { step_grid = 0; return true; }
        return build_step_plans() && build_step_maps();
    }

    // One phase table per row count (the geometry depends on the model, the grid and n only); rebuilt when the decoder
    // workspace moves.
    bool build_step_plans() {
        if (step_grid <= 0) return true;
        std::vector<StepLayerW> lw(hp.n_text_layer);
        for (int i = 0; i < hp.n_text_layer; ++i) {
            const DecLayerW & L = dec[i];
            lw[i] = StepLayerW{L.ln1_g, L.ln1_b, L.lnc_g, L.lnc_b, L.ln2_g, L.ln2_b, L.wqkv, L.bqkv, L.wo, L.bo, L.wcq, L.bcq, L.wco, L.bco,
                               L.w1, L.b1, L.w2, L.b2};
        }
        // plans[(G - 1)][n]: the table for n rows served by step_grid / G CTAs
        std::vector<StepPhase> plans((size_t) kStepMaxGroups * (kStepMaxRows + 1) * kStepMaxPhases);
        for (int G = 1; G <= step_groups_max; ++G) {
            for (int n = 1; n <= kStepMaxRows; ++n) {
                step_n_phases = decode_step_plan(lw.data(), hp.n_text_layer, hp.n_text_state, hp.n_text_head, hp.n_vocab, d_te, d_ln_g, d_ln_b,
                                                 dattn16.as<__half>(), dh16.as<__half>(), n, step_grid / G, step_slot,
                                                 plans.data() + ((size_t) (G - 1) * (kStepMaxRows + 1) + n) * kStepMaxPhases);
                if (step_n_phases <= 0) { step_grid = 0; return true; }
            }
        }
        if (!step_plans.ensure(plans.size() * sizeof(StepPhase)) || !step_records.ensure((size_t) step_grid * kStepMaxRows * 6 * sizeof(double)) ||
            !step_bar.ensure(256)) return false;
        CUDA_OK(cudaMemcpy(step_plans.p, plans.data(), plans.size() * sizeof(StepPhase), cudaMemcpyHostToDevice));
        if (const char * e = getenv("WHISPER_B200_STEP_TRACE")) {
            step_trace_groups = std::max(1, atoi(e));
            if (!step_trace.ensure((size_t) step_grid * kStepMaxPhases * 8 * 8)) return false;
        }
        return true;
    }

    // Bump allocator over one device arena; every tensor starts on a 256-byte boundary.
    struct Packer {
        std::vector<uint8_t> host;
        size_t add(const void * src, size_t bytes) {
            const size_t off = (size_t) align_up((int64_t) host.size(), 256);
            host.resize(off + bytes, 0);
            if (src) memcpy(host.data() + off, src, bytes);
            return off;
        }
    };

    bool upload_weights(const ModelFile & mf) {
        Packer pk;
        const int d = hp.n_audio_state, n_mels = hp.n_mels;
        auto get = [&](const std::string & name) -> const TensorView & { return mf.tensors.at(name); };
        auto raw = [&](const std::string & name) -> size_t {
            const TensorView & t = get(name);
            return pk.add(t.data, t.nbytes);                      // t.data == nullptr (weight-less test model) => zeros
        };
        // conv kernel [out][in][3] (file order) -> [out][3][in]
        auto conv = [&](const std::string & name, int n_in) -> size_t {
            const TensorView & t = get(name);
            std::vector<uint16_t> tmp((size_t) d * 3 * n_in, 0);
            if (t.data) {
                const uint16_t * src = (const uint16_t *) t.data;
                for (int o = 0; o < d; ++o)
                    for (int i = 0; i < n_in; ++i)
                        for (int k = 0; k < 3; ++k) tmp[((size_t) o * 3 + k) * n_in + i] = src[((size_t) o * n_in + i) * 3 + k];
            }
            return pk.add(tmp.data(), tmp.size() * 2);
        };
        auto stack = [&](std::initializer_list<std::string> names) -> size_t {
            std::vector<uint8_t> tmp;
            for (const auto & nm : names) {
                const TensorView & t = get(nm);
                const size_t o = tmp.size();
                tmp.resize(o + t.nbytes, 0);
                if (t.data) memcpy(tmp.data() + o, t.data, t.nbytes);
            }
            return pk.add(tmp.data(), tmp.size());
        };
        // stacked Q/K/V bias; the key projection has none in the reference (whisper.cpp:1838-1841) => zeros
        auto stack_bias = [&](const std::string & a, const std::string & c, int dd) -> size_t {
            std::vector<float> tmp((size_t) 3 * dd, 0.0f);
            const TensorView & ta = get(a);
            const TensorView & tc = get(c);
            if (ta.data) memcpy(tmp.data(), ta.data, (size_t) dd * 4);
            if (tc.data) memcpy(tmp.data() + 2 * dd, tc.data, (size_t) dd * 4);
            return pk.add(tmp.data(), tmp.size() * 4);
        };

        struct Off { size_t v[24]; };
        const size_t o_conv1w = conv("encoder.conv1.weight", n_mels), o_conv1b = raw("encoder.conv1.bias");
        const size_t o_conv2w = conv("encoder.conv2.weight", d),      o_conv2b = raw("encoder.conv2.bias");
        const size_t o_epe = raw("encoder.positional_embedding");
        const size_t o_elng = raw("encoder.ln_post.weight"), o_elnb = raw("encoder.ln_post.bias");
        std::vector<Off> eo(hp.n_audio_layer), dof(hp.n_text_layer);
        for (int i = 0; i < hp.n_audio_layer; ++i) {
            const std::string p = "encoder.blocks." + std::to_string(i) + ".";
            Off & o = eo[i];
            o.v[0] = raw(p + "attn_ln.weight"); o.v[1] = raw(p + "attn_ln.bias");
            o.v[2] = raw(p + "mlp_ln.weight");  o.v[3] = raw(p + "mlp_ln.bias");
            o.v[4] = stack({p + "attn.query.weight", p + "attn.key.weight", p + "attn.value.weight"});
            o.v[5] = stack_bias(p + "attn.query.bias", p + "attn.value.bias", d);
            o.v[6] = raw(p + "attn.out.weight"); o.v[7] = raw(p + "attn.out.bias");
            o.v[8] = raw(p + "mlp.0.weight");    o.v[9] = raw(p + "mlp.0.bias");
            o.v[10] = raw(p + "mlp.2.weight");   o.v[11] = raw(p + "mlp.2.bias");
        }
        const size_t o_dpe = raw("decoder.positional_embedding"), o_dte = raw("decoder.token_embedding.weight");
        const size_t o_dlng = raw("decoder.ln.weight"), o_dlnb = raw("decoder.ln.bias");
        for (int i = 0; i < hp.n_text_layer; ++i) {
            const std::string p = "decoder.blocks." + std::to_string(i) + ".";
            Off & o = dof[i];
            o.v[0] = raw(p + "attn_ln.weight");       o.v[1] = raw(p + "attn_ln.bias");
            o.v[2] = raw(p + "cross_attn_ln.weight"); o.v[3] = raw(p + "cross_attn_ln.bias");
            o.v[4] = raw(p + "mlp_ln.weight");        o.v[5] = raw(p + "mlp_ln.bias");
            o.v[6] = stack({p + "attn.query.weight", p + "attn.key.weight", p + "attn.value.weight"});
            o.v[7] = stack_bias(p + "attn.query.bias", p + "attn.value.bias", d);
            o.v[8] = raw(p + "attn.out.weight");        o.v[9] = raw(p + "attn.out.bias");
            o.v[10] = raw(p + "cross_attn.query.weight"); o.v[11] = raw(p + "cross_attn.query.bias");
            o.v[12] = stack({p + "cross_attn.key.weight", p + "cross_attn.value.weight"});
            {
                std::vector<float> tmp((size_t) 2 * d, 0.0f);
                const TensorView & t = get(p + "cross_attn.value.bias");
                if (t.data) memcpy(tmp.data() + d, t.data, (size_t) d * 4);
                o.v[13] = pk.add(tmp.data(), tmp.size() * 4);
            }
            o.v[14] = raw(p + "cross_attn.out.weight"); o.v[15] = raw(p + "cross_attn.out.bias");
            o.v[16] = raw(p + "mlp.0.weight"); o.v[17] = raw(p + "mlp.0.bias");
            o.v[18] = raw(p + "mlp.2.weight"); o.v[19] = raw(p + "mlp.2.bias");
        }
        // activation tables (tables.cpp; same construction as ggml.c:2218-2236)
        std::vector<uint16_t> lut_gelu(65536), lut_exp(65536);
        build_f16_tables(lut_gelu.data(), lut_exp.data());
        const size_t o_gelu = pk.add(lut_gelu.data(), 65536 * 2), o_exp = pk.add(lut_exp.data(), 65536 * 2);
        // the fused attention kernel stages only the head of the negative half of the exp table on chip: everything behind it
        // must be zero (exp(x) < 2^-25 for x < -17.33), -0 included as entry 0
        for (int i = attention_enc_table_entries() - 1; i <= 0x7C00 && fused_attn; ++i) {
            if (lut_exp[0x8000 + i] != 0) { WB_LOG_WARN("%s: exp table has a non-zero tail; fused attention disabled\n", __func__); fused_attn = false; }
        }
        // per-token class bits for the device-side greedy sampler: which of whisper_process_logits' unconditional /
        // flag-conditional suppressions apply to a token id (whisper.cpp:4527-4594; same sets as csrc/decode_host.cpp)
        std::vector<uint8_t> cls_h((size_t) hp.n_vocab, 0);
        {
            const Vocab & vc = mf.vocab;
            LogitsRules rules;
            rules.build(vc);
            auto mark = [&](int id, uint8_t bit) { if (id >= 0 && id < hp.n_vocab) cls_h[id] |= bit; };
            mark(vc.token_not, 1); mark(vc.token_sot, 1); mark(vc.token_nosp, 1); mark(vc.token_translate, 1);
            mark(vc.token_transcribe, 1); mark(vc.token_prev, 1);
            for (int i = 0; i < lang_count(); ++i) mark(vc.token_lang(i), 1);
            for (int32_t id : rules.non_speech) mark(id, 2);
            mark(vc.token_eot, 4); mark(rules.blank, 4);
            mark(vc.token_solm, 8);
            token_beg = vc.token_beg; token_eot = vc.token_eot;
        }
        const size_t o_cls = pk.add(cls_h.data(), cls_h.size());
        // log-mel tables: the bits the host transform multiplies by (csrc/mel.cpp), the model's filter bank and its non-zero spans
        size_t o_hann = 0, o_lc = 0, o_ls = 0, o_twr = 0, o_twi = 0, o_filt = 0, o_g0 = 0, o_g1 = 0;
        if (mf.filters.n_fft != 201 || mf.filters.n_mel <= 0 || mf.filters.n_mel > 128) mel_dev_on = false;
        if (mel_dev_on) {
            const MelTablesView tv = mel_tables_view();
            o_hann = pk.add(tv.hann, 400 * 4); o_lc = pk.add(tv.leaf_cos, 625 * 4); o_ls = pk.add(tv.leaf_sin, 625 * 4);
            o_twr = pk.add(tv.tw_re, 800 * 4); o_twi = pk.add(tv.tw_im, 800 * 4);
            o_filt = pk.add(mf.filters.data.data(), mf.filters.data.size() * 4);
            std::vector<int> g0, g1;
            mel_filter_spans(mf.filters, g0, g1);
            o_g0 = pk.add(g0.data(), g0.size() * 4); o_g1 = pk.add(g1.data(), g1.size() * 4);
            filt_n_mel = mf.filters.n_mel;
            mel_low = (float) log10(1e-10);
        }

        if (!wbuf.ensure(pk.host.size())) return false;
        CUDA_OK(cudaMemcpy(wbuf.p, pk.host.data(), pk.host.size(), cudaMemcpyHostToDevice));
        const uint8_t * base = (const uint8_t *) wbuf.p;
        auto H = [&](size_t o) { return (const __half *) (base + o); };
        auto F = [&](size_t o) { return (const float *) (base + o); };
        conv1_w = H(o_conv1w); conv1_b = F(o_conv1b); conv2_w = H(o_conv2w); conv2_b = F(o_conv2b);
        e_pe = F(o_epe); e_ln_g = F(o_elng); e_ln_b = F(o_elnb);
        d_pe = F(o_dpe); d_te = H(o_dte); d_ln_g = F(o_dlng); d_ln_b = F(o_dlnb);
        gelu_lut = (const uint16_t *) (base + o_gelu); exp_lut = (const uint16_t *) (base + o_exp);
        cls_tab = base + o_cls;
        if (mel_dev_on) {
            meltab.hann = F(o_hann); meltab.leaf_cos = F(o_lc); meltab.leaf_sin = F(o_ls); meltab.tw_re = F(o_twr); meltab.tw_im = F(o_twi);
            meltab.filt = F(o_filt); meltab.g0 = (const int *) (base + o_g0); meltab.g1 = (const int *) (base + o_g1); meltab.n_mel = filt_n_mel;
        }
        enc.resize(hp.n_audio_layer);
        for (int i = 0; i < hp.n_audio_layer; ++i) {
            const Off & o = eo[i];
            enc[i] = EncLayerW{F(o.v[0]), F(o.v[1]), F(o.v[2]), F(o.v[3]), H(o.v[4]), F(o.v[5]), H(o.v[6]), F(o.v[7]),
                               H(o.v[8]), F(o.v[9]), H(o.v[10]), F(o.v[11])};
        }
        dec.resize(hp.n_text_layer);
        for (int i = 0; i < hp.n_text_layer; ++i) {
            const Off & o = dof[i];
            dec[i] = DecLayerW{F(o.v[0]), F(o.v[1]), F(o.v[2]), F(o.v[3]), F(o.v[4]), F(o.v[5]), H(o.v[6]), F(o.v[7]),
                               H(o.v[8]), F(o.v[9]), H(o.v[10]), F(o.v[11]), H(o.v[12]), F(o.v[13]), H(o.v[14]), F(o.v[15]),
                               H(o.v[16]), F(o.v[17]), H(o.v[18]), F(o.v[19])};
        }
        WB_LOG_INFO("%s: %.2f MB of weights resident in HBM (device %d)\n", __func__, pk.host.size() / 1e6, device);
        return true;
    }

    bool ensure_slots(int n) override {
        if (n <= slots) return true;
        cudaSetDevice(device);
        cudaStreamSynchronize(st);
        const int d = hp.n_text_state, L = hp.n_text_layer;
        cross_k_slot = (int64_t) L * Tmax * d;
        cross_v_slot = (int64_t) L * d * Tpmax;
        self_k_slot  = (int64_t) L * kv_cells * d;
        self_v_slot  = (int64_t) L * d * kv_cells;
        // slot contents only live for the duration of one whisper_full call, so growing = fresh zeroed buffers
        cross_k.release(); cross_v.release(); self_k.release(); self_v.release();
        drop_graphs();
        gemm_tc_forget_maps(); gemm_enc_forget_maps();
        if (!cross_k.ensure((size_t) n * cross_k_slot * 2) || !cross_v.ensure((size_t) n * cross_v_slot * 2) ||
            !self_k.ensure((size_t) (n + 1) * self_k_slot * 2) || !self_v.ensure((size_t) (n + 1) * self_v_slot * 2)) return false;   // + one scratch slot (padding rows of wide passes)
        run_seqs.release(); run_tokens.release(); run_init_h.release(); run_fetch_h.release();
        if (!run_seqs.ensure((size_t) n * sizeof(RunSeq)) || !run_tokens.ensure((size_t) n * kRunTokenCap * 6 * 4) ||
            !run_init_h.ensure((size_t) n * sizeof(RunSeq)) || !run_fetch_h.ensure((size_t) n * (sizeof(RunSeq) + kRunTokenCap * 6 * 4))) return false;
        run_pos_ub.assign(n, 0);
        if (mel_dev_on) {
            raw_mel.release(); mel_max.release();
            if (!raw_mel.ensure((size_t) n * kMelFramesCap * filt_n_mel * 4) || !mel_max.ensure((size_t) n * 4)) return false;
            mel_n_calc.assign(n, 0); mel_n_len.assign(n, 0);
            // one pinned energy-envelope buffer per slot (1.9 MB each): the host reads the envelope in place; if the host cannot pin that
            // much the envelopes travel through the PCM staging buffers instead
            energy_pool.release();
            if (!energy_pool.ensure((size_t) n * kPcmCap * 4)) { cudaGetLastError(); WB_LOG_WARN("%s: no pinned memory for %d energy buffers\n", __func__, n); }
        }
        slots = n;
        slot_n_ctx.assign(n, 0);
        return build_step_maps();
    }

    // TMA views of the cross-attention caches for the decode-step kernel: K as [slots*Lt*Tmax rows][d], V^T as [slots*Lt*d rows][Tpmax]
    bool build_step_maps() {
        if (step_grid <= 0 || slots <= 0) return true;
        const int d = hp.n_text_state, L = hp.n_text_layer;
        step_chunk_keys_cross = (step_slot / 128) & ~127;
        if (step_chunk_keys_cross < 128) { step_grid = 0; return true; }
        if (!make_tensor_map_2d_f16(&step_tm_ck, cross_k.p, (uint64_t) d, (uint64_t) slots * L * Tmax, (uint64_t) d * 2, 64, 128) ||
            !make_tensor_map_2d_f16(&step_tm_cv, cross_v.p, (uint64_t) Tpmax, (uint64_t) slots * L * d, (uint64_t) Tpmax * 2, 64, 64)) {
            WB_LOG_WARN("%s: cannot encode the cross-attention tensor maps; decode-step kernel disabled\n", __func__);
            step_grid = 0;
            return true;
        }
        // token embedding for the logits phase: box rows = 16 * tj of that phase (same rule as decode_step_plan)
        int tj = 4;
        while (tj > 1 && tj * 16 * d * 2 > step_slot) tj >>= 1;
        if (!make_tensor_map_2d_f16(&step_tm_te, d_te, (uint64_t) d, (uint64_t) hp.n_vocab, (uint64_t) d * 2, 64, (uint32_t) (16 * tj))) {
            WB_LOG_WARN("%s: cannot encode the token-embedding tensor map; decode-step kernel disabled\n", __func__);
            step_grid = 0;
        }
        return true;
    }

    // ---- GEMM dispatch ---------------------------------------------------------------------------------------------

    // algorithmic traffic of one contraction: both operands once + every output once
    static double gemm_bytes(const GemmShape & sh, const GemmEpi & e) {
        const double nb = (double) sh.nb1 * sh.nb2;
        double out_b = 0.0;
        for (int i = 0; i < e.nseg; ++i) {
            const double cols = e.nseg > 1 ? e.seg_m : sh.M;
            out_b += (double) sh.N * cols * nb * ((e.seg[i].out32 ? 4 : 0) + (e.seg[i].out16 ? 2 : 0) + (e.seg[i].out16t ? 2 : 0) + (e.seg[i].res ? 4 : 0));
        }
        return ((double) sh.N * sh.K + (double) sh.M * sh.K) * 2.0 * nb + out_b;
    }

    bool gemm(const Operand & A, const Operand & W, const GemmShape & sh, GemmEpi epi, int kind = PROF_GEMM_ENC, cudaStream_t stream = nullptr) {
        if (!stream) stream = st;
        epi.gelu_lut = gelu_lut;
        ++launches;
        prof_begin(kind, 2.0 * sh.N * (double) sh.M * sh.K * sh.nb1 * sh.nb2, gemm_bytes(sh, epi));
        bool ok = true;
        if (engine == 1) launch_gemm_simt(A, W, sh, epi, stream);
        else ok = launch_gemm_tc(A, W, sh, epi, stream);
        prof_end();
        return ok;
    }

    // The encoder's large contractions on the TMA-store kernel (gemm_enc.cu); false = not applicable here (debug engine, odd shape,
    // WHISPER_B200_GEMM_V2=0): the caller then takes the first-generation kernel with the same arithmetic.
    bool gemm_v2_ok(const EncGemm & g) const { return engine == 0 && gemm_enc_usable(g); }
    bool gemm_v2(EncGemm & g, cudaStream_t stream) {
        g.gelu_lut = gelu_lut;
        ++launches;
        double out_b = 0.0;
        for (int i = 0; i < g.nseg; ++i) out_b += (double) g.N * (g.nseg > 1 ? g.seg_m : g.M) * g.nb * (g.res32 ? 8.0 : 2.0);
        prof_begin(PROF_GEMM_ENC, 2.0 * g.N * (double) g.M * g.K * g.nb, ((double) g.N * g.K * g.nb + (double) g.M * g.K) * 2.0 + out_b);
        const bool ok = launch_gemm_enc(g, stream);
        prof_end();
        if (!ok) WB_LOG_ERROR("%s: launch failed: %s\n", __func__, cudaGetErrorString(cudaGetLastError()));
        return ok;
    }

    static Operand op2d(const __half * p, int64_t ld, int rows) { Operand o; o.p = p; o.ld = ld; o.rows = rows; return o; }

    // ---- encoder ---------------------------------------------------------------------------------------------------

    bool ensure_enc(int B) {
        if (B <= enc_cap) return true;
        const int64_t d = hp.n_audio_state, T = Tmax, Tp = Tpmax, nm = hp.n_mels;
        gemm_tc_forget_maps(); gemm_enc_forget_maps();
        bool ok = mel_d.ensure((size_t) B * nm * 2 * T * 4) && melT.ensure((size_t) B * (2 * T + 2) * nm * 2) &&
                  act1.ensure((size_t) B * (2 * T + 1) * d * 2) && conv16.ensure((size_t) B * T * d * 2) &&
                  x32.ensure((size_t) B * T * d * 4) && xn16.ensure((size_t) B * T * d * 2) && q16.ensure((size_t) B * T * d * 2) &&
                  k16.ensure((size_t) B * T * d * 2) && vt16.ensure((size_t) B * d * Tp * 2) &&
                  attn16.ensure((size_t) B * T * d * 2) && h16.ensure((size_t) B * T * 4 * d * 2) &&
                  enc32.ensure((size_t) B * T * d * 4) && mel_h.ensure((size_t) B * nm * 2 * T * 4);
        for (int i = 0; i < 2 && ok; ++i) {
            ok = slotmap_h2[i].ensure((size_t) B * sizeof(int)) && slotmap_d2[i].ensure((size_t) B * sizeof(int));
            if (ok && mel_dev_on) ok = pcm_d2[i].ensure((size_t) B * kPcmCap * 4) && clips_d2[i].ensure((size_t) B * sizeof(MelClip)) && clips_h2[i].ensure((size_t) B * sizeof(MelClip)) &&
                                       wins_d2[i].ensure((size_t) B * sizeof(MelWindow)) && wins_h2[i].ensure((size_t) B * sizeof(MelWindow)) &&
                                       energy_d2[i].ensure((size_t) B * kPcmCap * 4) && eclips_d2[i].ensure((size_t) B * sizeof(EnergyClip)) &&
                                       eclips_h2[i].ensure((size_t) B * sizeof(EnergyClip));
        }
        if (ok) enc_cap = B;
        return ok;
    }

    // ---- PCM staging for the device log-mel ------------------------------------------------------------------------------------
    bool mel_on_device() const override { return mel_dev_on; }
    bool is_pinned_host(const void * p) const override {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    }
    float * energy_buffer(int slot) override {
        if (!mel_dev_on || slot < 0 || slot >= slots || !energy_pool.p) return nullptr;
        return energy_pool.as<float>() + (size_t) slot * kPcmCap;
    }
    int  pcm_stage_samples() const override { return mel_dev_on ? kPcmCap : 0; }
    float * pcm_stage_acquire(int n_samples) override {
        if (!mel_dev_on || n_samples <= 0 || n_samples > kPcmCap) return nullptr;
        std::unique_lock<std::mutex> lk(pool_mu);
        if (!pool_ready) {
            cudaSetDevice(device);
            if (!pcm_pool.ensure((size_t) kPcmStages * kPcmCap * 4)) return nullptr;
            for (int i = 0; i < kPcmStages; ++i) pool_free.push_back(pcm_pool.as<float>() + (size_t) i * kPcmCap);
            pool_ready = true;
        }
        pool_cv.wait(lk, [&] { return !pool_free.empty(); });
        float * b = pool_free.back();
        pool_free.pop_back();
        return b;
    }
    void pcm_stage_release(float * buf) override {
        if (!buf) return;
        { std::lock_guard<std::mutex> lk(pool_mu); pool_free.push_back(buf); }
        pool_cv.notify_one();
    }

    bool encode(const float * mel_window, int n_ctx) override {
        EncodeJob j; j.mel_window = mel_window; j.slot = 0;
        return encode_batch(&j, 1, n_ctx);
    }

    // Two encoder passes may be queued: everything pass k + 1 brings from the host (PCM, job tables) is copied on its own stream while
    // pass k computes, and the kernels of k + 1 follow those of k on the encoder stream.
    int encode_sets() const override { return encoder_concurrent() ? 2 : 1; }
    bool encode_batch(const EncodeJob * jobs, int B, int n_ctx) override { return encode_enqueue(jobs, B, n_ctx, 0) && encode_collect(0); }

    bool encode_collect(int set) override {
        if (set < 0 || set > 1 || !enc_pending[set]) { WB_LOG_ERROR("%s: nothing queued on encoder set %d\n", __func__, set); return false; }
        CUDA_OK(cudaSetDevice(device));
        enc_pending[set] = false;
        CUDA_OK(cudaEventSynchronize(ev_enc1s[set]));
        if (e2h_pending2[set]) { e2h_pending2[set] = false; CUDA_OK(cudaEventSynchronize(ev_e2h2[set])); }
        CUDA_OK(cudaGetLastError());
        // the pass in two parts, like the reference's own timers (t_mel_us / t_encode_us, whisper.cpp:3793-3815): spectrogram stage (log-mel,
        // energy envelope, window staging), then whisper_encode_internal's work (conv stem, encoder layers, cross K / V)
        { float ms = 0.0f; if (cudaEventElapsedTime(&ms, ev_enc0s[set], ev_mel1s[set]) == cudaSuccess) t_mel_ms += ms; }
        { float ms = 0.0f; if (cudaEventElapsedTime(&ms, ev_mel1s[set], ev_enc1s[set]) == cudaSuccess) { t_enc_ms += ms; ++n_enc_calls; } }
        note_busy(ev_enc0s[set], ev_enc1s[set]);
        if (!encoder_concurrent()) prof_collect();
        return true;
    }

    bool encode_enqueue(const EncodeJob * jobs, int B, int n_ctx, int set) override {
        CUDA_OK(cudaSetDevice(device));
        if (set < 0 || set > 1 || enc_pending[set]) { WB_LOG_ERROR("%s: encoder set %d is busy\n", __func__, set); return false; }
        // the encoder has its own (low-priority) stream: its passes overlap with decoder passes, which leave most SMs idle in
        // their small kernels; while profiling everything runs on one stream so that the event brackets do not interleave
        const cudaStream_t es = encoder_concurrent() ? st_enc : st;
        const cudaStream_t hs2 = encoder_concurrent() ? st_h2d : es;       // host -> device copies of this pass
        DevBuf & pcm_d = pcm_d2[set]; DevBuf & clips_d = clips_d2[set]; DevBuf & wins_d = wins_d2[set]; DevBuf & energy_d = energy_d2[set];
        DevBuf & eclips_d = eclips_d2[set]; DevBuf & slotmap_d = slotmap_d2[set];
        PinnedBuf & clips_h = clips_h2[set]; PinnedBuf & wins_h = wins_h2[set]; PinnedBuf & eclips_h = eclips_h2[set]; PinnedBuf & slotmap_h = slotmap_h2[set];
        cudaEvent_t ev_enc0 = ev_enc0s[set], ev_enc1 = ev_enc1s[set], ev_energy = ev_energy2[set], ev_e2h = ev_e2h2[set];
        bool & e2h_pending = e2h_pending2[set];
        if (n_ctx <= 0 || n_ctx > Tmax) { WB_LOG_ERROR("%s: n_ctx %d out of range\n", __func__, n_ctx); return false; }
        for (int b = 0; b < B; ++b) if (jobs[b].slot < 0 || jobs[b].slot >= slots) { WB_LOG_ERROR("%s: bad slot\n", __func__); return false; }
        if (!ensure_enc(B)) return false;
        const int d = hp.n_audio_state, h = hp.n_audio_head, nm = hp.n_mels, T = n_ctx, F = 2 * n_ctx;
        const int Tp = (int) align_up(T, 8);
        const int64_t BT = (int64_t) B * T;

        cudaEventRecord(ev_enc0, es);
        // the previous pass that used this set's PCM buffer has read it (mel + energy kernels), its energy copies have left energy_d
        if (hs2 != es) CUDA_OK(cudaStreamWaitEvent(hs2, ev_pcm_read[set], 0));
        const size_t mel_elems = (size_t) nm * F;
        const int64_t melT_chunk = (int64_t) (F + 2) * nm, act1_chunk = (int64_t) (F + 1) * d;
        {
            // spectrogram windows -> f16 token-major rows with the conv's zero padding.  Host-computed windows go through pinned staging;
            // PCM jobs get their log-mel spectrogram on the device (mel_kernels.cu) and every device job reads its window from the
            // slot's resident spectrogram.
            int n_host = 0, n_clips = 0, n_wins = 0, max_calc = 0, n_energy = 0, max_samples = 0;
            EnergyClip * eh = mel_dev_on ? eclips_h.as<EnergyClip>() : nullptr;
            MelClip * ch = clips_h.as<MelClip>();
            MelWindow * wh = wins_h.as<MelWindow>();
            for (int b = 0; b < B; ++b) {
                const EncodeJob & j = jobs[b];
                if (j.mel_offset < 0) {
                    if (!j.mel_window) { WB_LOG_ERROR("%s: job without input\n", __func__); return false; }
                    memcpy(mel_h.as<float>() + b * mel_elems, j.mel_window, mel_elems * 4);
                    CUDA_OK(cudaMemcpyAsync(mel_d.as<float>() + b * mel_elems, mel_h.as<float>() + b * mel_elems, mel_elems * 4, cudaMemcpyHostToDevice, es));
                    h2d_bytes_enc += (double) mel_elems * 4;
                    ++n_host;
                    continue;
                }
                if (!mel_dev_on) { WB_LOG_ERROR("%s: device log-mel is off\n", __func__); return false; }
                float * raw = raw_mel.as<float>() + (size_t) j.slot * kMelFramesCap * filt_n_mel;
                if (j.pcm) {
                    if (j.n_samples <= 0 || j.n_samples > kPcmCap) { WB_LOG_ERROR("%s: clip of %d samples\n", __func__, j.n_samples); return false; }
                    int n_len = 0, n_len_org = 0, n_calc = 0;
                    mel_shape(j.n_samples, n_len, n_len_org, n_calc);
                    if (n_calc > kMelFramesCap) return false;
                    mel_n_calc[j.slot] = n_calc; mel_n_len[j.slot] = n_len;
                    float * pd = pcm_d.as<float>() + (size_t) b * kPcmCap;
                    CUDA_OK(cudaMemcpyAsync(pd, j.pcm, (size_t) j.n_samples * 4, cudaMemcpyHostToDevice, hs2));
                    h2d_bytes_enc += (double) j.n_samples * 4;
                    ch[n_clips++] = MelClip{pd, raw, mel_max.as<int>() + j.slot, j.n_samples, n_calc};
                    max_calc = std::max(max_calc, n_calc);
                    if (j.energy_out) { eh[n_energy++] = EnergyClip{pd, energy_d.as<float>() + (size_t) b * kPcmCap, j.n_samples}; max_samples = std::max(max_samples, j.n_samples); }
                }
                if (mel_n_len[j.slot] <= 0) { WB_LOG_ERROR("%s: slot %d has no spectrogram\n", __func__, j.slot); return false; }
                wh[n_wins++] = MelWindow{raw, mel_max.as<int>() + j.slot, melT.as<__half>() + b * melT_chunk, mel_n_calc[j.slot], mel_n_len[j.slot], j.mel_offset};
            }
            if (n_clips > 0) CUDA_OK(cudaMemcpyAsync(clips_d.p, ch, (size_t) n_clips * sizeof(MelClip), cudaMemcpyHostToDevice, hs2));
            if (n_energy > 0) CUDA_OK(cudaMemcpyAsync(eclips_d.p, eh, (size_t) n_energy * sizeof(EnergyClip), cudaMemcpyHostToDevice, hs2));
            if (n_wins > 0) CUDA_OK(cudaMemcpyAsync(wins_d.p, wh, (size_t) n_wins * sizeof(MelWindow), cudaMemcpyHostToDevice, hs2));
            if (hs2 != es) {
                CUDA_OK(cudaEventRecord(ev_h2d[set], hs2));
                CUDA_OK(cudaStreamWaitEvent(es, ev_h2d[set], 0));
            }
            if (n_clips > 0) {
                prof_begin(PROF_MISC, 0.0, 0.0);
                launch_logmel_frames(meltab, clips_d.as<MelClip>(), n_clips, max_calc, es); launches += 2;
                prof_end();
            }
            if (n_energy > 0) {
                // energy envelope of the clips that asked for it, into the pinned buffer each names
                CUDA_OK(cudaStreamWaitEvent(es, ev_e2h, 0));             // (the copies that last read this set's energy buffer)
                prof_begin(PROF_MISC, 0.0, (double) n_energy * max_samples * 8.0);
                launch_signal_energy(eclips_d.as<EnergyClip>(), n_energy, max_samples, 32, es); ++launches;
                prof_end();
                // (on a side stream: the copies run under the encoder kernels that follow instead of in front of them)
                CUDA_OK(cudaEventRecord(ev_energy, es));
                CUDA_OK(cudaStreamWaitEvent(st_e2h, ev_energy, 0));
                for (int b = 0; b < B; ++b) {
                    const EncodeJob & j = jobs[b];
                    if (j.mel_offset < 0 || !j.pcm || !j.energy_out) continue;
                    CUDA_OK(cudaMemcpyAsync(j.energy_out, energy_d.as<float>() + (size_t) b * kPcmCap, (size_t) j.n_samples * 4, cudaMemcpyDeviceToHost, st_e2h));
                    d2h_bytes += (double) j.n_samples * 4;
                }
                CUDA_OK(cudaEventRecord(ev_e2h, st_e2h));
                e2h_pending = true;
            }
            if (hs2 != es) CUDA_OK(cudaEventRecord(ev_pcm_read[set], es));
            if (n_wins > 0) {
                prof_begin(PROF_MISC, 0.0, (double) n_wins * nm * F * 6);
                launch_mel_window(wins_d.as<MelWindow>(), n_wins, nm, F, mel_low, es); ++launches;
                prof_end();
            }
            for (int b = 0; b < B && n_host > 0; ++b) {
                if (jobs[b].mel_offset >= 0) continue;
                prof_begin(PROF_MISC, 0.0, (double) nm * F * 6);
                launch_mel_to_tokens(mel_d.as<float>() + b * mel_elems, melT.as<__half>() + b * melT_chunk, nm, F, es);
                prof_end();
                ++launches;
            }
        }
        cudaEventRecord(ev_mel1s[set], es);
        // row 0 of every act1 chunk is the left zero pad of conv2 (rows 1.. are rewritten below)
        CUDA_OK(cudaMemset2DAsync(act1.p, (size_t) act1_chunk * 2, 0, (size_t) d * 2, (size_t) B, es));

        // conv1 (k=3, s=1, p=1) + bias + GELU: implicit GEMM, row t = mel frames t-1..t+1 (whisper.cpp:1711-1714)
        {
            Operand A; A.p = melT.as<__half>(); A.ld = nm; A.bs2 = melT_chunk; A.rows = F;
            Operand W = op2d(conv1_w, 3 * nm, d);
            GemmShape sh; sh.N = F; sh.M = d; sh.K = 3 * nm; sh.nb2 = B;
            GemmEpi e; EpiSeg & s = e.seg[0];
            s.bias = conv1_b; s.gelu = 1;
            s.out16 = act1.as<__half>() + d; s.out16_ld = d; s.out16_bs2 = act1_chunk;
            if (!gemm(A, W, sh, e, PROF_GEMM_ENC, es)) return false;
        }
        // conv2 (k=3, s=2, p=1) + bias + GELU, then + positional embedding (whisper.cpp:1716-1719, 1803-1807)
        {
            Operand A; A.p = act1.as<__half>(); A.ld = 2 * d; A.bs2 = act1_chunk; A.rows = T;
            Operand W = op2d(conv2_w, 3 * d, d);
            GemmShape sh; sh.N = T; sh.M = d; sh.K = 3 * d; sh.nb2 = B;
            GemmEpi e; EpiSeg & s = e.seg[0];
            s.bias = conv2_b; s.gelu = 1;
            s.out16 = conv16.as<__half>(); s.out16_ld = d; s.out16_bs2 = (int64_t) T * d; s.out16_pre = 1;   // embd_conv (GELU output is f16-exact)
            s.res = e_pe; s.res_ld = d;
            s.out32 = x32.as<float>(); s.out32_ld = d; s.out32_bs2 = (int64_t) T * d;
            if (!gemm(A, W, sh, e, PROF_GEMM_ENC, es)) return false;
        }

        for (int il = 0; il < hp.n_audio_layer; ++il) {
            const EncLayerW & L = enc[il];
            prof_begin(PROF_LAYERNORM, 0.0, (double) BT * d * 6);
            launch_layernorm(x32.as<float>(), L.ln1_g, L.ln1_b, xn16.as<__half>(), nullptr, (int) BT, d, hp.eps, es); ++launches;
            prof_end();
            EncGemm gq;        // Q (+b), K, V (+b, stored transposed per chunk)   whisper.cpp:1831-1850, 1880-1909
            gq.A = xn16.as<__half>(); gq.a_ld = d; gq.a_bs = (int64_t) T * d; gq.a_rows = T; gq.W = L.wqkv; gq.w_ld = d;
            gq.N = T; gq.M = 3 * d; gq.K = d; gq.nb = B; gq.nseg = 3; gq.seg_m = d;
            gq.out[0].p = q16.p; gq.out[0].ld = d; gq.out[0].bs = (int64_t) T * d; gq.out[0].bias = L.bqkv;
            gq.out[1].p = k16.p; gq.out[1].ld = d; gq.out[1].bs = (int64_t) T * d;
            gq.out[2].p = vt16.p; gq.out[2].ld = Tp; gq.out[2].bs = (int64_t) d * Tp; gq.out[2].bias = L.bqkv + 2 * d; gq.out[2].transposed = 1;
            if (gemm_v2_ok(gq)) { if (!gemm_v2(gq, es)) return false; }
            else {
                Operand A; A.p = xn16.as<__half>(); A.ld = d; A.bs2 = (int64_t) T * d; A.rows = T;
                Operand W = op2d(L.wqkv, d, 3 * d);
                GemmShape sh; sh.N = T; sh.M = 3 * d; sh.K = d; sh.nb2 = B;
                GemmEpi e; e.nseg = 3; e.seg_m = d;
                e.seg[0].bias = L.bqkv;         e.seg[0].out16 = q16.as<__half>(); e.seg[0].out16_ld = d; e.seg[0].out16_bs2 = (int64_t) T * d;
                                                e.seg[1].out16 = k16.as<__half>(); e.seg[1].out16_ld = d; e.seg[1].out16_bs2 = (int64_t) T * d;
                e.seg[2].bias = L.bqkv + 2 * d; e.seg[2].out16t = vt16.as<__half>(); e.seg[2].out16t_ld = Tp; e.seg[2].out16t_bs2 = (int64_t) d * Tp;
                if (!gemm(A, W, sh, e, PROF_GEMM_ENC, es)) return false;
            }
            if (fused_attn && engine == 0) {
                // softmax(K q / sqrt(64)) V with the scores kept on chip (attn_enc.cu)    whisper.cpp:1880-1917
                ++launches;
                prof_begin(PROF_GEMM_ATTN, 4.0 * B * h * (double) T * T * 64.0, (double) BT * d * 2 * 4);
                const bool ok = launch_attention_enc(q16.as<__half>(), k16.as<__half>(), vt16.as<__half>(), attn16.as<__half>(), B, T, Tp, d, h, exp_lut, es);
                prof_end();
                if (!ok) return false;
            } else {
                // score / probability buffers of the three-launch path (debug engine, WHISPER_B200_FUSED_ATTN=0): allocated on first use
                if (!S32.ensure((size_t) enc_cap * h * Tmax * Tpmax * 4) || !P16.ensure((size_t) enc_cap * h * Tmax * Tpmax * 2)) return false;
                {   // S = (K q) / sqrt(64)     whisper.cpp:1894-1897
                    Operand A; A.p = q16.as<__half>(); A.ld = d; A.bs1 = 64; A.bs2 = (int64_t) T * d; A.rows = T;
                    Operand W; W.p = k16.as<__half>(); W.ld = d; W.bs1 = 64; W.bs2 = (int64_t) T * d; W.rows = T;
                    GemmShape sh; sh.N = T; sh.M = T; sh.K = 64; sh.nb1 = h; sh.nb2 = B;
                    GemmEpi e; EpiSeg & s = e.seg[0];
                    s.scale = 1.0f / sqrtf(float(d) / h);
                    s.out32 = S32.as<float>(); s.out32_ld = Tp; s.out32_bs1 = (int64_t) T * Tp; s.out32_bs2 = (int64_t) h * T * Tp;
                    if (!gemm(A, W, sh, e, PROF_GEMM_ATTN, es)) return false;
                }
                prof_begin(PROF_SOFTMAX, 0.0, (double) B * h * T * (double) T * 6);
                launch_softmax_rows(S32.as<float>(), P16.as<__half>(), (int64_t) B * h * T, T, Tp, Tp, exp_lut, es); ++launches;
                prof_end();
                {   // O = P V, heads merged back to [T][d]    whisper.cpp:1911-1917
                    Operand A; A.p = P16.as<__half>(); A.ld = Tp; A.bs1 = (int64_t) T * Tp; A.bs2 = (int64_t) h * T * Tp; A.rows = T;
                    Operand W; W.p = vt16.as<__half>(); W.ld = Tp; W.bs1 = (int64_t) 64 * Tp; W.bs2 = (int64_t) d * Tp; W.rows = 64;
                    GemmShape sh; sh.N = T; sh.M = 64; sh.K = T; sh.nb1 = h; sh.nb2 = B;
                    GemmEpi e; EpiSeg & s = e.seg[0];
                    s.out16 = attn16.as<__half>(); s.out16_ld = d; s.out16_bs1 = 64; s.out16_bs2 = (int64_t) T * d;
                    if (!gemm(A, W, sh, e, PROF_GEMM_ATTN, es)) return false;
                }
            }
            EncGemm go;        // out projection + bias + residual     whisper.cpp:1922-1930
            go.A = attn16.as<__half>(); go.a_ld = d; go.a_rows = (int) BT; go.W = L.wo; go.w_ld = d; go.N = (int) BT; go.M = d; go.K = d;
            go.res32 = true; go.out[0].p = x32.p; go.out[0].ld = d; go.out[0].bias = L.bo; go.res = x32.as<float>(); go.res_ld = d; go.res_rows = (int) BT;
            if (gemm_v2_ok(go)) { if (!gemm_v2(go, es)) return false; }
            else {
                GemmShape sh; sh.N = (int) BT; sh.M = d; sh.K = d;
                GemmEpi e; EpiSeg & s = e.seg[0];
                s.bias = L.bo; s.res = x32.as<float>(); s.res_ld = d; s.out32 = x32.as<float>(); s.out32_ld = d;
                if (!gemm(op2d(attn16.as<__half>(), d, (int) BT), op2d(L.wo, d, d), sh, e, PROF_GEMM_ENC, es)) return false;
            }
            prof_begin(PROF_LAYERNORM, 0.0, (double) BT * d * 6);
            launch_layernorm(x32.as<float>(), L.ln2_g, L.ln2_b, xn16.as<__half>(), nullptr, (int) BT, d, hp.eps, es); ++launches;
            prof_end();
            EncGemm g1;        // FC1 + bias + GELU     whisper.cpp:1952-1959
            g1.A = xn16.as<__half>(); g1.a_ld = d; g1.a_rows = (int) BT; g1.W = L.w1; g1.w_ld = d; g1.N = (int) BT; g1.M = 4 * d; g1.K = d;
            g1.out[0].p = h16.p; g1.out[0].ld = 4 * d; g1.out[0].bias = L.b1; g1.out[0].gelu = 1;
            if (gemm_v2_ok(g1)) { if (!gemm_v2(g1, es)) return false; }
            else {
                GemmShape sh; sh.N = (int) BT; sh.M = 4 * d; sh.K = d;
                GemmEpi e; EpiSeg & s = e.seg[0];
                s.bias = L.b1; s.gelu = 1; s.out16 = h16.as<__half>(); s.out16_ld = 4 * d;
                if (!gemm(op2d(xn16.as<__half>(), d, (int) BT), op2d(L.w1, d, 4 * d), sh, e, PROF_GEMM_ENC, es)) return false;
            }
            EncGemm g2;        // FC2 + bias + residual     whisper.cpp:1962-1970
            g2.A = h16.as<__half>(); g2.a_ld = 4 * d; g2.a_rows = (int) BT; g2.W = L.w2; g2.w_ld = 4 * d; g2.N = (int) BT; g2.M = d; g2.K = 4 * d;
            g2.res32 = true; g2.out[0].p = x32.p; g2.out[0].ld = d; g2.out[0].bias = L.b2; g2.res = x32.as<float>(); g2.res_ld = d; g2.res_rows = (int) BT;
            if (gemm_v2_ok(g2)) { if (!gemm_v2(g2, es)) return false; }
            else {
                GemmShape sh; sh.N = (int) BT; sh.M = d; sh.K = 4 * d;
                GemmEpi e; EpiSeg & s = e.seg[0];
                s.bias = L.b2; s.res = x32.as<float>(); s.res_ld = d; s.out32 = x32.as<float>(); s.out32_ld = d;
                if (!gemm(op2d(h16.as<__half>(), 4 * d, (int) BT), op2d(L.w2, 4 * d, d), sh, e, PROF_GEMM_ENC, es)) return false;
            }
        }
        // ln_post -> embd_enc (f32 for the stage probe, f16 as the operand of the cross projections)  whisper.cpp:1975-1983
        prof_begin(PROF_LAYERNORM, 0.0, (double) BT * d * 10);
        launch_layernorm(x32.as<float>(), e_ln_g, e_ln_b, xn16.as<__half>(), enc32.as<float>(), (int) BT, d, hp.eps, es); ++launches;
        prof_end();

        // cross-attention K (scaled) and V (+b, transposed) of every decoder layer into the chunk's slot  whisper.cpp:2038-2066
        const float kscale = (float) pow((double) ((float) hp.n_text_state / hp.n_text_head), -0.25);
        // one launch per decoder layer for the whole batch: chunk b writes into device slot jobs[b].slot (bmap2)
        for (int b = 0; b < B; ++b) { slot_n_ctx[jobs[b].slot] = T; slotmap_h.as<int>()[b] = jobs[b].slot; }
        CUDA_OK(cudaMemcpyAsync(slotmap_d.p, slotmap_h.p, (size_t) B * sizeof(int), cudaMemcpyHostToDevice, es));
        for (int il = 0; il < hp.n_text_layer; ++il) {
            const DecLayerW & L = dec[il];
            EncGemm gc;
            gc.A = xn16.as<__half>(); gc.a_ld = d; gc.a_bs = (int64_t) T * d; gc.a_rows = T; gc.W = L.wckv; gc.w_ld = d;
            gc.N = T; gc.M = 2 * d; gc.K = d; gc.nb = B; gc.nseg = 2; gc.seg_m = d;
            gc.out[0].p = cross_k.as<__half>() + (int64_t) il * Tmax * d; gc.out[0].ld = d; gc.out[0].bs = cross_k_slot; gc.out[0].n_batch_out = slots;
            gc.out[0].scale = kscale; gc.out[0].bmap = slotmap_d.as<int>();
            gc.out[1].p = cross_v.as<__half>() + (int64_t) il * d * Tpmax; gc.out[1].ld = Tpmax; gc.out[1].bs = cross_v_slot; gc.out[1].n_batch_out = slots;
            gc.out[1].bias = L.bckv + d; gc.out[1].transposed = 1; gc.out[1].bmap = slotmap_d.as<int>();
            if (gemm_v2_ok(gc)) { if (!gemm_v2(gc, es)) return false; continue; }
            Operand A; A.p = xn16.as<__half>(); A.ld = d; A.bs2 = (int64_t) T * d; A.rows = T;
            GemmShape sh; sh.N = T; sh.M = 2 * d; sh.K = d; sh.nb2 = B;
            GemmEpi e; e.nseg = 2; e.seg_m = d;
            e.seg[0].scale = kscale;
            e.seg[0].out16 = cross_k.as<__half>() + (int64_t) il * Tmax * d; e.seg[0].out16_ld = d; e.seg[0].out16_bs2 = cross_k_slot;
            e.seg[0].bmap2 = slotmap_d.as<int>();
            e.seg[1].bias = L.bckv + d;
            e.seg[1].out16t = cross_v.as<__half>() + (int64_t) il * d * Tpmax; e.seg[1].out16t_ld = Tpmax; e.seg[1].out16t_bs2 = cross_v_slot;
            e.seg[1].bmap2 = slotmap_d.as<int>();
            if (!gemm(A, op2d(L.wckv, d, 2 * d), sh, e, PROF_GEMM_ENC, es)) return false;
        }
        enc_last_B = B; enc_last_T = T;
        cudaEventRecord(ev_enc1, es);
        enc_pending[set] = true;
        return true;
    }
    int enc_last_B = 0, enc_last_T = 0;

    // ---- decoder ---------------------------------------------------------------------------------------------------

    // layout of the per-step staging block (one H2D copy): all arrays sized for `cap` rows
    static constexpr int kDrawsPerRow = 8;                   // average draws per row a pass can carry (a prompt row of a beam search takes beam x decoders)
    struct StageLayout {
        size_t nkv, token, pos, want, wslot, rule, drule, draws, rowmap_k, rowmap_v, koff_self, voff_self, koff_cross, voff_cross, mask, total;
        StageLayout(int cap, int kv) {
            size_t o = 0;
            auto take = [&](size_t bytes) { const size_t r = o; o = (size_t) align_up((int64_t) (o + bytes), 256); return r; };
            nkv = take(4);
            token = take((size_t) cap * 4); pos = take((size_t) cap * 4); want = take((size_t) cap * 4); wslot = take((size_t) cap * 4);
            rule = take((size_t) cap * 16);
            drule = take((size_t) cap * 32);                 // rows sampled from their distribution: 8 x int32 each (kernels.cuh launch_sample_dist)
            draws = take((size_t) cap * kDrawsPerRow * 8);   // ... and their uniform variates
            rowmap_k = take((size_t) cap * 4); rowmap_v = take((size_t) cap * 4);
            koff_self = take((size_t) cap * 8); voff_self = take((size_t) cap * 8);
            koff_cross = take((size_t) cap * 8); voff_cross = take((size_t) cap * 8);
            mask = take((size_t) cap * kv * 4);
            total = o;
        }
    };

    bool ensure_dec(int n) {
        if (n <= dec_cap) return true;
        drop_graphs();
        const int cap = (int) align_up(n, 64);
        const int64_t d = hp.n_text_state, V = hp.n_vocab;
        gemm_tc_forget_maps(); gemm_enc_forget_maps();
        StageLayout sl(cap, kv_cells);
        bool ok = dx32.ensure((size_t) cap * d * 4) && dxn16.ensure((size_t) cap * d * 2) && dq16.ensure((size_t) cap * d * 2) &&
                  dattn16.ensure((size_t) cap * d * 2) && dh16.ensure((size_t) cap * 4 * d * 2) && dxw32.ensure((size_t) cap * d * 4) &&
                  dlogits.ensure((size_t) cap * V * 4) && dstage.ensure(sl.total) && hstage.ensure(sl.total) &&
                  act_b.x32.ensure((size_t) cap * d * 4) && act_b.xn16.ensure((size_t) cap * d * 2) && act_b.q16.ensure((size_t) cap * d * 2) &&
                  act_b.attn16.ensure((size_t) cap * d * 2) && act_b.h16.ensure((size_t) cap * 4 * d * 2) && act_b.xw32.ensure((size_t) cap * d * 4) &&
                  act_b.logits.ensure((size_t) cap * V * 4) &&
                  hlogits.ensure((size_t) cap * V * 4) && dsampled.ensure((size_t) cap * 24) && hsampled.ensure((size_t) cap * 24) &&
                  dstage2.ensure(sl.total) && hstage2.ensure(sl.total) && dsampled2.ensure((size_t) cap * 24) && hsampled2.ensure((size_t) cap * 24);
        ok = ok && ddist.ensure((size_t) cap * kDrawsPerRow * 24) && hdist.ensure((size_t) cap * kDrawsPerRow * 24);
        ok = ok && dstage_run.ensure(sl.total) && dsampled_run.ensure((size_t) cap * 24) && run_rows_d.ensure((size_t) cap * 4) &&
             run_status_d.ensure((size_t) cap * 4);
        for (int i = 0; i < kRunRing && ok; ++i) ok = run_rows_h[i].ensure((size_t) cap * 4) && run_status_h[i].ensure((size_t) cap * 4);
        if (ok) dec_cap = cap;
        if (ok && step_grid > 0) ok = build_step_plans();
        return ok;
    }

    bool decode(const DecodeInput & in, int n_audio_ctx, float * logits_out) override {
        DecodeJob j; j.in = in; j.slot = 0; j.logits_out = logits_out;
        return decode_batch(&j, 1, n_audio_ctx);
    }

    // linear map on n decoder rows: skinny kernel for n <= 8, tensor cores above
    bool dec_linear(const float * x32_in, const float * g, const float * b, const __half * x16_in, int64_t x16_ld,
                    const __half * W, int n, int M, int K, GemmEpi e, cudaStream_t ds = nullptr, __half * xn = nullptr) {
        if (!ds) ds = st;
        if (!xn) xn = dxn16.as<__half>();
        e.gelu_lut = gelu_lut;
        if (n <= 8 && engine != 1) {
            SkinnyIn in;
            if (x32_in) { in.x32 = x32_in; in.x32_ld = K; in.gamma = g; in.beta = b; in.eps = hp.eps; }
            else        { in.x16 = x16_in; in.x16_ld = x16_ld; }
            prof_begin(PROF_SKINNY, 2.0 * n * (double) M * K, (double) M * K * 2 + (double) n * (K + M) * 4);
            launch_gemm_skinny(in, W, n, M, K, e, ds); ++launches;
            prof_end();
            return true;
        }
        const __half * a = x16_in;
        int64_t ld = x16_ld;
        if (x32_in) {
            prof_begin(PROF_LAYERNORM, 0.0, (double) n * K * 6);
            launch_layernorm(x32_in, g, b, xn, nullptr, n, K, hp.eps, ds); ++launches;
            prof_end();
            a = xn; ld = K;
        }
        if (n <= 32 && engine != 1 && (K % 32) == 0 && (size_t) 32 * (K + 32) * 2 + 8192 <= 200 * 1024) {
            // 9..32 rows: weights streamed once through mma.sync fragments (kernels.cu), HBM-bound
            prof_begin(PROF_SKINNY, 2.0 * n * (double) M * K, (double) M * K * 2 + (double) n * (K + M) * 4);
            launch_gemm_skinny_mma(a, ld, W, n, M, K, e, ds); ++launches;
            prof_end();
            return true;
        }
        GemmShape sh; sh.N = n; sh.M = M; sh.K = K;
        return gemm(op2d(a, ld, n), op2d(W, K, M), sh, e, PROF_GEMM_DEC, ds);
    }

    // ---- one decoder step as a fixed launch sequence (captured into CUDA graphs by decode_batch) ------------------------
    struct DecodeShape {
        int n, n_full, n_samp, kvb, n_audio_ctx, engine, set;
        bool operator<(const DecodeShape & o) const {
            return std::tie(n, n_full, n_samp, kvb, n_audio_ctx, engine, set) < std::tie(o.n, o.n_full, o.n_samp, o.kvb, o.n_audio_ctx, o.engine, o.set);
        }
    };
    std::map<DecodeShape, cudaGraphExec_t> graphs;
    std::map<DecodeShape, int64_t> graph_nodes;
    std::map<DecodeShape, int> graph_seen;
    bool use_graphs = true;
    bool force_multi = false;      // test hook: route every step through the multi-kernel path
    void drop_graphs() {
        for (auto & kv : graphs) cudaGraphExecDestroy(kv.second);
        graphs.clear(); graph_nodes.clear(); graph_seen.clear();
    }

    bool enqueue_decode(int n, int n_full, int n_samp, int kvb, int ld_mask, int n_audio_ctx, const StageLayout & sl, int set, cudaStream_t dst, int aset, int n_kv_live) {
        const int n_want = n_full + n_samp;
        const int d = hp.n_text_state, h = hp.n_text_head, V = hp.n_vocab, Lt = hp.n_text_layer;
        const uint8_t * ds = (set == 2 ? dstage_run : set ? dstage2 : dstage).as<uint8_t>();
        float * sampled_out = (set == 2 ? dsampled_run : set ? dsampled2 : dsampled).as<float>();
        const int * d_token = (const int *) (ds + sl.token), * d_pos = (const int *) (ds + sl.pos), * d_want = (const int *) (ds + sl.want);
        const int * d_rk = (const int *) (ds + sl.rowmap_k), * d_rv = (const int *) (ds + sl.rowmap_v);
        const int64_t * d_ks = (const int64_t *) (ds + sl.koff_self), * d_vs = (const int64_t *) (ds + sl.voff_self);
        const int64_t * d_kc = (const int64_t *) (ds + sl.koff_cross), * d_vc = (const int64_t *) (ds + sl.voff_cross);
        const float * d_mask = (const float *) (ds + sl.mask);
        const int * d_nkv = (const int *) (ds + sl.nkv);
        const int n_kv = kvb;   // upper bound baked into the launches; the live key count is read from *d_nkv on the device

        // activations of this pass: set 0 shares them with the decode-step kernel, set 1 (second decoder stream) has its own
        __half * a_xn = aset ? act_b.xn16.as<__half>() : dxn16.as<__half>();
        __half * a_q = aset ? act_b.q16.as<__half>() : dq16.as<__half>();
        __half * a_attn = aset ? act_b.attn16.as<__half>() : dattn16.as<__half>();
        __half * a_h = aset ? act_b.h16.as<__half>() : dh16.as<__half>();
        float * a_xw = aset ? act_b.xw32.as<float>() : dxw32.as<float>();
        float * a_logits = aset ? act_b.logits.as<float>() : dlogits.as<float>();
        float * x = aset ? act_b.x32.as<float>() : dx32.as<float>();
        prof_begin(PROF_MISC, 0.0, (double) n * d * 10);
        launch_embed(d_te, d_pe, d_token, d_pos, x, n, d, dst); ++launches;
        prof_end();
        const float qscale = (float) pow((double) ((float) d / h), -0.25);

        for (int il = 0; il < Lt; ++il) {
            const DecLayerW & L = dec[il];
            {   // self-attention projections; K / V go straight into their cache cells   whisper.cpp:2240-2288
                GemmEpi e; e.nseg = 3; e.seg_m = d;
                e.seg[0].bias = L.bqkv; e.seg[0].scale = qscale; e.seg[0].out16 = a_q; e.seg[0].out16_ld = d;
                e.seg[1].scale = qscale;
                e.seg[1].out16 = self_k.as<__half>() + (int64_t) il * kv_cells * d; e.seg[1].out16_ld = d; e.seg[1].rowmap16 = d_rk;
                e.seg[2].bias = L.bqkv + 2 * d;
                e.seg[2].out16t = self_v.as<__half>() + (int64_t) il * d * kv_cells; e.seg[2].out16t_ld = kv_cells; e.seg[2].rowmap16t = d_rv;
                if (!dec_linear(x, L.ln1_g, L.ln1_b, nullptr, 0, L.wqkv, n, 3 * d, d, e, dst, a_xn)) return false;
            }
            {   // softmax(K q + mask) V   whisper.cpp:2291-2330
                AttnArgs a; a.q = a_q;
                a.K = self_k.as<__half>() + (int64_t) il * kv_cells * d; a.koff = d_ks;
                a.Vt = self_v.as<__half>() + (int64_t) il * d * kv_cells; a.voff = d_vs; a.ld_v = kv_cells;
                a.mask = d_mask; a.ld_mask = ld_mask; a.out = a_attn;
                a.n = n; a.d = d; a.n_head = h; a.n_keys = n_kv; a.n_keys_dev = d_nkv; a.exp_lut = exp_lut;
                prof_begin(PROF_DEC_ATTN, 4.0 * n * h * 64.0 * n_kv_live, (double) n * h * 64.0 * n_kv_live * 4);   // (the live key count, not its bucket)
                launch_decode_attention(a, dst); ++launches;
                prof_end();
            }
            {   // out projection + residual   whisper.cpp:2333-2345
                GemmEpi e; e.seg[0].bias = L.bo; e.seg[0].res = x; e.seg[0].res_ld = d; e.seg[0].out32 = x; e.seg[0].out32_ld = d;
                if (!dec_linear(nullptr, nullptr, nullptr, a_attn, d, L.wo, n, d, d, e, dst, a_xn)) return false;
            }
            {   // cross-attention query   whisper.cpp:2349-2370
                GemmEpi e; e.seg[0].bias = L.bcq; e.seg[0].scale = qscale; e.seg[0].out16 = a_q; e.seg[0].out16_ld = d;
                if (!dec_linear(x, L.lnc_g, L.lnc_b, nullptr, 0, L.wcq, n, d, d, e, dst, a_xn)) return false;
            }
            {   // softmax(Kc q) Vc, no mask   whisper.cpp:2372-2423
                AttnArgs a; a.q = a_q;
                a.K = cross_k.as<__half>() + (int64_t) il * Tmax * d; a.koff = d_kc;
                a.Vt = cross_v.as<__half>() + (int64_t) il * d * Tpmax; a.voff = d_vc; a.ld_v = Tpmax;
                a.out = a_attn;
                a.n = n; a.d = d; a.n_head = h; a.n_keys = n_audio_ctx; a.exp_lut = exp_lut;
                prof_begin(PROF_DEC_ATTN, 4.0 * n * h * 64.0 * n_audio_ctx, (double) n * h * 64.0 * n_audio_ctx * 4);
                launch_decode_attention(a, dst); ++launches;
                prof_end();
            }
            {   // cross out projection + residual   whisper.cpp:2426-2438
                GemmEpi e; e.seg[0].bias = L.bco; e.seg[0].res = x; e.seg[0].res_ld = d; e.seg[0].out32 = x; e.seg[0].out32_ld = d;
                if (!dec_linear(nullptr, nullptr, nullptr, a_attn, d, L.wco, n, d, d, e, dst, a_xn)) return false;
            }
            {   // FFN   whisper.cpp:2443-2478
                GemmEpi e; e.seg[0].bias = L.b1; e.seg[0].gelu = 1; e.seg[0].out16 = a_h; e.seg[0].out16_ld = 4 * d;
                if (!dec_linear(x, L.ln2_g, L.ln2_b, nullptr, 0, L.w1, n, 4 * d, d, e, dst, a_xn)) return false;
                GemmEpi e2; e2.seg[0].bias = L.b2; e2.seg[0].res = x; e2.seg[0].res_ld = d; e2.seg[0].out32 = x; e2.seg[0].out32_ld = d;
                if (!dec_linear(nullptr, nullptr, nullptr, a_h, 4 * d, L.w2, n, d, 4 * d, e2, dst, a_xn)) return false;
            }
        }
        if (n_want > 0) {
            // final LN + logits against the token embedding, only for the rows that were asked for (whisper.cpp:2484-2498)
            prof_begin(PROF_MISC, 0.0, (double) n_want * d * 8);
            launch_gather_rows(x, d_want, a_xw, n_want, d, dst); ++launches;
            prof_end();
            GemmEpi e; e.seg[0].out32 = a_logits; e.seg[0].out32_ld = V;
            if (!dec_linear(a_xw, d_ln_g, d_ln_b, nullptr, 0, d_te, n_want, V, d, e, dst, a_xn)) return false;
            if (n_samp > 0) {
                // rules + log-softmax + greedy pick for the rows that asked for it (the last n_samp wanted rows)
                prof_begin(PROF_MISC, 0.0, (double) n_samp * V * 4 * 5);
                launch_sample_greedy(a_logits + (size_t) n_full * V, n_samp, V, (const int *) (ds + sl.rule), cls_tab, token_beg,
                                     token_eot, sampled_out, dst); ++launches;
                prof_end();
            }
        }
        return true;
    }

    // One launch of the persistent decode-step kernel over the rows described by the staging block at `ds` (decode_step.cu).
    bool launch_step(const uint8_t * ds, float * sampled_dev, const StageLayout & sl, int n, int n_full, int n_groups, const int * n_grp,
                     int n_audio_ctx, int ld_mask) {
        const int V = hp.n_vocab, Lt = hp.n_text_layer;
        StepArgs a;
        a.d = hp.n_text_state; a.n_head = hp.n_text_head; a.n_layer = Lt; a.n_vocab = V;
        const StepPhase * plan_base = step_plans.as<StepPhase>() + (size_t) (n_groups - 1) * (kStepMaxRows + 1) * kStepMaxPhases;
        a.phases = plan_base + (size_t) std::min(n, kStepMaxRows) * kStepMaxPhases; a.n_phases = step_n_phases;
        a.n_groups = n_groups;
        for (int g = 0; g < n_groups; ++g) { a.n_grp[g] = n_grp[g]; a.phases_grp[g] = plan_base + (size_t) n_grp[g] * kStepMaxPhases; }
        a.te = d_te; a.pe = d_pe;
        a.gelu_lut = gelu_lut; a.exp_lut = exp_lut; a.cls = cls_tab; a.token_beg = token_beg; a.token_eot = token_eot;
        a.eps = hp.eps; a.qscale = (float) pow((double) ((float) hp.n_text_state / hp.n_text_head), -0.25);
        a.self_k = self_k.as<__half>(); a.self_v = self_v.as<__half>(); a.cross_k = cross_k.as<__half>(); a.cross_v = cross_v.as<__half>();
        a.kv_cells = kv_cells; a.Tmax = Tmax; a.Tpmax = Tpmax;
        a.n = n; a.n_full = n_full; a.n_audio_ctx = n_audio_ctx; a.ld_mask = ld_mask;
        a.token = (const int *) (ds + sl.token); a.pos = (const int *) (ds + sl.pos); a.wslot = (const int *) (ds + sl.wslot);
        a.rule = (const int *) (ds + sl.rule); a.rowmap_k = (const int *) (ds + sl.rowmap_k); a.rowmap_v = (const int *) (ds + sl.rowmap_v);
        a.koff_self = (const int64_t *) (ds + sl.koff_self); a.voff_self = (const int64_t *) (ds + sl.voff_self);
        a.koff_cross = (const int64_t *) (ds + sl.koff_cross); a.voff_cross = (const int64_t *) (ds + sl.voff_cross);
        a.mask = (const float *) (ds + sl.mask); a.n_kv_dev = (const int *) (ds + sl.nkv);
        a.x32 = dx32.as<float>(); a.q16 = dq16.as<__half>(); a.attn16 = dattn16.as<__half>(); a.h16 = dh16.as<__half>();
        a.logits = dlogits.as<float>(); a.sampled = sampled_dev;
        a.records = step_records.as<double>(); a.bar = step_bar.as<unsigned long long>();
        a.xs_bytes = step_xs; a.slot_bytes = step_slot; a.chunk_keys = step_chunk_keys;
        a.tm_cross_k = step_tm_ck; a.tm_cross_v = step_tm_cv; a.tm_te = step_tm_te; a.chunk_keys_cross = step_chunk_keys_cross;
        // diagnostics: WHISPER_B200_STEP_TRACE=<g> keeps the barrier trace of the most recent launch that had g row groups
        a.trace = (step_trace.p && n_groups == step_trace_groups) ? step_trace.as<unsigned long long>() : nullptr;
        const double w_bytes = 2.0 * ((double) Lt * 14.0 * a.d * a.d + (double) V * a.d) + (double) n * Lt * 4.0 * n_audio_ctx * a.d;
        prof_begin(PROF_STEP, 2.0 * n * ((double) Lt * 14.0 * a.d * a.d + (double) V * a.d), w_bytes);
        const bool ok = launch_decode_step(a, step_grid, step_smem, st);
        prof_end();
        if (!ok) return false;
        ++launches; ++n_step_launches; step_bytes_total += w_bytes;
        return true;
    }

    // Row groups of a decode-step launch over n rows that may be split anywhere (one row per sequence): as even as possible.
    bool split_step_groups(int n, int & n_groups, int * n_grp) const {
        n_groups = (n + kStepMaxRows - 1) / kStepMaxRows;
        if (n_groups > step_groups_max) return false;
        for (int i = 0; i < kStepMaxGroups; ++i) n_grp[i] = 0;
        for (int g = 0; g < n_groups; ++g) n_grp[g] = n / n_groups + (g < n % n_groups ? 1 : 0);
        return true;
    }

    // ---- device-resident greedy runs (Forward::run_*) -------------------------------------------------------------------------

    bool supports_runs() const override { return runs_on && engine != 1; }
    int  run_rows_max() const override { return std::min(run_rows, dec_cap); }
    int  run_depth() const override { return run_depth_; }

    bool run_start(int slot, const RunSeq & init) override {
        CUDA_OK(cudaSetDevice(device));
        if (slot < 0 || slot >= slots || init.pos < 0 || init.pos >= kv_cells || init.token < 0 || init.token >= hp.n_vocab) {
            WB_LOG_ERROR("%s: bad run (slot %d, pos %d, token %d)\n", __func__, slot, init.pos, init.token);
            return false;
        }
        RunSeq * h = run_init_h.as<RunSeq>() + slot;      // (one pinned entry per slot: rewritten only after the slot's previous run is over)
        *h = init;
        CUDA_OK(cudaMemcpyAsync(run_seqs.as<RunSeq>() + slot, h, sizeof(RunSeq), cudaMemcpyHostToDevice, st));
        h2d_bytes += (double) sizeof(RunSeq);
        run_pos_ub[slot] = init.pos;
        return true;
    }

    int run_step_enqueue(const int * row_slots, int n, int n_audio_ctx) override {
        if (cudaSetDevice(device) != cudaSuccess) return -1;
        if (n <= 0 || n > run_rows_max()) { WB_LOG_ERROR("%s: %d rows (max %d)\n", __func__, n, run_rows_max()); return -1; }
        const int Lt = hp.n_text_layer;
        int pos_ub = 0;
        for (int r = 0; r < n; ++r) {
            const int sidx = row_slots[r];
            if (sidx < 0 || sidx >= slots) { WB_LOG_ERROR("%s: bad slot %d\n", __func__, sidx); return -1; }
            if (slot_n_ctx[sidx] != n_audio_ctx) { WB_LOG_ERROR("%s: slot %d was encoded with n_ctx %d, the run asks for %d\n", __func__, sidx, slot_n_ctx[sidx], n_audio_ctx); return -1; }
            pos_ub = std::max(pos_ub, run_pos_ub[sidx]);
        }
        if (pos_ub + 1 > kv_cells) { WB_LOG_ERROR("%s: sequence longer than the cache (%d cells)\n", __func__, kv_cells); return -1; }
        if ((int64_t) (slots + 1) * self_v_slot + kv_cells >= ((int64_t) 1 << 31)) { WB_LOG_ERROR("%s: cache too large for 32-bit row maps\n", __func__); return -1; }
        const bool step_ok = step_usable() && n <= step_rows_max;
        const int n_pad = (step_ok || n <= 32) ? n : std::min(dec_cap, (int) align_up(n, 32));
        const int kvb = std::min(kv_cells, (int) align_up(pos_ub + 1, 128));
        const int ld_mask = kvb;
        const StageLayout sl(dec_cap, kv_cells);
        const int k = (int) (run_ticket % kRunRing);
        int * rows_h = run_rows_h[k].as<int>();
        for (int r = 0; r < n_pad; ++r) rows_h[r] = r < n ? row_slots[r] : -1;
        for (int r = 0; r < n; ++r) run_pos_ub[row_slots[r]] += 1;
        cudaEventRecord(run_ev0[k], st);
        if (cudaMemcpyAsync(run_rows_d.p, rows_h, (size_t) n_pad * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        h2d_bytes += (double) n_pad * 4;
        uint8_t * ds = dstage_run.as<uint8_t>();
        auto body = [&]() -> bool {
            if (cudaMemsetAsync(ds + sl.nkv, 0, 4, st) != cudaSuccess) return false;
            RunPrepArgs pa;
            pa.seqs = run_seqs.as<RunSeq>(); pa.row_slot = run_rows_d.as<int>();
            pa.token = (int *) (ds + sl.token); pa.pos = (int *) (ds + sl.pos); pa.want = (int *) (ds + sl.want); pa.wslot = (int *) (ds + sl.wslot);
            pa.rule = (int *) (ds + sl.rule); pa.rowmap_k = (int *) (ds + sl.rowmap_k); pa.rowmap_v = (int *) (ds + sl.rowmap_v);
            pa.koff_self = (int64_t *) (ds + sl.koff_self); pa.voff_self = (int64_t *) (ds + sl.voff_self);
            pa.koff_cross = (int64_t *) (ds + sl.koff_cross); pa.voff_cross = (int64_t *) (ds + sl.voff_cross);
            pa.mask = (float *) (ds + sl.mask); pa.ld_mask = ld_mask; pa.n_kv = (int *) (ds + sl.nkv);
            pa.n_slots = slots; pa.n_layer = Lt; pa.kv_cells = kv_cells; pa.token_beg = token_beg;
            pa.self_k_slot = self_k_slot; pa.self_v_slot = self_v_slot; pa.cross_k_slot = cross_k_slot; pa.cross_v_slot = cross_v_slot;
            prof_begin(PROF_MISC, 0.0, (double) n_pad * (ld_mask * 4.0 + 128.0));
            launch_run_prep(pa, n_pad, st); ++launches;
            prof_end();
            if (step_ok) {
                int n_groups = 1, n_grp[kStepMaxGroups];
                if (!split_step_groups(n_pad, n_groups, n_grp)) return false;
                if (!launch_step(ds, dsampled_run.as<float>(), sl, n_pad, 0, n_groups, n_grp, n_audio_ctx, ld_mask)) return false;
            } else if (!enqueue_decode(n_pad, 0, n_pad, kvb, ld_mask, n_audio_ctx, sl, 2, st, 0, pos_ub + 1)) return false;
            prof_begin(PROF_MISC, 0.0, (double) n_pad * 160.0);
            launch_run_advance(run_seqs.as<RunSeq>(), run_rows_d.as<int>(), n_pad, dsampled_run.as<float>(), run_tokens.as<float>(),
                               run_status_d.as<int>(), token_beg, token_eot, st); ++launches;
            prof_end();
            return true;
        };
        // the kernels of a step never change between steps of the same shape (everything that does lives in device memory): wide steps
        // are captured once per (rows, key bucket, audio ctx) and replayed; decode-step launches are cooperative and go out directly
        const DecodeShape shape{n_pad, 0, n_pad, kvb, n_audio_ctx, engine, 2};
        bool done = false;
        if (use_graphs && !prof_on && !step_ok) {
            auto it = graphs.find(shape);
            if (it != graphs.end()) {
                if (cudaGraphLaunch(it->second, st) != cudaSuccess) return -1;
                launches += graph_nodes[shape];
                done = true;
            } else if (++graph_seen[shape] >= 2) {
                const int64_t l0 = launches.load();
                cudaGraph_t g = nullptr;
                if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return -1;
                const bool ok = body();
                const cudaError_t ce = cudaStreamEndCapture(st, &g);       // (always ended: a failed body must not leave the stream capturing)
                if (!ok || ce != cudaSuccess || !g) {
                    if (g) cudaGraphDestroy(g);
                    WB_LOG_ERROR("%s: graph capture failed: %s\n", __func__, cudaGetErrorString(ce));
                    return -1;
                }
                cudaGraphExec_t ge = nullptr;
                if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { cudaGraphDestroy(g); return -1; }
                cudaGraphDestroy(g);
                graphs[shape] = ge;
                graph_nodes[shape] = launches - l0;
                launches = l0;
                if (cudaGraphLaunch(ge, st) != cudaSuccess) return -1;
                launches += graph_nodes[shape];
                done = true;
            }
        }
        if (!done && !body()) return -1;
        if (cudaMemcpyAsync(run_status_h[k].p, run_status_d.p, (size_t) n * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
        d2h_bytes += (double) n * 4;
        cudaEventRecord(run_ev1[k], st);
        run_rows_of[k] = n;
        return (int) (run_ticket++ & 0x3fffffff);
    }

    bool run_step_wait(int ticket, int32_t * status) override {
        CUDA_OK(cudaSetDevice(device));
        const int k = ticket % kRunRing;
        CUDA_OK(cudaEventSynchronize(run_ev1[k]));
        CUDA_OK(cudaGetLastError());
        { float ms = 0.0f; if (cudaEventElapsedTime(&ms, run_ev0[k], run_ev1[k]) == cudaSuccess) { t_dec_ms += ms; ++n_dec_calls; } }
        note_busy(run_ev0[k], run_ev1[k]);
        memcpy(status, run_status_h[k].p, (size_t) run_rows_of[k] * 4);
        if (prof_on && !pend[0].active && !pend[1].active) { cudaStreamSynchronize(st); prof_collect(); }
        return true;
    }

    bool run_fetch(int slot, int ticket, RunSeq & out, std::vector<whisper_token_data> & tokens) override {
        CUDA_OK(cudaSetDevice(device));
        if (slot < 0 || slot >= slots) return false;
        const size_t entry = sizeof(RunSeq) + (size_t) kRunTokenCap * 6 * 4;
        uint8_t * h = run_fetch_h.as<uint8_t>() + (size_t) slot * entry;
        // on the copy stream, behind the step in which the sequence finished — not behind the steps queued after it
        CUDA_OK(cudaStreamWaitEvent(st_copy, run_ev1[ticket % kRunRing], 0));
        CUDA_OK(cudaMemcpyAsync(h, run_seqs.as<RunSeq>() + slot, sizeof(RunSeq), cudaMemcpyDeviceToHost, st_copy));
        CUDA_OK(cudaMemcpyAsync(h + sizeof(RunSeq), run_tokens.as<float>() + (size_t) slot * kRunTokenCap * 6, (size_t) kRunTokenCap * 6 * 4,
                                cudaMemcpyDeviceToHost, st_copy));
        CUDA_OK(cudaEventRecord(ev_copy, st_copy));
        CUDA_OK(cudaEventSynchronize(ev_copy));
        memcpy(&out, h, sizeof(RunSeq));
        const int n_out = std::min(out.n_out, (int32_t) kRunTokenCap);
        d2h_bytes += (double) entry;
        const float * o = (const float *) (h + sizeof(RunSeq));
        tokens.clear();
        tokens.reserve(n_out);
        for (int i = 0; i < n_out; ++i, o += 6) {
            whisper_token_data td = { 0, 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, 0.0f };
            memcpy(&td.id, &o[0], 4); memcpy(&td.tid, &o[1], 4);
            td.p = o[2]; td.plog = o[3]; td.pt = o[4]; td.ptsum = o[5];
            tokens.push_back(td);
        }
        return true;
    }

    bool decode_batch(const DecodeJob * jobs, int n_jobs, int n_audio_ctx) override {
        return decode_enqueue(jobs, n_jobs, n_audio_ctx, 0) && decode_collect(0);
    }
    bool step_usable() const { return use_step && step_grid > 0 && engine == 0 && !force_multi && step_rows_max > 0; }
    int decode_rows_per_pass() const override {
        const int r = step_usable() ? step_rows_max : kStepMaxRows;
        return (engine == 0 && !force_multi) ? std::max(r, wide_rows) : r;
    }
    int decode_sets() const override { return (engine == 0 && !force_multi && !prof_on && (step_usable() || wide_rows > 0)) ? 2 : 1; }

    // Stages one decoder pass and queues it on the stream (host->device copy, kernels, device->host copy of the results, event).
    // Set 1 has its own staging buffers, so it can be filled while the pass of set 0 still runs (and vice versa); it only takes
    // passes the persistent decode-step kernel serves with device-side sampling.
    bool decode_enqueue(const DecodeJob * jobs, int n_jobs, int n_audio_ctx, int set) override {
        CUDA_OK(cudaSetDevice(device));
        if (set < 0 || set > 1 || pend[set].active) { WB_LOG_ERROR("%s: staging set %d is busy\n", __func__, set); return false; }
        PinnedBuf & hstage_s = set ? hstage2 : hstage;  DevBuf & dstage_s = set ? dstage2 : dstage;
        PinnedBuf & hsampled_s = set ? hsampled2 : hsampled;  DevBuf & dsampled_s = set ? dsampled2 : dsampled;
        cudaEvent_t ev0 = set ? ev2_call0 : ev_call0, ev1 = set ? ev2_call1 : ev_call1;
        const int V = hp.n_vocab, Lt = hp.n_text_layer;
        int n = 0, n_kv = 0, n_full = 0, n_samp = 0, n_dist = 0, n_draws_total = 0;
        for (int j = 0; j < n_jobs; ++j) {
            const DecodeInput & in = jobs[j].in;
            if (jobs[j].slot < 0 || jobs[j].slot >= slots) { WB_LOG_ERROR("%s: bad slot %d\n", __func__, jobs[j].slot); return false; }
            if (in.n_tokens <= 0 || in.kv_head + in.n_tokens > kv_cells || in.n_kv > kv_cells) { WB_LOG_ERROR("%s: bad batch\n", __func__); return false; }
            if (slot_n_ctx[jobs[j].slot] != n_audio_ctx) {
                WB_LOG_ERROR("%s: slot %d was encoded with n_ctx %d, decode asks for %d\n", __func__, jobs[j].slot, slot_n_ctx[jobs[j].slot], n_audio_ctx);
                return false;
            }
            n += in.n_tokens;
            n_kv = std::max(n_kv, in.n_kv);
            for (int i = 0; i < in.n_tokens; ++i) if (in.want_logits[i]) {
                if (in.sample && in.n_draws && in.n_draws[i] > 0) { ++n_dist; n_draws_total += in.n_draws[i]; }
                else if (in.sample) ++n_samp;
                else ++n_full;
            }
        }
        if (!ensure_dec(n)) return false;
        if (n_dist > 0 && (set != 0 || n_draws_total > dec_cap * kDrawsPerRow)) { WB_LOG_ERROR("%s: %d draws do not fit this pass\n", __func__, n_draws_total); return false; }
        // Rows sampled from their distribution look like rows that want full logits to the decoder kernels (the logits stay on the device,
        // launch_sample_dist reads them there): from here on n_full counts both, n_host only the rows whose logits travel to the host.
        const int n_host = n_full;
        n_full += n_dist;
        const StageLayout sl(dec_cap, kv_cells);
        uint8_t * hs = hstage_s.as<uint8_t>();
        int32_t * h_token = (int32_t *) (hs + sl.token), * h_pos = (int32_t *) (hs + sl.pos), * h_want = (int32_t *) (hs + sl.want);
        int32_t * h_rk = (int32_t *) (hs + sl.rowmap_k), * h_rv = (int32_t *) (hs + sl.rowmap_v);
        int64_t * h_ks = (int64_t *) (hs + sl.koff_self), * h_vs = (int64_t *) (hs + sl.voff_self);
        int64_t * h_kc = (int64_t *) (hs + sl.koff_cross), * h_vc = (int64_t *) (hs + sl.voff_cross);
        float * h_mask = (float *) (hs + sl.mask);
        const int kvb = std::min(kv_cells, (int) align_up(std::max(n_kv, 1), 128));   // key-count bucket (part of the graph shape)
        const int ld_mask = kvb;
        *(int32_t *) (hs + sl.nkv) = n_kv;
        const int n_want = n_full + n_samp;
        int32_t * h_rule = (int32_t *) (hs + sl.rule);
        int32_t * h_wslot = (int32_t *) (hs + sl.wslot);
        // the persistent step kernel serves steps in which every row is the single new token of its own sequence
        bool step_ok = step_usable() && n <= step_rows_max && n_want == n;
        // row groups of the launch: whole jobs, at most kStepMaxRows rows each, as evenly as the job sizes allow
        int n_groups = 1, n_grp[kStepMaxGroups] = {0, 0, 0, 0};
        if (step_ok) {
            n_groups = (n + kStepMaxRows - 1) / kStepMaxRows;
            for (;; ++n_groups) {
                if (n_groups > step_groups_max) { step_ok = false; break; }
                const int target = (n + n_groups - 1) / n_groups;
                int g = 0; bool fits = true;
                for (int i = 0; i < kStepMaxGroups; ++i) n_grp[i] = 0;
                for (int j = 0; j < n_jobs && fits; ++j) {
                    const int t = jobs[j].in.n_tokens;
                    if (n_grp[g] > 0 && n_grp[g] + t > std::min(kStepMaxRows, std::max(target, t))) ++g;
                    if (g >= n_groups || t > kStepMaxRows) { fits = false; break; }
                    n_grp[g] += t;
                }
                if (fits) { n_groups = g + 1; break; }
            }
        }
        {
            // wanted rows: those that need full logits first, then the ones sampled on the device
            int r = 0, w = 0, ws = 0, wd = 0, draw_off = 0;
            int32_t * h_drule = (int32_t *) (hs + sl.drule);
            double * h_draws = (double *) (hs + sl.draws);
            for (int j = 0; j < n_jobs; ++j) {
                const DecodeInput & in = jobs[j].in;
                const int64_t slot = jobs[j].slot;
                int job_draw = 0;
                for (int i = 0; i < in.n_tokens; ++i, ++r) {
                    h_token[r] = in.token[i]; h_pos[r] = in.pos[i];
                    h_wslot[r] = -1;
                    for (int i2 = 0; i2 < i; ++i2) if (in.seq[i2] == in.seq[i]) step_ok = false;
                    if (in.token[i] < 0 || in.token[i] >= V || in.pos[i] < 0 || in.pos[i] >= hp.n_text_ctx) {
                        WB_LOG_ERROR("%s: token %d / position %d out of range\n", __func__, in.token[i], in.pos[i]);
                        return false;
                    }
                    if (in.want_logits[i] && in.sample && in.n_draws && in.n_draws[i] > 0) {
                        const int idx = n_host + wd;
                        h_want[idx] = r; h_wslot[r] = idx;
                        int32_t * dr = h_drule + 8 * wd;
                        dr[0] = in.sample[i].flags; dr[1] = in.sample[i].tid0_initial; dr[2] = in.sample[i].tid0_seek; dr[3] = in.n_draws[i];
                        memcpy(&dr[4], &in.temperature, 4); dr[5] = draw_off; dr[6] = in.tid_default; dr[7] = 0;
                        memcpy(h_draws + draw_off, in.draws + job_draw, sizeof(double) * (size_t) in.n_draws[i]);
                        draw_off += in.n_draws[i]; job_draw += in.n_draws[i];
                        ++wd;
                    } else if (in.want_logits[i]) {
                        if (in.sample) {
                            h_want[n_full + ws] = r;
                            h_wslot[r] = n_full + ws;
                            memcpy(h_rule + 4 * ws, &in.sample[i], 16);
                            ++ws;
                        } else {
                            h_wslot[r] = w;
                            h_want[w++] = r;
                        }
                    }
                    // K rows: [slot][layer][cell][d] -> row index relative to the layer base; V^T columns likewise
                    h_rk[r] = (int32_t) (slot * (int64_t) Lt * kv_cells + in.kv_head + i);
                    h_rv[r] = (int32_t) (slot * self_v_slot + in.kv_head + i);
                    h_ks[r] = slot * self_k_slot;  h_vs[r] = slot * self_v_slot;
                    h_kc[r] = slot * cross_k_slot; h_vc[r] = slot * cross_v_slot;
                    // visibility mask (whisper.cpp:2203-2226): cell must hold the row's sequence and not lie in its future
                    float * m = h_mask + (size_t) r * ld_mask;
                    for (int c = 0; c < n_kv; ++c) {
                        const bool vis = c < in.n_kv && in.cells[c].has_seq(in.seq[i]) && in.cells[c].pos <= in.pos[i];
                        m[c] = vis ? 0.0f : -INFINITY;
                    }
                }
            }
        }
        if ((int64_t) (slots + 1) * self_v_slot + kv_cells >= ((int64_t) 1 << 31)) { WB_LOG_ERROR("%s: cache too large for 32-bit row maps\n", __func__); return false; }
        // Wide passes (every row the single new token of a sequence sampled on the device, more rows than the decode-step kernel
        // takes) are padded to a multiple of 32 rows so that a handful of CUDA graphs serves every pass: a padding row repeats
        // row 0 and writes its K / V into the scratch slot behind the last real one; its sample is never read.
        const int n_real = n, n_samp_real = n_samp;
        if (!step_ok && n_full == 0 && n_samp == n && n > 32) {
            const int n_pad = std::min(dec_cap, (int) align_up(n, 32));
            for (int r = n; r < n_pad; ++r) {
                h_token[r] = h_token[0]; h_pos[r] = h_pos[0];
                h_want[r] = r; h_wslot[r] = r;
                memcpy(h_rule + 4 * r, h_rule, 16);
                h_rk[r] = (int32_t) ((int64_t) slots * Lt * kv_cells + (r - n));
                h_rv[r] = (int32_t) ((int64_t) slots * self_v_slot + (r - n));
                h_ks[r] = h_ks[0]; h_vs[r] = h_vs[0]; h_kc[r] = h_kc[0]; h_vc[r] = h_vc[0];
                memcpy(h_mask + (size_t) r * ld_mask, h_mask, (size_t) n_kv * 4);
            }
            n = n_pad; n_samp = n_pad;
        }
        const size_t stage_bytes = sl.mask + (size_t) n * ld_mask * 4;
        if (set != 0 && n_full != 0) { WB_LOG_ERROR("%s: staging set 1 only takes passes that are sampled on the device\n", __func__); return false; }
        // stream of this pass: wide passes of staging set 1 go to the second decoder stream (with their own activations); decode-step
        // launches (cooperative, one CTA per SM, set-0 activations in their plans) and everything of set 0 stay on the first
        const int aset = (set == 1 && !step_ok && st_dec2) ? 1 : 0;
        const cudaStream_t ps = aset ? st_dec2 : st;
        cudaEventRecord(ev0, ps);
        CUDA_OK(cudaMemcpyAsync(dstage_s.p, hs, stage_bytes, cudaMemcpyHostToDevice, ps));
        h2d_bytes += (double) stage_bytes;
        // The kernels of one decode step.  Everything that changes from step to step (tokens, positions, cache cells, mask, live
        // key count) lives in the staging block, not in launch arguments, so a step shape (rows, wanted rows, key bucket,
        // audio ctx) that has been seen twice is captured once and replayed as a CUDA graph.
        if (step_ok) {
            if (!launch_step(dstage_s.as<uint8_t>(), dsampled_s.as<float>(), sl, n, n_full, n_groups, n_grp, n_audio_ctx, ld_mask)) return false;
        } else {
            const DecodeShape shape{n, n_full, n_samp, kvb, n_audio_ctx, engine, set};   // (the set fixes stream, staging block and activations)
            bool replayed = false;
            if (use_graphs && !prof_on) {
                auto it = graphs.find(shape);
                if (it != graphs.end()) {
                    CUDA_OK(cudaGraphLaunch(it->second, ps));
                    launches += graph_nodes[shape];
                    replayed = true;
                } else if (++graph_seen[shape] >= 2) {
                    const int64_t l0 = launches.load();
                    cudaGraph_t g = nullptr;
                    CUDA_OK(cudaStreamBeginCapture(ps, cudaStreamCaptureModeThreadLocal));
                    const bool ok = enqueue_decode(n, n_full, n_samp, kvb, ld_mask, n_audio_ctx, sl, set, ps, aset, n_kv);
                    const cudaError_t ce = cudaStreamEndCapture(ps, &g);
                    if (!ok || ce != cudaSuccess || !g) {      // (the capture is always ended, so the stream stays usable; the partial graph is dropped)
                        if (g) cudaGraphDestroy(g);
                        WB_LOG_ERROR("%s: graph capture failed: %s\n", __func__, cudaGetErrorString(ce));
                        return false;
                    }
                    cudaGraphExec_t ge = nullptr;
                    CUDA_OK(cudaGraphInstantiate(&ge, g, 0));
                    cudaGraphDestroy(g);
                    graphs[shape] = ge;
                    graph_nodes[shape] = launches - l0;
                    launches = l0;
                    CUDA_OK(cudaGraphLaunch(ge, ps));
                    launches += graph_nodes[shape];
                    replayed = true;
                }
            }
            if (!replayed && !enqueue_decode(n, n_full, n_samp, kvb, ld_mask, n_audio_ctx, sl, set, ps, aset, n_kv)) return false;
        }
        if (n_want > 0) {
            if (n_dist > 0) {
                // rules + temperature + one token per uniform variate, on the logits where they lie; 24 bytes per draw travel back
                const uint8_t * ds = dstage_s.as<uint8_t>();
                prof_begin(PROF_MISC, 0.0, (double) n_dist * V * 4 * 6);
                launch_sample_dist((aset ? act_b.logits : dlogits).as<float>() + (size_t) n_host * V, n_dist, V, (const int *) (ds + sl.drule),
                                   (const double *) (ds + sl.draws), cls_tab, token_beg, token_eot, ddist.as<float>(), ps); ++launches;
                prof_end();
                CUDA_OK(cudaMemcpyAsync(hdist.p, ddist.p, (size_t) n_draws_total * 24, cudaMemcpyDeviceToHost, ps));
                d2h_bytes += (double) n_draws_total * 24;
            }
            if (n_host > 0) {
                CUDA_OK(cudaMemcpyAsync(hlogits.p, (aset ? act_b.logits : dlogits).p, (size_t) n_host * V * 4, cudaMemcpyDeviceToHost, ps));
                d2h_bytes += (double) n_host * V * 4;
            }
            if (n_samp > 0) {
                CUDA_OK(cudaMemcpyAsync(hsampled_s.p, dsampled_s.p, (size_t) n_samp * 24, cudaMemcpyDeviceToHost, ps));
                d2h_bytes += (double) n_samp * 24;
            }
        }
        cudaEventRecord(ev1, ps);
        PendingPass & pp = pend[set];
        pp.active = true; pp.jobs.assign(jobs, jobs + n_jobs); pp.n_full = n_host; pp.n_samp = n_samp_real; pp.n_dist = n_dist;
        (void) n_real;
        return true;
    }

    // Waits for the pass queued on `set` and hands its results (sampled tokens / logits rows) to the jobs.
    bool decode_collect(int set) override {
        if (set < 0 || set > 1 || !pend[set].active) { WB_LOG_ERROR("%s: nothing queued on staging set %d\n", __func__, set); return false; }
        CUDA_OK(cudaSetDevice(device));
        PendingPass & pp = pend[set];
        pp.active = false;
        PinnedBuf & hsampled_s = set ? hsampled2 : hsampled;
        cudaEvent_t ev0 = set ? ev2_call0 : ev_call0, ev1 = set ? ev2_call1 : ev_call1;
        const int V = hp.n_vocab;
        CUDA_OK(cudaEventSynchronize(ev1));
        CUDA_OK(cudaGetLastError());
        { float ms = 0.0f; if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) { t_dec_ms += ms; ++n_dec_calls; } }
        note_busy(ev0, ev1);
        if (!pend[0].active && !pend[1].active) prof_collect();
        int w = 0, ws = 0, wdraw = 0;
        auto unpack = [](const float * o) {
            whisper_token_data td = { 0, 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, 0.0f };
            memcpy(&td.id, &o[0], 4); memcpy(&td.tid, &o[1], 4);
            td.p = o[2]; td.plog = o[3]; td.pt = o[4]; td.ptsum = o[5];
            return td;
        };
        for (const DecodeJob & job : pp.jobs) {
            const DecodeInput & in = job.in;
            int job_draw = 0;
            for (int i = 0; i < in.n_tokens; ++i) {
                if (!in.want_logits[i]) continue;
                if (in.sample && in.n_draws && in.n_draws[i] > 0) {
                    for (int dd = 0; dd < in.n_draws[i]; ++dd, ++wdraw, ++job_draw)
                        if (job.dist_out) job.dist_out[job_draw] = unpack(hdist.as<float>() + 6 * (size_t) wdraw);
                } else if (in.sample) {
                    const float * o = hsampled_s.as<float>() + 6 * (size_t) ws++;
                    whisper_token_data td = { 0, 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, 0.0f };
                    memcpy(&td.id, &o[0], 4); memcpy(&td.tid, &o[1], 4);
                    td.p = o[2]; td.plog = o[3]; td.pt = o[4]; td.ptsum = o[5];
                    if (job.sampled_out) job.sampled_out[i] = td;
                } else {
                    memcpy(job.logits_out + (size_t) i * V, hlogits.as<float>() + (size_t) w * V, (size_t) V * 4);
                    ++w;
                }
            }
        }
        return true;
    }

    // ---- stage probes ------------------------------------------------------------------------------------------------

    long long read_stage(int what, void * dst, long long cap) override { return read_stage_slot(0, what, dst, cap); }

    long long read_stage_slot(int slot, int what, void * dst, long long cap) override {
        cudaSetDevice(device);
        cudaStreamSynchronize(st);
        const int d = hp.n_audio_state, T = enc_last_T, Lt = hp.n_text_layer;
        if (slot < 0 || slot >= slots) return -1;
        // encoder-side stages address the chunk that was encoded into `slot` most recently: chunk index == position in
        // the last batch; only batch position 0 is exposed for slot 0 (tests use single-chunk encodes for these probes)
        auto copy_out = [&](const void * src_host, long long nbytes) {
            if (dst) memcpy(dst, src_host, (size_t) std::min(nbytes, cap));
            return nbytes;
        };
        switch (what) {
            case STAGE_MEL_WINDOW: {
                if (T == 0) return -1;
                const long long nb = (long long) hp.n_mels * 2 * T * 4;
                std::vector<uint8_t> tmp(nb);
                cudaMemcpy(tmp.data(), mel_d.p, nb, cudaMemcpyDeviceToHost);
                return copy_out(tmp.data(), nb);
            }
            case STAGE_EMBD_CONV: {
                if (T == 0) return -1;
                std::vector<uint16_t> hbuf((size_t) T * d);
                cudaMemcpy(hbuf.data(), conv16.p, hbuf.size() * 2, cudaMemcpyDeviceToHost);
                std::vector<float> f(hbuf.size());
                for (size_t i = 0; i < f.size(); ++i) f[i] = f16_to_f32(hbuf[i]);
                return copy_out(f.data(), (long long) f.size() * 4);
            }
            case STAGE_EMBD_ENC: {
                if (T == 0) return -1;
                const long long nb = (long long) T * d * 4;
                std::vector<uint8_t> tmp(nb);
                cudaMemcpy(tmp.data(), enc32.p, nb, cudaMemcpyDeviceToHost);
                return copy_out(tmp.data(), nb);
            }
            case STAGE_CROSS_K: {     // -> [Lt][T][d]
                const int Ts = slot_n_ctx[slot];
                if (Ts == 0) return -1;
                std::vector<uint16_t> out((size_t) Lt * Ts * d);
                for (int il = 0; il < Lt; ++il)
                    cudaMemcpy(out.data() + (size_t) il * Ts * d, cross_k.as<__half>() + slot * cross_k_slot + (int64_t) il * Tmax * d,
                               (size_t) Ts * d * 2, cudaMemcpyDeviceToHost);
                return copy_out(out.data(), (long long) out.size() * 2);
            }
            case STAGE_CROSS_V: {     // -> [Lt][d][T]
                const int Ts = slot_n_ctx[slot];
                if (Ts == 0) return -1;
                std::vector<uint16_t> out((size_t) Lt * d * Ts);
                cudaMemcpy2D(out.data(), (size_t) Ts * 2, cross_v.as<__half>() + slot * cross_v_slot, (size_t) Tpmax * 2, (size_t) Ts * 2,
                             (size_t) Lt * d, cudaMemcpyDeviceToHost);
                return copy_out(out.data(), (long long) out.size() * 2);
            }
            case STAGE_SELF_K: {      // [Lt][cells][d]
                const long long nb = self_k_slot * 2;
                std::vector<uint8_t> tmp(nb);
                cudaMemcpy(tmp.data(), self_k.as<__half>() + slot * self_k_slot, nb, cudaMemcpyDeviceToHost);
                return copy_out(tmp.data(), nb);
            }
            case STAGE_SELF_V: {      // [Lt][d][cells]
                const long long nb = self_v_slot * 2;
                std::vector<uint8_t> tmp(nb);
                cudaMemcpy(tmp.data(), self_v.as<__half>() + slot * self_v_slot, nb, cudaMemcpyDeviceToHost);
                return copy_out(tmp.data(), nb);
            }
            case 9: {                 // the slot's device-computed spectrogram, normalised like whisper.cpp:2856-2871 -> f32 [n_mel][n_len]
                if (!mel_dev_on || mel_n_len[slot] <= 0) return -1;
                const int n_calc = mel_n_calc[slot], n_len = mel_n_len[slot], nm = filt_n_mel;
                std::vector<float> raw((size_t) n_calc * nm);
                int mx = 0;
                cudaStreamSynchronize(st_enc);
                cudaMemcpy(raw.data(), raw_mel.as<float>() + (size_t) slot * kMelFramesCap * nm, raw.size() * 4, cudaMemcpyDeviceToHost);
                cudaMemcpy(&mx, mel_max.as<int>() + slot, 4, cudaMemcpyDeviceToHost);
                float fmax = mel_ordered_to_float(mx);
                if (n_calc < n_len) fmax = std::max(fmax, mel_low);
                const double mmax = (double) fmax - 8.0;
                const float fclamp = (float) mmax;
                std::vector<float> out((size_t) nm * n_len);
                for (int m = 0; m < nm; ++m)
                    for (int i = 0; i < n_len; ++i) {
                        float v = i < n_calc ? raw[(size_t) i * nm + m] : mel_low;
                        if (v < mmax) v = fclamp;
                        out[(size_t) m * n_len + i] = (float) ((v + 4.0) / 4.0);
                    }
                return copy_out(out.data(), (long long) out.size() * 4);
            }
            case 8: {                 // barrier trace of the most recent decode-step launch: u64 [grid][kStepMaxPhases][8]
                if (!step_trace.p) return -1;
                const long long nb = (long long) step_grid * kStepMaxPhases * 8 * 8;
                std::vector<uint8_t> tmp(nb);
                cudaMemcpy(tmp.data(), step_trace.p, nb, cudaMemcpyDeviceToHost);
                return copy_out(tmp.data(), nb);
            }
            default: return -1;
        }
    }
};

}  // namespace

void * host_alloc_pinned(size_t bytes) {
    void * p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void host_free_pinned(void * p) { if (p) cudaFreeHost(p); }

Forward * create_forward(const ModelFile & model, int kv_self_cells, int device) {
    CudaForward * f = new CudaForward;
    if (!f->init(model, kv_self_cells, device)) {
        delete f;
        return nullptr;
    }
    return f;
}

}  // namespace wb200
