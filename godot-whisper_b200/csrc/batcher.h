// Batching of device passes across chunk workers, and the scheduler of the device-resident greedy runs.
//
// whisper_b200_full_batch runs one host thread per in-flight chunk; each thread executes the ordinary whisper_full() state machine
// (csrc/full.cpp) on its own whisper_state and device slot.  What those threads ask of the device meets here:
//   * encoder requests are merged into batched encoder passes by the encoder driver thread (own CUDA stream);
//   * a greedy t = 0 sequence is ONE request — a run (run_state.h): the decoder driver keeps the steps of all live runs queued on the
//     device a few steps ahead, learns from the status words of older steps which sequences have finished, hands those their
//     tokens and admits waiting runs in their place.  The thread of a chunk sleeps through its whole token loop; nothing on the
//     host happens per token;
//   * everything else (prompt prefill, beam search, best-of sampling at t > 0, whisper_decode) arrives as ordinary decoder requests
//     that are merged into passes as before: a pass goes when it is full or nobody could add to it.
// The CPU analogue in the reference is whisper_full_parallel (/root/reference/thirdparty/whisper.cpp/whisper.cpp:5817-5930): one
// state per worker, shared read-only weights — but there each worker computes alone.
#pragma once

#include "forward.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace wb200 {

class Batcher {
public:
    explicit Batcher(Forward * fwd) : fwd_(fwd) {}
    ~Batcher();

    // Worker registration (a thread that will issue passes through this batcher until it calls worker_end).
    void add_workers(int n);     // called once by the spawning thread: n workers are about to attach
    void worker_attach();        // called by each worker thread itself
    void worker_begin() { add_workers(1); worker_attach(); }
    void worker_end();
    // A registered worker entering / leaving a long host-only phase (log-mel of its next chunk): while there, nobody's batch
    // waits for it; at most max_host_ workers are there at once.  No-ops on threads that are not registered workers.
    void host_phase_begin();
    void host_phase_end();
    // A registered worker entering / leaving the decoding part of its chunk.  At most max_decode_workers_ workers decode at
    // once; the others — already encoded — wait here without holding up anybody's batch.  No-ops on unregistered threads.
    void decode_phase_begin();
    void decode_phase_end();
    void set_max_host(int n) { std::lock_guard<std::mutex> lk(mu_); max_host_ = n < 1 ? 1 : n; }

    // Blocking; safe to call from an unregistered thread when no workers exist (runs immediately, batch of one).
    bool encode(int slot, const float * mel_window, int n_ctx);
    // ... from the slot's device-resident spectrogram at frame mel_offset; pcm != nullptr (staged, Forward::pcm_stage_acquire): compute it first
    bool encode_pcm(int slot, const float * pcm, int n_samples, int mel_offset, int n_ctx, float * energy_out = nullptr);
    bool decode(int slot, const DecodeInput & in, int n_audio_ctx, float * logits_out, whisper_token_data * sampled_out = nullptr,
                whisper_token_data * dist_out = nullptr);
    // One greedy run (Forward::run_*): returns when the sequence has completed / failed / run out of steps, with its final state and tokens.
    bool run(int slot, const RunSeq & init, int n_audio_ctx, RunSeq & final_state, std::vector<whisper_token_data> & tokens);

    struct RequestView { int kind; int n_tokens; bool sampled; };

    // statistics: device passes issued / requests served (requests / passes = achieved batching factor)
    std::atomic<int64_t> n_passes{0}, n_requests{0}, n_run_steps{0}, n_run_rows{0};
    // where the driver thread spends its time (us): staging + queueing passes, waiting for the device, waking workers, idle
    int64_t t_stage_us = 0, t_device_wait_us = 0, t_complete_us = 0, t_idle_us = 0, t_run_us = 0;

private:
    struct Request {
        int kind = 0;                 // 0 encode, 1 decode, 2 run
        int slot = 0;
        int n_ctx = 0;
        const float * mel = nullptr;
        const float * pcm = nullptr;
        int n_samples = 0, mel_offset = -1;
        float * energy_out = nullptr;
        DecodeInput in;
        float * logits = nullptr;
        whisper_token_data * sampled = nullptr;
        whisper_token_data * dist = nullptr;
        RunSeq run_init;
        RunSeq * run_final = nullptr;
        std::vector<whisper_token_data> * run_tokens = nullptr;
        bool ok = false;
        std::chrono::steady_clock::time_point t_submit;
        // completion is signalled per request: a served thread wakes without touching the shared lock
        std::mutex m;
        std::condition_variable cv;
        bool done = false;
    };
    bool submit(Request & r);
    bool pick(std::vector<Request *> & batch);      // the batching policy; called with mu_ held
    bool pick_encode(std::vector<Request *> & batch);
    bool pick_decode(std::vector<Request *> & batch);
    void encoder_loop();                            // second driver thread: encoder passes on the forward's encoder stream
    void run(std::vector<Request *> & batch);
    void complete(std::vector<Request *> & batch);
    void driver_loop();
    bool run_alone(Request & r);                    // a run from a thread that is not a worker: driven right there
    void wake_driver() { cv_drv_.notify_one(); cv_enc_.notify_one(); }

    Forward * fwd_;
    std::mutex mu_;
    std::condition_variable cv_host_, cv_dec_, cv_drv_;
    std::thread driver_, enc_driver_;
    std::condition_variable cv_enc_;
    int inflight_enc_ = 0, inflight_dec_ = 0;     // requests inside a device pass right now
    int live_runs_ = 0;                           // runs admitted to the device and not yet handed back
    bool driver_started_ = false, stop_ = false;
    int active_ = 0;                  // registered workers that are neither on the host nor waiting for a decode seat
    int in_host_ = 0;                 // workers inside a host-only phase
    int in_decode_ = 0;               // workers inside the decoding part of a chunk
    int max_decode_workers_ = 48;
    int max_host_ = 1 << 30;          // how many may be in a host phase at once (set_max_host: one per core)
    std::vector<Request *> pending_enc_, pending_dec_, pending_run_;
    bool is_worker() const;           // the caller is a registered worker thread of this batcher
    int max_encode_batch_ = 64;       // chunks per encoder pass (measured: 32 beats 16 by 10 % of encoder time, 64 beats 32 by another 8 %; activations 3.6 GB)
    int encode_batch_target_ = 32;    // hold encode requests until this many wait (or nothing else can run)
    int encode_grace_us_ = 5000;      // ... but never longer than this
    int pass_split_ = 2;              // decoding workers are served as this many alternating passes (WHISPER_B200_PASS_SPLIT)
    bool host_batch_policy_ = true;   // log-mel phases run under SCHED_BATCH (WHISPER_B200_HOST_BATCH_POLICY=0: leave the policy alone)
    int pass_min_rows_ = 16;          // ... but a pass is never cut below this many rows while more could come
    int run_min_rows_ = 1;            // run steps wait for this many live runs while more are on their way (WHISPER_B200_RUN_MIN_ROWS)
    int max_decode_rows_ = 16;        // rows per decoder pass: what the persistent decode-step kernel takes in one launch
};

}  // namespace wb200
