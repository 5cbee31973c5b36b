// Leader/follower batching of device passes across chunk workers.
//
// whisper_b200_full_batch runs one host thread per in-flight chunk; each thread executes the ordinary whisper_full()
// state machine (csrc/full.cpp) on its own whisper_state and device slot.  Their encoder / decoder passes meet here:
// a request blocks until every active worker has one pending, then the last arriver executes all of them as ONE
// batched pass (Forward::encode_batch / decode_batch) and wakes the others.  The CPU analogue in the reference is
// whisper_full_parallel (/root/reference/thirdparty/whisper.cpp/whisper.cpp:5817-5930): one state per worker, shared
// read-only weights — but there each worker computes alone.
#pragma once

#include "forward.h"

#include <condition_variable>
#include <mutex>
#include <vector>

namespace wb200 {

class Batcher {
public:
    explicit Batcher(Forward * fwd) : fwd_(fwd) {}

    // Worker registration (a thread that will issue passes through this batcher until it calls worker_end).
    void add_workers(int n);     // called once by the spawning thread: n workers are about to attach
    void worker_attach();        // called by each worker thread itself
    void worker_begin() { add_workers(1); worker_attach(); }
    void worker_end();
    // A registered worker entering / leaving a long host-only phase (log-mel of its next chunk): while paused the
    // other workers' batches do not wait for it.  No-ops on threads that are not registered workers.
    void host_phase_begin();
    void host_phase_end();

    // Blocking; safe to call from an unregistered thread when no workers exist (runs immediately, batch of one).
    bool encode(int slot, const float * mel_window, int n_ctx);
    bool decode(int slot, const DecodeInput & in, int n_audio_ctx, float * logits_out, whisper_token_data * sampled_out = nullptr);

    // statistics: device passes issued / requests served (requests / passes = achieved batching factor)
    int64_t n_passes = 0, n_requests = 0;

private:
    struct Request {
        int kind = 0;                 // 0 encode, 1 decode
        int slot = 0;
        int n_ctx = 0;
        const float * mel = nullptr;
        DecodeInput in;
        float * logits = nullptr;
        whisper_token_data * sampled = nullptr;
        bool done = false, ok = false;
    };
    bool submit(Request & r);
    void flush(std::unique_lock<std::mutex> & lk);
    void run(std::vector<Request *> & batch);

    Forward * fwd_;
    std::mutex mu_;
    std::condition_variable cv_;
    int active_ = 0;
    std::vector<Request *> pending_;
    int max_encode_batch_ = 16;
};

}  // namespace wb200
