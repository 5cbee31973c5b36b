#include "batcher.h"

#include <algorithm>
#include <cstdlib>

#include <sched.h>
#include <map>

namespace wb200 {

namespace { thread_local Batcher * tl_worker_of = nullptr; }

Batcher::~Batcher() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_drv_.notify_all();
    cv_enc_.notify_all();
    if (driver_.joinable()) driver_.join();
    if (enc_driver_.joinable()) enc_driver_.join();
}

void Batcher::add_workers(int n) {
    std::lock_guard<std::mutex> lk(mu_);
    active_ += n;
    max_decode_rows_ = fwd_->decode_rows_per_pass();
    max_decode_workers_ = 3 * max_decode_rows_;       // one pass on the device, one queued behind it, one doing its host bookkeeping
    if (const char * e = getenv("WHISPER_B200_PASS_SPLIT")) pass_split_ = std::max(1, atoi(e));
    if (const char * e = getenv("WHISPER_B200_HOST_BATCH_POLICY")) host_batch_policy_ = atoi(e) != 0;
    if (const char * e = getenv("WHISPER_B200_ENC_BATCH")) { max_encode_batch_ = std::max(1, atoi(e)); encode_batch_target_ = std::max(1, max_encode_batch_ / 2); }
    if (const char * e = getenv("WHISPER_B200_PASS_MIN_ROWS")) pass_min_rows_ = std::max(1, atoi(e));
    if (!driver_started_) {
        driver_started_ = true;
        driver_ = std::thread([this] { driver_loop(); });
        enc_driver_ = std::thread([this] { encoder_loop(); });
    }
}

void Batcher::worker_attach() { if (Fiber * f = FiberPool::current()) f->owner = this; else tl_worker_of = this; }

bool Batcher::is_worker() const {
    if (Fiber * f = FiberPool::current()) return f->owner == this;
    return tl_worker_of == this;
}

void Batcher::host_phase_begin() {
    if (!is_worker()) return;
    Fiber * f = FiberPool::current();
    {
        std::unique_lock<std::mutex> lk(mu_);
        --active_;
        wake_driver();
        // at most one host-bound worker per core: more of them would only slow each other down and delay the first encoder pass
        if (!f) {
            cv_host_.wait(lk, [&] { return in_host_ < max_host_; });
            ++in_host_;
        } else if (in_host_ < max_host_) {
            ++in_host_;
            // seat taken, no wait — but a long phase starts: step to the back of the ready queue first, flagged heavy, so that the
            // fibers this pool thread took together with this one do not sit behind a spectrogram
            f->heavy = true;
            FiberPool::prepare_block(f);
            FiberPool::wake(f);
        } else {
            FiberPool::prepare_block(f);
            host_waiters_.push_back(f);           // host_phase_end of another worker takes the seat on this fiber's behalf and wakes it
        }
    }
    if (f) FiberPool::suspend(f);
    // log-mel is pure number crunching: SCHED_BATCH tells the kernel so, and the workers it wakes with sampled tokens (a few
    // microseconds of bookkeeping each, on the critical path of the next decoder pass) get a core ahead of it
    // (a fiber stays on its pool thread for the whole phase: there is no block() inside log-mel)
    if (host_batch_policy_) { sched_param sp{}; sched_setscheduler(0, SCHED_BATCH, &sp); }
}

void Batcher::host_phase_end() {
    if (!is_worker()) return;
    if (host_batch_policy_) { sched_param sp{}; sched_setscheduler(0, SCHED_OTHER, &sp); }
    Fiber * next = nullptr;
    {
        std::lock_guard<std::mutex> lk(mu_);
        ++active_;
        if (!host_waiters_.empty() && in_host_ <= max_host_) { next = host_waiters_.front(); host_waiters_.pop_front(); next->heavy = true; }   // the seat changes hands
        else --in_host_;
        cv_host_.notify_one();
    }
    if (next) FiberPool::wake(next);
}

void Batcher::decode_phase_begin() {
    if (!is_worker()) return;
    Fiber * f = FiberPool::current();
    {
        std::unique_lock<std::mutex> lk(mu_);
        if (in_decode_ < max_decode_workers_) { ++in_decode_; return; }
        --active_;                                // waiting for a seat: nobody's batch depends on this worker
        wake_driver();
        if (!f) {
            cv_dec_.wait(lk, [&] { return in_decode_ < max_decode_workers_; });
            ++in_decode_;
            ++active_;
            return;
        }
        FiberPool::prepare_block(f);
        dec_waiters_.push_back(f);                // decode_phase_end hands its seat over (in_decode_ and active_ adjusted there)
    }
    FiberPool::suspend(f);
}

void Batcher::decode_phase_end() {
    if (!is_worker()) return;
    Fiber * next = nullptr;
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (!dec_waiters_.empty()) { next = dec_waiters_.front(); dec_waiters_.pop_front(); ++active_; }   // the seat changes hands
        else --in_decode_;
        cv_dec_.notify_one();
    }
    if (next) FiberPool::wake(next);
}

void Batcher::worker_end() {
    std::lock_guard<std::mutex> lk(mu_);
    if (Fiber * f = FiberPool::current()) f->owner = nullptr; else tl_worker_of = nullptr;
    --active_;
    wake_driver();                            // the workers that remain may all be waiting already
}

bool Batcher::encode(int slot, const float * mel_window, int n_ctx) {
    Request r;
    r.kind = 0; r.slot = slot; r.n_ctx = n_ctx; r.mel = mel_window;
    return submit(r);
}

bool Batcher::decode(int slot, const DecodeInput & in, int n_audio_ctx, float * logits_out, whisper_token_data * sampled_out) {
    Request r;
    r.kind = 1; r.slot = slot; r.n_ctx = n_audio_ctx; r.in = in; r.logits = logits_out; r.sampled = sampled_out;
    return submit(r);
}

bool Batcher::submit(Request & r) {
    {
        std::unique_lock<std::mutex> lk(mu_);
        if (!driver_started_ || !is_worker()) {
            // plain whisper_full() from a host thread that is not a chunk worker: run right here, batch of one
            lk.unlock();
            std::vector<Request *> one{&r};
            run(one);
            return r.ok;
        }
        r.t_submit = std::chrono::steady_clock::now();
        r.fiber = FiberPool::current();
        if (r.fiber) FiberPool::prepare_block(r.fiber);        // (before the request becomes visible: completion may come at once)
        (r.kind == 0 ? pending_enc_ : pending_dec_).push_back(&r);
        wake_driver();
    }
    if (r.fiber) {
        FiberPool::suspend(r.fiber);                           // resumed by complete(), possibly on another pool thread
        return r.ok;
    }
    std::unique_lock<std::mutex> lr(r.m);
    r.cv.wait(lr, [&] { return r.done; });
    return r.ok;
}

// The batching policy (mu_ held).
//   * decoder rows go as soon as a full pass worth of them waits, or every active worker is waiting (on a decode or on an
//     encode) so that nobody could add one: decoding workers are never held back by a worker that is busy on the host;
//   * encoder passes are cheaper per chunk when several chunks share them, so encode requests are held until
//     encode_batch_target_ of them wait, the oldest has waited encode_grace_us_, or nobody else could join (every active
//     worker waits, nobody decodes, nobody is on the host).
bool Batcher::pick_encode(std::vector<Request *> & batch) {
    const int n_enc = (int) pending_enc_.size(), n_dec = (int) pending_dec_.size();
    if (n_enc == 0) return false;
    // nobody else could join: every active worker waits for a pass or sits in one, and nobody is on the host
    const bool all_waiting = n_enc + n_dec + inflight_enc_ + inflight_dec_ >= active_;
    const auto waited = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - pending_enc_.front()->t_submit).count();
    if (n_enc >= encode_batch_target_ || waited >= encode_grace_us_ || (all_waiting && in_host_ == 0)) {
        const size_t take = std::min((size_t) max_encode_batch_, pending_enc_.size());
        batch.assign(pending_enc_.begin(), pending_enc_.begin() + take);
        pending_enc_.erase(pending_enc_.begin(), pending_enc_.begin() + take);
        return true;
    }
    return false;
}

bool Batcher::pick_decode(std::vector<Request *> & batch) {
    const int n_enc = (int) pending_enc_.size(), n_dec = (int) pending_dec_.size();
    if (n_dec == 0) return false;
    int rows = 0;
    for (Request * q : pending_dec_) rows += q->in.n_tokens;
    // nobody could add a row: every active worker has one queued, sits in a pass, or waits for the encoder
    const bool all_waiting = n_enc + n_dec + inflight_enc_ + inflight_dec_ >= active_;
    // a pass goes when it is full, or holds its share of the decoding workers (pass_split_ passes alternate: one on the
    // device while the workers of the other do their host bookkeeping), or nobody could add a row
    const int target = std::min(max_decode_rows_, std::max(pass_min_rows_, (in_decode_ + pass_split_ - 1) / pass_split_));
    if (rows >= target || all_waiting) {
        // one pass: requests in arrival order while they fit (a request is never split)
        size_t take = 0;
        int r = 0;
        while (take < pending_dec_.size() && (take == 0 || r + pending_dec_[take]->in.n_tokens <= max_decode_rows_)) r += pending_dec_[take++]->in.n_tokens;
        batch.assign(pending_dec_.begin(), pending_dec_.begin() + take);
        pending_dec_.erase(pending_dec_.begin(), pending_dec_.begin() + take);
        return true;
    }
    return false;
}

// The policy of the decoder driver (mu_ held).  Encoder requests are its business only while the forward pass cannot run
// them next to decoder passes (no encoder stream / profiling): then they go first, as one pass through run().
bool Batcher::pick(std::vector<Request *> & batch) {
    if (!fwd_->encoder_concurrent() && pick_encode(batch)) return true;
    return pick_decode(batch);
}

// Encoder passes from their own host thread: staging of the mel windows, the H2D copy and the encoder kernels (own stream)
// overlap with the decoder passes the other driver keeps in flight.
void Batcher::encoder_loop() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
        std::vector<Request *> batch;
        while (!stop_ && !(fwd_->encoder_concurrent() && pick_encode(batch))) {
            if (!pending_enc_.empty()) cv_enc_.wait_for(lk, std::chrono::microseconds(200));   // grace period / mode change
            else cv_enc_.wait(lk);
        }
        if (stop_) return;
        inflight_enc_ += (int) batch.size();
        lk.unlock();
        run(batch);
        complete(batch);
        lk.lock();
        inflight_enc_ -= (int) batch.size();
        wake_driver();
    }
}

// True if the batch is a plain greedy decoder step for every request (one new token, sampled on the device): the shape the
// forward pass can keep two of in flight.
static bool pipelinable(const std::vector<Batcher::RequestView> & v) {
    for (const auto & q : v) if (q.kind != 1 || q.n_tokens != 1 || !q.sampled) return false;
    return !v.empty();
}

void Batcher::complete(std::vector<Request *> & batch) {
    std::vector<Fiber *> fibers;
    for (Request * q : batch) {
        if (q->fiber) { fibers.push_back(q->fiber); continue; }      // (q lives on the fiber's stack: not touched after the wake below)
        std::lock_guard<std::mutex> g(q->m);      // notify under the request's lock: it may be destroyed right after done is seen
        q->done = true;
        q->cv.notify_one();
    }
    if (!fibers.empty()) FiberPool::wake_many(fibers.data(), (int) fibers.size());    // one queue operation for the whole pass
}

// The driver keeps up to two decoder passes queued on the device: while pass A runs, the rows of the other group of workers
// are staged and queued behind it (pass B); then A's results are handed out and its workers woken while B runs.  Anything
// that is not a plain greedy step (encoder passes, prompts, beam search, host-side logits) drains the queue and runs alone.
void Batcher::driver_loop() {
    std::unique_lock<std::mutex> lk(mu_);
    struct InFlight { std::vector<Request *> batch; int set; };
    std::vector<InFlight> fly;                    // oldest first
    const int max_fly = fwd_->decode_sets();
    auto now_us = [] { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto collect_oldest = [&] {          // (called without mu_)
        InFlight f = std::move(fly.front());
        fly.erase(fly.begin());
        const int64_t t0 = now_us();
        const bool ok = fwd_->decode_collect(f.set);
        const int64_t t1 = now_us();
        for (Request * q : f.batch) q->ok = ok;
        n_requests += (int64_t) f.batch.size();
        complete(f.batch);
        t_device_wait_us += t1 - t0; t_complete_us += now_us() - t1;
        std::lock_guard<std::mutex> g(mu_);
        inflight_dec_ -= (int) f.batch.size();
    };
    for (;;) {
        std::vector<Request *> batch;
        while (!stop_ && !pick(batch)) {
            if (!fly.empty()) break;              // nothing new to queue: go hand out the oldest pass
            const int64_t t0 = now_us();
            if (!pending_enc_.empty() && !fwd_->encoder_concurrent()) cv_drv_.wait_for(lk, std::chrono::microseconds(200));   // the grace period of a waiting encode runs out
            else cv_drv_.wait(lk);
            t_idle_us += now_us() - t0;
        }
        if (stop_) { lk.unlock(); while (!fly.empty()) collect_oldest(); return; }
        inflight_dec_ += (int) batch.size();
        lk.unlock();
        if (batch.empty()) {
            collect_oldest();
        } else {
            std::vector<RequestView> view;
            for (Request * q : batch) view.push_back(RequestView{q->kind, q->in.n_tokens, q->in.sample != nullptr && q->sampled != nullptr});
            const int n_ctx0 = batch.front()->n_ctx;
            bool same_ctx = true;
            for (Request * q : batch) same_ctx = same_ctx && q->n_ctx == n_ctx0;
            if (max_fly > 1 && same_ctx && pipelinable(view)) {
                if ((int) fly.size() >= max_fly) collect_oldest();
                int set = 0;
                for (const InFlight & f : fly) if (f.set == set) set = 1 - set;
                std::vector<DecodeJob> jobs;
                for (Request * q : batch) { DecodeJob j; j.in = q->in; j.slot = q->slot; j.logits_out = q->logits; j.sampled_out = q->sampled; jobs.push_back(j); }
                const int64_t t0 = now_us();
                const bool queued = fwd_->decode_enqueue(jobs.data(), (int) jobs.size(), n_ctx0, set);
                t_stage_us += now_us() - t0;
                if (queued) {
                    fly.push_back(InFlight{batch, set});
                    ++n_passes;
                } else {
                    for (Request * q : batch) q->ok = false;
                    complete(batch);
                    std::lock_guard<std::mutex> g(mu_);
                    inflight_dec_ -= (int) batch.size();
                }
            } else {
                while (!fly.empty()) collect_oldest();
                const int64_t t0 = now_us();
                run(batch);
                const int64_t t1 = now_us();
                complete(batch);
                t_run_us += t1 - t0; t_complete_us += now_us() - t1;
                std::lock_guard<std::mutex> g(mu_);
                inflight_dec_ -= (int) batch.size();
            }
        }
        lk.lock();
    }
}

void Batcher::run(std::vector<Request *> & batch) {
    // group by (kind, n_ctx): one device pass per group
    std::map<std::pair<int, int>, std::vector<Request *>> groups;
    for (Request * q : batch) groups[{q->kind, q->n_ctx}].push_back(q);
    for (auto & g : groups) {
        std::vector<Request *> & v = g.second;
        if (g.first.first == 0) {
            std::vector<EncodeJob> jobs;
            for (Request * q : v) { EncodeJob j; j.mel_window = q->mel; j.slot = q->slot; jobs.push_back(j); }
            const bool ok = fwd_->encode_batch(jobs.data(), (int) jobs.size(), g.first.second);
            for (Request * q : v) q->ok = ok;
        } else {
            std::vector<DecodeJob> jobs;
            for (Request * q : v) { DecodeJob j; j.in = q->in; j.slot = q->slot; j.logits_out = q->logits; j.sampled_out = q->sampled; jobs.push_back(j); }
            const bool ok = fwd_->decode_batch(jobs.data(), (int) jobs.size(), g.first.second);
            for (Request * q : v) q->ok = ok;
        }
        ++n_passes;
        n_requests += (int64_t) v.size();
    }
}

}  // namespace wb200
