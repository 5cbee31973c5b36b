#include "batcher.h"

#include <algorithm>
#include <cstdlib>
#include <deque>
#include <map>
#include <unordered_map>

#include <sched.h>

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
    max_decode_workers_ = std::max(3 * max_decode_rows_, 2 * fwd_->run_rows_max());
    if (const char * e = getenv("WHISPER_B200_PASS_SPLIT")) pass_split_ = std::max(1, atoi(e));
    if (const char * e = getenv("WHISPER_B200_HOST_BATCH_POLICY")) host_batch_policy_ = atoi(e) != 0;
    if (const char * e = getenv("WHISPER_B200_ENC_BATCH")) { max_encode_batch_ = std::max(1, atoi(e)); encode_batch_target_ = std::max(1, max_encode_batch_ / 2); }
    if (const char * e = getenv("WHISPER_B200_PASS_MIN_ROWS")) pass_min_rows_ = std::max(1, atoi(e));
    run_min_rows_ = std::min(fwd_->run_rows_max(), 320);      // (measured sweep: 128 / 192 / 320 -> 320)
    if (const char * e = getenv("WHISPER_B200_RUN_MIN_ROWS")) run_min_rows_ = std::max(1, atoi(e));
    if (!driver_started_) {
        driver_started_ = true;
        driver_ = std::thread([this] { driver_loop(); });
        enc_driver_ = std::thread([this] { encoder_loop(); });
    }
}

void Batcher::worker_attach() { tl_worker_of = this; }

bool Batcher::is_worker() const { return tl_worker_of == this; }

void Batcher::host_phase_begin() {
    if (!is_worker()) return;
    {
        std::unique_lock<std::mutex> lk(mu_);
        --active_;
        wake_driver();
        // at most one host-bound worker per core: more of them would only slow each other down and delay the first encoder pass
        cv_host_.wait(lk, [&] { return in_host_ < max_host_; });
        ++in_host_;
    }
    // a long stretch of pure number crunching: SCHED_BATCH tells the kernel so, and the threads the drivers wake (a few
    // microseconds of bookkeeping each, on the critical path of the next device pass) get a core ahead of it
    if (host_batch_policy_) { sched_param sp{}; sched_setscheduler(0, SCHED_BATCH, &sp); }
}

void Batcher::host_phase_end() {
    if (!is_worker()) return;
    if (host_batch_policy_) { sched_param sp{}; sched_setscheduler(0, SCHED_OTHER, &sp); }
    std::lock_guard<std::mutex> lk(mu_);
    ++active_;
    --in_host_;
    cv_host_.notify_one();
}

void Batcher::decode_phase_begin() {
    if (!is_worker()) return;
    std::unique_lock<std::mutex> lk(mu_);
    if (in_decode_ < max_decode_workers_) { ++in_decode_; return; }
    --active_;                                // waiting for a seat: nobody's batch depends on this worker
    wake_driver();
    cv_dec_.wait(lk, [&] { return in_decode_ < max_decode_workers_; });
    ++in_decode_;
    ++active_;
}

void Batcher::decode_phase_end() {
    if (!is_worker()) return;
    std::lock_guard<std::mutex> lk(mu_);
    --in_decode_;
    cv_dec_.notify_one();
}

void Batcher::worker_end() {
    std::lock_guard<std::mutex> lk(mu_);
    tl_worker_of = nullptr;
    --active_;
    wake_driver();                            // the workers that remain may all be waiting already
}

bool Batcher::encode(int slot, const float * mel_window, int n_ctx) {
    Request r;
    r.kind = 0; r.slot = slot; r.n_ctx = n_ctx; r.mel = mel_window;
    return submit(r);
}

bool Batcher::encode_pcm(int slot, const float * pcm, int n_samples, int mel_offset, int n_ctx, float * energy_out) {
    Request r;
    r.kind = 0; r.slot = slot; r.n_ctx = n_ctx; r.pcm = pcm; r.n_samples = n_samples; r.mel_offset = mel_offset; r.energy_out = energy_out;
    return submit(r);
}

bool Batcher::decode(int slot, const DecodeInput & in, int n_audio_ctx, float * logits_out, whisper_token_data * sampled_out, whisper_token_data * dist_out) {
    Request r;
    r.kind = 1; r.slot = slot; r.n_ctx = n_audio_ctx; r.in = in; r.logits = logits_out; r.sampled = sampled_out; r.dist = dist_out;
    return submit(r);
}

bool Batcher::run(int slot, const RunSeq & init, int n_audio_ctx, RunSeq & final_state, std::vector<whisper_token_data> & tokens) {
    Request r;
    r.kind = 2; r.slot = slot; r.n_ctx = n_audio_ctx; r.run_init = init; r.run_final = &final_state; r.run_tokens = &tokens;
    return submit(r);
}

bool Batcher::submit(Request & r) {
    {
        std::unique_lock<std::mutex> lk(mu_);
        if (!driver_started_ || !is_worker()) {
            // plain whisper_full() from a host thread that is not a chunk worker: run right here, batch of one
            lk.unlock();
            if (r.kind == 2) return run_alone(r);
            std::vector<Request *> one{&r};
            run(one);
            return r.ok;
        }
        r.t_submit = std::chrono::steady_clock::now();
        (r.kind == 0 ? pending_enc_ : r.kind == 1 ? pending_dec_ : pending_run_).push_back(&r);
        wake_driver();
    }
    std::unique_lock<std::mutex> lr(r.m);
    r.cv.wait(lr, [&] { return r.done; });
    return r.ok;
}

// A run driven by the calling thread (single whisper_full call): run_depth() steps stay queued on the device, the status word of
// the oldest tells when the sequence is over.  Steps queued behind the last real one idle through (the device checks the status).
bool Batcher::run_alone(Request & r) {
    if (!fwd_->run_start(r.slot, r.run_init)) return false;
    std::deque<int> tickets;
    const int depth = std::max(1, fwd_->run_depth());
    int32_t status = RUN_LIVE;
    for (;;) {
        while (status == RUN_LIVE && (int) tickets.size() < depth) {
            const int t = fwd_->run_step_enqueue(&r.slot, 1, r.n_ctx);
            if (t < 0) return false;
            tickets.push_back(t);
            ++n_run_steps; ++n_run_rows;
        }
        if (tickets.empty()) return false;
        const int t = tickets.front();
        tickets.pop_front();
        if (!fwd_->run_step_wait(t, &status)) return false;
        if (status != RUN_LIVE) return fwd_->run_fetch(r.slot, t, *r.run_final, *r.run_tokens);
    }
}

// The batching policy (mu_ held).
//   * decoder rows go as soon as a full pass worth of them waits, or every active worker is waiting (on a decode, an encode or a
//     run) so that nobody could add one: decoding workers are never held back by a worker that is busy on the host;
//   * encoder passes are cheaper per chunk when several chunks share them, so encode requests are held until
//     encode_batch_target_ of them wait, the oldest has waited encode_grace_us_, or nobody else could join (every active
//     worker waits, nobody is on the host).
bool Batcher::pick_encode(std::vector<Request *> & batch) {
    const int n_enc = (int) pending_enc_.size(), n_dec = (int) pending_dec_.size(), n_run = (int) pending_run_.size() + live_runs_;
    if (n_enc == 0) return false;
    // nobody else could join: every active worker waits for a pass or sits in one, and nobody is on the host
    const bool all_waiting = n_enc + n_dec + n_run + inflight_enc_ + inflight_dec_ >= active_;
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
    const int n_enc = (int) pending_enc_.size(), n_dec = (int) pending_dec_.size(), n_run = (int) pending_run_.size() + live_runs_;
    if (n_dec == 0) return false;
    int rows = 0;
    for (Request * q : pending_dec_) rows += q->in.n_tokens;
    // nobody could add a row: every active worker has one queued, sits in a pass or a run, or waits for the encoder
    const bool all_waiting = n_enc + n_dec + n_run + inflight_enc_ + inflight_dec_ >= active_;
    // a pass goes when it is full, or holds its share of the decoding workers (pass_split_ passes alternate: one on the
    // device while the workers of the other do their host bookkeeping), or nobody could add a row
    const int target = std::min(max_decode_rows_, std::max(pass_min_rows_, (in_decode_ - live_runs_ + pass_split_ - 1) / pass_split_));
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

// Encoder passes from their own host thread: staging of the inputs, the H2D copy and the encoder kernels (own stream)
// overlap with the decoder work the other driver keeps in flight.
void Batcher::encoder_loop() {
    std::unique_lock<std::mutex> lk(mu_);
    struct Fly { std::vector<Request *> batch; int set; };
    std::deque<Fly> fly;                           // queued passes, oldest first
    auto finish = [&](Fly & f, bool ok) {          // (called without mu_)
        for (Request * q : f.batch) q->ok = ok;
        n_requests += (int64_t) f.batch.size();
        { std::lock_guard<std::mutex> g(mu_); inflight_enc_ -= (int) f.batch.size(); }   // counters first, wake-ups second
        complete(f.batch);
    };
    for (;;) {
        std::vector<Request *> batch;
        const int max_fly = std::max(1, fwd_->encode_sets());
        for (;;) {
            if (stop_) break;
            if ((int) fly.size() < max_fly && fwd_->encoder_concurrent() && pick_encode(batch)) break;
            if (!fly.empty()) break;              // nothing (more) to queue right now: hand out the oldest pass
            if (!pending_enc_.empty()) cv_enc_.wait_for(lk, std::chrono::microseconds(200));   // grace period / mode change
            else cv_enc_.wait(lk);
        }
        if (stop_ && batch.empty() && fly.empty()) return;
        inflight_enc_ += (int) batch.size();
        lk.unlock();
        if (!batch.empty()) {
            const int n_ctx0 = batch.front()->n_ctx;
            bool same_ctx = true;
            for (Request * q : batch) same_ctx = same_ctx && q->n_ctx == n_ctx0;
            if (!same_ctx) {
                // mixed audio contexts: one pass per context, run to completion (rare: the realtime loop calls from one thread)
                while (!fly.empty()) { Fly f = std::move(fly.front()); fly.pop_front(); finish(f, fwd_->encode_collect(f.set)); }
                run(batch);
                { std::lock_guard<std::mutex> g(mu_); inflight_enc_ -= (int) batch.size(); }
                complete(batch);
            } else {
                int set = 0;
                for (const Fly & f : fly) if (f.set == set) set = 1 - set;
                std::vector<EncodeJob> jobs;
                for (Request * q : batch) { EncodeJob j; j.mel_window = q->mel; j.pcm = q->pcm; j.n_samples = q->n_samples; j.mel_offset = q->mel_offset; j.energy_out = q->energy_out; j.slot = q->slot; jobs.push_back(j); }
                Fly f{batch, set};
                if (fwd_->encode_enqueue(jobs.data(), (int) jobs.size(), n_ctx0, set)) { fly.push_back(std::move(f)); ++n_passes; }
                else finish(f, false);
            }
        } else if (!fly.empty()) {
            Fly f = std::move(fly.front());
            fly.pop_front();
            finish(f, fwd_->encode_collect(f.set));
        }
        lk.lock();
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
    for (Request * q : batch) {
        std::lock_guard<std::mutex> g(q->m);      // notify under the request's lock: it may be destroyed right after done is seen
        q->done = true;
        q->cv.notify_one();
    }
}

// The decoder driver.  Two kinds of work share it (and the decoder stream):
//   runs     — every loop turn queues one more step for all live runs until run_depth() steps wait on the device, then waits for
//              the oldest step, hands finished sequences back (their tokens come over a copy stream) and admits waiting runs;
//   requests — ordinary decoder passes (prefill, beam search, t > 0): up to two plain greedy passes queued (while pass A runs,
//              the rows of the other group of workers are staged behind it); anything else drains the queue and runs alone.
void Batcher::driver_loop() {
    std::unique_lock<std::mutex> lk(mu_);
    struct InFlight { std::vector<Request *> batch; int set; };
    std::vector<InFlight> fly;                    // oldest first
    const int max_fly = fwd_->decode_sets();
    struct Live { Request * q; uint64_t serial; };
    std::vector<Live> live;                       // the rows of the next run step
    struct Step { int ticket; int n_ctx; std::vector<uint64_t> serials; };
    std::deque<Step> steps;                       // queued run steps, oldest first
    uint64_t next_serial = 1;
    auto now_us = [] { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto collect_oldest = [&] {          // (called without mu_)
        InFlight f = std::move(fly.front());
        fly.erase(fly.begin());
        const int64_t t0 = now_us();
        const bool ok = fwd_->decode_collect(f.set);
        const int64_t t1 = now_us();
        for (Request * q : f.batch) q->ok = ok;
        n_requests += (int64_t) f.batch.size();
        { std::lock_guard<std::mutex> g(mu_); inflight_dec_ -= (int) f.batch.size(); }
        complete(f.batch);
        t_device_wait_us += t1 - t0; t_complete_us += now_us() - t1;
    };
    // hands a run back to its thread (called without mu_)
    auto retire = [&](size_t idx, bool ok, int ticket) {
        Request * q = live[idx].q;
        q->ok = ok && fwd_->run_fetch(q->slot, ticket, *q->run_final, *q->run_tokens);
        live[idx] = live.back();
        live.pop_back();
        { std::lock_guard<std::mutex> g(mu_); --live_runs_; }
        std::vector<Request *> one{q};
        ++n_requests;
        complete(one);
    };
    auto fail_all_runs = [&] {
        while (!live.empty()) retire(live.size() - 1, false, 0);
        steps.clear();
    };
    for (;;) {
        std::vector<Request *> batch, joiners;
        const bool runs_ok = fwd_->supports_runs();
        for (;;) {
            if (pick(batch)) break;
            const size_t room = runs_ok ? (size_t) std::max(0, fwd_->run_rows_max() - (int) live.size()) : pending_run_.size();
            const size_t take = std::min(room, pending_run_.size());
            if (take > 0) {
                joiners.assign(pending_run_.begin(), pending_run_.begin() + take);
                pending_run_.erase(pending_run_.begin(), pending_run_.begin() + take);
                live_runs_ += (int) take;
                break;
            }
            if (!fly.empty() || !steps.empty() || stop_) break;
            // Wide steps are cheaper per sequence (the decoder weights are read once per step, and the small kernels of a step cost
            // the same for 32 rows as for 512): while more runs are on their way — chunks in the encoder or in front of it — a
            // thin population waits for them instead of stepping alone.
            const bool more_coming = !pending_enc_.empty() || inflight_enc_ > 0 || in_host_ > 0;
            if (!live.empty() && !((int) live.size() < run_min_rows_ && more_coming)) break;
            const int64_t t0 = now_us();
            if (!live.empty()) { cv_drv_.wait_for(lk, std::chrono::microseconds(500)); t_idle_us += now_us() - t0; continue; }
            if (!pending_enc_.empty() && !fwd_->encoder_concurrent()) cv_drv_.wait_for(lk, std::chrono::microseconds(200));   // the grace period of a waiting encode runs out
            else cv_drv_.wait(lk);
            t_idle_us += now_us() - t0;
        }
        if (stop_ && batch.empty() && joiners.empty() && live.empty() && steps.empty()) { lk.unlock(); while (!fly.empty()) collect_oldest(); return; }
        inflight_dec_ += (int) batch.size();
        lk.unlock();

        // ---- ordinary decoder requests ----
        if (!batch.empty()) {
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
                for (Request * q : batch) { DecodeJob j; j.in = q->in; j.slot = q->slot; j.logits_out = q->logits; j.sampled_out = q->sampled; j.dist_out = q->dist; jobs.push_back(j); }
                const int64_t t0 = now_us();
                const bool queued = fwd_->decode_enqueue(jobs.data(), (int) jobs.size(), n_ctx0, set);
                t_stage_us += now_us() - t0;
                if (queued) {
                    fly.push_back(InFlight{batch, set});
                    ++n_passes;
                } else {
                    for (Request * q : batch) q->ok = false;
                    { std::lock_guard<std::mutex> g(mu_); inflight_dec_ -= (int) batch.size(); }
                    complete(batch);
                }
            } else {
                while (!fly.empty()) collect_oldest();
                const int64_t t0 = now_us();
                run(batch);
                const int64_t t1 = now_us();
                { std::lock_guard<std::mutex> g(mu_); inflight_dec_ -= (int) batch.size(); }
                complete(batch);
                t_run_us += t1 - t0; t_complete_us += now_us() - t1;
            }
        } else if (!fly.empty()) {
            collect_oldest();
        }

        // ---- runs ----
        for (Request * q : joiners) {
            if (runs_ok && fwd_->run_start(q->slot, q->run_init)) { live.push_back(Live{q, next_serial++}); continue; }
            q->ok = false;
            { std::lock_guard<std::mutex> g(mu_); --live_runs_; }
            std::vector<Request *> one{q};
            complete(one);
        }
        bool hold = false;
        {
            std::lock_guard<std::mutex> g(mu_);
            hold = (int) live.size() < run_min_rows_ && (!pending_enc_.empty() || inflight_enc_ > 0 || in_host_ > 0);
        }
        if (!live.empty() && !hold && (int) steps.size() < std::max(1, fwd_->run_depth())) {
            // one more step for every live run (sequences that finished in a step not yet waited for idle through it on the device)
            const int64_t t0 = now_us();
            const int n_ctx0 = live.front().q->n_ctx;
            std::vector<int> slots;
            Step s;
            for (const Live & l : live) if (l.q->n_ctx == n_ctx0) { slots.push_back(l.q->slot); s.serials.push_back(l.serial); }   // (runs of another audio context wait for the next step)
            s.n_ctx = n_ctx0;
            s.ticket = fwd_->run_step_enqueue(slots.data(), (int) slots.size(), n_ctx0);
            t_stage_us += now_us() - t0;
            if (s.ticket < 0) { fail_all_runs(); }
            else {
                ++n_run_steps; n_run_rows += (int64_t) slots.size(); ++n_passes;
                steps.push_back(std::move(s));
                // a mixed population: rotate so that the other audio contexts get their turn
                if (slots.size() < live.size()) std::rotate(live.begin(), live.begin() + 1, live.end());
            }
        } else if (!steps.empty()) {
            Step s = std::move(steps.front());
            steps.pop_front();
            std::vector<int32_t> status(s.serials.size(), RUN_LIVE);
            const int64_t t0 = now_us();
            const bool ok = fwd_->run_step_wait(s.ticket, status.data());
            const int64_t t1 = now_us();
            t_device_wait_us += t1 - t0;
            if (!ok) { fail_all_runs(); }
            else {
                std::unordered_map<uint64_t, int32_t> st;
                for (size_t i = 0; i < s.serials.size(); ++i) if (status[i] != RUN_LIVE) st[s.serials[i]] = status[i];
                if (!st.empty()) {
                    for (size_t i = live.size(); i-- > 0;) if (st.count(live[i].serial)) retire(i, true, s.ticket);
                }
            }
            t_complete_us += now_us() - t1;
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
            for (Request * q : v) { EncodeJob j; j.mel_window = q->mel; j.pcm = q->pcm; j.n_samples = q->n_samples; j.mel_offset = q->mel_offset; j.energy_out = q->energy_out; j.slot = q->slot; jobs.push_back(j); }
            const bool ok = fwd_->encode_batch(jobs.data(), (int) jobs.size(), g.first.second);
            for (Request * q : v) q->ok = ok;
        } else {
            std::vector<DecodeJob> jobs;
            for (Request * q : v) { DecodeJob j; j.in = q->in; j.slot = q->slot; j.logits_out = q->logits; j.sampled_out = q->sampled; j.dist_out = q->dist; jobs.push_back(j); }
            const bool ok = fwd_->decode_batch(jobs.data(), (int) jobs.size(), g.first.second);
            for (Request * q : v) q->ok = ok;
        }
        ++n_passes;
        n_requests += (int64_t) v.size();
    }
}

}  // namespace wb200
