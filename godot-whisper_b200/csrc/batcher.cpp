#include "batcher.h"

#include <algorithm>
#include <map>

namespace wb200 {

namespace { thread_local Batcher * tl_worker_of = nullptr; }

void Batcher::add_workers(int n) {
    std::lock_guard<std::mutex> lk(mu_);
    active_ += n;
}

void Batcher::worker_attach() { tl_worker_of = this; }

void Batcher::host_phase_begin() {
    if (tl_worker_of != this) return;
    std::unique_lock<std::mutex> lk(mu_);
    --active_;
    if (!pending_.empty() && (int) pending_.size() >= active_) flush(lk);
}

void Batcher::host_phase_end() {
    if (tl_worker_of != this) return;
    std::lock_guard<std::mutex> lk(mu_);
    ++active_;
}

void Batcher::worker_end() {
    std::unique_lock<std::mutex> lk(mu_);
    tl_worker_of = nullptr;
    --active_;
    // the workers that remain may all be waiting already: this thread executes their batch before it leaves
    if (!pending_.empty() && (int) pending_.size() >= active_) flush(lk);
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
    std::unique_lock<std::mutex> lk(mu_);
    if (active_ == 0) {                       // plain whisper_full() from a single host thread
        lk.unlock();
        std::vector<Request *> one{&r};
        run(one);
        return r.ok;
    }
    pending_.push_back(&r);
    if ((int) pending_.size() >= active_) {
        flush(lk);                            // last arriver leads
    } else {
        cv_.wait(lk, [&] { return r.done; });
    }
    return r.ok;
}

void Batcher::flush(std::unique_lock<std::mutex> & lk) {
    std::vector<Request *> batch;
    batch.swap(pending_);
    lk.unlock();
    run(batch);
    lk.lock();
    for (Request * q : batch) q->done = true;
    cv_.notify_all();
}

void Batcher::run(std::vector<Request *> & batch) {
    // group by (kind, n_ctx): one device pass per group
    std::map<std::pair<int, int>, std::vector<Request *>> groups;
    for (Request * q : batch) groups[{q->kind, q->n_ctx}].push_back(q);
    for (auto & g : groups) {
        std::vector<Request *> & v = g.second;
        if (g.first.first == 0) {
            for (size_t i0 = 0; i0 < v.size(); i0 += max_encode_batch_) {
                const size_t i1 = std::min(v.size(), i0 + (size_t) max_encode_batch_);
                std::vector<EncodeJob> jobs;
                for (size_t i = i0; i < i1; ++i) { EncodeJob j; j.mel_window = v[i]->mel; j.slot = v[i]->slot; jobs.push_back(j); }
                const bool ok = fwd_->encode_batch(jobs.data(), (int) jobs.size(), g.first.second);
                for (size_t i = i0; i < i1; ++i) v[i]->ok = ok;
                ++n_passes;
            }
        } else {
            std::vector<DecodeJob> jobs;
            for (Request * q : v) { DecodeJob j; j.in = q->in; j.slot = q->slot; j.logits_out = q->logits; j.sampled_out = q->sampled; jobs.push_back(j); }
            const bool ok = fwd_->decode_batch(jobs.data(), (int) jobs.size(), g.first.second);
            for (Request * q : v) q->ok = ok;
            ++n_passes;
        }
        n_requests += (int64_t) v.size();
    }
}

}  // namespace wb200
