// The chunk workers of whisper_b200_full_batch as a pool that lives as long as the context: a call with 512 chunks used to create and
// join 512 threads (about 8 ms of a 196 ms call); now it wakes the threads of the previous call.  A worker is parked on a condition
// variable between calls; run() hands every worker w < n the same job and returns when all n have finished it.
#pragma once

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace wb200 {

class WorkerPool {
public:
    ~WorkerPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_job_.notify_all();
        for (auto & t : threads_) t.join();
    }
    // Runs job(w) for w = 0 .. n - 1 on n threads of the pool (grown on demand) and waits for all of them.
    void run(int n, const std::function<void(int)> & job) {
        if (n <= 0) return;
        std::unique_lock<std::mutex> lk(m_);
        job_ = &job; n_active_ = n; n_done_ = 0; ++generation_;
        while ((int) threads_.size() < n) {
            const int w = (int) threads_.size();
            threads_.emplace_back([this, w] { loop(w); });
        }
        cv_job_.notify_all();
        cv_done_.wait(lk, [&] { return n_done_ == n_active_; });
        job_ = nullptr;
    }

private:
    void loop(int w) {
        long long seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_job_.wait(lk, [&] { return stop_ || (generation_ != seen && w < n_active_); });
            if (stop_) return;
            seen = generation_;
            const std::function<void(int)> * job = job_;
            lk.unlock();
            (*job)(w);
            lk.lock();
            if (++n_done_ == n_active_) cv_done_.notify_all();
        }
    }
    std::mutex m_;
    std::condition_variable cv_job_, cv_done_;
    std::vector<std::thread> threads_;
    const std::function<void(int)> * job_ = nullptr;
    long long generation_ = 0;
    int n_active_ = 0, n_done_ = 0;
    bool stop_ = false;
};

}  // namespace wb200
