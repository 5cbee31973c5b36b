#include "fiber.h"

#include <cstdint>

namespace wb200 {

namespace {
thread_local Fiber * tl_fiber = nullptr;     // fiber running on this pool thread right now
}

// (noinline: the thread_local must be looked up afresh by every caller — a fiber may have moved to another thread since its last look)
__attribute__((noinline)) Fiber * FiberPool::current() { return tl_fiber; }

FiberPool::FiberPool(int n_threads, size_t stack_bytes) : stack_bytes_(stack_bytes) {
    for (int i = 0; i < n_threads; ++i) threads_.emplace_back([this] { thread_main(); });
}

FiberPool::~FiberPool() {
    wait_all();
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto & t : threads_) t.join();
}

void FiberPool::trampoline(unsigned lo, unsigned hi) {
    Fiber * f = (Fiber *) (((uintptr_t) hi << 32) | (uintptr_t) lo);
    f->fn();
    f->finished = true;
    // back to the pool thread for good (f->back was set by whoever resumed us last)
    setcontext(f->back);
}

void FiberPool::spawn(std::function<void()> fn, void * owner) {
    std::unique_ptr<Fiber> f(new Fiber);
    f->fn = std::move(fn);
    f->pool = this;
    f->owner = owner;
    f->heavy = true;                       // (a worker starts with the log-mel of its first chunk)
    f->stack.reset(new char[stack_bytes_]);
    getcontext(&f->ctx);
    f->ctx.uc_stack.ss_sp = f->stack.get();
    f->ctx.uc_stack.ss_size = stack_bytes_;
    f->ctx.uc_link = nullptr;
    const uintptr_t p = (uintptr_t) f.get();
    makecontext(&f->ctx, (void (*)()) trampoline, 2, (unsigned) (p & 0xffffffffu), (unsigned) (p >> 32));
    Fiber * raw = f.get();
    {
        std::lock_guard<std::mutex> lk(mu_);
        fibers_.push_back(std::move(f));
        ++live_;
        ready_.push_back(raw);
    }
    cv_.notify_one();
}

void FiberPool::wait_all() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return live_ == 0; });
    fibers_.clear();
}

void FiberPool::make_ready(Fiber * const * fibers, int n) {
    if (n <= 0) return;
    int to_wake = 0;
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (int i = 0; i < n; ++i) ready_.push_back(fibers[i]);
        to_wake = std::min(n, sleeping_);
    }
    if (to_wake >= (int) threads_.size() / 2) cv_.notify_all();
    else for (int i = 0; i < to_wake; ++i) cv_.notify_one();
}

void FiberPool::wake_many(Fiber * const * fibers, int n) {
    // fibers that have switched away become ready now; one that is still on its way out is flagged and re-queued by its pool thread
    std::vector<Fiber *> ready;
    ready.reserve((size_t) n);
    FiberPool * pool = nullptr;
    for (int i = 0; i < n; ++i) {
        Fiber * f = fibers[i];
        pool = f->pool;
        for (;;) {
            int s = f->state.load(std::memory_order_acquire);
            if (s == Fiber::SUSPENDED) {
                if (f->state.compare_exchange_weak(s, Fiber::RUNNING, std::memory_order_acq_rel)) { ready.push_back(f); break; }
            } else if (s == Fiber::BLOCKING) {
                if (f->state.compare_exchange_weak(s, Fiber::WOKEN_EARLY, std::memory_order_acq_rel)) break;
            } else {
                break;          // RUNNING / WOKEN_EARLY: a wake without a matching wait — nothing to do
            }
        }
    }
    if (pool) pool->make_ready(ready.data(), (int) ready.size());
}

void FiberPool::wake(Fiber * f) { wake_many(&f, 1); }

__attribute__((noinline)) void FiberPool::suspend(Fiber * f) {
    // (f->back belongs to the pool thread we are on; after the switch back nothing of this thread's state is touched)
    swapcontext(&f->ctx, f->back);
}

void FiberPool::thread_main() {
    ucontext_t here;
    std::vector<Fiber *> mine;                     // fibers taken from the ready queue in one go (one lock operation per batch)
    for (;;) {
        mine.clear();
        {
            std::unique_lock<std::mutex> lk(mu_);
            ++sleeping_;
            cv_.wait(lk, [&] { return stop_ || !ready_.empty(); });
            --sleeping_;
            if (ready_.empty()) return;       // stop_
            // a fair share of what is ready, at most 16; a heavy fiber (log-mel ahead) closes the batch so that nothing waits behind it
            const size_t share = std::min<size_t>(16, std::max<size_t>(1, ready_.size() / threads_.size()));
            while (mine.size() < share && !ready_.empty()) {
                Fiber * f = ready_.front();
                ready_.pop_front();
                mine.push_back(f);
                if (f->heavy) break;
            }
        }
        for (Fiber * f : mine) {
            f->heavy = false;
            f->back = &here;
            tl_fiber = f;
            swapcontext(&here, &f->ctx);
            tl_fiber = nullptr;
            if (f->finished) {
                std::lock_guard<std::mutex> lk(mu_);
                if (--live_ == 0) cv_done_.notify_all();
                continue;
            }
            // the fiber announced a wait (prepare_block) and switched away: from now on a wake may queue it; if the wake came first,
            // queue it here
            int expect = Fiber::BLOCKING;
            if (!f->state.compare_exchange_strong(expect, Fiber::SUSPENDED, std::memory_order_acq_rel)) {
                f->state.store(Fiber::RUNNING, std::memory_order_release);
                make_ready(&f, 1);
            }
        }
    }
}

}  // namespace wb200
