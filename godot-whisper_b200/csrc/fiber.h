// Cooperative fibers for the chunk workers of whisper_b200_full_batch.
//
// Each in-flight chunk runs the ordinary, sequential whisper_full() state machine (csrc/full.cpp) — on its own small stack instead
// of its own OS thread.  Where the state machine would block (an encoder / decoder request handed to the Batcher, a seat for a host
// phase) the fiber switches back to the pool thread that was running it, which picks the next ready fiber; the pass driver makes the
// fibers of a finished pass ready again with one queue operation.  A token step of 256 live sequences then costs a few pool threads
// a tight loop over 256 continuations instead of 256 futex wake-ups and context switches.
//
// Rules for code that runs on a fiber: no mutex held and no thread_local touched across a block(); a fiber may resume on a different
// pool thread.  CUDA is never called from a fiber (device work belongs to the Batcher's driver threads).
#pragma once

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <ucontext.h>

namespace wb200 {

class FiberPool;

struct Fiber {
    enum State : int { RUNNING = 0, BLOCKING = 1, SUSPENDED = 2, WOKEN_EARLY = 3 };
    ucontext_t ctx;
    ucontext_t * back = nullptr;          // context of the pool thread that resumed this fiber most recently
    std::unique_ptr<char[]> stack;
    std::function<void()> fn;
    std::atomic<int> state{RUNNING};
    bool finished = false;
    bool heavy = false;                   // the next run of this fiber is long (a log-mel phase): pool threads do not queue others behind it
    FiberPool * pool = nullptr;
    void * owner = nullptr;               // the Batcher this fiber works for (what thread_local tl_worker_of is for worker threads)
};

class FiberPool {
public:
    explicit FiberPool(int n_threads, size_t stack_bytes = 1 << 20);
    ~FiberPool();                         // waits for every spawned fiber to finish

    void spawn(std::function<void()> fn, void * owner);
    void wait_all();

    static Fiber * current();             // the fiber the calling code runs on, or nullptr on a plain thread
    // Called on a fiber: announces that it is about to wait.  From here on wake() may be called for it (from any thread, even before
    // suspend() has switched away).  No blocking calls between prepare_block() and suspend().
    static void prepare_block(Fiber * f) { f->state.store(Fiber::BLOCKING, std::memory_order_release); }
    static void suspend(Fiber * f);       // switch to the pool thread; returns after wake(f)
    static void wake(Fiber * f);
    static void wake_many(Fiber * const * fibers, int n);

private:
    void thread_main();
    void make_ready(Fiber * const * fibers, int n);
    static void trampoline(unsigned lo, unsigned hi);

    size_t stack_bytes_;
    std::mutex mu_;
    std::condition_variable cv_, cv_done_;
    std::deque<Fiber *> ready_;
    std::vector<std::unique_ptr<Fiber>> fibers_;
    std::vector<std::thread> threads_;
    int live_ = 0, sleeping_ = 0;
    bool stop_ = false;
};

}  // namespace wb200
