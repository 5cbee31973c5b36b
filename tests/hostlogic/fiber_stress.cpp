// Stand-alone stress test of the fiber pool (godot-whisper_b200/csrc/fiber.h): N fibers block K times each; a waker thread makes them
// ready again in batches — sometimes before the fiber has switched away (the WOKEN_EARLY path), sometimes long after.  Every fiber
// must finish with exactly K resumes, whatever pool thread it lands on.
#include "../../godot-whisper_b200/csrc/fiber.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

using namespace wb200;

int main(int argc, char ** argv) {
    const int n_threads = argc > 1 ? atoi(argv[1]) : 8, n_fibers = argc > 2 ? atoi(argv[2]) : 512, n_rounds = argc > 3 ? atoi(argv[3]) : 200;
    std::mutex mu;
    std::vector<Fiber *> waiting;
    std::atomic<long> resumes{0}, finished{0};
    std::atomic<bool> stop{false};
    std::thread waker([&] {
        std::vector<Fiber *> batch;
        unsigned spin = 0;
        while (!stop.load()) {
            {
                std::lock_guard<std::mutex> lk(mu);
                batch.swap(waiting);
            }
            if (batch.empty()) { if ((++spin & 63) == 0) std::this_thread::yield(); continue; }
            FiberPool::wake_many(batch.data(), (int) batch.size());
            batch.clear();
        }
    });
    {
        FiberPool pool(n_threads, 64 * 1024);
        for (int i = 0; i < n_fibers; ++i) {
            pool.spawn([&, i] {
                volatile long local = 0;                 // lives on the fiber's stack across switches
                for (int k = 0; k < n_rounds; ++k) {
                    Fiber * f = FiberPool::current();
                    if (!f) { fprintf(stderr, "no current fiber\n"); abort(); }
                    FiberPool::prepare_block(f);
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        waiting.push_back(f);
                    }
                    if ((i + k) % 7 == 0) std::this_thread::yield();      // give the waker a chance to come first
                    FiberPool::suspend(f);
                    local = local + 1;
                    resumes.fetch_add(1, std::memory_order_relaxed);
                }
                if (local != n_rounds) { fprintf(stderr, "fiber %d: %ld resumes\n", i, (long) local); abort(); }
                finished.fetch_add(1);
            }, nullptr);
        }
        pool.wait_all();
    }
    stop.store(true);
    waker.join();
    const bool ok = finished.load() == n_fibers && resumes.load() == (long) n_fibers * n_rounds;
    printf("%s: %ld fibers finished, %ld resumes\n", ok ? "ok" : "FAILED", finished.load(), resumes.load());
    return ok ? 0 : 1;
}
