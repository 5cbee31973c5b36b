// Stand-alone stress test of csrc/worker_pool.h: many generations with growing and shrinking worker counts, jobs of uneven length,
// a job that finishes before the other workers have woken up.  Prints "ok <calls>" or aborts.
#include "worker_pool.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

int main() {
    std::mt19937 rng(7);
    long long calls = 0;
    for (int round = 0; round < 3; ++round) {
        wb200::WorkerPool pool;
        for (int gen = 0; gen < 400; ++gen) {
            const int n = 1 + (int) (rng() % (gen % 50 == 0 ? 300 : 40));
            std::vector<std::atomic<int>> hit(n);
            for (auto & h : hit) h = 0;
            std::atomic<int> next{0}, total{0};
            const int items = n * 3 + (int) (rng() % 7);
            const std::function<void(int)> job = [&](int w) {
                if (w < 0 || w >= n) abort();
                hit[w].fetch_add(1);
                for (;;) {                                        // the chunk-dealing loop of whisper_b200_full_batch
                    const int c = next.fetch_add(1);
                    if (c >= items) break;
                    if ((c & 15) == 0) std::this_thread::sleep_for(std::chrono::microseconds(50));
                    total.fetch_add(1);
                }
            };
            pool.run(n, job);
            ++calls;
            if (total.load() != items) { fprintf(stderr, "generation %d: %d of %d items\n", gen, total.load(), items); return 1; }
            for (int w = 0; w < n; ++w) if (hit[w].load() != 1) { fprintf(stderr, "generation %d: worker %d ran %d times\n", gen, w, hit[w].load()); return 1; }
        }
    }
    printf("ok %lld\n", calls);
    return 0;
}
