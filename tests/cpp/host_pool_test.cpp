// Exercises gact::HostPool (csrc/host_pool.h): every part runs exactly once, calls of different widths follow each
// other, and a second caller thread falls back to running its parts itself.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "host_pool.h"

int main()
{
    gact::HostPool pool;
    int bad = 0;
    for (int round = 0; round < 2000; round++) {
        const int parts = 1 + (round * 7) % 8;
        std::vector<std::atomic<int>> hit(8);
        for (auto &h : hit) h = 0;
        pool.run(parts, [&](int p) { hit[(size_t)p]++; });
        for (int p = 0; p < 8; p++) if (hit[(size_t)p] != (p < parts ? 1 : 0)) bad++;
    }
    // threaded copy: the shape par_memcpy uses
    std::vector<char> src(37 << 20), dst(37 << 20);
    for (size_t i = 0; i < src.size(); i++) src[i] = (char)(i * 2654435761u >> 13);
    const size_t bytes = src.size(), per = ((bytes / 5) + 63) & ~(size_t)63;
    pool.run(5, [&](int p) {
        const size_t a = per * p, b = (p == 4) ? bytes : std::min(bytes, per * (p + 1));
        if (a < b) memcpy(dst.data() + a, src.data() + a, b - a);
    });
    if (memcmp(src.data(), dst.data(), bytes) != 0) bad++;
    // two callers at once: both complete, every part once
    std::atomic<int> total{0};
    auto caller = [&] { for (int r = 0; r < 500; r++) pool.run(4, [&](int) { total++; }); };
    std::thread t1(caller), t2(caller);
    t1.join(); t2.join();
    if (total != 2 * 500 * 4) bad++;
    printf(bad ? "FAIL %d\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
