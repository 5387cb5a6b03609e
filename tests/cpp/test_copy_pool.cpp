// test_copy_pool.cpp -- CPU test of include/frame-copy-pool.hpp: many back-to-back and spaced copies of odd sizes are
// byte-exact, small copies take the plain path, the pool shuts down cleanly; prints the measured copy rates.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "frame-copy-pool.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); return 1; } } while (0)

int main()
{
    const size_t big = 1199988;                 // one frame of the 100k-atom box
    std::vector<char> src(big + 4096), dst(big + 4096);
    for (size_t i = 0; i < src.size(); ++i) src[i] = (char)(i * 2654435761u >> 13);
    for (int helpers = 0; helpers <= 3; ++helpers) {
        FrameCopyPool pool(helpers);
        CHECK(pool.helpers() == helpers);
        const size_t sizes[] = {0, 1, 4095, 65536, 262143, 262144 * 4 + 1, big, big + 4096};
        for (int it = 0; it < 300; ++it) {
            const size_t n = sizes[it % 8];
            std::fill(dst.begin(), dst.end(), (char)it);
            pool.copy(dst.data(), src.data() + (it % 3), n ? n - (it % 3 < (int)n ? it % 3 : 0) : 0);
            const size_t m = n ? n - (it % 3 < (int)n ? it % 3 : 0) : 0;
            for (size_t i = 0; i < m; ++i) CHECK(dst[i] == src[i + it % 3]);
            for (size_t i = m; i < m + 64 && i < dst.size(); ++i) CHECK(dst[i] == (char)it);     // nothing past the end
            if (it % 50 == 49) std::this_thread::sleep_for(std::chrono::milliseconds(2));      // let the helpers fall asleep
        }
        const auto t0 = std::chrono::steady_clock::now();
        const int reps = 400;
        for (int it = 0; it < reps; ++it) pool.copy(dst.data(), src.data(), big);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("helpers %d: %.2f GB/s (%.0f frames/s of the 100k-atom box)\n", helpers, big * (double)reps / s / 1e9, reps / s);
    }
    std::printf("test_copy_pool OK\n");
    return 0;
}
