// test_findpairs_mock.cpp -- CudaFindPairs' result hand-over (find by value, findInto, findRaw) on lists large
// enough for its parallel paths (huge-page hint + first touch from several threads, shared copy into a kept vector),
// over the stand-in C ABI of tests/mock.  Built by tests/mock/Makefile; run by tests/test_host_mock_cpu.py.
#include <chrono>
#include <cstdio>
#include <memory>
#include <random>
#include <vector>

#include "cuda-find-pairs.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); return 1; } } while (0)

namespace {
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

std::vector<std::pair<int, int>> brute(const std::vector<Point>& pts)
{
    std::vector<std::pair<int, int>> want;
    for (unsigned i = 0; i < pts.size(); ++i)
        for (unsigned j = i + 1; j < pts.size(); ++j) {
            const double dx = pts[i].x() - pts[j].x(), dy = pts[i].y() - pts[j].y(), dz = pts[i].z() - pts[j].z();
            if (10.0 * (dx * dx + dy * dy + dz * dz) < 42.25) want.push_back(std::make_pair((int)i, (int)j));
        }
    return want;
}
}  // namespace

int main()
{
    try {
        std::mt19937 rng(5);
        std::shared_ptr<FindPairs> mp = std::make_shared<CudaFindPairs>(0);
        CudaFindPairs* cf = static_cast<CudaFindPairs*>(mp.get());
        std::vector<std::pair<int, int>> kept;                     // the vector a caller keeps from frame to frame
        // sizes: large (parallel paths), larger (kept vector grows), small (single-thread path, kept vector shrinks), empty
        const int sizes[] = {2600, 3400, 300, 0, 2600};
        const double edge[] = {1.6, 1.9, 3.0, 1.0, 1.6};
        for (int c = 0; c < 5; ++c) {
            std::uniform_real_distribution<double> u(0.0, edge[c]);
            std::vector<Point> pts;
            for (int i = 0; i < sizes[c]; ++i) pts.push_back(Point(u(rng), u(rng), u(rng)));
            const std::vector<std::pair<int, int>> want = brute(pts);
            if (c < 2) CHECK(want.size() * sizeof(want[0]) > ((size_t)8 << 20));        // really above the 4 MB threshold
            const double t0 = now();
            std::vector<std::pair<int, int>> got = mp->find(pts);                    // by value, the reference's interface
            const double t1 = now();
            CHECK(got == want);
            cf->findInto(pts, kept);
            const double t2 = now();
            CHECK(kept == want);
            uint64_t m = 0;
            const std::pair<int, int>* raw = cf->findRaw(pts, &m);
            CHECK(m == want.size());
            for (uint64_t k = 0; k < m; k += 997) CHECK(raw[k] == want[(size_t)k]);
            std::printf("%d points, %zu pairs: find %.1f ms, findInto %.1f ms (stand-in search included)\n", sizes[c], want.size(),
                        (t1 - t0) * 1e3, (t2 - t1) * 1e3);
        }
        std::printf("test_findpairs_mock OK\n");
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
}
