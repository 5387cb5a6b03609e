// cuda-find-pairs.cpp -- CudaFindPairs over the C ABI (see include/cuda-find-pairs.hpp).
#include "cuda-find-pairs.hpp"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <system_error>
#include <thread>
#ifdef __linux__
#include <sys/mman.h>
#endif

namespace {
[[noreturn]] void raise(const char* what, int rc, const wn_ctx* ctx)
{
    throw std::runtime_error(std::string(what) + " failed (status " + std::to_string(rc) + "): " +
                             wn_last_error(ctx));
}
}  // namespace

CudaFindPairs::CudaFindPairs(int device)
    : ctx_(nullptr), first_limit_(-1), ties_{0, 0, 0, 0, 0}, box_{0.f, 0.f, 0.f}, periodic_(false)
{
    const int rc = wn_create(device, &ctx_);
    if (rc != WN_OK) raise("CudaFindPairs: wn_create", rc, nullptr);
}

CudaFindPairs::~CudaFindPairs() { wn_destroy(ctx_); }

// ---- handing the pair list to the caller ------------------------------------------------------------------
// At the reference's 2.0555 nm cutoff the list is large (15.9 M pairs = 127 MB for the 100k-atom box) and
// FindPairs::find returns it BY VALUE: the kernels take 0.7 ms and the copy out of pinned memory 2.3 ms, but a freshly
// allocated std::vector of that size costs 31 k first-touch page faults on the calling thread before a byte is copied
// (the range constructor measured 50-100 ms).  So: the allocation is reserved untouched, marked for transparent huge
// pages and first touched by a few threads at once, then filled in one pass; a caller that keeps its vector
// (findInto) pays no faults after the first frame and the copy itself is shared between the threads.
namespace {

typedef std::pair<int, int> IntPair;
static_assert(sizeof(IntPair) == 2 * sizeof(int32_t), "pair<int,int> layout");     // two packed ints: bulk copies

unsigned helper_threads(size_t bytes)
{
    if (bytes < ((size_t)4 << 20)) return 1;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    return std::min(8u, hw);
}

template <typename F>
void split_bytes(size_t bytes, unsigned parts, F fn)       // fn(lo, hi) on equal 2 MB-aligned parts, this thread takes part 0
{
    if (parts <= 1) { fn((size_t)0, bytes); return; }
    const size_t each = (bytes / parts + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    std::vector<std::thread> ts;
    for (unsigned k = 1; k < parts; ++k) {
        const size_t lo = std::min(bytes, each * k), hi = k + 1 == parts ? bytes : std::min(bytes, each * (k + 1));
        if (hi <= lo) continue;
        try { ts.emplace_back([=]() { fn(lo, hi); }); }
        catch (const std::system_error&) { fn(lo, hi); }       // no thread to be had: this thread takes the part as well
    }
    fn((size_t)0, std::min(bytes, each));
    for (std::thread& t : ts) t.join();
}

void fill_pairs(std::vector<IntPair>& out, const IntPair* first, uint64_t m)
{
    if (m == 0) { out.clear(); return; }
    const size_t bytes = (size_t)m * sizeof(IntPair);
    const unsigned parts = helper_threads(bytes);
    if (out.capacity() < m) {
        std::vector<IntPair>().swap(out);
        out.reserve((size_t)m);                                  // raw storage, no page touched yet
        char* raw = reinterpret_cast<char*>(out.data());
#ifdef __linux__
        if (parts > 1) {
            const uintptr_t a = ((uintptr_t)raw + ((uintptr_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1);
            const uintptr_t e = ((uintptr_t)raw + bytes) & ~(((uintptr_t)2 << 20) - 1);
            if (e > a) madvise(reinterpret_cast<void*>(a), (size_t)(e - a), MADV_HUGEPAGE);      // a hint; failure is harmless
        }
#endif
        // first touch from several threads (one byte per 4 KB page of raw storage; assign() below overwrites it)
        if (parts > 1)
            split_bytes(bytes, parts, [raw](size_t lo, size_t hi) { for (size_t o = lo; o < hi; o += 4096) raw[o] = 0; });
        out.assign(first, first + m);                            // one pass (memmove for a trivially copyable type)
        return;
    }
    // the caller's storage is already mapped: set the size (value-initialises only a growing tail), copy in parallel
    out.resize((size_t)m);
    char* dst = reinterpret_cast<char*>(out.data());
    const char* src = reinterpret_cast<const char*>(first);
    split_bytes(bytes, parts, [dst, src](size_t lo, size_t hi) { std::memcpy(dst + lo, src + lo, hi - lo); });
}

}  // namespace

const std::pair<int, int>* CudaFindPairs::findRaw(std::vector<Point>& points, uint64_t* m)
{
    // std::vector<Point> storage is AoS double[n][3] (static_assert in wn_point.hpp)
    const int32_t n = (int32_t)points.size();
    const double* xyz = n ? reinterpret_cast<const double*>(points.data()) : nullptr;
    const int32_t* pairs = nullptr;
    *m = 0;
    const int32_t fl = first_limit_ < 0 ? n : first_limit_;
    const int rc = periodic_ ? wn_find_pairs_pbc(ctx_, xyz, n, fl, box_, &pairs, m, &ties_)
                             : wn_find_pairs(ctx_, xyz, n, fl, &pairs, m, &ties_);
    if (rc != WN_OK) raise("CudaFindPairs::find: wn_find_pairs", rc, ctx_);
    return reinterpret_cast<const IntPair*>(pairs);
}

std::vector<std::pair<int, int>> CudaFindPairs::find(std::vector<Point>& points)
{
    uint64_t m = 0;
    const IntPair* first = findRaw(points, &m);
    std::vector<IntPair> out;
    fill_pairs(out, first, m);
    return out;
}

void CudaFindPairs::findInto(std::vector<Point>& points, std::vector<std::pair<int, int>>& pairs)
{
    uint64_t m = 0;
    const IntPair* first = findRaw(points, &m);
    fill_pairs(pairs, first, m);
}
