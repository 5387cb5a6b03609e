// cuda-find-pairs.cpp -- CudaFindPairs over the C ABI (see include/cuda-find-pairs.hpp).
#include "cuda-find-pairs.hpp"

#include <stdexcept>
#include <string>

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

std::vector<std::pair<int, int>> CudaFindPairs::find(std::vector<Point>& points)
{
    // std::vector<Point> storage is AoS double[n][3] (static_assert in wn_point.hpp)
    const int32_t n = (int32_t)points.size();
    const double* xyz = n ? reinterpret_cast<const double*>(points.data()) : nullptr;
    const int32_t* pairs = nullptr;
    uint64_t m = 0;
    const int32_t fl = first_limit_ < 0 ? n : first_limit_;
    const int rc = periodic_ ? wn_find_pairs_pbc(ctx_, xyz, n, fl, box_, &pairs, &m, &ties_)
                             : wn_find_pairs(ctx_, xyz, n, fl, &pairs, &m, &ties_);
    if (rc != WN_OK) raise("CudaFindPairs::find: wn_find_pairs", rc, ctx_);
    // std::pair<int,int> is two packed ints: one bulk range copy instead of m push_backs
    static_assert(sizeof(std::pair<int, int>) == 2 * sizeof(int32_t), "pair<int,int> layout");
    const std::pair<int, int>* first = reinterpret_cast<const std::pair<int, int>*>(pairs);
    return std::vector<std::pair<int, int>>(first, first + m);      // one pass, no zero-fill
}
