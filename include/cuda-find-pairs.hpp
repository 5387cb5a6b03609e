// cuda-find-pairs.hpp -- CudaFindPairs: GPU drop-in for BruteFindPairs
// (reference include/brute-find-pairs.hpp:7-11, src/brute-find-pairs.cpp:7-32).
//
// find() returns exactly what BruteFindPairs::find returns on the same points:
// every (i, j), i < j, with 10.0 * squared_distance(p_i, p_j) < 42.25 evaluated in
// fp64 without FMA, in lexicographic order, as 32-bit ints indexing the caller's
// order.  The work runs on one B200 through the C ABI (wn_find_pairs); there is no
// CPU fallback: construction or find() throws std::runtime_error when the device or
// the kernels are unusable (GROMACS' runAsMain catches std::exception, as it does for
// the reference's std::out_of_range from .at()).
//
// One instance = one wn_ctx = one GPU, single caller (find is called from the single
// GROMACS analysis thread).  Device buffers are allocated lazily, grown geometrically
// and freed in the destructor.
#ifndef WN_CUDAFINDPAIRS_HPP
#define WN_CUDAFINDPAIRS_HPP

#include <cstdint>
#include <utility>
#include <vector>

#include "find-pairs.hpp"
#include "wn_b200.h"

class CudaFindPairs : public FindPairs
{
public:
    explicit CudaFindPairs(int device = 0);
    ~CudaFindPairs() override;
    CudaFindPairs(const CudaFindPairs&) = delete;
    CudaFindPairs& operator=(const CudaFindPairs&) = delete;

    std::vector<std::pair<int, int>> find(std::vector<Point>& points) override;
    // The same list into a vector the caller keeps from frame to frame: its storage stays mapped, so the 127 MB of
    // the 100k-atom box cost one shared copy instead of a fresh allocation's page faults (not in the reference's
    // interface; find() above is).
    void findInto(std::vector<Point>& points, std::vector<std::pair<int, int>>& pairs);
    // ... and without any copy: the library-owned pinned list itself, valid until the next call on this finder.
    const std::pair<int, int>* findRaw(std::vector<Point>& points, uint64_t* n_pairs);

    // Asymmetric search (BASELINE config 5): only rows i < first_limit; <0 = all rows.
    void setFirstLimit(int first_limit) { first_limit_ = first_limit; }
    // Periodic search (extension; the reference has none): rectangular box (Lx,Ly,Lz),
    // minimum-image displacements as defined in wn_b200.h.  clearBox() restores the
    // reference behaviour (open boundary).
    void setBox(float lx, float ly, float lz) { box_[0] = lx; box_[1] = ly; box_[2] = lz; periodic_ = true; }
    void clearBox() { periodic_ = false; }
    // Exact-comparison ties seen by the last find() (10*d2 == 42.25).
    const wn_tie_report& lastTies() const { return ties_; }
    wn_ctx* context() { return ctx_; }

private:
    wn_ctx* ctx_;
    int first_limit_;
    wn_tie_report ties_;
    float box_[3];
    bool periodic_;
};

#endif
