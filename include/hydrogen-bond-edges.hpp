// hydrogen-bond-edges.hpp -- the hydrogen-bond geometry/energy stage as an object.
//
// Replaces, many frames per call, what WaterNetwork::analyzeFrame does per frame
// (reference src/waternetwork.cpp): fromGmxtoCgalPosition (:68-81), the site gather
// (:538-577), mp_->find (:583), the std::async fan-out of make_edges (:590-633,
// :182-263), computeEnergy (:49-66) and the statistics row (:654-666).
//
// Configuration surface = the setters of the reference's CalculateHydrogenBondEnergy
// (include/calculate-hydrogen-bond-energy.hpp:9-10).  Defaults reproduce every live
// call of computeEnergy (r_on 5.5, r_off 6.5, theta_on 0.25, theta_off 0.0,
// src/waternetwork.cpp:51-54), including the caller-side swap of the angle arguments
// (:59-60).  Unlike that (dead) class the setters take the plain values, not
// squares; squaring happens where computeEnergy does it (:57).
//
// FrameBatcher is the batching shim for analyzeFrame: push one frame per call, get the
// per-frame edge lists back in frame order when a batch is full or on finish().
#ifndef WN_HYDROGENBONDEDGES_HPP
#define WN_HYDROGENBONDEDGES_HPP

#include <cstddef>
#include <cstdint>
#include <functional>
#include <vector>

#include "wn_b200.h"
#include "wn_types.hpp"

// One processed batch, host side.  edges of frame f: [offsets[f], offsets[f+1]).
struct HydrogenBondBatch {
    const wn_edge* edges = nullptr;
    const uint64_t* offsets = nullptr;
    const wn_frame_stats* stats = nullptr;
    int n_frames = 0;
    uint64_t n_edges = 0;
    uint64_t n_pair_evals = 0;
    wn_tie_report ties = {0, 0, 0, 0, 0};
};

class HydrogenBondEdges
{
public:
    explicit HydrogenBondEdges(int device = 0);
    ~HydrogenBondEdges();
    HydrogenBondEdges(const HydrogenBondEdges&) = delete;
    HydrogenBondEdges& operator=(const HydrogenBondEdges&) = delete;

    // include/calculate-hydrogen-bond-energy.hpp:9-10
    void setRadiusShift(float r_on, float r_off);
    void setAngleShift(float a_on, float a_off);

    // Site table in the reference's terms.  sites[s].atmIndex is the atom of site s in
    // the frame; hydrogens[s] are the atom indices of its hydrogen list in list order
    // (for waters the reference uses {3w, 3w+1}, src/waternetwork.cpp:564-570).
    void setSites(const std::vector<SiteInfo>& sites, const std::vector<std::vector<int>>& hydrogens,
                  int first_limit = -1);
    // Reference water table for n_waters molecules (OW,HW1,HW2 atom order).
    void setWaterSites(int n_waters, int first_limit = -1);

    // frames: float32 [n_frames][natoms][3] on the host (rvec layout).  The result
    // points into library-owned pinned memory, valid until the next calculate().
    HydrogenBondBatch calculate(const float* frames, int n_frames, int natoms);
    // Periodic variant (extension, defined in wn_b200.h): boxes = [n_frames][3] box lengths,
    // the diagonal of each frame's t_trxframe::box.
    HydrogenBondBatch calculate(const float* frames, const float* boxes, int n_frames, int natoms);
    // General boxes: boxes9 = [n_frames][9], t_trxframe::box row-major (triclinic frames run the
    // all-pairs kernel; a batch of rectangular frames is routed to the cell-list sweep).
    HydrogenBondBatch calculateBox9(const float* frames, const float* boxes9, int n_frames, int natoms);

    // Per-frame site lists (what as_->locate() leaves, src/waternetwork.cpp:529-533): frame f uses the
    // table rows sites[offsets[f] .. offsets[f+1]); edge indices are positions in that list.
    HydrogenBondBatch calculate(const float* frames, const float* boxes, int n_frames, int natoms,
                                const int64_t* subset_offsets, const int32_t* subset_sites);

    // Text of edges-frames.xvg.dat for the LAST calculate() (frame lines + 37-byte rows, the bytes
    // the reference's iostream code writes, src/waternetwork.cpp:638-652), formatted on the GPU.
    // ids[s] = SiteInfo::id of table row s (empty: the index).  Returns false when a field does not
    // fit the fixed-width row (ids over 6 digits, values over 8 characters); use EdgeFileWriter then.
    bool formatLastBatch(const std::vector<int>& frameNumbers, const std::vector<unsigned>& ids,
                         const char** text, uint64_t* bytes);

    // Connected components of the networks of the LAST calculate() (python/network-analysis.py:141-171:
    // edges with |energy| <= weightThreshold dropped, then sites without an edge dropped).
    // labels[f * nSites + s] = smallest site of s's component in frame f, -1 without a kept edge;
    // library-owned pinned memory, valid until the next call.
    void componentsOfLastBatch(float weightThreshold, const int32_t** labels, const wn_component_stats** stats);

    // Convenience: one frame's result as the reference's std::vector<HydrogenBond>
    // (pair order, direction FORWARD), given the SiteInfo_ptr table of the caller.
    static std::vector<HydrogenBond> toHydrogenBonds(const HydrogenBondBatch& batch, int frame,
                                                     const std::vector<SiteInfo_ptr>& sites);

    wn_ctx* context() { return ctx_; }

private:
    void pushParams();
    wn_ctx* ctx_;
    double r_on_, r_off_, a_on_, a_off_;
};

// Frame batching for WaterNetwork::analyzeFrame (src/waternetwork.cpp:503): buffer
// frames on the host, run one batch per `batch_frames`, deliver results per frame in
// frame order through the sink (which writes the edge rows :638-652 and the data-set
// points :654-666).  finish() flushes the tail (call from finishAnalysis, :670).
class FrameBatcher
{
public:
    typedef std::function<void(int frnr, const wn_edge* edges, size_t n_edges,
                               const wn_frame_stats& stats)> Sink;

    FrameBatcher(HydrogenBondEdges& engine, int natoms, int batch_frames, Sink sink);
    void push(int frnr, const float* xyz /* natoms*3 */);
    // periodic: box = GROMACS matrix fr.box (rectangular or triclinic)
    void push(int frnr, const float* xyz, const float box[3][3]);
    // with the frame's own site list (rows of the site table, e.g. protein sites + the buried
    // waters of this frame); box may be nullptr.  All frames of a run must use the same variant.
    void push(int frnr, const float* xyz, const float (*box)[3], const std::vector<int32_t>& frameSites);
    void finish();
    uint64_t framesDone() const { return frames_done_; }
    // Optional: called once per processed batch, before the per-frame sink calls, while the batch is
    // still the engine's "last batch" (so formatLastBatch / componentsOfLastBatch apply to it).
    typedef std::function<void(const std::vector<int>& frameNumbers, const HydrogenBondBatch& batch)> BatchHook;
    void setBatchHook(BatchHook hook) { batch_hook_ = hook; }

private:
    void flush();
    HydrogenBondEdges& engine_;
    int natoms_, batch_frames_;
    Sink sink_;
    BatchHook batch_hook_;
    std::vector<float> buf_;
    std::vector<float> boxes_;
    std::vector<int64_t> sub_off_;
    std::vector<int32_t> sub_sites_;
    std::vector<int> frnr_;
    uint64_t frames_done_;
};

#endif
