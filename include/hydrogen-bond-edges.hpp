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
#include <memory>
#include <future>
#include <utility>
#include <vector>

#include "wn_b200.h"
#include "wn_types.hpp"

// One processed batch, host side.  edges of frame f: [offsets[f], offsets[f+1]).
// layout WN_LAYOUT_EDGES: `edges` are wn_edge records.  Compact layouts (WN_LAYOUT_CSR16 / _CSR12): the same
// memory holds wn_edge16 / wn_edge12 records (use edges16() / edges12()) and the row index i comes from
// row_offsets: edges of site i in frame f = [offsets[f] + row_offsets[f*n_sites + i], ... next row).
struct HydrogenBondBatch {
    const wn_edge* edges = nullptr;
    const uint64_t* offsets = nullptr;
    const wn_frame_stats* stats = nullptr;
    int n_frames = 0;
    uint64_t n_edges = 0;
    uint64_t n_pair_evals = 0;
    wn_tie_report ties = {0, 0, 0, 0, 0};
    const uint32_t* row_offsets = nullptr;
    int n_sites = 0;
    int layout = WN_LAYOUT_EDGES;
    const wn_edge16* edges16() const { return reinterpret_cast<const wn_edge16*>(edges); }
    const wn_edge12* edges12() const { return reinterpret_cast<const wn_edge12*>(edges); }
};

// One frame of a batch as the compact sink sees it (any layout).
struct FrameResult {
    int frnr = 0;
    int layout = WN_LAYOUT_EDGES;
    const void* edges = nullptr;           // wn_edge / wn_edge16 / wn_edge12 records of this frame
    size_t n_edges = 0;
    const uint32_t* row_offsets = nullptr; // [n_sites] offsets of the rows inside this frame's records
    int n_sites = 0;
    wn_frame_stats stats = {0, 0, 0, 0, 0};
    int device = 0;                        // GPU that processed the frame
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

    // make_edges of the reference (src/waternetwork.cpp:182-263) on ONE frame and a caller-supplied pair list, in
    // the list's own order and (first, second) orientation -- what the live module does with the list of its
    // DelaunayFindPairs (:399,:583,:590-633).  Keep the host finder, hand its pairs over, get the same
    // std::vector<HydrogenBond> the reference's chunked std::async fan-out joins (:621-633).
    // frame: float32 [natoms][3]; box: NULL (reference behaviour) or the GROMACS box matrix (periodic extension).
    std::vector<HydrogenBond> makeEdges(const float* frame, int natoms, const std::vector<std::pair<int, int>>& pairs,
                                        const std::vector<SiteInfo_ptr>& sites, const float (*box)[3] = nullptr);
    // the same, raw: library-owned wn_edge records (i = first, j = second), valid until the next call
    const wn_edge* makeEdgesRaw(const float* frame, int natoms, const std::pair<int, int>* pairs, size_t n_pairs,
                                size_t* n_edges, const float (*box)[3] = nullptr);

    // Convenience: one frame's result as the reference's std::vector<HydrogenBond>
    // (pair order, direction FORWARD), given the SiteInfo_ptr table of the caller.
    static std::vector<HydrogenBond> toHydrogenBonds(const HydrogenBondBatch& batch, int frame,
                                                     const std::vector<SiteInfo_ptr>& sites);

    // The general form (wn_hbond_batch_ex): any combination of boxes, site lists, result layout.
    HydrogenBondBatch calculateEx(const wn_batch_args& args);
    // 2: consecutive calculate() results alternate between two host buffer sets (a result stays valid
    // while the next batch runs); FrameBatcher switches this on for pipelined batching.
    void setResultSets(int n);
    // Pin the calling thread next to this engine's GPU (see wn_bind_thread_to_device); returns the CPUs in the set.
    int bindThreadToDevice();
    // Pinned host memory (for staging frames): faster H2D, required for asynchronous overlap.
    float* allocPinnedFloats(size_t count);
    void freePinned(void* p);
    // Several GPUs: per-frame statistics rows are summed over the shards with NCCL (the only collective).
    void ncclInit(const void* id128, int rank, int nranks);
    void statsReduce(wn_frame_stats* table, int64_t n_frames_total);
    int device() const { return device_; }

    wn_ctx* context() { return ctx_; }

private:
    void pushParams();
    wn_ctx* ctx_;
    int device_;
    double r_on_, r_off_, a_on_, a_off_;
};

// Frame batching for WaterNetwork::analyzeFrame (src/waternetwork.cpp:503): frames are staged in PINNED host
// memory, one batch per `batch_frames`, results delivered per frame in frame order through the sink (which
// writes the edge rows :638-652 and the data-set points :654-666).  finish() flushes the tail (call from
// finishAnalysis, :670).
//
// Pipelined: while batch k is on the GPU (a worker thread inside the C ABI call), the caller keeps pushing
// batch k+1 into the second staging buffer; batch k is handed to the sink -- on the CALLER's thread, in frame
// order -- when batch k+1 is submitted, i.e. while the GPU already works on k+1 (two host result sets).
// Several engines (one per GPU): batches go round-robin to the GPUs, i.e. every GPU owns contiguous blocks of
// batch_frames frames; delivery stays in frame order; the only cross-GPU exchange is the NCCL sum
// of the statistics table at finish() (reducedStats()), which serves the .xvg rows.
class FrameBatcher
{
public:
    typedef std::function<void(int frnr, const wn_edge* edges, size_t n_edges,
                               const wn_frame_stats& stats)> Sink;
    typedef std::function<void(const FrameResult&)> FrameSink;

    // classic form: 20-byte wn_edge records ({i, j, energy, length, cosine})
    FrameBatcher(HydrogenBondEdges& engine, int natoms, int batch_frames, Sink sink);
    // compact form (default layout WN_LAYOUT_CSR16: 16 bytes per edge over PCIe), one or several GPUs
    FrameBatcher(const std::vector<HydrogenBondEdges*>& engines, int natoms, int batch_frames, FrameSink sink,
                 int layout = WN_LAYOUT_CSR16);
    ~FrameBatcher();
    FrameBatcher(const FrameBatcher&) = delete;
    FrameBatcher& operator=(const FrameBatcher&) = delete;

    void push(int frnr, const float* xyz /* natoms*3 */);
    // periodic: box = GROMACS matrix fr.box (rectangular or triclinic)
    void push(int frnr, const float* xyz, const float box[3][3]);
    // with the frame's own site list (rows of the site table, e.g. protein sites + the buried
    // waters of this frame); box may be nullptr.  All frames of a run must use the same variant.
    void push(int frnr, const float* xyz, const float (*box)[3], const std::vector<int32_t>& frameSites);
    // zero-copy producers: where the next frame's natoms*3 floats go (pinned memory); then commit(frnr)
    float* stage();
    void commit(int frnr);
    // submit the partial batch and deliver everything in flight (the pipeline is empty afterwards)
    void flush();
    void finish();
    uint64_t framesDone() const { return frames_done_; }
    // Optional: called once per processed batch, before the per-frame sink calls, while the batch is
    // still the engine's "last batch" (so formatLastBatch / componentsOfLastBatch apply to it).  A hook
    // switches the pipelining off (batch k is fully delivered before batch k+1 is submitted).
    typedef std::function<void(const std::vector<int>& frameNumbers, const HydrogenBondBatch& batch)> BatchHook;
    void setBatchHook(BatchHook hook) { batch_hook_ = hook; }
    // push() copies every frame into pinned staging memory; one host thread moves ~7 GB/s (~6 k frames/s of the
    // 100k-atom box).  n > 1: n - 1 helper threads share each copy (include/frame-copy-pool.hpp).  Default 1.
    void setCopyThreads(int n);
    void setPipelined(bool on) { pipelined_ = on; }
    // Several GPUs: false (default) -- batches go round-robin; true -- the next batch is staged for the GPU expected to
    // be free first (GPUs behind a shared PCIe uplink are slower end to end, round-robin waits for the slowest).
    // Results are delivered in frame order either way.  Opt-in: measured on 2 GPUs only.
    void setDynamicGpuChoice(bool on) { dynamic_ = on; }
    // After finish(): one row per pushed frame, in push order.  With several GPUs this is the NCCL sum of
    // the per-GPU zero-padded tables (wn_stats_reduce), with one GPU the rows as delivered.
    const std::vector<wn_frame_stats>& reducedStats() const { return reduced_; }
    int gpus() const { return (int)lanes_.size(); }
    // wall-clock seconds every batch spent inside the C ABI call (H2D + kernels + D2H), in delivery order
    const std::vector<double>& batchSeconds() const { return batch_seconds_; }

private:
    struct Lane;
    int chooseLane() const;        // GPU expected to be free first (setDynamicGpuChoice)
    void init(int natoms, int batch_frames, int layout);
    void destroy();                     // workers joined, staging freed (destructor, and a constructor that fails half way)
    void submit();                      // current lane's filling buffer -> GPU
    void deliverOldest();               // wait for the oldest in-flight batch, hand it to the sink
    std::vector<Lane*> lanes_;
    std::vector<int> pending_;          // lanes with a batch in flight, oldest first
    int cur_;                           // lane being filled
    int natoms_, batch_frames_, layout_;
    Sink sink_;
    FrameSink frame_sink_;
    BatchHook batch_hook_;
    bool pipelined_;
    bool dynamic_ = false;
    uint64_t frames_done_, frames_pushed_;
    std::vector<wn_frame_stats> reduced_;
    std::vector<double> batch_seconds_;
    std::unique_ptr<class FrameCopyPool> copy_pool_;
};

#endif
