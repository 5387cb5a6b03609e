// hydrogen-bond-edges.cpp -- HydrogenBondEdges and FrameBatcher over the C ABI
// (see include/hydrogen-bond-edges.hpp).
#include "hydrogen-bond-edges.hpp"
#include "frame-copy-pool.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <condition_variable>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <thread>

namespace {
[[noreturn]] void raise(const char* what, int rc, const wn_ctx* ctx)
{
    throw std::runtime_error(std::string(what) + " failed (status " + std::to_string(rc) + "): " +
                             wn_last_error(ctx));
}

HydrogenBondBatch wrap(const wn_batch_result& r)
{
    HydrogenBondBatch b;
    b.edges = r.edges; b.offsets = r.frame_offsets; b.stats = r.stats;
    b.n_frames = r.n_frames; b.n_edges = r.n_edges; b.n_pair_evals = r.n_pair_evals; b.ties = r.ties;
    b.row_offsets = r.row_offsets; b.n_sites = r.n_sites; b.layout = r.layout;
    return b;
}
}  // namespace

HydrogenBondEdges::HydrogenBondEdges(int device)
    : ctx_(nullptr), device_(device), r_on_(5.5), r_off_(6.5), a_on_(0.25), a_off_(0.0)
{
    const int rc = wn_create(device, &ctx_);
    if (rc != WN_OK) raise("HydrogenBondEdges: wn_create", rc, nullptr);
}

HydrogenBondEdges::~HydrogenBondEdges() { wn_destroy(ctx_); }

void HydrogenBondEdges::pushParams()
{
    const int rc = wn_set_energy_params(ctx_, r_on_, r_off_, a_on_, a_off_);
    if (rc != WN_OK) raise("HydrogenBondEdges: wn_set_energy_params", rc, ctx_);
}

void HydrogenBondEdges::setRadiusShift(float r_on, float r_off)
{
    r_on_ = r_on; r_off_ = r_off;
    pushParams();
}

void HydrogenBondEdges::setAngleShift(float a_on, float a_off)
{
    a_on_ = a_on; a_off_ = a_off;
    pushParams();
}

void HydrogenBondEdges::setSites(const std::vector<SiteInfo>& sites,
                                 const std::vector<std::vector<int>>& hydrogens, int first_limit)
{
    if (sites.size() != hydrogens.size())
        throw std::invalid_argument("HydrogenBondEdges::setSites: sites and hydrogens differ in length");
    const int32_t n = (int32_t)sites.size();
    size_t hmax = 1;
    for (const auto& h : hydrogens) hmax = std::max(hmax, h.size());
    std::vector<int32_t> site_atom((size_t)n), h_atom((size_t)n * hmax, -1);
    std::vector<uint8_t> type((size_t)n);
    for (int32_t s = 0; s < n; ++s) {
        site_atom[s] = (int32_t)sites[s].atmIndex;
        type[s] = (uint8_t)sites[s].type;
        for (size_t k = 0; k < hydrogens[s].size(); ++k) h_atom[(size_t)s * hmax + k] = hydrogens[s][k];
    }
    const int rc = wn_set_sites(ctx_, n, site_atom.data(), h_atom.data(), (int32_t)hmax, type.data(),
                                first_limit < 0 ? n : first_limit);
    if (rc != WN_OK) raise("HydrogenBondEdges::setSites: wn_set_sites", rc, ctx_);
}

void HydrogenBondEdges::setWaterSites(int n_waters, int first_limit)
{
    const int rc = wn_set_water_sites(ctx_, n_waters, first_limit < 0 ? n_waters : first_limit);
    if (rc != WN_OK) raise("HydrogenBondEdges::setWaterSites: wn_set_water_sites", rc, ctx_);
}

HydrogenBondBatch HydrogenBondEdges::calculate(const float* frames, int n_frames, int natoms)
{
    return calculate(frames, nullptr, n_frames, natoms);
}

HydrogenBondBatch HydrogenBondEdges::calculate(const float* frames, const float* boxes, int n_frames, int natoms)
{
    wn_batch_result r;
    const int rc = boxes ? wn_hbond_batch_pbc(ctx_, frames, boxes, n_frames, natoms, &r)
                         : wn_hbond_batch(ctx_, frames, n_frames, natoms, &r);
    if (rc != WN_OK) raise("HydrogenBondEdges::calculate: wn_hbond_batch", rc, ctx_);
    return wrap(r);
}

HydrogenBondBatch HydrogenBondEdges::calculate(const float* frames, const float* boxes, int n_frames, int natoms,
                                               const int64_t* subset_offsets, const int32_t* subset_sites)
{
    wn_batch_result r;
    const int rc = wn_hbond_batch_subset(ctx_, frames, boxes, n_frames, natoms, subset_offsets, subset_sites, &r);
    if (rc != WN_OK) raise("HydrogenBondEdges::calculate: wn_hbond_batch_subset", rc, ctx_);
    return wrap(r);
}

HydrogenBondBatch HydrogenBondEdges::calculateBox9(const float* frames, const float* boxes9, int n_frames, int natoms)
{
    wn_batch_result r;
    const int rc = wn_hbond_batch_box9(ctx_, frames, boxes9, n_frames, natoms, &r);
    if (rc != WN_OK) raise("HydrogenBondEdges::calculateBox9: wn_hbond_batch_box9", rc, ctx_);
    return wrap(r);
}

HydrogenBondBatch HydrogenBondEdges::calculateEx(const wn_batch_args& args)
{
    wn_batch_result r;
    const int rc = wn_hbond_batch_ex(ctx_, &args, &r);
    if (rc != WN_OK) raise("HydrogenBondEdges::calculateEx: wn_hbond_batch_ex", rc, ctx_);
    return wrap(r);
}

void HydrogenBondEdges::setResultSets(int n)
{
    const int rc = wn_set_result_sets(ctx_, n);
    if (rc != WN_OK) raise("HydrogenBondEdges::setResultSets", rc, ctx_);
}

int HydrogenBondEdges::bindThreadToDevice()
{
    int32_t n = 0;
    const int rc = wn_bind_thread_to_device(ctx_, &n);
    if (rc != WN_OK) raise("HydrogenBondEdges::bindThreadToDevice", rc, ctx_);
    return n;
}

float* HydrogenBondEdges::allocPinnedFloats(size_t count)
{
    void* p = nullptr;
    const int rc = wn_host_alloc_pinned(ctx_, count * sizeof(float), &p);
    if (rc != WN_OK) raise("HydrogenBondEdges::allocPinnedFloats: wn_host_alloc_pinned", rc, ctx_);
    return static_cast<float*>(p);
}

void HydrogenBondEdges::freePinned(void* p) { if (p) wn_host_free_pinned(ctx_, p); }

void HydrogenBondEdges::ncclInit(const void* id128, int rank, int nranks)
{
    const int rc = wn_nccl_init(ctx_, id128, rank, nranks);
    if (rc != WN_OK) raise("HydrogenBondEdges::ncclInit: wn_nccl_init", rc, ctx_);
}

void HydrogenBondEdges::statsReduce(wn_frame_stats* table, int64_t n_frames_total)
{
    const int rc = wn_stats_reduce(ctx_, table, n_frames_total);
    if (rc != WN_OK) raise("HydrogenBondEdges::statsReduce: wn_stats_reduce", rc, ctx_);
}

bool HydrogenBondEdges::formatLastBatch(const std::vector<int>& frameNumbers, const std::vector<unsigned>& ids,
                                        const char** text, uint64_t* bytes)
{
    const int rc = wn_format_edges(ctx_, frameNumbers.data(), ids.empty() ? nullptr : ids.data(), (int32_t)ids.size(),
                                   text, bytes);
    if (rc == WN_ERR_OVERFLOW) return false;
    if (rc != WN_OK) raise("HydrogenBondEdges::formatLastBatch: wn_format_edges", rc, ctx_);
    return true;
}

void HydrogenBondEdges::componentsOfLastBatch(float weightThreshold, const int32_t** labels, const wn_component_stats** stats)
{
    const int rc = wn_components(ctx_, weightThreshold, labels, stats);
    if (rc != WN_OK) raise("HydrogenBondEdges::componentsOfLastBatch: wn_components", rc, ctx_);
}

const wn_edge* HydrogenBondEdges::makeEdgesRaw(const float* frame, int natoms, const std::pair<int, int>* pairs,
                                              size_t n_pairs, size_t* n_edges, const float (*box)[3])
{
    static_assert(sizeof(std::pair<int, int>) == 2 * sizeof(int32_t), "pair<int,int> layout");
    pushParams();
    float b9[9];
    if (box)
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) b9[3 * a + c] = box[a][c];
    const wn_edge* e = nullptr;
    uint64_t n = 0;
    const int rc = wn_make_edges(ctx_, frame, natoms, box ? b9 : nullptr, reinterpret_cast<const int32_t*>(pairs),
                                 (uint64_t)n_pairs, &e, &n, nullptr, nullptr);
    if (rc != WN_OK) {
        // an index outside the site table is what the reference's .at() reports as std::out_of_range
        const std::string msg = std::string("HydrogenBondEdges::makeEdges: wn_make_edges failed (status ") + std::to_string(rc) +
                                "): " + wn_last_error(ctx_);
        if (rc == WN_ERR_BAD_ARG && msg.find("outside the table") != std::string::npos) throw std::out_of_range(msg);
        throw std::runtime_error(msg);
    }
    if (n_edges) *n_edges = (size_t)n;
    return e;
}

std::vector<HydrogenBond> HydrogenBondEdges::makeEdges(const float* frame, int natoms,
                                                       const std::vector<std::pair<int, int>>& pairs,
                                                       const std::vector<SiteInfo_ptr>& sites, const float (*box)[3])
{
    size_t n = 0;
    const wn_edge* e = makeEdgesRaw(frame, natoms, pairs.data(), pairs.size(), &n, box);
    std::vector<HydrogenBond> out;
    out.reserve(n);
    for (size_t k = 0; k < n; ++k)
        out.push_back(HydrogenBond{SiteInfoPair(sites.at(e[k].i), sites.at(e[k].j)), FORWARD,
                                   (double)e[k].length, (double)e[k].cosine, (double)e[k].energy});
    return out;
}

std::vector<HydrogenBond> HydrogenBondEdges::toHydrogenBonds(const HydrogenBondBatch& batch, int frame,
                                                             const std::vector<SiteInfo_ptr>& sites)
{
    std::vector<HydrogenBond> out;
    if (frame < 0 || frame >= batch.n_frames) throw std::out_of_range("toHydrogenBonds: frame");
    out.reserve((size_t)(batch.offsets[frame + 1] - batch.offsets[frame]));
    for (uint64_t k = batch.offsets[frame]; k < batch.offsets[frame + 1]; ++k) {
        const wn_edge& e = batch.edges[k];
        out.push_back(HydrogenBond{SiteInfoPair(sites.at(e.i), sites.at(e.j)), FORWARD,
                                   (double)e.length, (double)e.cosine, (double)e.energy});
    }
    return out;
}

// ---------------------------------------------------------------- FrameBatcher
// One persistent worker thread per GPU: it is bound next to its device once, takes one batch call at a time and
// hands the result (or the exception) back.  (A std::async per batch paid a thread start, an affinity call and the
// per-thread CUDA set-up every 35 ms: ~4.8 ms of idle GPU between two batches.)
class BatchWorker
{
public:
    explicit BatchWorker(HydrogenBondEdges* eng) : eng_(eng), thread_([this]() { run(); }) {}
    ~BatchWorker()
    {
        { std::lock_guard<std::mutex> g(m_); quit_ = true; }
        cv_.notify_all();
        thread_.join();
    }
    void start(const wn_batch_args& a, double* seconds)
    {
        { std::lock_guard<std::mutex> g(m_); args_ = a; seconds_ = seconds; has_job_ = true; done_ = false; err_ = nullptr; }
        cv_.notify_all();
    }
    bool finished() { std::lock_guard<std::mutex> g(m_); return done_; }      // the batch handed over last has left the GPU
    HydrogenBondBatch get()
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this]() { return done_; });
        started_ = false;
        if (err_) { std::exception_ptr e = err_; err_ = nullptr; std::rethrow_exception(e); }
        return result_;
    }

private:
    void run()
    {
        try { eng_->bindThreadToDevice(); } catch (...) {}
        for (;;) {
            wn_batch_args a;
            double* secs;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [this]() { return has_job_ || quit_; });
                if (quit_) return;
                a = args_; secs = seconds_; has_job_ = false; started_ = true;
            }
            HydrogenBondBatch r;
            std::exception_ptr err;
            const auto t0 = std::chrono::steady_clock::now();
            try { r = eng_->calculateEx(a); } catch (...) { err = std::current_exception(); }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            {
                std::lock_guard<std::mutex> g(m_);
                result_ = r; err_ = err; done_ = true;
                if (secs) *secs = dt;
            }
            cv_.notify_all();
        }
    }
    HydrogenBondEdges* eng_;
    std::mutex m_;
    std::condition_variable cv_;
    wn_batch_args args_;
    double* seconds_ = nullptr;
    HydrogenBondBatch result_;
    std::exception_ptr err_;
    bool has_job_ = false, started_ = false, done_ = false, quit_ = false;
    std::thread thread_;          // last: starts when everything above exists
};

struct FrameBatcher::Lane {
    HydrogenBondEdges* eng = nullptr;
    float* stage[2] = {nullptr, nullptr};      // pinned staging buffers: one fills while the other is on the GPU
    int fill = 0;
    std::vector<int> frnr[2];
    std::vector<uint64_t> seq[2];              // push sequence number of every frame = its row of the statistics table
    std::vector<float> boxes[2];
    std::vector<int64_t> sub_off[2];
    std::vector<int32_t> sub_sites[2];
    std::vector<float> b3;                     // rectangular box lengths of a site-list batch
    std::unique_ptr<BatchWorker> worker;       // created with the first pipelined batch
    double seconds[2] = {0.0, 0.0};            // wall time of the batch of each staging buffer inside the C ABI
    double t_start = 0.0;                      // when the batch in flight was handed to the worker (steady clock, s)
    double typical = 0.0;                      // this GPU's last batch time (0: not known yet)
    int flying = -1;                           // staging buffer of the batch in flight, -1: none
    bool have_ready = false;                   // a finished batch waits for delivery
    HydrogenBondBatch ready;
    int ready_buf = 0;
    std::vector<wn_frame_stats> table;         // this GPU's rows of the run's statistics table (zero elsewhere)
};

void FrameBatcher::init(int natoms, int batch_frames, int layout)
{
    natoms_ = natoms; batch_frames_ = std::max(batch_frames, 1); layout_ = layout;
    cur_ = 0; frames_done_ = 0; frames_pushed_ = 0; pipelined_ = true;
    const size_t count = (size_t)batch_frames_ * (size_t)natoms_ * 3;
    for (Lane* l : lanes_) {
        // allocate from a thread pinned next to the GPU: pinned pages are touched (placed) by the allocating thread
        std::exception_ptr err;
        std::thread t([&]() {
            try {
                l->eng->bindThreadToDevice();
                l->stage[0] = l->eng->allocPinnedFloats(std::max<size_t>(count, 1));
                l->stage[1] = l->eng->allocPinnedFloats(std::max<size_t>(count, 1));
                l->eng->setResultSets(2);
            } catch (...) { err = std::current_exception(); }
        });
        t.join();
        if (err) std::rethrow_exception(err);
    }
    if (lanes_.size() > 1) {
        // one NCCL communicator over the GPUs of this process (statistics table only)
        unsigned char id[128];
        if (wn_nccl_unique_id(id) != WN_OK)
            throw std::runtime_error(std::string("FrameBatcher: wn_nccl_unique_id failed: ") + wn_last_error(nullptr));
        std::vector<std::thread> ts;
        std::vector<std::exception_ptr> errs(lanes_.size());
        for (size_t g = 0; g < lanes_.size(); ++g)
            ts.emplace_back([&, g]() {
                try { lanes_[g]->eng->ncclInit(id, (int)g, (int)lanes_.size()); } catch (...) { errs[g] = std::current_exception(); }
            });
        for (auto& t : ts) t.join();
        for (auto& e : errs) if (e) std::rethrow_exception(e);
    }
}

FrameBatcher::FrameBatcher(HydrogenBondEdges& engine, int natoms, int batch_frames, Sink sink)
    : sink_(sink)
{
    Lane* l = new Lane();
    l->eng = &engine;
    lanes_.push_back(l);
    try { init(natoms, batch_frames, WN_LAYOUT_EDGES); } catch (...) { destroy(); throw; }
}

FrameBatcher::FrameBatcher(const std::vector<HydrogenBondEdges*>& engines, int natoms, int batch_frames, FrameSink sink,
                           int layout)
    : frame_sink_(sink)
{
    if (engines.empty()) throw std::invalid_argument("FrameBatcher: no engine");
    for (HydrogenBondEdges* e : engines) {
        Lane* l = new Lane();
        l->eng = e;
        lanes_.push_back(l);
    }
    try { init(natoms, batch_frames, layout); } catch (...) { destroy(); throw; }      // a constructor that throws runs no destructor
}

FrameBatcher::~FrameBatcher() { destroy(); }

void FrameBatcher::destroy()
{
    for (Lane* l : lanes_) {
        if (l->flying >= 0 && l->worker) { try { l->worker->get(); } catch (...) {} }
        l->worker.reset();
        l->eng->freePinned(l->stage[0]);
        l->eng->freePinned(l->stage[1]);
        try { l->eng->setResultSets(1); } catch (...) {}
        delete l;
    }
    lanes_.clear();
}

float* FrameBatcher::stage()
{
    Lane& l = *lanes_[cur_];
    return l.stage[l.fill] + l.frnr[l.fill].size() * (size_t)natoms_ * 3;
}

void FrameBatcher::commit(int frnr)
{
    Lane& l = *lanes_[cur_];
    l.frnr[l.fill].push_back(frnr);
    l.seq[l.fill].push_back(frames_pushed_++);
    if ((int)l.frnr[l.fill].size() >= batch_frames_) submit();
}

void FrameBatcher::setCopyThreads(int n)
{
    copy_pool_.reset(n > 1 ? new FrameCopyPool(n - 1) : nullptr);
}

void FrameBatcher::push(int frnr, const float* xyz)
{
    const size_t bytes = (size_t)natoms_ * 3 * sizeof(float);
    if (bytes) {                                    // a frame without atoms: nothing to copy (and xyz may be null)
        if (copy_pool_) copy_pool_->copy(stage(), xyz, bytes);
        else std::memcpy(stage(), xyz, bytes);
    }
    commit(frnr);
}

void FrameBatcher::push(int frnr, const float* xyz, const float box[3][3])
{
    Lane& l = *lanes_[cur_];
    if (!l.frnr[l.fill].empty() && l.boxes[l.fill].size() != 9 * l.frnr[l.fill].size())
        throw std::invalid_argument("FrameBatcher::push: periodic and open frames mixed in one batch");
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) l.boxes[l.fill].push_back(box[a][b]);
    push(frnr, xyz);
}

void FrameBatcher::push(int frnr, const float* xyz, const float (*box)[3], const std::vector<int32_t>& frameSites)
{
    Lane& l = *lanes_[cur_];
    const int b = l.fill;
    if (!l.frnr[b].empty() && l.sub_off[b].size() != l.frnr[b].size() + 1)
        throw std::invalid_argument("FrameBatcher::push: frames with and without site lists mixed in one batch");
    if (l.sub_off[b].empty()) l.sub_off[b].push_back(0);
    l.sub_sites[b].insert(l.sub_sites[b].end(), frameSites.begin(), frameSites.end());
    l.sub_off[b].push_back((int64_t)l.sub_sites[b].size());
    if (box) {
        for (int a = 0; a < 3; ++a)
            for (int c = 0; c < 3; ++c) {
                if (a != c && box[a][c] != 0.0f)
                    throw std::invalid_argument("FrameBatcher::push: per-frame site lists need a rectangular box");
                l.boxes[b].push_back(box[a][c]);
            }
    }
    push(frnr, xyz);
}

void FrameBatcher::deliverOldest()
{
    if (pending_.empty()) return;
    Lane& l = *lanes_[pending_.front()];
    pending_.erase(pending_.begin());
    if (!l.have_ready) {                        // still in flight: wait for it
        l.ready = l.worker->get();
        l.ready_buf = l.flying;
        l.typical = l.seconds[l.flying];
        l.flying = -1;
    }
    l.have_ready = false;
    const HydrogenBondBatch& b = l.ready;
    const int buf = l.ready_buf;
    batch_seconds_.push_back(l.seconds[buf]);
    const size_t esz = b.layout == WN_LAYOUT_CSR16 ? sizeof(wn_edge16) : b.layout == WN_LAYOUT_CSR12 ? sizeof(wn_edge12)
                                                                                                    : sizeof(wn_edge);
    for (int f = 0; f < b.n_frames; ++f) {
        const uint64_t seq = l.seq[buf][f];
        if (l.table.size() <= seq) l.table.resize(seq + 1, wn_frame_stats{0, 0, 0, 0, 0});
        l.table[seq] = b.stats[f];
        const size_t n = (size_t)(b.offsets[f + 1] - b.offsets[f]);
        if (frame_sink_) {
            FrameResult r;
            r.frnr = l.frnr[buf][f]; r.layout = b.layout;
            r.edges = reinterpret_cast<const char*>(b.edges) + (size_t)b.offsets[f] * esz;
            r.n_edges = n;
            r.row_offsets = b.row_offsets ? b.row_offsets + (size_t)f * b.n_sites : nullptr;
            r.n_sites = b.n_sites; r.stats = b.stats[f]; r.device = l.eng->device();
            frame_sink_(r);
        } else if (sink_) {
            sink_(l.frnr[buf][f], b.edges + b.offsets[f], n, b.stats[f]);
        }
    }
    frames_done_ += (uint64_t)b.n_frames;
    l.frnr[buf].clear(); l.seq[buf].clear(); l.boxes[buf].clear(); l.sub_off[buf].clear(); l.sub_sites[buf].clear();
}

namespace {
double steady_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

// The GPU the next batch is staged for: an idle one if there is any, otherwise the one whose batch in flight is
// expected to finish first (start time + that GPU's typical batch time).  GPUs of one host are not equally fast end to
// end -- several may share a PCIe uplink -- and fixed round-robin makes everyone wait for the slowest.
int FrameBatcher::chooseLane() const
{
    const int n = (int)lanes_.size();
    int best = (cur_ + 1) % n;
    double best_t = 1.0e300;
    for (int k = 1; k <= n; ++k) {
        const int c = (cur_ + k) % n;
        const Lane& l = *lanes_[c];
        // free now / free once its finished batch has gone to the sink (frame order may make that wait) / busy
        double t;
        if (l.flying < 0) t = l.have_ready ? 1.0 : 0.0;
        else if (l.worker && l.worker->finished()) t = l.have_ready ? 1.0 : 0.5;      // done, result not collected yet
        else t = l.typical > 0.0 ? l.t_start + l.typical : 1.0e299;      // no batch of this GPU timed yet: rotation order
        if (t < best_t) { best_t = t; best = c; }
    }
    return best;
}

void FrameBatcher::submit()
{
    Lane& l = *lanes_[cur_];
    const int buf = l.fill;
    if (l.frnr[buf].empty()) return;
    // a finished batch of this GPU that has not reached the sink yet must go first (frame order; the GPU's two host
    // result sets hold the batch in flight and that one)
    while (l.have_ready && !pending_.empty()) deliverOldest();
    // one call at a time per context: the lane's previous batch must have left the GPU (its result stays
    // valid in the other host result set and is handed to the sink below, while the new batch runs)
    if (l.flying >= 0) {
        l.ready = l.worker->get();
        l.ready_buf = l.flying;
        l.typical = l.seconds[l.flying];                  // the last batch (the first ones of a run allocate: no average)
        l.flying = -1;
        l.have_ready = true;
    }
    wn_batch_args a;
    a.frames = l.stage[buf]; a.frames_on_device = 0; a.n_frames = (int32_t)l.frnr[buf].size(); a.natoms = natoms_;
    a.box_stride = 0; a.boxes = nullptr; a.subset_offsets = nullptr; a.subset_sites = nullptr;
    a.result_layout = layout_; a.results_on_device = 0;
    if (!l.sub_off[buf].empty()) {
        a.subset_offsets = l.sub_off[buf].data(); a.subset_sites = l.sub_sites[buf].data();
        if (!l.boxes[buf].empty()) {              /* site lists: rectangular boxes only (checked in push) */
            l.b3.clear();
            for (size_t f = 0; 9 * f < l.boxes[buf].size(); ++f) {
                l.b3.push_back(l.boxes[buf][9 * f]); l.b3.push_back(l.boxes[buf][9 * f + 4]); l.b3.push_back(l.boxes[buf][9 * f + 8]);
            }
            a.box_stride = 3; a.boxes = l.b3.data();
        }
    } else if (!l.boxes[buf].empty()) {
        a.box_stride = 9; a.boxes = l.boxes[buf].data();
    }
    HydrogenBondEdges* eng = l.eng;
    const bool async = pipelined_ && !batch_hook_;
    if (async) {
        if (!l.worker) l.worker.reset(new BatchWorker(eng));
        l.worker->start(a, &l.seconds[buf]);
        l.t_start = steady_now();
        l.flying = buf;
        pending_.push_back(cur_);
        l.fill = 1 - buf;
        // everything older than the batch just submitted goes to the sink now, in frame order.  Round-robin: at most one
        // batch per GPU stays in flight behind it.  Dynamic choice: only what has already left its GPU is delivered here
        // (waiting for the oldest batch would idle the fast GPUs behind a slow one); a GPU's own older result is
        // delivered before that GPU is handed its next batch (top of submit / after the choice below), which bounds
        // what is in flight to the two staging buffers and two result sets per GPU.
        if (dynamic_) {
            while (!pending_.empty()) {
                const Lane& o = *lanes_[pending_.front()];
                if (!(o.have_ready || (o.flying >= 0 && o.worker && o.worker->finished()))) break;
                deliverOldest();
            }
        } else {
            while (pending_.size() > lanes_.size() || (!pending_.empty() && lanes_[pending_.front()]->have_ready)) deliverOldest();
        }
    } else {
        while (!pending_.empty()) deliverOldest();
        const auto t0 = std::chrono::steady_clock::now();
        l.ready = eng->calculateEx(a);
        l.seconds[buf] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        l.ready_buf = buf;
        l.have_ready = true;
        if (batch_hook_) batch_hook_(l.frnr[buf], l.ready);
        pending_.push_back(cur_);
        deliverOldest();
    }
    // Round-robin over the GPUs by default.  Handing the next batch to the GPU expected to be free first (chooseLane,
    // setDynamicGpuChoice) exists because GPUs behind a shared PCIe uplink are ~1.5x slower end to end: fine on 2 GPUs
    // and on the CPU stand-in with uneven speeds (tests/cpp/test_batcher_mock.cpp).  Its one 8-GPU run showed single
    // batch calls of 88 s -- since explained as the benchmark committing never-written (all-zero) staging buffers, i.e.
    // frames with every site at one point, not this logic -- but it has not been measured again,
    // so the static order measured at 2/4/8 GPUs stays the default.
    cur_ = dynamic_ && async ? chooseLane() : (cur_ + 1) % (int)lanes_.size();
    while (lanes_[cur_]->have_ready && !pending_.empty()) deliverOldest();
}

void FrameBatcher::flush()
{
    submit();                                   // the partial tail batch, if any
    while (!pending_.empty()) deliverOldest();
}

void FrameBatcher::finish()
{
    flush();
    // statistics rows of the whole run: per GPU a zero-padded table (its own rows only), summed with NCCL
    // when there are several GPUs -- every row is owned by exactly one GPU, so the sum is exact
    const size_t F = (size_t)frames_pushed_;
    if (lanes_.size() > 1 && F > 0) {
        std::vector<std::vector<wn_frame_stats> > tmp(lanes_.size());
        for (size_t g = 0; g < lanes_.size(); ++g) {
            tmp[g] = lanes_[g]->table;
            tmp[g].resize(F, wn_frame_stats{0, 0, 0, 0, 0});
        }
        std::vector<std::thread> ts;
        std::vector<std::exception_ptr> errs(lanes_.size());
        for (size_t g = 0; g < lanes_.size(); ++g)
            ts.emplace_back([&, g]() {
                try { lanes_[g]->eng->statsReduce(tmp[g].data(), (int64_t)F); } catch (...) { errs[g] = std::current_exception(); }
            });
        for (auto& t : ts) t.join();
        for (auto& e : errs) if (e) std::rethrow_exception(e);
        reduced_.swap(tmp[0]);
    } else {
        reduced_ = lanes_[0]->table;
        reduced_.resize(F, wn_frame_stats{0, 0, 0, 0, 0});
    }
}
