// hydrogen-bond-edges.cpp -- HydrogenBondEdges and FrameBatcher over the C ABI
// (see include/hydrogen-bond-edges.hpp).
#include "hydrogen-bond-edges.hpp"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

namespace {
[[noreturn]] void raise(const char* what, int rc, const wn_ctx* ctx)
{
    throw std::runtime_error(std::string(what) + " failed (status " + std::to_string(rc) + "): " +
                             wn_last_error(ctx));
}
}  // namespace

HydrogenBondEdges::HydrogenBondEdges(int device)
    : ctx_(nullptr), r_on_(5.5), r_off_(6.5), a_on_(0.25), a_off_(0.0)
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
    HydrogenBondBatch b;
    b.edges = r.edges; b.offsets = r.frame_offsets; b.stats = r.stats;
    b.n_frames = r.n_frames; b.n_edges = r.n_edges; b.n_pair_evals = r.n_pair_evals; b.ties = r.ties;
    return b;
}

HydrogenBondBatch HydrogenBondEdges::calculate(const float* frames, const float* boxes, int n_frames, int natoms,
                                               const int64_t* subset_offsets, const int32_t* subset_sites)
{
    wn_batch_result r;
    const int rc = wn_hbond_batch_subset(ctx_, frames, boxes, n_frames, natoms, subset_offsets, subset_sites, &r);
    if (rc != WN_OK) raise("HydrogenBondEdges::calculate: wn_hbond_batch_subset", rc, ctx_);
    HydrogenBondBatch b;
    b.edges = r.edges; b.offsets = r.frame_offsets; b.stats = r.stats;
    b.n_frames = r.n_frames; b.n_edges = r.n_edges; b.n_pair_evals = r.n_pair_evals; b.ties = r.ties;
    return b;
}

HydrogenBondBatch HydrogenBondEdges::calculateBox9(const float* frames, const float* boxes9, int n_frames, int natoms)
{
    wn_batch_result r;
    const int rc = wn_hbond_batch_box9(ctx_, frames, boxes9, n_frames, natoms, &r);
    if (rc != WN_OK) raise("HydrogenBondEdges::calculateBox9: wn_hbond_batch_box9", rc, ctx_);
    HydrogenBondBatch b;
    b.edges = r.edges; b.offsets = r.frame_offsets; b.stats = r.stats;
    b.n_frames = r.n_frames; b.n_edges = r.n_edges; b.n_pair_evals = r.n_pair_evals; b.ties = r.ties;
    return b;
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
FrameBatcher::FrameBatcher(HydrogenBondEdges& engine, int natoms, int batch_frames, Sink sink)
    : engine_(engine), natoms_(natoms), batch_frames_(std::max(batch_frames, 1)), sink_(sink),
      frames_done_(0)
{
    buf_.reserve((size_t)batch_frames_ * natoms_ * 3);
}

void FrameBatcher::push(int frnr, const float* xyz)
{
    buf_.insert(buf_.end(), xyz, xyz + (size_t)natoms_ * 3);
    frnr_.push_back(frnr);
    if ((int)frnr_.size() >= batch_frames_) flush();
}

void FrameBatcher::push(int frnr, const float* xyz, const float box[3][3])
{
    if (!frnr_.empty() && boxes_.size() != 9 * frnr_.size())
        throw std::invalid_argument("FrameBatcher::push: periodic and open frames mixed in one batch");
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) boxes_.push_back(box[a][b]);
    push(frnr, xyz);
}

void FrameBatcher::push(int frnr, const float* xyz, const float (*box)[3], const std::vector<int32_t>& frameSites)
{
    if (!frnr_.empty() && sub_off_.size() != frnr_.size() + 1)
        throw std::invalid_argument("FrameBatcher::push: frames with and without site lists mixed in one batch");
    if (sub_off_.empty()) sub_off_.push_back(0);
    sub_sites_.insert(sub_sites_.end(), frameSites.begin(), frameSites.end());
    sub_off_.push_back((int64_t)sub_sites_.size());
    if (box) {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                if (a != b && box[a][b] != 0.0f)
                    throw std::invalid_argument("FrameBatcher::push: per-frame site lists need a rectangular box");
                boxes_.push_back(box[a][b]);
            }
    }
    push(frnr, xyz);
}

void FrameBatcher::finish() { flush(); }

void FrameBatcher::flush()
{
    if (frnr_.empty()) return;
    HydrogenBondBatch b;
    if (sub_off_.empty()) {
        b = boxes_.empty() ? engine_.calculate(buf_.data(), (int)frnr_.size(), natoms_)
                           : engine_.calculateBox9(buf_.data(), boxes_.data(), (int)frnr_.size(), natoms_);
    } else {
        std::vector<float> b3;                    /* site lists: rectangular boxes only (checked in push) */
        for (size_t f = 0; 9 * f < boxes_.size(); ++f) { b3.push_back(boxes_[9 * f]); b3.push_back(boxes_[9 * f + 4]); b3.push_back(boxes_[9 * f + 8]); }
        b = engine_.calculate(buf_.data(), b3.empty() ? nullptr : b3.data(), (int)frnr_.size(), natoms_, sub_off_.data(),
                              sub_sites_.data());
    }
    if (batch_hook_) batch_hook_(frnr_, b);
    for (int f = 0; f < b.n_frames; ++f)
        sink_(frnr_[f], b.edges + b.offsets[f], (size_t)(b.offsets[f + 1] - b.offsets[f]), b.stats[f]);
    frames_done_ += (uint64_t)b.n_frames;
    buf_.clear();
    boxes_.clear();
    sub_off_.clear();
    sub_sites_.clear();
    frnr_.clear();
}
