// wn_mock_abi.cpp -- TEST INFRASTRUCTURE ONLY: a CPU stand-in for the C ABI of include/wn_b200.h.
//
// The C++ host mirror (CudaFindPairs, HydrogenBondEdges, FrameBatcher, the standalone driver) is host logic over
// the C ABI: batching, two staging buffers and two result sets per GPU, one worker thread per GPU, frame-order
// delivery, the statistics table summed over the GPUs.  On the CPU-only build box none of it can run against
// libwn_b200.so (no device: every entry point fails loudly, as it must).  This file implements the entry points
// the host mirror calls with the test oracle (oracle/wn_oracle.c) behind them, so that tests/ can run the host
// logic -- including the several-GPU paths, with artificial per-"device" delays, under ThreadSanitizer -- on a CPU.
//
// It is linked ONLY into test binaries built by tests/mock/Makefile (never into libwn_host.so, the driver or the
// bench that ship), it is not a fallback of the product, and nothing measured or shipped touches it
// (tests/test_host_cpu.py::test_product_never_touches_the_oracle covers the product tree).
//
// A contract check the real library cannot afford is made here: a context is single-caller (two threads inside
// one context at once abort the test).  A batch result is only valid until the next call (or the one after it
// with two result sets): the result buffers are re-assigned by the call that expires them, so a host that reads
// too late trips AddressSanitizer / reads another batch and fails the comparison.
#include "wn_b200.h"
#include "wn_mock.h"

#include "../../oracle/wn_oracle.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Comm {
    int nranks = 0;
    std::mutex m;
    std::condition_variable cv;
    std::vector<double> sum;
    int arrived = 0, left = 0;
    uint64_t generation = 0;
};

std::mutex g_comm_m;
std::map<std::string, std::shared_ptr<Comm> > g_comms;
std::atomic<uint64_t> g_next_id{1};
std::atomic<long> g_delay_us[64];
std::atomic<long> g_calls[64];
std::atomic<int> g_fail_device{-1};
std::atomic<long> g_fail_after{0};
std::atomic<int> g_alloc_fail_device{-1};
std::atomic<long> g_alloc_fail_after{0};
thread_local std::string t_create_error;

struct ResultSet {
    std::vector<unsigned char> edges;     // records of the call's layout
    std::vector<uint64_t> off;
    std::vector<wn_frame_stats> stats;
    std::vector<uint32_t> row_off;
};

}  // namespace

struct wn_ctx {
    int device = 0;
    int n_sites = -1, hmax = 1, first_limit = 0;
    std::vector<int32_t> site_atom, h_atom;
    std::vector<uint8_t> site_type;
    double r_on = 5.5, r_off = 6.5, t_on = 0.25, t_off = 0.0;
    int n_sets = 1, cur = 0;
    ResultSet set[2];
    // last batch, for wn_format_edges / wn_components
    int last_frames = -1, last_layout = WN_LAYOUT_EDGES, last_set = 0;
    bool last_subset = false;
    std::vector<int32_t> pairs;                 // wn_find_pairs result
    std::vector<wn_edge> me;                    // wn_make_edges result
    std::string text;                           // wn_format_edges result
    std::vector<int32_t> labels;
    std::vector<wn_component_stats> cstats;
    std::string err;
    std::shared_ptr<Comm> comm;
    int rank = 0;
    std::atomic<int> inside{0};
    wn_timings tm{};
};

namespace {

int fail(wn_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }

// single-caller contract of a context
struct Enter {
    wn_ctx* c;
    explicit Enter(wn_ctx* ctx) : c(ctx)
    {
        if (c && c->inside.fetch_add(1) != 0) {
            std::fprintf(stderr, "wn_mock: two callers inside one context (device %d) at the same time\n", c->device);
            std::abort();
        }
    }
    ~Enter() { if (c) c->inside.fetch_sub(1); }
};

bool default_params(const wn_ctx* c) { return c->r_on == 5.5 && c->r_off == 6.5 && c->t_on == 0.25 && c->t_off == 0.0; }

void re_energy(const wn_ctx* c, wno_edge* e, size_t n)
{
    if (default_params(c)) return;
    for (size_t k = 0; k < n; ++k)
        e[k].energy = (float)wno_compute_energy_ex((double)e[k].length, (double)e[k].cosine, c->r_on, c->r_off, c->t_on, c->t_off);
}

int batch(wn_ctx* c, const wn_batch_args* a, wn_batch_result* out)
{
    if (!c) return WN_ERR_BAD_ARG;
    Enter guard(c);
    if (!a || !out || a->n_frames < 0 || a->natoms < 0 || (!a->frames && a->n_frames > 0 && a->natoms > 0))
        return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch: bad argument");
    if (a->frames_on_device || a->results_on_device) return fail(c, WN_ERR_BAD_ARG, "wn_mock: no device memory in the CPU stand-in");
    if (c->n_sites < 0) return fail(c, WN_ERR_NO_SITES, "wn_hbond_batch: call wn_set_sites first");
    if (a->box_stride != 0 && a->box_stride != 3 && a->box_stride != 9) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: box_stride");
    if ((a->box_stride != 0) != (a->boxes != nullptr)) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: boxes / box_stride");
    if (a->result_layout < WN_LAYOUT_EDGES || a->result_layout > WN_LAYOUT_CSR12) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_ex: result_layout");
    for (int s = 0; s < c->n_sites; ++s)
        if (c->site_atom[(size_t)s] >= a->natoms) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch: site table refers to atoms beyond natoms");
    const long call = g_calls[c->device & 63].fetch_add(1);
    if (g_fail_device.load() == c->device && call >= g_fail_after.load())
        return fail(c, WN_ERR_CUDA, "wn_mock: injected failure");
    c->err.clear();
    if (c->n_sets == 2) c->cur ^= 1;
    ResultSet& R = c->set[c->cur];
    const int F = a->n_frames, N = c->n_sites;
    const size_t esz = a->result_layout == WN_LAYOUT_CSR16 ? sizeof(wn_edge16) : a->result_layout == WN_LAYOUT_CSR12 ? sizeof(wn_edge12) : sizeof(wn_edge);
    R.edges.clear(); R.off.assign((size_t)F + 1, 0); R.stats.assign((size_t)std::max(F, 1), wn_frame_stats{0, 0, 0, 0, 0});
    R.row_off.assign((size_t)F * (size_t)std::max(N, 0) + 1, 0);
    std::memset(out, 0, sizeof *out);
    wno_ties ties{0, 0, 0, 0, 0};
    uint64_t evals_all = 0;
    std::vector<wno_edge> tmp;
    std::vector<int32_t> sa, ha; std::vector<uint8_t> ty;
    const auto t0 = std::chrono::steady_clock::now();
    for (int f = 0; f < F; ++f) {
        const float* frame = a->frames + (size_t)f * a->natoms * 3;
        const int32_t* psa = c->site_atom.data(); const int32_t* pha = c->h_atom.data(); const uint8_t* pty = c->site_type.data();
        int n = N, fl = c->first_limit;
        if (a->subset_offsets) {
            const int64_t b = a->subset_offsets[f], e = a->subset_offsets[f + 1];
            if (b < 0 || e < b || e - b > N) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_subset: bad offsets");
            n = (int)(e - b); fl = n;
            sa.resize((size_t)n); ha.resize((size_t)n * c->hmax); ty.resize((size_t)n);
            for (int k = 0; k < n; ++k) {
                const int32_t s = a->subset_sites[b + k];
                if (s < 0 || s >= N) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_subset: site outside the table");
                sa[(size_t)k] = c->site_atom[(size_t)s]; ty[(size_t)k] = c->site_type[(size_t)s];
                for (int h = 0; h < c->hmax; ++h) ha[(size_t)k * c->hmax + h] = c->h_atom[(size_t)s * c->hmax + h];
            }
            psa = sa.data(); pha = ha.data(); pty = ty.data();
        }
        size_t cap = (size_t)n * (size_t)std::max(n - 1, 0) / 2 + 1;
        tmp.resize(cap);
        double st[4] = {0, 0, 0, 0};
        uint64_t np = 0, ev = 0;
        size_t ne;
        if (a->box_stride == 3) {
            const float* b = a->boxes + 3 * (size_t)f;
            if (!(b[0] > 0.f && b[1] > 0.f && b[2] > 0.f)) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_pbc: box lengths must be positive and finite");
            ne = wno_frame_hbonds_pbc(frame, a->natoms, psa, pha, c->hmax, pty, n, fl, b, tmp.data(), cap, st, &np, &ev, &ties);
        } else if (a->box_stride == 9) {
            // a call whose frames are all rectangular uses the rectangular definition (wn_b200.h)
            bool rect = true;
            for (int g = 0; g < F && rect; ++g) {
                const float* b = a->boxes + 9 * (size_t)g;
                rect = b[1] == 0.f && b[2] == 0.f && b[3] == 0.f && b[5] == 0.f && b[6] == 0.f && b[7] == 0.f;
            }
            const float* b = a->boxes + 9 * (size_t)f;
            if (!(b[0] > 0.f && b[4] > 0.f && b[8] > 0.f) || b[1] != 0.f || b[2] != 0.f || b[5] != 0.f)
                return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_box9: box must be lower-triangular with a positive diagonal (GROMACS convention)");
            if (rect) {
                const float b3[3] = {b[0], b[4], b[8]};
                ne = wno_frame_hbonds_pbc(frame, a->natoms, psa, pha, c->hmax, pty, n, fl, b3, tmp.data(), cap, st, &np, &ev, &ties);
            } else {
                ne = wno_frame_hbonds_tric(frame, a->natoms, psa, pha, c->hmax, pty, n, fl, b, tmp.data(), cap, st, &np, &ev, &ties);
            }
        } else {
            ne = wno_frame_hbonds(frame, a->natoms, psa, pha, c->hmax, pty, n, fl, tmp.data(), cap, st, &np, &ev, &ties);
        }
        if (ne == (size_t)-1) return fail(c, WN_ERR_OVERFLOW, "wn_mock: oracle capacity");
        re_energy(c, tmp.data(), ne);
        double sum_e = 0.0;
        for (size_t k = 0; k < ne; ++k) sum_e += (double)tmp[k].energy;
        const size_t base = R.edges.size();
        R.edges.resize(base + ne * esz);
        // row offsets of the frame (stride N, rows past the frame's list end at the frame's edge count)
        {
            size_t k = 0;
            for (int i = 0; i < N; ++i) {
                while (k < ne && tmp[k].i < i) ++k;
                R.row_off[(size_t)f * N + i] = (uint32_t)k;
            }
        }
        for (size_t k = 0; k < ne; ++k) {
            unsigned char* p = R.edges.data() + base + k * esz;
            if (a->result_layout == WN_LAYOUT_CSR16) { wn_edge16 r{tmp[k].j, tmp[k].energy, tmp[k].length, tmp[k].cosine}; std::memcpy(p, &r, esz); }
            else if (a->result_layout == WN_LAYOUT_CSR12) { wn_edge12 r{tmp[k].j, tmp[k].length, tmp[k].cosine}; std::memcpy(p, &r, esz); }
            else { wn_edge r{tmp[k].i, tmp[k].j, tmp[k].energy, tmp[k].length, tmp[k].cosine}; std::memcpy(p, &r, esz); }
        }
        R.off[(size_t)f + 1] = R.off[(size_t)f] + ne;
        R.stats[(size_t)f] = wn_frame_stats{st[0], st[1], st[2], sum_e, (double)ev};
        evals_all += ev;
    }
    // a "device" that is slower than the others: sleep per frame
    const long us = g_delay_us[c->device & 63].load();
    if (us > 0 && F > 0) std::this_thread::sleep_for(std::chrono::microseconds(us * F));
    c->tm = wn_timings{};
    c->tm.total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->tm.passes = 1;
    out->edges = reinterpret_cast<const wn_edge*>(R.edges.data());
    out->frame_offsets = R.off.data();
    out->stats = R.stats.data();
    out->n_edges = R.off[(size_t)F];
    out->n_pair_evals = evals_all;
    out->ties = wn_tie_report{ties.pair_cutoff_equal, ties.length_equal_cut, ties.cosine_argmax_equal, ties.pos_zero, ties.pair_cutoff_near};
    out->n_frames = F; out->on_device = 0;
    out->row_offsets = R.row_off.data(); out->n_sites = N; out->layout = a->result_layout;
    c->last_frames = F; c->last_layout = a->result_layout; c->last_set = c->cur; c->last_subset = a->subset_offsets != nullptr;
    return WN_OK;
}

int shorthand(wn_ctx* c, const float* frames, const float* boxes, int stride, int nf, int natoms, const int64_t* so,
              const int32_t* ss, wn_batch_result* out)
{
    wn_batch_args a;
    a.frames = frames; a.frames_on_device = 0; a.n_frames = nf; a.natoms = natoms; a.box_stride = boxes ? stride : 0; a.boxes = boxes;
    a.subset_offsets = so; a.subset_sites = ss; a.result_layout = WN_LAYOUT_EDGES; a.results_on_device = 0;
    return batch(c, &a, out);
}

}  // namespace

extern "C" {

int wn_create(int device, wn_ctx** out)
{
    if (!out) return WN_ERR_BAD_ARG;
    *out = nullptr;
    if (device < 0 || device >= wn_device_count()) { t_create_error = "wn_mock: no such device"; return WN_ERR_NO_DEVICE; }
    wn_ctx* c = new wn_ctx();
    c->device = device;
    *out = c;
    return WN_OK;
}
void wn_destroy(wn_ctx* c) { delete c; }
const char* wn_last_error(const wn_ctx* c) { return c ? c->err.c_str() : t_create_error.c_str(); }
const char* wn_version(void) { return "wn_mock (CPU stand-in for tests; not wn_b200)"; }
int wn_device_count(void)
{
    const char* e = std::getenv("WN_MOCK_DEVICES");
    const int n = e ? std::atoi(e) : 8;
    return n < 0 ? 0 : (n > 64 ? 64 : n);
}

int wn_set_sites(wn_ctx* c, int32_t n, const int32_t* site_atom, const int32_t* h_atom, int32_t hmax, const uint8_t* type, int32_t first_limit)
{
    if (!c) return WN_ERR_BAD_ARG;
    Enter guard(c);
    if (n < 0 || hmax < 1 || (n > 0 && (!site_atom || !h_atom || !type))) return fail(c, WN_ERR_BAD_ARG, "wn_set_sites: bad argument");
    c->site_atom.assign(site_atom, site_atom + n);
    c->h_atom.assign(h_atom, h_atom + (size_t)n * hmax);
    c->site_type.assign(type, type + n);
    for (int s = 0; s < n; ++s)
        if (site_atom[s] < 0 || type[s] > WN_WATER) return fail(c, WN_ERR_BAD_ARG, "wn_set_sites: bad table");
    c->n_sites = n; c->hmax = hmax; c->first_limit = (first_limit < 0 || first_limit > n) ? n : first_limit;
    c->last_frames = -1;
    return WN_OK;
}
int wn_set_water_sites(wn_ctx* c, int32_t n_waters, int32_t first_limit)
{
    if (!c || n_waters < 0) return WN_ERR_BAD_ARG;
    std::vector<int32_t> sa((size_t)n_waters), ha((size_t)n_waters * 2);
    std::vector<uint8_t> ty((size_t)n_waters);
    wno_water_site_table(n_waters, sa.data(), ha.data(), ty.data());
    return wn_set_sites(c, n_waters, sa.data(), ha.data(), 2, ty.data(), first_limit);
}
int wn_set_energy_params(wn_ctx* c, double r_on, double r_off, double t_on, double t_off)
{
    if (!c) return WN_ERR_BAD_ARG;
    Enter guard(c);
    c->r_on = r_on; c->r_off = r_off; c->t_on = t_on; c->t_off = t_off;
    return WN_OK;
}
int wn_set_result_sets(wn_ctx* c, int32_t n)
{
    if (!c || (n != 1 && n != 2)) return c ? fail(c, WN_ERR_BAD_ARG, "wn_set_result_sets: 1 or 2") : WN_ERR_BAD_ARG;
    Enter guard(c);
    c->n_sets = n; if (n == 1) c->cur = 0;
    return WN_OK;
}
int wn_set_batch_frames(wn_ctx* c, int32_t) { return c ? WN_OK : WN_ERR_BAD_ARG; }
int wn_bind_thread_to_device(wn_ctx* c, int32_t* n_cpus) { if (!c) return WN_ERR_BAD_ARG; if (n_cpus) *n_cpus = 0; return WN_OK; }

int wn_hbond_batch_ex(wn_ctx* c, const wn_batch_args* a, wn_batch_result* out) { return batch(c, a, out); }
int wn_hbond_batch(wn_ctx* c, const float* fr, int32_t nf, int32_t natoms, wn_batch_result* out) { return shorthand(c, fr, nullptr, 0, nf, natoms, nullptr, nullptr, out); }
int wn_hbond_batch_pbc(wn_ctx* c, const float* fr, const float* boxes, int32_t nf, int32_t natoms, wn_batch_result* out)
{
    if (c && !boxes) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_pbc: boxes is NULL");
    return shorthand(c, fr, boxes, 3, nf, natoms, nullptr, nullptr, out);
}
int wn_hbond_batch_box9(wn_ctx* c, const float* fr, const float* boxes9, int32_t nf, int32_t natoms, wn_batch_result* out)
{
    if (c && !boxes9) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_box9: boxes is NULL");
    return shorthand(c, fr, boxes9, 9, nf, natoms, nullptr, nullptr, out);
}
int wn_hbond_batch_subset(wn_ctx* c, const float* fr, const float* boxes, int32_t nf, int32_t natoms, const int64_t* so, const int32_t* ss, wn_batch_result* out)
{
    if (c && (!so || (!ss && nf > 0 && so[nf] > 0))) return fail(c, WN_ERR_BAD_ARG, "wn_hbond_batch_subset: bad argument");
    return shorthand(c, fr, boxes, 3, nf, natoms, so, ss, out);
}

static int find_common(wn_ctx* c, const double* xyz, int32_t n, int32_t fl, const float* box3, const int32_t** pairs, uint64_t* m, wn_tie_report* ties)
{
    if (!c) return WN_ERR_BAD_ARG;
    Enter guard(c);
    if (n < 0 || (n > 0 && !xyz) || !pairs || !m) return fail(c, WN_ERR_BAD_ARG, "wn_find_pairs: bad argument");
    if (fl < 0 || fl > n) fl = n;
    wno_ties t{0, 0, 0, 0, 0};
    const size_t cnt = box3 ? wno_brute_find_pairs_pbc(xyz, n, fl, box3, nullptr, 0, nullptr) : wno_brute_find_pairs_limited(xyz, n, fl, nullptr, 0, nullptr);
    c->pairs.resize(2 * cnt + 2);
    if (box3) wno_brute_find_pairs_pbc(xyz, n, fl, box3, c->pairs.data(), cnt, &t);
    else wno_brute_find_pairs_limited(xyz, n, fl, c->pairs.data(), cnt, &t);
    *pairs = c->pairs.data(); *m = cnt;
    if (ties) *ties = wn_tie_report{t.pair_cutoff_equal, t.length_equal_cut, t.cosine_argmax_equal, t.pos_zero, t.pair_cutoff_near};
    return WN_OK;
}
int wn_find_pairs(wn_ctx* c, const double* xyz, int32_t n, int32_t fl, const int32_t** pairs, uint64_t* m, wn_tie_report* ties) { return find_common(c, xyz, n, fl, nullptr, pairs, m, ties); }
int wn_find_pairs_pbc(wn_ctx* c, const double* xyz, int32_t n, int32_t fl, const float box[3], const int32_t** pairs, uint64_t* m, wn_tie_report* ties)
{
    if (c && !box) return fail(c, WN_ERR_BAD_ARG, "wn_find_pairs_pbc: box is NULL");
    return find_common(c, xyz, n, fl, box, pairs, m, ties);
}

int wn_make_edges(wn_ctx* c, const float* frame, int32_t natoms, const float* box9, const int32_t* pairs, uint64_t m,
                  const wn_edge** edges, uint64_t* n_edges, uint64_t* n_pair_evals, wn_tie_report* ties)
{
    if (!c) return WN_ERR_BAD_ARG;
    Enter guard(c);
    if (!edges || !n_edges || (m > 0 && (!pairs || !frame))) return fail(c, WN_ERR_BAD_ARG, "wn_make_edges: bad argument");
    if (c->n_sites < 0) return fail(c, WN_ERR_NO_SITES, "wn_make_edges: call wn_set_sites first");
    if (box9) return fail(c, WN_ERR_BAD_ARG, "wn_mock: periodic wn_make_edges is not in the CPU stand-in");
    const int N = c->n_sites;
    for (uint64_t k = 0; k < 2 * m; ++k)
        if (pairs[k] < 0 || pairs[k] >= N) return fail(c, WN_ERR_BAD_ARG, "wn_make_edges: pair index outside the table");
    std::vector<double> sx((size_t)std::max(N, 1) * 3), hx((size_t)std::max(N, 1) * c->hmax * 3);
    std::vector<int32_t> hoff((size_t)N + 1);
    wno_gather_sites(frame, natoms, c->site_atom.data(), c->h_atom.data(), c->hmax, N, sx.data(), hx.data(), hoff.data());
    std::vector<wno_edge> tmp((size_t)m + 1);
    uint64_t ev = 0; wno_ties t{0, 0, 0, 0, 0};
    const size_t ne = wno_make_edges(pairs, (size_t)m, sx.data(), N, hx.data(), hoff.data(), c->site_type.data(), tmp.data(), &ev, &t);
    re_energy(c, tmp.data(), ne);
    c->me.resize(ne + 1);
    for (size_t k = 0; k < ne; ++k) c->me[k] = wn_edge{tmp[k].i, tmp[k].j, tmp[k].energy, tmp[k].length, tmp[k].cosine};
    *edges = c->me.data(); *n_edges = ne;
    if (n_pair_evals) *n_pair_evals = ev;
    if (ties) *ties = wn_tie_report{t.pair_cutoff_equal, t.length_equal_cut, t.cosine_argmax_equal, t.pos_zero, t.pair_cutoff_near};
    c->last_frames = -1;
    return WN_OK;
}

int wn_format_edges(wn_ctx* c, const int32_t* frame_numbers, const uint32_t* ids, int32_t n_ids, const char** text, uint64_t* bytes)
{
    if (!c || !text || !bytes) return WN_ERR_BAD_ARG;
    Enter guard(c);
    if (c->last_frames < 0) return fail(c, WN_ERR_BAD_ARG, "wn_format_edges: no batch result to format");
    if (c->last_layout != WN_LAYOUT_EDGES) return fail(c, WN_ERR_BAD_ARG, "wn_format_edges: needs a WN_LAYOUT_EDGES result");
    if (c->last_subset && ids) return fail(c, WN_ERR_BAD_ARG, "wn_format_edges: the last batch used per-frame site lists");
    if (c->last_frames > 0 && !frame_numbers) return fail(c, WN_ERR_BAD_ARG, "wn_format_edges: frame_numbers is NULL");
    if (ids && n_ids < c->n_sites) return fail(c, WN_ERR_BAD_ARG, "wn_format_edges: site_ids shorter than the site table");
    const ResultSet& R = c->set[c->last_set];
    const wn_edge* e = reinterpret_cast<const wn_edge*>(R.edges.data());
    c->text.clear();
    char row[128];
    for (int f = 0; f < c->last_frames; ++f) {
        std::snprintf(row, sizeof row, "#--- frame %d ---\n", frame_numbers[f]);
        c->text += row;
        for (uint64_t k = R.off[(size_t)f]; k < R.off[(size_t)f + 1]; ++k) {
            const unsigned a = ids ? ids[e[k].i] : (unsigned)e[k].i, b = ids ? ids[e[k].j] : (unsigned)e[k].j;
            const int len = std::snprintf(row, sizeof row, "%6u%6u%8.4f%8.4f%8.4f\n", a, b, (double)e[k].energy, (double)e[k].length, (double)e[k].cosine);
            if (len != 37) return WN_ERR_OVERFLOW;
            c->text += row;
        }
    }
    *text = c->text.data(); *bytes = c->text.size();
    return WN_OK;
}

int wn_components(wn_ctx* c, float thr, const int32_t** labels, const wn_component_stats** stats)
{
    if (!c || !labels || !stats) return WN_ERR_BAD_ARG;
    Enter guard(c);
    if (c->last_frames < 0) return fail(c, WN_ERR_BAD_ARG, "wn_components: no batch result to analyse");
    if (c->last_layout != WN_LAYOUT_EDGES) return fail(c, WN_ERR_BAD_ARG, "wn_mock: components need a WN_LAYOUT_EDGES result in the CPU stand-in");
    const ResultSet& R = c->set[c->last_set];
    const int N = c->n_sites, F = c->last_frames;
    c->labels.assign((size_t)F * std::max(N, 1), -1);
    c->cstats.assign((size_t)std::max(F, 1), wn_component_stats{0, 0, 0, 0});
    for (int f = 0; f < F; ++f) {
        int32_t st[4];
        wno_components(reinterpret_cast<const wno_edge*>(R.edges.data()) + R.off[(size_t)f], (size_t)(R.off[(size_t)f + 1] - R.off[(size_t)f]), N, thr,
                       c->labels.data() + (size_t)f * N, st);
        c->cstats[(size_t)f] = wn_component_stats{st[0], st[1], st[2], st[3]};
    }
    *labels = c->labels.data(); *stats = c->cstats.data();
    return WN_OK;
}

int wn_host_alloc_pinned(wn_ctx* c, size_t bytes, void** p)
{
    if (!c || !p) return WN_ERR_BAD_ARG;
    *p = nullptr;
    if (g_alloc_fail_device.load() == c->device && g_alloc_fail_after.fetch_sub(1) <= 0)
        return fail(c, WN_ERR_OOM, "wn_mock: injected pinned allocation failure");
    *p = std::malloc(bytes ? bytes : 1);
    if (!*p) return fail(c, WN_ERR_OOM, "wn_mock: malloc");
    std::memset(*p, 0xCD, bytes);
    return WN_OK;
}
int wn_host_free_pinned(wn_ctx* c, void* p) { if (!c) return WN_ERR_BAD_ARG; std::free(p); return WN_OK; }
int wn_get_timings(wn_ctx* c, wn_timings* out) { if (!c || !out) return WN_ERR_BAD_ARG; *out = c->tm; return WN_OK; }

int wn_nccl_unique_id(void* id128)
{
    if (!id128) return WN_ERR_BAD_ARG;
    std::memset(id128, 0, 128);
    const uint64_t v = g_next_id.fetch_add(1);
    std::memcpy(id128, &v, sizeof v);
    return WN_OK;
}
int wn_nccl_init(wn_ctx* c, const void* id128, int32_t rank, int32_t nranks)
{
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return c ? fail(c, WN_ERR_BAD_ARG, "wn_nccl_init: bad argument") : WN_ERR_BAD_ARG;
    const std::string key(static_cast<const char*>(id128), 128);
    std::lock_guard<std::mutex> g(g_comm_m);
    std::shared_ptr<Comm>& cm = g_comms[key];
    if (!cm) { cm.reset(new Comm()); cm->nranks = nranks; }
    if (cm->nranks != nranks) return fail(c, WN_ERR_NCCL, "wn_mock: ranks disagree on the communicator size");
    c->comm = cm; c->rank = rank;
    return WN_OK;
}
int wn_stats_reduce(wn_ctx* c, wn_frame_stats* table, int64_t n)
{
    if (!c || n < 0 || (n > 0 && !table)) return c ? fail(c, WN_ERR_BAD_ARG, "wn_stats_reduce: bad argument") : WN_ERR_BAD_ARG;
    Enter guard(c);
    if (!c->comm) return fail(c, WN_ERR_NCCL, "wn_stats_reduce: communicator not initialised (wn_nccl_init)");
    Comm& cm = *c->comm;
    const size_t words = (size_t)n * 5;
    std::unique_lock<std::mutex> g(cm.m);
    cm.cv.wait(g, [&]() { return cm.left == 0; });          // every rank has copied the previous collective's sum out
    if (cm.arrived == 0) cm.sum.assign(words, 0.0);
    if (cm.sum.size() != words) return fail(c, WN_ERR_NCCL, "wn_mock: ranks disagree on the table size");
    const double* src = reinterpret_cast<const double*>(table);
    for (size_t k = 0; k < words; ++k) cm.sum[k] += src[k];
    const uint64_t gen = cm.generation;
    if (++cm.arrived == cm.nranks) { cm.arrived = 0; cm.left = cm.nranks; ++cm.generation; cm.cv.notify_all(); }
    else cm.cv.wait(g, [&]() { return cm.generation != gen; });
    std::memcpy(table, cm.sum.data(), words * sizeof(double));
    if (--cm.left == 0) cm.cv.notify_all();
    return WN_OK;
}

// ---- controls of the stand-in (tests/mock/wn_mock.h)
void wn_mock_set_delay_us(int device, long us_per_frame) { g_delay_us[device & 63].store(us_per_frame); }
long wn_mock_batch_calls(int device) { return g_calls[device & 63].load(); }
void wn_mock_fail_batches(int device, long after_calls) { g_fail_after.store(after_calls + g_calls[device < 0 ? 0 : (device & 63)].load()); g_fail_device.store(device); }

void wn_mock_fail_pinned_alloc(int device, long after_allocs) { g_alloc_fail_after.store(after_allocs); g_alloc_fail_device.store(device); }

}  // extern "C"
