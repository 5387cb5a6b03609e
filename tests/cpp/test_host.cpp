// test_host.cpp -- the C++ host mirror on a GPU, written like the reference's caller:
// a FindPairs pointer that happens to be a CudaFindPairs, and a frame loop that
// pushes frames into the batcher the way analyzeFrame would.  Checked against an
// O(N^2) loop in this file (for pairs) and against self-consistency (batched ==
// unbatched); bit-exact parity with the oracle is covered by the pytest suite.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <random>
#include <stdexcept>

#include "cuda-find-pairs.hpp"
#include "hydrogen-bond-edges.hpp"
#include "edge-file-writer.hpp"
#include "wn_energy.h"
#include <algorithm>
#include <cstring>
#include <sstream>

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); return 1; } } while (0)

int main()
{
    try {
        std::mt19937 rng(7);
        std::uniform_real_distribution<float> u(0.f, 3.f);
        // ---- FindPairs drop-in
        std::shared_ptr<FindPairs> mp_ = std::make_shared<CudaFindPairs>(0);
        std::vector<Point> pts;
        for (int i = 0; i < 500; ++i) pts.push_back(Point(u(rng), u(rng), u(rng)));
        std::vector<Point> before = pts;
        std::vector<std::pair<int, int>> pairs = mp_->find(pts);
        std::vector<std::pair<int, int>> want;
        for (unsigned i = 0; i < pts.size(); ++i)
            for (unsigned j = i + 1; j < pts.size(); ++j) {
                const double dx = pts[i].x() - pts[j].x(), dy = pts[i].y() - pts[j].y(), dz = pts[i].z() - pts[j].z();
                if (10.0 * (dx * dx + dy * dy + dz * dz) < 42.25) want.push_back(std::make_pair((int)i, (int)j));
            }
        CHECK(pairs == want);
        for (size_t i = 0; i < pts.size(); ++i) CHECK(pts[i].x() == before[i].x());   // input not reordered
        std::vector<Point> none;
        CHECK(mp_->find(none).empty());

        // ---- batched edges: FrameBatcher(3 per batch) == one calculate() over all frames
        const int n_waters = 300, natoms = 900, n_frames = 7;
        std::vector<float> frames((size_t)n_frames * natoms * 3);
        std::uniform_real_distribution<float> ub(0.f, 2.1f), uh(-0.06f, 0.06f);
        for (int f = 0; f < n_frames; ++f)
            for (int w = 0; w < n_waters; ++w) {
                float* p = &frames[((size_t)f * natoms + 3 * w) * 3];
                for (int c = 0; c < 3; ++c) { p[c] = ub(rng); p[3 + c] = p[c] + uh(rng); p[6 + c] = p[c] + uh(rng); }
            }
        HydrogenBondEdges hb(0);
        hb.setRadiusShift(5.5f, 6.5f);
        hb.setAngleShift(0.25f, 0.0f);
        hb.setWaterSites(n_waters);
        HydrogenBondBatch all = hb.calculate(frames.data(), n_frames, natoms);
        CHECK(all.n_frames == n_frames && all.n_edges > 0);
        std::vector<wn_edge> ref_edges(all.edges, all.edges + all.n_edges);
        std::vector<uint64_t> ref_off(all.offsets, all.offsets + n_frames + 1);
        std::vector<wn_edge> got;
        std::vector<int> order;
        FrameBatcher batcher(hb, natoms, 3, [&](int frnr, const wn_edge* e, size_t n, const wn_frame_stats& st) {
            order.push_back(frnr);
            if (st.num_edges != (double)n || st.num_sites != n_waters) throw std::runtime_error("stats mismatch");
            got.insert(got.end(), e, e + n);
        });
        for (int f = 0; f < n_frames; ++f) batcher.push(100 + f, &frames[(size_t)f * natoms * 3]);
        batcher.finish();
        CHECK(batcher.framesDone() == (uint64_t)n_frames);
        CHECK(order.size() == (size_t)n_frames);
        for (int f = 0; f < n_frames; ++f) CHECK(order[f] == 100 + f);
        CHECK(got.size() == ref_edges.size());
        for (size_t k = 0; k < got.size(); ++k)
            CHECK(got[k].i == ref_edges[k].i && got[k].j == ref_edges[k].j && got[k].energy == ref_edges[k].energy &&
                  got[k].length == ref_edges[k].length && got[k].cosine == ref_edges[k].cosine);
        // lexicographic order inside a frame
        for (int f = 0; f < n_frames; ++f)
            for (uint64_t k = ref_off[f] + 1; k < ref_off[f + 1]; ++k)
                CHECK(ref_edges[k - 1].i < ref_edges[k].i ||
                      (ref_edges[k - 1].i == ref_edges[k].i && ref_edges[k - 1].j < ref_edges[k].j));
        // the reference's container type
        std::vector<SiteInfo_ptr> infos;
        for (int w = 0; w < n_waters; ++w) {
            SiteInfo s; s.type = WATER; s.id = (unsigned)(w + 1); s.atmIndex = 3 * w; s.resIndex = w; s.nbHydrogen = 2;
            infos.push_back(std::make_shared<SiteInfo>(s));
        }
        HydrogenBondBatch again = hb.calculate(frames.data(), 1, natoms);
        std::vector<HydrogenBond> hbs = HydrogenBondEdges::toHydrogenBonds(again, 0, infos);
        CHECK(hbs.size() == ref_off[1] && hbs[0].direction == FORWARD);
        // ---- make_edges on the finder's own pair list (what analyzeFrame does, :583-633): the oxygens of frame 0
        // through FindPairs::find, the list handed to makeEdges in order -> the batched path's frame 0;
        // the same list reversed and flipped -> the same number of edges, every one with first/second swapped
        {
            std::vector<Point> oxy;
            for (int w = 0; w < n_waters; ++w) {
                const float* p = &frames[(size_t)(3 * w) * 3];
                oxy.push_back(Point((double)p[0], (double)p[1], (double)p[2]));
            }
            std::vector<std::pair<int, int>> pl = mp_->find(oxy);
            std::vector<HydrogenBond> me = hb.makeEdges(frames.data(), natoms, pl, infos);
            CHECK(me.size() == hbs.size());
            for (size_t k = 0; k < me.size(); ++k)
                CHECK(me[k].sites.first->id == hbs[k].sites.first->id && me[k].sites.second->id == hbs[k].sites.second->id &&
                      me[k].length == hbs[k].length && me[k].angle == hbs[k].angle && me[k].energy == hbs[k].energy);
            std::vector<std::pair<int, int>> rev;
            for (size_t k = pl.size(); k-- > 0;) rev.push_back(std::make_pair(pl[k].second, pl[k].first));
            std::vector<HydrogenBond> mr = hb.makeEdges(frames.data(), natoms, rev, infos);
            CHECK(mr.size() == me.size());                     // water-water, one real hydrogen each: symmetric decision
            for (size_t k = 0; k < mr.size(); ++k) {
                const HydrogenBond& a = mr[k];
                const HydrogenBond& b = me[me.size() - 1 - k];
                CHECK(a.sites.first->id == b.sites.second->id && a.sites.second->id == b.sites.first->id && a.length == b.length);
            }
            bool threw_oor = false;
            try { hb.makeEdges(frames.data(), natoms, {std::make_pair(0, n_waters)}, infos); } catch (const std::out_of_range&) { threw_oor = true; }
            CHECK(threw_oor);                                  // the reference's .at() throws the same
        }
        // periodic extension through the same classes: box through the batcher == calculate()
        {
            const float box[3][3] = {{2.1f, 0.f, 0.f}, {0.f, 2.1f, 0.f}, {0.f, 0.f, 2.1f}};
            std::vector<float> boxes;
            for (int f = 0; f < n_frames; ++f) { boxes.push_back(2.1f); boxes.push_back(2.1f); boxes.push_back(2.1f); }
            HydrogenBondBatch per = hb.calculate(frames.data(), boxes.data(), n_frames, natoms);
            CHECK(per.n_edges > all.n_edges);                       // pairs through the faces
            std::vector<wn_edge> pref(per.edges, per.edges + per.n_edges);
            std::vector<wn_edge> pgot;
            FrameBatcher pb(hb, natoms, 4, [&](int, const wn_edge* e, size_t n, const wn_frame_stats&) {
                pgot.insert(pgot.end(), e, e + n);
            });
            for (int f = 0; f < n_frames; ++f) pb.push(f, &frames[(size_t)f * natoms * 3], box);
            pb.finish();
            CHECK(pgot.size() == pref.size());
            for (size_t k = 0; k < pgot.size(); ++k) CHECK(pgot[k].i == pref[k].i && pgot[k].j == pref[k].j && pgot[k].cosine == pref[k].cosine);
            // a triclinic matrix is accepted too (all-pairs kernel): same frames, skewed cell
            const float tric[3][3] = {{2.1f, 0.f, 0.f}, {0.5f, 2.1f, 0.f}, {-0.3f, 0.4f, 2.1f}};
            size_t n_tric = 0;
            FrameBatcher tb(hb, natoms, 4, [&](int, const wn_edge*, size_t n, const wn_frame_stats&) { n_tric += n; });
            for (int f = 0; f < 2; ++f) tb.push(f, &frames[(size_t)f * natoms * 3], tric);
            tb.finish();
            CHECK(n_tric > 0);
            CudaFindPairs* cf = static_cast<CudaFindPairs*>(mp_.get());
            cf->setBox(3.0f, 3.0f, 3.0f);
            std::vector<std::pair<int, int>> pp = mp_->find(pts);
            cf->clearBox();
            CHECK(pp.size() >= pairs.size());
        }
        // GPU-formatted edge file text == host EdgeFileWriter (which is checked against iostreams).
        // Random frames may hold near-coincident sites whose energy needs more than 8 characters: then
        // the GPU formatter must refuse (fixed-width rows) and the host writer is the fallback.
        {
            auto host_text = [&](const HydrogenBondBatch& b, const std::vector<unsigned>& ids, int first_frnr) {
                std::ostringstream os;
                EdgeFileWriter w(os, ids);
                for (int f = 0; f < b.n_frames; ++f)
                    w.writeFrame(first_frnr + f, b.edges + b.offsets[f], (size_t)(b.offsets[f + 1] - b.offsets[f]));
                return os.str();
            };
            std::vector<unsigned> ids(n_waters);
            for (int s = 0; s < n_waters; ++s) ids[s] = (unsigned)(s + 41);
            HydrogenBondBatch b3 = hb.calculate(frames.data(), 3, natoms);
            const std::string ht = host_text(b3, ids, 5);
            bool fits = true;
            for (size_t p = 0, q; (q = ht.find('\n', p)) != std::string::npos; p = q + 1)
                if (ht[p] != '#' && q - p != 36) fits = false;
            const char* txt = nullptr; uint64_t nbytes = 0;
            const bool gpu_ok = hb.formatLastBatch({5, 6, 7}, ids, &txt, &nbytes);
            CHECK(gpu_ok == fits);
            if (gpu_ok) CHECK(ht == std::string(txt, (size_t)nbytes));
            // a regular lattice of waters: every value fits, the GPU text must equal the host text
            std::vector<float> lat((size_t)natoms * 3);
            for (int w = 0; w < n_waters; ++w) {
                float* p = &lat[(size_t)w * 9];
                p[0] = 0.31f * (float)(w % 7); p[1] = 0.29f * (float)((w / 7) % 7); p[2] = 0.33f * (float)(w / 49);
                p[3] = p[0] + 0.08f; p[4] = p[1] + 0.05f; p[5] = p[2];
                p[6] = p[0]; p[7] = p[1] - 0.09f; p[8] = p[2] + 0.03f;
            }
            HydrogenBondBatch bl = hb.calculate(lat.data(), 1, natoms);
            CHECK(bl.n_edges > 100);
            const std::string hl = host_text(bl, ids, 42);
            CHECK(hb.formatLastBatch({42}, ids, &txt, &nbytes));
            CHECK(hl == std::string(txt, (size_t)nbytes));
            std::vector<unsigned> wide(ids);
            for (unsigned& v : wide) v += 10000000u;
            CHECK(!hb.formatLastBatch({42}, wide, &txt, &nbytes));
            // connected components of that frame against a plain union-find on the host
            {
                const int32_t* labels = nullptr; const wn_component_stats* cs = nullptr;
                const float thr = 0.5f;
                hb.componentsOfLastBatch(thr, &labels, &cs);
                std::vector<int> par((size_t)n_waters), has((size_t)n_waters, 0);
                for (int s = 0; s < n_waters; ++s) par[s] = s;
                auto root = [&](int v) { while (par[v] != v) v = par[v] = par[par[v]]; return v; };
                int kept = 0;
                for (uint64_t k = bl.offsets[0]; k < bl.offsets[1]; ++k) {
                    const wn_edge& e = bl.edges[k];
                    if (std::fabs(e.energy) <= thr) continue;
                    ++kept; has[e.i] = has[e.j] = 1;
                    const int a = root(e.i), b = root(e.j);
                    if (a < b) par[b] = a; else if (b < a) par[a] = b;
                }
                int comps = 0, nodes = 0;
                for (int s = 0; s < n_waters; ++s) {
                    const int want = has[s] ? root(s) : -1;
                    CHECK(labels[s] == want);
                    nodes += has[s]; comps += (want == s);
                }
                CHECK(cs[0].n_edges == kept && cs[0].n_nodes == nodes && cs[0].n_components == comps);
            }
        }
        // per-frame site lists through the batcher: a frame restricted to the first 100 waters gives the
        // edges of the full frame whose endpoints are both < 100 (indices are list positions)
        {
            std::vector<int32_t> first100(100);
            for (int s = 0; s < 100; ++s) first100[s] = s;
            std::vector<wn_edge> sub;
            FrameBatcher sb(hb, natoms, 2, [&](int frnr, const wn_edge* e, size_t n, const wn_frame_stats& st) {
                if (frnr == 0) { sub.assign(e, e + n); if (st.num_sites != 100) throw std::runtime_error("num_sites"); }
            });
            for (int f = 0; f < 3; ++f) sb.push(f, &frames[(size_t)f * natoms * 3], nullptr, first100);
            sb.finish();
            std::vector<wn_edge> want100;
            for (uint64_t k = ref_off[0]; k < ref_off[1]; ++k)
                if (ref_edges[k].i < 100 && ref_edges[k].j < 100) want100.push_back(ref_edges[k]);
            CHECK(sub.size() == want100.size() && !sub.empty());
            for (size_t k = 0; k < sub.size(); ++k)
                CHECK(sub[k].i == want100[k].i && sub[k].j == want100[k].j && sub[k].energy == want100[k].energy);
        }
        // compact layouts + pipelined batching (+ several GPUs when the box has them): the same edges as the
        // single calculate() above, delivered in frame order; the NCCL-summed statistics table has every row
        {
            auto run_compact = [&](const std::vector<HydrogenBondEdges*>& engs, int layout, int per_batch, bool staged,
                                   std::vector<wn_edge>& out, std::vector<wn_frame_stats>& rows) -> bool {
                out.clear();
                std::vector<int> seen;
                const wn_energy_params ep = wn_energy_defaults();
                FrameBatcher cb(engs, natoms, per_batch, [&](const FrameResult& r) {
                    seen.push_back(r.frnr);
                    if (r.n_sites != n_waters || r.layout != layout) throw std::runtime_error("frame result header");
                    for (int i = 0; i < n_waters; ++i) {
                        const size_t b = r.row_offsets[i], e = i + 1 < n_waters ? r.row_offsets[i + 1] : r.n_edges;
                        for (size_t k = b; k < e; ++k) {
                            wn_edge w; w.i = i;
                            if (layout == WN_LAYOUT_CSR16) {
                                const wn_edge16& c = static_cast<const wn_edge16*>(r.edges)[k];
                                w.j = c.j; w.energy = c.energy; w.length = c.length; w.cosine = c.cosine;
                            } else if (layout == WN_LAYOUT_CSR12) {
                                const wn_edge12& c = static_cast<const wn_edge12*>(r.edges)[k];
                                w.j = c.j; w.length = c.length; w.cosine = c.cosine;
                                w.energy = wn_edge_energy(c.length, c.cosine, &ep);        // on demand, on the host
                            } else {
                                w = static_cast<const wn_edge*>(r.edges)[k];
                            }
                            out.push_back(w);
                        }
                    }
                }, layout);
                for (int f = 0; f < n_frames; ++f) {
                    if (staged) { std::memcpy(cb.stage(), &frames[(size_t)f * natoms * 3], (size_t)natoms * 3 * sizeof(float)); cb.commit(100 + f); }
                    else cb.push(100 + f, &frames[(size_t)f * natoms * 3]);
                }
                cb.finish();
                rows = cb.reducedStats();
                if (seen.size() != (size_t)n_frames) return false;
                for (int f = 0; f < n_frames; ++f) if (seen[f] != 100 + f) return false;
                return true;
            };
            auto same = [&](const std::vector<wn_edge>& a, bool exact_energy) {
                if (a.size() != ref_edges.size()) return false;
                for (size_t k = 0; k < a.size(); ++k) {
                    if (a[k].i != ref_edges[k].i || a[k].j != ref_edges[k].j || a[k].length != ref_edges[k].length ||
                        a[k].cosine != ref_edges[k].cosine) return false;
                    const double d = std::fabs((double)a[k].energy - (double)ref_edges[k].energy);
                    if (exact_energy ? d != 0.0 : d > 1e-6 * std::fabs((double)ref_edges[k].energy)) return false;
                }
                return true;
            };
            std::vector<wn_frame_stats> ref_stats(all.stats, all.stats + n_frames);
            HydrogenBondBatch again0 = hb.calculate(frames.data(), n_frames, natoms);       // `all` was overwritten long ago
            ref_stats.assign(again0.stats, again0.stats + n_frames);
            std::vector<wn_edge> got2;
            std::vector<wn_frame_stats> rows;
            std::vector<HydrogenBondEdges*> one{&hb};
            CHECK(run_compact(one, WN_LAYOUT_CSR16, 2, false, got2, rows) && same(got2, true));
            CHECK(rows.size() == (size_t)n_frames);
            for (int f = 0; f < n_frames; ++f)
                CHECK(rows[f].num_edges == ref_stats[f].num_edges && rows[f].sum_energy == ref_stats[f].sum_energy);
            CHECK(run_compact(one, WN_LAYOUT_CSR12, 3, true, got2, rows) && same(got2, false));
            for (int f = 0; f < n_frames; ++f) CHECK(rows[f].sum_energy == ref_stats[f].sum_energy);
            CHECK(run_compact(one, WN_LAYOUT_EDGES, 1, false, got2, rows) && same(got2, true));
            const int ndev = wn_device_count();
            if (ndev >= 2) {
                std::vector<std::unique_ptr<HydrogenBondEdges> > more;
                std::vector<HydrogenBondEdges*> engs{&hb};
                for (int g = 1; g < std::min(ndev, 4); ++g) {
                    more.emplace_back(new HydrogenBondEdges(g));
                    more.back()->setWaterSites(n_waters);
                    engs.push_back(more.back().get());
                }
                CHECK(run_compact(engs, WN_LAYOUT_CSR16, 2, false, got2, rows) && same(got2, true));
                CHECK(rows.size() == (size_t)n_frames);
                for (int f = 0; f < n_frames; ++f)      // NCCL sum of zero-padded per-GPU tables == the rows themselves
                    CHECK(rows[f].num_sites == n_waters && rows[f].num_edges == ref_stats[f].num_edges &&
                          rows[f].ratio == ref_stats[f].ratio && rows[f].sum_energy == ref_stats[f].sum_energy &&
                          rows[f].pair_evals == ref_stats[f].pair_evals);
                std::printf("multi-GPU batcher OK on %zu GPUs\n", engs.size());
            }
        }
        // error behaviour: no exception escapes the C ABI; the C++ layer throws
        bool threw = false;
        try { hb.calculate(frames.data(), 1, 10); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw);
        std::printf("test_host OK: %zu pairs, %llu edges in %d frames\n", pairs.size(),
                    (unsigned long long)all.n_edges, n_frames);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
}
