// test_batcher_mock.cpp -- FrameBatcher (include/hydrogen-bond-edges.hpp) on a CPU box, over the stand-in C ABI of
// tests/mock (oracle behind it): the batching state machine the reference-side analyzeFrame patch relies on --
// two staging buffers and two result sets per GPU, one worker thread per GPU, round-robin frame blocks over
// several GPUs (contiguous frame blocks, no data-path exchange), delivery in frame order on the
// caller's thread, the statistics table summed over the GPUs at finish() -- with uneven "GPU" speeds, partial
// tails, flushes, periodic / site-list batches, helper copy threads and failing batch calls.
// Built by tests/mock/Makefile (also under ThreadSanitizer and AddressSanitizer); run by tests/test_host_mock_cpu.py.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <stdexcept>
#include <thread>
#include <vector>

#include "hydrogen-bond-edges.hpp"
#include "wn_energy.h"
#include "wn_mock.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); return 1; } } while (0)

namespace {

const int kWaters = 60, kAtoms = 180;

struct Reference {
    std::vector<std::vector<wn_edge> > edges;      // per frame
    std::vector<wn_frame_stats> stats;
};

bool same_stats(const wn_frame_stats& a, const wn_frame_stats& b)
{
    return a.num_sites == b.num_sites && a.num_edges == b.num_edges && a.ratio == b.ratio && a.sum_energy == b.sum_energy &&
           a.pair_evals == b.pair_evals;
}

// every frame result back to wn_edge records
std::vector<wn_edge> decode(const FrameResult& r)
{
    std::vector<wn_edge> out;
    const wn_energy_params ep = wn_energy_defaults();
    for (int i = 0; i < r.n_sites; ++i) {
        const size_t b = r.row_offsets[i], e = i + 1 < r.n_sites ? r.row_offsets[i + 1] : r.n_edges;
        for (size_t k = b; k < e; ++k) {
            wn_edge w; w.i = i;
            if (r.layout == WN_LAYOUT_CSR16) {
                const wn_edge16& c = static_cast<const wn_edge16*>(r.edges)[k];
                w.j = c.j; w.energy = c.energy; w.length = c.length; w.cosine = c.cosine;
            } else if (r.layout == WN_LAYOUT_CSR12) {
                const wn_edge12& c = static_cast<const wn_edge12*>(r.edges)[k];
                w.j = c.j; w.length = c.length; w.cosine = c.cosine; w.energy = wn_edge_energy(c.length, c.cosine, &ep);
            } else {
                w = static_cast<const wn_edge*>(r.edges)[k];
            }
            out.push_back(w);
        }
    }
    return out;
}

bool same_edges(const std::vector<wn_edge>& a, const std::vector<wn_edge>& b, bool exact_energy)
{
    if (a.size() != b.size()) return false;
    for (size_t k = 0; k < a.size(); ++k) {
        if (a[k].i != b[k].i || a[k].j != b[k].j || a[k].length != b[k].length || a[k].cosine != b[k].cosine) return false;
        const double d = std::fabs((double)a[k].energy - (double)b[k].energy);
        if (exact_energy ? d != 0.0 : d > 1e-6 * std::fabs((double)b[k].energy)) return false;
    }
    return true;
}

struct Run {
    int gpus, per_batch, layout;
    bool staged, pipelined;
    int copy_threads;
    int flush_at;      // flush() after this many frames (-1: never)
    int flush2_at;     // a second flush (-1: never)
    int n_frames;      // frames pushed (<= the 23 prepared)
    bool dynamic;      // setDynamicGpuChoice
};

}  // namespace

int main()
{
    try {
        std::mt19937 rng(11);
        const int F = 23;
        std::vector<float> frames((size_t)F * kAtoms * 3);
        // waters on a jittered 4 x 4 x 4 lattice of 0.32 nm (no overlaps: every energy fits the 8-character text field)
        std::uniform_real_distribution<float> ub(-0.05f, 0.05f), uh(-0.06f, 0.06f);
        for (int f = 0; f < F; ++f)
            for (int w = 0; w < kWaters; ++w) {
                float* p = &frames[((size_t)f * kAtoms + 3 * w) * 3];
                const int cell[3] = {w % 4, (w / 4) % 4, w / 16};
                for (int c = 0; c < 3; ++c) { p[c] = 0.05f + 0.32f * (float)cell[c] + ub(rng); p[3 + c] = p[c] + uh(rng); p[6 + c] = p[c] + uh(rng); }
            }
        CHECK(wn_device_count() == 8);
        std::vector<std::unique_ptr<HydrogenBondEdges> > engines;
        for (int g = 0; g < 8; ++g) {
            engines.emplace_back(new HydrogenBondEdges(g));
            engines.back()->setWaterSites(kWaters);
        }
        // the frames in one unbatched call on one context
        Reference ref;
        {
            HydrogenBondBatch all = engines[0]->calculate(frames.data(), F, kAtoms);
            CHECK(all.n_frames == F && all.n_edges > 200);
            for (int f = 0; f < F; ++f) {
                ref.edges.emplace_back(all.edges + all.offsets[f], all.edges + all.offsets[f + 1]);
                ref.stats.push_back(all.stats[f]);
            }
        }
        const std::thread::id me = std::this_thread::get_id();

        // ---- every shape of a run: GPUs x batch size x layout x push/stage x pipelined x copy threads x a flush in the middle
        std::vector<Run> runs;
        const int gpus_list[] = {1, 2, 3, 5, 8};
        const int batch_list[] = {1, 2, 3, 7, 40};
        int variant = 0;
        for (int g : gpus_list)
            for (int b : batch_list) {
                Run r;
                r.gpus = g; r.per_batch = b; r.layout = variant % 3; r.staged = (variant / 3) % 2 == 1; r.pipelined = variant % 5 != 4;
                r.copy_threads = variant % 4 == 1 ? 3 : 1; r.flush_at = variant % 3 == 2 ? 5 + variant % 7 : -1;
                r.dynamic = variant % 2 == 1 && g > 1;
                r.flush2_at = -1; r.n_frames = F;
                runs.push_back(r);
                ++variant;
            }
        // ... and a seeded random sweep over the same dimensions, with shorter runs, empty runs and two flushes
        {
            std::mt19937 fz(2024);
            auto pick = [&](int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(fz); };
            for (int k = 0; k < 150; ++k) {
                Run r;
                r.gpus = pick(1, 8); r.per_batch = pick(1, 9); r.layout = pick(0, 2); r.staged = pick(0, 1) == 1; r.pipelined = pick(0, 4) != 0;
                r.copy_threads = pick(1, 3); r.dynamic = pick(0, 1) == 1 && r.gpus > 1;
                r.n_frames = pick(0, F);
                r.flush_at = pick(0, 2) == 0 && r.n_frames > 0 ? pick(1, r.n_frames) : -1;
                r.flush2_at = r.flush_at > 0 && pick(0, 1) == 1 && r.flush_at < r.n_frames ? pick(r.flush_at + 1, r.n_frames) : -1;
                runs.push_back(r);
            }
        }
        std::uniform_int_distribution<int> delay(0, 400);
        for (const Run& r : runs) {
            const int F = r.n_frames;
            for (int g = 0; g < 8; ++g) wn_mock_set_delay_us(g, delay(rng));       // GPUs of a host are not equally fast
            std::vector<HydrogenBondEdges*> engs;
            for (int g = 0; g < r.gpus; ++g) engs.push_back(engines[(size_t)g].get());
            std::vector<int> seen, dev;
            std::vector<wn_frame_stats> seen_stats;
            bool ok_thread = true, ok_edges = true;
            FrameBatcher fb(engs, kAtoms, r.per_batch, [&](const FrameResult& fr) {
                if (std::this_thread::get_id() != me) ok_thread = false;
                const int f = fr.frnr - 1000;
                seen.push_back(f); dev.push_back(fr.device); seen_stats.push_back(fr.stats);
                if (f < 0 || f >= F || fr.layout != r.layout || fr.n_sites != kWaters ||
                    !same_edges(decode(fr), ref.edges[(size_t)f], r.layout != WN_LAYOUT_CSR12)) ok_edges = false;
            }, r.layout);
            CHECK(fb.gpus() == r.gpus);
            fb.setPipelined(r.pipelined);
            fb.setCopyThreads(r.copy_threads);
            fb.setDynamicGpuChoice(r.dynamic);
            for (int f = 0; f < F; ++f) {
                const float* src = &frames[(size_t)f * kAtoms * 3];
                if (r.staged) { std::memcpy(fb.stage(), src, (size_t)kAtoms * 3 * sizeof(float)); fb.commit(1000 + f); }
                else fb.push(1000 + f, src);
                if (f + 1 == r.flush_at || f + 1 == r.flush2_at) {
                    fb.flush();
                    CHECK(fb.framesDone() == (uint64_t)(f + 1) && (int)seen.size() == f + 1);       // nothing left in flight
                }
                CHECK(seen.size() <= (size_t)(f + 1));
            }
            fb.finish();
            CHECK(ok_thread && ok_edges);
            CHECK(fb.framesDone() == (uint64_t)F && (int)seen.size() == F);
            for (int f = 0; f < F; ++f) CHECK(seen[(size_t)f] == f);                                 // frame order
            for (int f = 0; f < F; ++f) CHECK(same_stats(seen_stats[(size_t)f], ref.stats[(size_t)f]));
            // the statistics table of the run (summed over the GPUs): one row per pushed frame, each filled by exactly one GPU
            const std::vector<wn_frame_stats>& rows = fb.reducedStats();
            CHECK(rows.size() == (size_t)F);
            for (int f = 0; f < F; ++f) CHECK(same_stats(rows[(size_t)f], ref.stats[(size_t)f]));
            // contiguous blocks of per_batch frames go round-robin to the GPUs (a flush ends a block early)
            if (!r.dynamic) {
                int lane = 0, in_block = 0;
                for (int f = 0; f < F; ++f) {
                    CHECK(dev[(size_t)f] == lane);
                    ++in_block;
                    if (in_block == r.per_batch || f + 1 == r.flush_at || f + 1 == r.flush2_at) { in_block = 0; lane = (lane + 1) % r.gpus; }
                }
            }
            CHECK(fb.batchSeconds().size() >= (size_t)((F + r.per_batch - 1) / r.per_batch));
        }
        for (int g = 0; g < 8; ++g) wn_mock_set_delay_us(g, 0);

        // ---- the GPUs really work at the same time: 4 GPUs, 8 batches of 30 ms each -> far less than 240 ms
        {
            for (int g = 0; g < 4; ++g) wn_mock_set_delay_us(g, 15000);
            std::vector<HydrogenBondEdges*> engs;
            for (int g = 0; g < 4; ++g) engs.push_back(engines[(size_t)g].get());
            int n_seen = 0;
            FrameBatcher fb(engs, kAtoms, 2, [&](const FrameResult&) { ++n_seen; });
            const auto t0 = std::chrono::steady_clock::now();
            for (int f = 0; f < 16; ++f) fb.push(f, &frames[(size_t)f * kAtoms * 3]);
            fb.finish();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            CHECK(n_seen == 16);
            CHECK(dt < 0.75 * 8 * 0.030);
            // one GPU, pipelined: the caller fills batch k+1 while batch k is on the GPU -- the run takes the GPU time,
            // not GPU time + producer time
            std::vector<HydrogenBondEdges*> one{engines[0].get()};
            FrameBatcher f1(one, kAtoms, 2, [&](const FrameResult&) {});
            const auto t1 = std::chrono::steady_clock::now();
            for (int f = 0; f < 8; ++f) {
                std::this_thread::sleep_for(std::chrono::milliseconds(12));        // a producer that needs 12 ms per frame
                f1.push(f, &frames[(size_t)f * kAtoms * 3]);
            }
            f1.finish();
            const double d1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
            CHECK(d1 < 0.85 * (8 * 0.012 + 4 * 0.030));
            for (int g = 0; g < 4; ++g) wn_mock_set_delay_us(g, 0);
        }

        // ---- opt-in dynamic GPU choice (setDynamicGpuChoice): with one GPU three times slower than the other two the
        // next batch goes to whichever GPU is expected to be free first; same frames, same order, same table, less time
        {
            std::vector<HydrogenBondEdges*> engs{engines[0].get(), engines[1].get(), engines[2].get()};
            double secs[2] = {0, 0};
            for (int dyn = 0; dyn < 2; ++dyn) {
                wn_mock_set_delay_us(0, 15000); wn_mock_set_delay_us(1, 5000); wn_mock_set_delay_us(2, 5000);
                std::vector<int> dev;
                bool ok = true; int n_seen = 0;
                FrameBatcher fb(engs, kAtoms, 2, [&](const FrameResult& fr) {
                    if (fr.frnr != n_seen || !same_edges(decode(fr), ref.edges[(size_t)(fr.frnr % F)], true)) ok = false;
                    dev.push_back(fr.device); ++n_seen;
                }, WN_LAYOUT_CSR16);
                fb.setDynamicGpuChoice(dyn == 1);
                const auto t0 = std::chrono::steady_clock::now();
                for (int f = 0; f < 48; ++f) fb.push(f, &frames[(size_t)(f % F) * kAtoms * 3]);
                fb.finish();
                secs[dyn] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                CHECK(ok && n_seen == 48 && fb.reducedStats().size() == 48);
                for (int f = 0; f < 48; ++f) CHECK(same_stats(fb.reducedStats()[(size_t)f], ref.stats[(size_t)(f % F)]));
                int on_slow = 0;
                for (int d : dev) on_slow += d == 0;
                if (dyn == 0) CHECK(on_slow == 16);                 // round-robin: a third of the frames on the slow GPU
                else CHECK(on_slow < 16 && on_slow > 0);
            }
            CHECK(secs[1] < 0.85 * secs[0]);
            for (int g = 0; g < 3; ++g) wn_mock_set_delay_us(g, 0);
        }

        // ---- periodic and triclinic boxes, and per-frame site lists, over three GPUs == one GPU unbatched
        {
            std::vector<HydrogenBondEdges*> engs{engines[0].get(), engines[1].get(), engines[2].get()};
            const float rect[3][3] = {{1.3f, 0.f, 0.f}, {0.f, 1.3f, 0.f}, {0.f, 0.f, 1.3f}};      // lattice spans 0..1.06 nm
            const float tric[3][3] = {{1.3f, 0.f, 0.f}, {0.4f, 1.3f, 0.f}, {-0.2f, 0.3f, 1.3f}};
            for (int which = 0; which < 2; ++which) {
                const float (*box)[3] = which ? tric : rect;
                std::vector<float> b9;
                for (int f = 0; f < F; ++f)
                    for (int a = 0; a < 3; ++a)
                        for (int c = 0; c < 3; ++c) b9.push_back(box[a][c]);
                HydrogenBondBatch want = engines[7]->calculateBox9(frames.data(), b9.data(), F, kAtoms);
                std::vector<std::vector<wn_edge> > we;
                for (int f = 0; f < F; ++f) we.emplace_back(want.edges + want.offsets[f], want.edges + want.offsets[f + 1]);
                size_t plain = 0, per = 0;
                for (int f = 0; f < F; ++f) { plain += ref.edges[(size_t)f].size(); per += we[(size_t)f].size(); }
                CHECK(per > plain);                                                  // pairs through the faces
                bool ok = true; int n_seen = 0;
                FrameBatcher fb(engs, kAtoms, 4, [&](const FrameResult& fr) {
                    if (fr.frnr != n_seen || !same_edges(decode(fr), we[(size_t)fr.frnr], true)) ok = false;
                    ++n_seen;
                }, WN_LAYOUT_CSR16);
                for (int f = 0; f < F; ++f) fb.push(f, &frames[(size_t)f * kAtoms * 3], box);
                fb.finish();
                CHECK(ok && n_seen == F);
            }
            // site lists: frame f keeps the waters w with (w + f) % 3 != 0
            std::vector<std::vector<int32_t> > lists((size_t)F);
            std::vector<int64_t> off{0};
            std::vector<int32_t> flat;
            for (int f = 0; f < F; ++f) {
                for (int w = 0; w < kWaters; ++w) if ((w + f) % 3 != 0) lists[(size_t)f].push_back(w);
                flat.insert(flat.end(), lists[(size_t)f].begin(), lists[(size_t)f].end());
                off.push_back((int64_t)flat.size());
            }
            HydrogenBondBatch want = engines[7]->calculate(frames.data(), nullptr, F, kAtoms, off.data(), flat.data());
            std::vector<std::vector<wn_edge> > we;
            std::vector<wn_frame_stats> ws(want.stats, want.stats + F);
            for (int f = 0; f < F; ++f) we.emplace_back(want.edges + want.offsets[f], want.edges + want.offsets[f + 1]);
            bool ok = true; int n_seen = 0;
            FrameBatcher fb(engs, kAtoms, 3, [&](const FrameResult& fr) {
                if (fr.frnr != n_seen || fr.stats.num_sites != (double)lists[(size_t)fr.frnr].size() ||
                    !same_edges(decode(fr), we[(size_t)fr.frnr], true)) ok = false;
                ++n_seen;
            }, WN_LAYOUT_EDGES);
            for (int f = 0; f < F; ++f) fb.push(f, &frames[(size_t)f * kAtoms * 3], nullptr, lists[(size_t)f]);
            fb.finish();
            CHECK(ok && n_seen == F);
            for (int f = 0; f < F; ++f) CHECK(same_stats(fb.reducedStats()[(size_t)f], ws[(size_t)f]));
            // mixing frames with and without boxes / lists inside one batch is refused
            FrameBatcher mix(engs, kAtoms, 4, [&](const FrameResult&) {}, WN_LAYOUT_EDGES);
            mix.push(0, frames.data());
            bool threw = false;
            try { mix.push(1, frames.data(), rect); } catch (const std::invalid_argument&) { threw = true; }
            CHECK(threw);
        }

        // ---- classic single-GPU form with a batch hook (what the standalone driver uses): the hook sees every batch
        // before its frames reach the sink, formatLastBatch applies to that batch
        {
            HydrogenBondEdges& hb = *engines[0];
            std::vector<int> order;
            std::vector<std::vector<int> > hook_batches;
            size_t text_bytes = 0;
            FrameBatcher fb(hb, kAtoms, 4, [&](int frnr, const wn_edge* e, size_t n, const wn_frame_stats& st) {
                order.push_back(frnr);
                if (st.num_edges != (double)n || !same_edges(std::vector<wn_edge>(e, e + n), ref.edges[(size_t)frnr], true))
                    throw std::runtime_error("classic sink: wrong edges");
            });
            fb.setBatchHook([&](const std::vector<int>& frnrs, const HydrogenBondBatch& b) {
                if ((int)frnrs.size() != b.n_frames) throw std::runtime_error("hook: frame numbers");
                if (!order.empty() && order.back() >= frnrs.front()) throw std::runtime_error("hook: a later batch before an earlier frame");
                hook_batches.push_back(frnrs);
                const char* text = nullptr; uint64_t bytes = 0;
                if (!hb.formatLastBatch(frnrs, std::vector<unsigned>(), &text, &bytes)) throw std::runtime_error("hook: format");
                text_bytes += (size_t)bytes;
            });
            for (int f = 0; f < 10; ++f) fb.push(f, &frames[(size_t)f * kAtoms * 3]);
            fb.finish();
            CHECK(order.size() == 10 && hook_batches.size() == 3 && hook_batches[2].size() == 2);
            size_t rows = 0;
            for (int f = 0; f < 10; ++f) rows += ref.edges[(size_t)f].size();
            CHECK(text_bytes > 37 * rows && text_bytes < 37 * rows + 10 * 32);
        }

        // ---- nothing pushed; a batcher destroyed with batches in flight
        {
            std::vector<HydrogenBondEdges*> engs{engines[0].get(), engines[1].get()};
            FrameBatcher empty(engs, kAtoms, 4, [&](const FrameResult&) {}, WN_LAYOUT_CSR16);
            empty.finish();
            CHECK(empty.reducedStats().empty() && empty.framesDone() == 0);
            wn_mock_set_delay_us(0, 20000); wn_mock_set_delay_us(1, 20000);
            {
                FrameBatcher dropped(engs, kAtoms, 1, [&](const FrameResult&) {}, WN_LAYOUT_CSR16);
                for (int f = 0; f < 4; ++f) dropped.push(f, &frames[(size_t)f * kAtoms * 3]);
            }                                                                        // waits for the workers, no sink call needed
            wn_mock_set_delay_us(0, 0); wn_mock_set_delay_us(1, 0);
        }

        // ---- a failing batch call surfaces as an exception on the caller's thread; the batcher can be destroyed,
        // the engines stay usable
        {
            std::vector<HydrogenBondEdges*> engs{engines[0].get(), engines[1].get(), engines[2].get()};
            wn_mock_fail_batches(1, 1);                     // GPU 1: its second batch call from now on fails
            bool threw = false;
            int delivered = 0;
            try {
                FrameBatcher fb(engs, kAtoms, 2, [&](const FrameResult& fr) { if (fr.frnr != delivered) throw std::logic_error("order"); ++delivered; },
                                WN_LAYOUT_CSR16);
                for (int f = 0; f < F; ++f) fb.push(f, &frames[(size_t)f * kAtoms * 3]);
                fb.finish();
            } catch (const std::runtime_error& e) {
                threw = std::string(e.what()).find("injected failure") != std::string::npos;
            }
            CHECK(threw);
            CHECK(delivered >= 6 && delivered < F);         // batches 0..3 (GPU 0,1,2,0) were fine; batch 4 (GPU 1) failed
            wn_mock_fail_batches(-1, 0);
            HydrogenBondBatch again = engines[1]->calculate(frames.data(), 2, kAtoms);
            CHECK(again.n_frames == 2 && same_edges(std::vector<wn_edge>(again.edges + again.offsets[1], again.edges + again.offsets[2]), ref.edges[1], true));
        }
        // ---- a constructor that fails half way (pinned staging of the second GPU cannot be allocated) throws and
        // leaves nothing behind (the AddressSanitizer build checks the leak); the engines stay usable
        {
            std::vector<HydrogenBondEdges*> engs{engines[0].get(), engines[1].get(), engines[2].get()};
            wn_mock_fail_pinned_alloc(1, 1);                // GPU 1: its second staging buffer
            bool threw = false;
            try { FrameBatcher fb(engs, kAtoms, 2, [&](const FrameResult&) {}, WN_LAYOUT_CSR16); } catch (const std::runtime_error&) { threw = true; }
            wn_mock_fail_pinned_alloc(-1, 0);
            CHECK(threw);
            int n_seen = 0;
            FrameBatcher fb(engs, kAtoms, 2, [&](const FrameResult&) { ++n_seen; }, WN_LAYOUT_CSR16);
            for (int f = 0; f < 5; ++f) fb.push(f, &frames[(size_t)f * kAtoms * 3]);
            fb.finish();
            CHECK(n_seen == 5);
        }
        std::printf("test_batcher_mock OK: %zu run shapes, up to %d frames each\n", runs.size(), F);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
}
