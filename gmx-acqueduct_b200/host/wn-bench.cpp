// wn-bench.cpp -- end-to-end throughput of the C++ drop-in itself (no Python in the loop).
//
// What the reference's maintainer would run after applying INTEGRATION.md: frames arrive one by one in the
// caller's own (pageable) buffers, as GROMACS hands them to WaterNetwork::analyzeFrame (reference
// src/waternetwork.cpp:503), go through FrameBatcher::push -> pinned staging -> H2D -> kernels -> D2H, and
// come back per frame, in frame order, to a sink that reads them.  One process, one FrameBatcher over
// -gpus N contexts (frame blocks round-robin over the GPUs, NCCL only for the statistics table).
// Also times CudaFindPairs::find behind the FindPairs interface (include/find-pairs.hpp:11).
//
//   wn_bench [-waters 33333] [-batch 250] [-batches 12] [-warmup 2] [-gpus 1] [-layout 16|12|20]
//            [-mode push|stage] [-ring 250] [-find 0|1] [-threads T] [-pipelined 1|0] [-copy-threads 4] [-dynamic 0|1]
// -dynamic 1: FrameBatcher::setDynamicGpuChoice (next batch to the GPU expected to be free first; default round-robin).
// Prints ONE JSON line.  Frames: the synthetic water box of include/wn_synth.h (seed 1234), a ring of
// -ring distinct frames generated before the clock starts.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "cuda-find-pairs.hpp"
#include "hydrogen-bond-edges.hpp"
#include "wn_energy.h"
#include "wn_synth.h"

namespace {
double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

int main(int argc, char** argv)
{
    int waters = 33333, batch = 250, batches = 12, warmup = 2, gpus = 1, layout_b = 16, ring = 250, do_find = 1, pipelined = 1, copy_threads = 4, dynamic = 0;
    unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    std::string mode = "push";
    for (int k = 1; k + 1 < argc; k += 2) {
        const std::string a = argv[k];
        const char* v = argv[k + 1];
        if (a == "-waters") waters = std::atoi(v);
        else if (a == "-batch") batch = std::atoi(v);
        else if (a == "-batches") batches = std::atoi(v);
        else if (a == "-warmup") warmup = std::atoi(v);
        else if (a == "-gpus") gpus = std::atoi(v);
        else if (a == "-layout") layout_b = std::atoi(v);
        else if (a == "-mode") mode = v;
        else if (a == "-ring") ring = std::atoi(v);
        else if (a == "-find") do_find = std::atoi(v);
        else if (a == "-pipelined") pipelined = std::atoi(v);
        else if (a == "-copy-threads") copy_threads = std::atoi(v);
        else if (a == "-dynamic") dynamic = std::atoi(v);
        else if (a == "-threads") threads = (unsigned)std::atoi(v);
        else { std::fprintf(stderr, "wn_bench: unknown option %s\n", a.c_str()); return 2; }
    }
    const int layout = layout_b == 12 ? WN_LAYOUT_CSR12 : layout_b == 20 ? WN_LAYOUT_EDGES : WN_LAYOUT_CSR16;
    const int natoms = 3 * waters;
    const size_t frame_floats = (size_t)natoms * 3;
    try {
        // ---- the caller's frames: a ring of distinct frames in ordinary memory
        ring = std::max(1, std::min(ring, batch * (batches + warmup)));
        std::vector<float> frames((size_t)ring * frame_floats);
        {
            const wn_synth_box box = wn_synth_make_box(waters, 1234, WN_MODEL_TIP3P);
            std::atomic<int> next(0);
            std::vector<std::thread> ts;
            for (unsigned t = 0; t < threads; ++t)
                ts.emplace_back([&]() {
                    for (int f; (f = next.fetch_add(1)) < ring;)
                        for (int w = 0; w < waters; ++w) wn_synth_water(&box, (uint64_t)f, w, &frames[(size_t)f * frame_floats + 9 * (size_t)w]);
                });
            for (auto& t : ts) t.join();
        }
        std::vector<std::unique_ptr<HydrogenBondEdges> > engines;
        std::vector<HydrogenBondEdges*> eng;
        for (int g = 0; g < gpus; ++g) {
            engines.emplace_back(new HydrogenBondEdges(g));
            engines.back()->setWaterSites(waters);
            eng.push_back(engines.back().get());
        }
        // ---- the sink reads what a consumer must at least read: the frame's statistics row, its row
        // offsets' end and its last edge record; with the 12-byte layout it also recomputes the energy of that
        // edge on the host (the on-demand path of include/wn_energy.h)
        uint64_t edges_seen = 0, frames_seen = 0;
        double checksum = 0.0;
        int last_frnr = -1, first_timed = 1 << 30;
        bool in_order = true;
        const wn_energy_params ep = wn_energy_defaults();
        std::vector<uint64_t> per_gpu((size_t)gpus, 0);
        auto sink = [&](const FrameResult& r) {
            if (r.frnr >= first_timed && r.device >= 0 && r.device < gpus) ++per_gpu[(size_t)r.device];
            in_order = in_order && r.frnr == last_frnr + 1;
            last_frnr = r.frnr;
            if (r.frnr >= first_timed) { edges_seen += r.n_edges; ++frames_seen; }
            if (r.n_edges) {
                if (r.layout == WN_LAYOUT_CSR12) {
                    const wn_edge12& e = static_cast<const wn_edge12*>(r.edges)[r.n_edges - 1];
                    checksum += (double)wn_edge_energy(e.length, e.cosine, &ep);
                } else if (r.layout == WN_LAYOUT_CSR16) {
                    checksum += (double)static_cast<const wn_edge16*>(r.edges)[r.n_edges - 1].energy;
                } else {
                    checksum += (double)static_cast<const wn_edge*>(r.edges)[r.n_edges - 1].energy;
                }
            }
            checksum += r.stats.num_edges;
        };
        double t_run = 0.0, t_first = 0.0;
        uint64_t timed_frames = 0, timed_edges = 0;
        size_t reduced_rows = 0;
        std::vector<double> batch_s;
        wn_timings last_tm;
        std::memset(&last_tm, 0, sizeof last_tm);
        {
            FrameBatcher fb(eng, natoms, batch, sink, layout);
            fb.setPipelined(pipelined != 0);
            fb.setCopyThreads(copy_threads);               // push(): the copy into pinned staging shared by a few threads
            int frnr = 0;
            auto feed = [&](int n_batches) {
                for (int b = 0; b < n_batches; ++b)
                    for (int k = 0; k < batch; ++k, ++frnr) {
                        const float* src = &frames[(size_t)(frnr % ring) * frame_floats];
                        if (mode == "stage") fb.commit(frnr);      // the producer wrote the frame in place (see below)
                        else fb.push(frnr, src);
                    }
            };
            // -mode stage: a zero-copy producer (a trajectory reader that decodes straight into stage()) pays no
            // host copy.  Emulated here: the warm-up fills both staging buffers of every GPU with frames, the timed
            // loop commits them again as they are -- the batcher, PCIe and the kernels do the same work.
            if (mode == "stage" && warmup < 2) warmup = 2;
            // warm-up (allocations, row capacity, pinned result buffers), not timed
            {
                const double t0 = now();
                if (mode == "stage")
                    for (int b = 0; b < warmup * gpus; ++b)
                        for (int k = 0; k < batch; ++k, ++frnr) { std::memcpy(fb.stage(), &frames[(size_t)(frnr % ring) * frame_floats], frame_floats * sizeof(float)); fb.commit(frnr); }
                else
                    feed(warmup * gpus);
                t_first = now() - t0;
            }
            // empty the pipeline, then start the clock: the timed frames are pushed, processed and delivered inside it
            fb.flush();
            fb.setDynamicGpuChoice(dynamic != 0);          // after the warm-up: round-robin filled both staging buffers of every GPU
            first_timed = frnr;
            const double t0 = now();
            feed(batches * gpus);
            fb.finish();
            t_run = now() - t0;
            timed_frames = frames_seen;
            timed_edges = edges_seen;
            reduced_rows = fb.reducedStats().size();
            batch_s = fb.batchSeconds();
            wn_get_timings(eng[0]->context(), &last_tm);
        }
        // ---- FindPairs drop-in: CudaFindPairs::find on the oxygens of frame 0 (std::vector<Point> in, vector of pairs out)
        double find_ms = 0.0, find_into_ms = 0.0;
        size_t n_pairs = 0;
        if (do_find) {
            std::shared_ptr<FindPairs> mp = std::make_shared<CudaFindPairs>(0);
            std::vector<Point> pts;
            pts.reserve((size_t)waters);
            for (int w = 0; w < waters; ++w) pts.push_back(Point(frames[9 * (size_t)w], frames[9 * (size_t)w + 1], frames[9 * (size_t)w + 2]));
            n_pairs = mp->find(pts).size();                      // warm-up
            const int reps = 5;
            const double t0 = now();
            for (int r = 0; r < reps; ++r) n_pairs = mp->find(pts).size();
            find_ms = (now() - t0) / reps * 1e3;
            // the same list into a vector the caller keeps (no fresh 127 MB allocation per frame)
            CudaFindPairs* cf = static_cast<CudaFindPairs*>(mp.get());
            std::vector<std::pair<int, int> > kept;
            cf->findInto(pts, kept);
            const double t1 = now();
            for (int r = 0; r < reps; ++r) cf->findInto(pts, kept);
            find_into_ms = (now() - t1) / reps * 1e3;
            if (kept.size() != n_pairs) throw std::runtime_error("findInto and find disagree");
        }
        const size_t esz = layout == WN_LAYOUT_CSR12 ? 12 : layout == WN_LAYOUT_CSR16 ? 16 : 20;
        std::string pg;
        for (int g = 0; g < gpus; ++g) pg += (g ? ", " : "") + std::to_string(per_gpu[(size_t)g]);
        std::string bs;
        for (size_t k = 0; k < batch_s.size(); ++k) { char t[32]; std::snprintf(t, sizeof t, "%s%.1f", k ? ", " : "", batch_s[k] * 1e3); bs += t; }
        std::printf("{\"bench\": \"wn_bench (C++ drop-in: FrameBatcher over the C ABI)\", \"e2e_cpp\": %.1f, \"unit\": \"frames/s\", "
                    "\"gpus\": %d, \"mode\": \"%s\", \"copy_threads\": %d, \"dynamic_gpu_choice\": %d, \"layout_bytes_per_edge\": %zu, \"n_waters\": %d, \"batch_frames\": %d, "
                    "\"timed_frames\": %llu, \"timed_s\": %.4f, \"edges_per_frame\": %.1f, \"in_frame_order\": %s, "
                    "\"h2d_bytes_per_frame\": %zu, \"d2h_bytes_per_frame\": %.0f, \"warmup_s\": %.3f, \"stats_rows_reduced\": %zu, "
                    "\"find_wall_ms\": %.3f, \"find_into_wall_ms\": %.3f, \"find_pairs\": %zu, \"checksum\": %.6g, \"frames_per_gpu\": [%s], \"batch_call_ms\": [%s], "
                    "\"last_call\": {\"h2d_ms\": %.2f, \"kernels_ms\": %.2f, \"d2h_ms\": %.2f, \"passes\": %d}}\n",
                    (double)timed_frames / t_run, gpus, mode.c_str(), copy_threads, dynamic, esz, waters, batch, (unsigned long long)timed_frames, t_run,
                    timed_frames ? (double)timed_edges / (double)timed_frames : 0.0, in_order ? "true" : "false",
                    frame_floats * sizeof(float),
                    timed_frames ? (double)timed_edges * (double)esz / (double)timed_frames + 4.0 * waters + 48.0 : 0.0,
                    t_first, reduced_rows, find_ms, find_into_ms, n_pairs, checksum, pg.c_str(), bs.c_str(), last_tm.h2d_ms, last_tm.total_ms, last_tm.d2h_ms,
                    last_tm.passes);
        return in_order ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "wn_bench: %s\n", e.what());
        return 1;
    }
}
