// waternetwork-driver.cpp -- the waternetwork analysis as a standalone program, without GROMACS.
//
// Walks a multi-frame .gro trajectory through the flow of the reference's trajectory-analysis module
// (reference src/waternetwork.cpp): initAnalysis (:392-500: makeSites for the protein and the solvent
// selection, solvent ids offset, nodes file, edges file header), analyzeFrame (:503-667: site list =
// protein sites then waters, pair search, make_edges, edge rows, the two statistics rows) and
// finishAnalysis (:670-676), with the search/geometry/energy stage on the GPU through
// CudaFindPairs' sibling HydrogenBondEdges + FrameBatcher.  It exists because the real
// gmx::TrajectoryAnalysisModule glue cannot be built where GROMACS is absent; the files it writes
// are the ones python/network-analysis.py reads (nodes-list.xvg.dat, edges-frames.xvg.dat).
//
// What it does not do (out of scope here): GROMACS selections (the solvent is chosen by residue
// name, the protein is everything else), the alpha-shape filter of buried waters (as_->locate, :533 --
// a per-frame list of water indices can be supplied instead with -buried), trajectory formats other
// than .gro, the xvg legend headers of AnalysisDataPlotModule (data rows only, " %11.3f" time and
// " %8.3f" values, the plot module's default formats).
//
//   wn_waternetwork -f traj.gro [-edges edges-frames.xvg] [-nodes nodes-list.xvg]
//                   [-ohb hydrogen-bonds-num.xvg] [-obw buried-waters-num.xvg]
//                   [-solvent SOL,WAT,HOH,TIP3] [-buried lists.txt] [-pbc] [-batch 256] [-device 0] [-gpus N]
//                   [-don 5.5] [-doff 6.5] [-aon 0.25] [-aoff 0.0] [-bug-compat 0|1] [-gpu-text 1|0]
//                   [-components threshold [-ocomp components-num.xvg]]
//                   [-b t0] [-e t1] [-dt step] [-frnr0 n] [-append 1]
// -b/-e/-dt: the frame selection GROMACS' runner gives every analysis tool (first / last time to analyse, only
// frames whose time is a multiple of -dt after the first frame's); frames are numbered among the ANALYSED frames, from
// -frnr0 (default 0), as analyzeFrame's frnr is.  -append 1 continues the files of an earlier run (no headers, no
// nodes file, rows appended): outputs are per frame, so "-e t" followed by "-b t' -frnr0 n -append 1" writes the same
// bytes as one run over everything (the restart lever the reference inherits from the runner).
// -components: per frame, the connected components of the network after dropping the edges with
// |energy| <= threshold (the first step of python/network-analysis.py, computed on the GPU): one row
// "time components largest sites edges" per frame.
// -gpus N (N > 1): the trajectory is sharded frame-wise over GPUs 0..N-1 of this box -- consecutive blocks of
// -batch frames go round-robin to the GPUs (one context, one staging pipeline and one worker thread per GPU),
// the edge rows are written in frame order from the per-GPU results, and the rows of hydrogen-bonds-num.xvg
// come from the NCCL sum of the per-GPU statistics tables (the only exchange between the GPUs).  The files are
// byte-identical to a one-GPU run.  -components and the GPU text formatter are one-GPU features.
// File name options take the reference's base names: ".dat" is appended to -edges/-nodes (:461,:485).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "edge-file-writer.hpp"
#include "hydrogen-bond-edges.hpp"
#include "make-sites.hpp"

namespace {

struct GroFrame {
    std::vector<float> xyz;          // natoms * 3, nm
    float box[3][3];                 // GROMACS matrix (rows = box vectors)
    double time;
    bool has_time;
};

std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\n')) --b;
    return s.substr(a, b - a);
}

// One frame of a .gro file.  Fixed columns: resnr 0-5, resname 5-10, atom name 10-15, atom nr 15-20,
// then the coordinates in equal-width fields (8.3 by default; the width is taken from the distance
// between the decimal points, as GROMACS does).  names/resnr are filled from the first frame only.
bool readGroFrame(std::istream& in, GroFrame& fr, std::vector<std::string>* atomName, std::vector<std::string>* resName,
                  std::vector<int>* resNr)
{
    std::string title, line;
    if (!std::getline(in, title)) return false;
    if (trim(title).empty() && in.peek() == EOF) return false;
    if (!std::getline(in, line)) return false;
    const long natoms = std::strtol(line.c_str(), nullptr, 10);
    if (natoms < 0) throw std::runtime_error(".gro: bad atom count line: " + line);
    fr.has_time = false; fr.time = 0.0;
    const size_t tp = title.rfind("t=");
    if (tp != std::string::npos) { fr.time = std::strtod(title.c_str() + tp + 2, nullptr); fr.has_time = true; }
    fr.xyz.assign((size_t)natoms * 3, 0.f);
    if (atomName) { atomName->clear(); resName->clear(); resNr->clear(); }
    for (long a = 0; a < natoms; ++a) {
        if (!std::getline(in, line)) throw std::runtime_error(".gro: file ends inside a frame");
        if (line.size() < 44) throw std::runtime_error(".gro: short atom line: " + line);
        if (atomName) {
            resNr->push_back((int)std::strtol(line.substr(0, 5).c_str(), nullptr, 10));
            resName->push_back(trim(line.substr(5, 5)));
            atomName->push_back(trim(line.substr(10, 5)));
        }
        const size_t p1 = line.find('.', 20), p2 = p1 == std::string::npos ? p1 : line.find('.', p1 + 1);
        const size_t w = (p1 != std::string::npos && p2 != std::string::npos) ? p2 - p1 : 8;
        for (int c = 0; c < 3; ++c)
            fr.xyz[(size_t)a * 3 + c] = (float)std::strtod(line.substr(20 + (size_t)c * w, w).c_str(), nullptr);
    }
    if (!std::getline(in, line)) throw std::runtime_error(".gro: missing box line");
    double b[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::istringstream bs(line);
    for (int k = 0; k < 9 && (bs >> b[k]); ++k) {}
    // v1(x) v2(y) v3(z) v1(y) v1(z) v2(x) v2(z) v3(x) v3(y)
    fr.box[0][0] = (float)b[0]; fr.box[1][1] = (float)b[1]; fr.box[2][2] = (float)b[2];
    fr.box[0][1] = (float)b[3]; fr.box[0][2] = (float)b[4];
    fr.box[1][0] = (float)b[5]; fr.box[1][2] = (float)b[6];
    fr.box[2][0] = (float)b[7]; fr.box[2][1] = (float)b[8];
    return true;
}

// element string of an atom from its name (a .gro file carries no elements): leading digits skipped,
// then the first letter -- "HW1", "1HH1" -> "H", "OW" -> "O".
std::string elementOf(const std::string& name)
{
    for (char ch : name)
        if (!(ch >= '0' && ch <= '9')) return std::string(1, ch);
    return std::string();
}

struct Options {
    std::string traj, edges = "edges-frames.xvg", nodes = "nodes-list.xvg", ohb = "hydrogen-bonds-num.xvg",
                obw = "buried-waters-num.xvg", buried, ocomp = "components-num.xvg";
    std::set<std::string> solvent{"SOL", "WAT", "HOH", "TIP3", "SPC"};
    bool pbc = false, bugCompat = true, gpuText = true;
    int batch = 256, device = 0, gpus = 1;
    double don = 5.5, doff = 6.5, aon = 0.25, aoff = 0.0;
    bool components = false;
    double compThreshold = 0.0;
    double begin = -1.0e300, end = 1.0e300, dt = 0.0;
    int frnr0 = 0;
    bool append = false;
};

Options parse(int argc, char** argv)
{
    Options o;
    for (int k = 1; k < argc; ++k) {
        const std::string a = argv[k];
        auto val = [&]() -> std::string {
            if (k + 1 >= argc) throw std::runtime_error("missing value after " + a);
            return argv[++k];
        };
        if (a == "-f") o.traj = val();
        else if (a == "-edges") o.edges = val();
        else if (a == "-nodes") o.nodes = val();
        else if (a == "-ohb") o.ohb = val();
        else if (a == "-obw") o.obw = val();
        else if (a == "-buried") o.buried = val();
        else if (a == "-pbc") o.pbc = true;
        else if (a == "-batch") o.batch = std::atoi(val().c_str());
        else if (a == "-device") o.device = std::atoi(val().c_str());
        else if (a == "-gpus") o.gpus = std::atoi(val().c_str());
        else if (a == "-don") o.don = std::atof(val().c_str());
        else if (a == "-doff") o.doff = std::atof(val().c_str());
        else if (a == "-aon") o.aon = std::atof(val().c_str());
        else if (a == "-aoff") o.aoff = std::atof(val().c_str());
        else if (a == "-bug-compat") o.bugCompat = std::atoi(val().c_str()) != 0;
        else if (a == "-gpu-text") o.gpuText = std::atoi(val().c_str()) != 0;
        else if (a == "-components") { o.components = true; o.compThreshold = std::atof(val().c_str()); }
        else if (a == "-ocomp") o.ocomp = val();
        else if (a == "-b") o.begin = std::atof(val().c_str());
        else if (a == "-e") o.end = std::atof(val().c_str());
        else if (a == "-dt") o.dt = std::atof(val().c_str());
        else if (a == "-frnr0") o.frnr0 = std::atoi(val().c_str());
        else if (a == "-append") o.append = std::atoi(val().c_str()) != 0;
        else if (a == "-solvent") {
            o.solvent.clear();
            std::istringstream ss(val());
            for (std::string t; std::getline(ss, t, ',');) if (!t.empty()) o.solvent.insert(t);
        } else throw std::runtime_error("unknown option " + a);
    }
    if (o.traj.empty()) throw std::runtime_error("usage: wn_waternetwork -f traj.gro [options] (see the file header)");
    if (o.batch < 1) o.batch = 1;
    if (o.gpus < 1) o.gpus = 1;
    if (o.gpus > 1 && o.components) throw std::runtime_error("-components needs -gpus 1 (it works on one context's last batch)");
    return o;
}

}  // namespace

int main(int argc, char** argv)
{
    try {
        const Options opt = parse(argc, argv);
        std::ifstream in(opt.traj);
        if (!in) throw std::runtime_error("cannot open " + opt.traj);

        // ---- topology from the first frame (initAnalysis, :392-425)
        GroFrame fr;
        std::vector<std::string> atomName, resNameAtom;
        std::vector<int> resNr;
        if (!readGroFrame(in, fr, &atomName, &resNameAtom, &resNr)) throw std::runtime_error("empty trajectory");
        const int natoms = (int)atomName.size();
        if (natoms == 0) throw std::runtime_error(".gro: the first frame has no atoms");
        TopologyView top;
        top.atomName = atomName;
        top.element.resize((size_t)natoms);
        top.resIndex.resize((size_t)natoms);
        for (int a = 0; a < natoms; ++a) {
            top.element[a] = elementOf(atomName[a]);
            if (a == 0 || resNr[a] != resNr[a - 1] || resNameAtom[a] != resNameAtom[a - 1]) top.resName.push_back(resNameAtom[a]);
            top.resIndex[a] = (int)top.resName.size() - 1;
        }
        std::vector<int> proteinSel, solventSel;
        for (int a = 0; a < natoms; ++a) (opt.solvent.count(resNameAtom[a]) ? solventSel : proteinSel).push_back(a);
        if (solventSel.size() % 3) throw std::runtime_error("solvent selection is not a whole number of 3-site waters");

        std::vector<SiteInfo> proteinSites, solventSites;
        makeSites(proteinSel, top, proteinSites);
        std::clog << "Protein has : " << proteinSites.size() << " Sites of out " << proteinSel.size() << " atoms." << std::endl;
        // The reference's makeSites walks TOPOLOGY indices (:150-166) but analyzeFrame reads the water w at
        // SELECTION position 3w (:564-575): the solvent sites are the oxygens of the selection in order.
        {
            std::vector<SiteInfo> all;
            makeSites(solventSel, top, all);
            solventSites = all;
        }
        std::clog << "Solvent has : " << solventSites.size() << " Sites of out " << solventSel.size() << " atoms." << std::endl;
        if (solventSites.size() * 3 != solventSel.size())
            throw std::runtime_error("solvent: expected one site (OW/O) per 3-atom water");
        offsetSolventIds(solventSites, proteinSites.size());
        std::clog << "Total number of sites : " << proteinSites.size() + solventSites.size() << std::endl;

        // ---- site table: protein sites, then every water (:547-576); atom indices into the frame as read
        const size_t nP = proteinSites.size(), nW = solventSites.size();
        std::vector<SiteInfo> table;
        std::vector<std::vector<int> > hydrogens;
        for (const SiteInfo& s : proteinSites) {
            SiteInfo t = s;
            std::vector<int> h;
            for (int hi : proteinHydrogenAtoms(s, opt.bugCompat)) {          // positions in the protein selection (:554)
                if (hi < 0 || (size_t)hi >= proteinSel.size()) throw std::out_of_range("protein hydrogen index (proteinPoints.at)");
                h.push_back(proteinSel[(size_t)hi]);
            }
            if ((size_t)s.atmIndex >= proteinSel.size()) throw std::out_of_range("protein site index (proteinPoints.at)");
            t.atmIndex = (unsigned)proteinSel[(size_t)s.atmIndex];             // proteinPoints.at(info.atmIndex) (:559)
            table.push_back(t);
            hydrogens.push_back(h);
        }
        for (size_t w = 0; w < nW; ++w) {
            SiteInfo t = solventSites[w];
            std::vector<int> h;
            for (int hi : waterHydrogenAtoms((int)w, solventSites[w])) h.push_back(solventSel[(size_t)hi]);   // :569
            t.atmIndex = (unsigned)solventSel[3 * w];                                                        // :575
            table.push_back(t);
            hydrogens.push_back(h);
        }
        std::vector<unsigned> ids;
        for (const SiteInfo& s : table) ids.push_back(s.id);

        // ---- nodes file (:461-482) and edges header (:485-499)
        const std::ios::openmode mode = opt.append ? std::ios::out | std::ios::app : std::ios::out;
        if (!opt.nodes.empty() && !opt.append) {
            std::ofstream nf(opt.nodes + ".dat");
            EdgeFileWriter::writeNodesHeader(nf);
            for (const SiteInfo& s : proteinSites) EdgeFileWriter::writeNode(nf, s);
            for (const SiteInfo& s : solventSites) EdgeFileWriter::writeNode(nf, s);
        }
        std::ofstream ef;
        std::unique_ptr<EdgeFileWriter> writer;
        if (!opt.edges.empty()) {
            ef.open(opt.edges + ".dat", mode);
            writer.reset(new EdgeFileWriter(ef, ids));
            if (!opt.append) writer->writeHeader();
        }
        std::ofstream ohb, obw;
        if (!opt.ohb.empty()) ohb.open(opt.ohb, mode);
        if (!opt.obw.empty()) obw.open(opt.obw, mode);

        // ---- optional per-frame water lists (what as_->locate leaves in filteredPoints, :533)
        std::ifstream buried;
        if (!opt.buried.empty()) { buried.open(opt.buried); if (!buried) throw std::runtime_error("cannot open " + opt.buried); }

        // one engine (context) per GPU; the site table and the energy parameters are replicated
        std::vector<std::unique_ptr<HydrogenBondEdges> > engines;
        for (int g = 0; g < opt.gpus; ++g) {
            engines.emplace_back(new HydrogenBondEdges(opt.gpus > 1 ? g : opt.device));
            engines.back()->setRadiusShift((float)opt.don, (float)opt.doff);
            engines.back()->setAngleShift((float)opt.aon, (float)opt.aoff);
            engines.back()->setSites(table, hydrogens);
        }
        HydrogenBondEdges& hb = *engines[0];

        std::vector<double> times;
        std::vector<size_t> filtered;
        std::vector<std::vector<int32_t> > frameRows;      // -buried: the site list of every pending frame
        char row[256];
        uint64_t total_edges = 0;
        bool batch_text_done = false;      // the batch's rows were formatted on the GPU and written in one piece
        const bool sharded = opt.gpus > 1;
        auto frameSink = [&](int frnr, const wn_edge* e, size_t n, const wn_frame_stats& st) {
                if (writer && !batch_text_done) {                                   // :638-652 (host formatter)
                    if (buried.is_open()) {
                        /* edge indices are positions in this frame's site list: back to table rows for the ids */
                        const std::vector<int32_t>& rows = frameRows[(size_t)(frnr - opt.frnr0)];
                        std::vector<wn_edge> mapped(e, e + n);
                        for (wn_edge& m : mapped) { m.i = rows[(size_t)m.i]; m.j = rows[(size_t)m.j]; }
                        writer->writeFrame(frnr, mapped.data(), n);
                        std::vector<int32_t>().swap(frameRows[(size_t)(frnr - opt.frnr0)]);
                    } else {
                        writer->writeFrame(frnr, e, n);
                    }
                }
                total_edges += n;
                const double t = times[(size_t)(frnr - opt.frnr0)];
                if (obw.is_open()) {                                                 // :654-657
                    std::snprintf(row, sizeof row, " %11.3f %8.3f %8.3f\n", t, (double)filtered[(size_t)(frnr - opt.frnr0)], 0.0);
                    obw << row;
                }
                if (ohb.is_open() && !sharded) {                                     // :659-666
                    std::snprintf(row, sizeof row, " %11.3f %8.3f %8.3f %8.3f %8.3f %8.3f %8.3f\n", t, st.num_sites,
                                  st.num_edges, st.ratio, 0.0, 0.0, 0.0);
                    ohb << row;
                }
            };
        std::unique_ptr<FrameBatcher> batcherPtr;
        if (!sharded) {
            batcherPtr.reset(new FrameBatcher(hb, natoms, opt.batch, frameSink));
        } else {
            std::vector<HydrogenBondEdges*> eng;
            for (auto& e : engines) eng.push_back(e.get());
            batcherPtr.reset(new FrameBatcher(eng, natoms, opt.batch, [&](const FrameResult& r) {
                frameSink(r.frnr, static_cast<const wn_edge*>(r.edges), r.n_edges, r.stats);
            }, WN_LAYOUT_EDGES));
        }
        FrameBatcher& batcher = *batcherPtr;

        // edge rows of a whole batch from the GPU formatter (same bytes); a batch with a field wider than its
        // column (ids over 6 digits, values over 8 characters) goes through the host writer instead
        std::ofstream ocomp;
        if (opt.components && !opt.ocomp.empty()) ocomp.open(opt.ocomp, mode);
        if (!sharded) batcher.setBatchHook([&](const std::vector<int>& frameNumbers, const HydrogenBondBatch&) {
            if (ocomp.is_open()) {
                const int32_t* labels = nullptr; const wn_component_stats* cs = nullptr;
                hb.componentsOfLastBatch((float)opt.compThreshold, &labels, &cs);
                for (size_t k = 0; k < frameNumbers.size(); ++k) {
                    std::snprintf(row, sizeof row, " %11.3f %8d %8d %8d %8d\n", times[(size_t)(frameNumbers[k] - opt.frnr0)], cs[k].n_components,
                                  cs[k].largest, cs[k].n_nodes, cs[k].n_edges);
                    ocomp << row;
                }
            }
            batch_text_done = false;
            if (!writer || !opt.gpuText || buried.is_open()) return;     // site lists: ids differ per frame, host path
            const char* text = nullptr; uint64_t bytes = 0;
            if (hb.formatLastBatch(frameNumbers, ids, &text, &bytes)) {
                ef.write(text, (std::streamsize)bytes);
                ef.flush();
                batch_text_done = true;
            }
        });

        int frnr = opt.frnr0;
        long fileFrame = -1;
        double tFirst = 0.0;
        do {
            if ((int)(fr.xyz.size() / 3) != natoms) throw std::runtime_error(".gro: atom count changes between frames");
            ++fileFrame;
            // frame selection of the GROMACS runner (-b, -e, -dt); a file without times counts frames
            const double tFrame = fr.has_time ? fr.time : (double)fileFrame;
            if (fileFrame == 0) tFirst = tFrame;
            std::string buriedLine;
            if (buried.is_open() && !std::getline(buried, buriedLine)) throw std::runtime_error("-buried: fewer lines than frames");
            if (tFrame > opt.end) break;
            bool use = tFrame >= opt.begin;
            if (use && opt.dt > 0.0) {
                const double q = (tFrame - tFirst) / opt.dt;
                use = std::fabs(q - std::floor(q + 0.5)) < 1e-6;
            }
            if (!use) continue;
            times.push_back(tFrame);
            const float (*box)[3] = opt.pbc ? fr.box : nullptr;
            if (buried.is_open()) {
                const std::string& line = buriedLine;
                std::vector<int32_t> rows;
                for (size_t s = 0; s < nP; ++s) rows.push_back((int32_t)s);
                std::istringstream ls(line);
                size_t cnt = 0;
                for (long w; ls >> w; ++cnt) {
                    if (w < 0 || (size_t)w >= nW) throw std::out_of_range("-buried: water index (solventSites_.at)");
                    rows.push_back((int32_t)(nP + (size_t)w));
                }
                filtered.push_back(cnt);
                frameRows.resize((size_t)(frnr - opt.frnr0) + 1);
                frameRows[(size_t)(frnr - opt.frnr0)] = rows;
                batcher.push(frnr, fr.xyz.data(), box, rows);
            } else {
                filtered.push_back(nW);
                if (box) batcher.push(frnr, fr.xyz.data(), box); else batcher.push(frnr, fr.xyz.data());
            }
            ++frnr;
        } while (readGroFrame(in, fr, nullptr, nullptr, nullptr));
        batcher.finish();                                                            // finishAnalysis (:670)
        if (sharded && ohb.is_open()) {
            // rows of every frame from the NCCL-summed statistics table (each row was filled by exactly one GPU)
            const std::vector<wn_frame_stats>& all = batcher.reducedStats();
            for (size_t f = 0; f < all.size(); ++f) {
                std::snprintf(row, sizeof row, " %11.3f %8.3f %8.3f %8.3f %8.3f %8.3f %8.3f\n", times[f], all[f].num_sites,
                              all[f].num_edges, all[f].ratio, 0.0, 0.0, 0.0);
                ohb << row;
            }
        }
        std::clog << frnr - opt.frnr0 << " frames, " << total_edges << " hydrogen bonds" << std::endl;
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "wn_waternetwork: " << e.what() << std::endl;
        return 1;
    }
}
