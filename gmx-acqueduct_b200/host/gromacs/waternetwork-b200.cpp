// waternetwork-b200.cpp -- the waternetwork trajectory-analysis module on the B200 path.
//
// COMPILE-GATED: built only with -DWN_WITH_GROMACS against a GROMACS 2019.x installation
// (`make -C gmx-acqueduct_b200 gromacs-module GMX_PREFIX=/path/to/gromacs`).  GROMACS is not present in
// the container this repository is developed in, so THIS FILE HAS NOT BEEN COMPILED AGAINST GROMACS.  What the test
// suite does instead: it compiles, links and RUNS this source against a small working miniature of the API calls it
// makes (tests/shim/gromacs_run, written from the public API documentation) and the CPU stand-in of the C ABI, on a
// protein-in-water .gro trajectory, and compares its nodes / edges files and data-set rows with those of
// host/waternetwork-driver.cpp (wn_waternetwork) -- the same flow without GROMACS, whose output files are
// byte-compared with the oracle (tests/test_host_mock_cpu.py, tests/test_gpu_host_cpp.py).
//
// It is the module of the reference (src/waternetwork.cpp: class WaterNetwork, include/waternetwork.hpp)
// with the per-frame body of analyzeFrame (:538-652) replaced by frame batching:
//   initOptions     same option names and defaults (:343-379).  -don/-doff/-aon/-aoff are parsed and IGNORED, as in
//                   the reference (computeEnergy is always called with its defaults 5.5/6.5/0.25/0.0),
//                   unless the new switch -switchopts is given -- then they reach the energy
//   differences from the reference binary a user must know: (1) the pair list is Brute's, not Delaunay's (twice the
//                   edges, lexicographic order: INTEGRATION.md section 0); (2) there is no alpha-shape filter, so
//                   every solvent water is a site and the "buried" data set holds the solvent count, not num_filtered;
//                   (3) EXPERIMENTAL until built and run against GROMACS 2019.x
//   initAnalysis    makeSites for both selections through the GROMACS-free restatement (include/make-sites.hpp),
//                   solvent ids offset (:420-423), nodes file and edges header (:461-499), site table upload
//   analyzeFrame    the frame's coordinates (protein selection, then solvent selection) are pushed into a
//                   FrameBatcher; with an alpha-shape filter the frame's site list is protein sites followed
//                   by the waters the filter keeps (:533, :562-576) -- the filter itself (CGAL alpha shapes) is
//                   outside this path, so this module uses every solvent water (FrameBatcher::push with a
//                   frameSites list is the hook for a filter)
//   ModuleData::finish / finishAnalysis  flush the last batch and its data-set rows, close the edges file (:670)
// Edge rows come from the GPU formatter once per batch (wn_format_edges), with the host writer as fallback.
#ifdef WN_WITH_GROMACS

#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include <gromacs/trajectoryanalysis.h>

#include "cuda-find-pairs.hpp"
#include "edge-file-writer.hpp"
#include "hydrogen-bond-edges.hpp"
#include "make-sites.hpp"

class WaterNetworkB200 : public gmx::TrajectoryAnalysisModule
{
public:
    WaterNetworkB200() : lengthOn_(5.5), lengthOff_(6.5), angleOn_(0.25), angleOff_(0.0), batchFrames_(256), usePbc_(false)
    {
        registerAnalysisDataset(&buriedData_, "buried");
        registerAnalysisDataset(&hydroBondData_, "hbond");
    }

    void initOptions(gmx::IOptionsContainer* options, gmx::TrajectoryAnalysisSettings* settings) override
    {
        static const char* const desc[] = {
            "Hydrogen-bond network of the solvent around a protein: per frame, the pairs of donor/acceptor",
            "sites within reach, the donor-hydrogen-acceptor geometry and a 6-4 energy for each bond.",
            "The pair search and the geometry run on the GPU, many frames per launch."};
        settings->setHelpText(desc);
        options->addOption(gmx::FileNameOption("obw").filetype(gmx::eftPlot).outputFile().store(&fnBuried_)
                               .defaultBasename("buried-waters-num").description("Waters used per frame"));
        options->addOption(gmx::FileNameOption("ohb").filetype(gmx::eftPlot).outputFile().store(&fnHydroBond_)
                               .defaultBasename("hydrogen-bonds-num").description("Sites, bonds and their ratio per frame"));
        options->addOption(gmx::FileNameOption("nodes").filetype(gmx::eftPlot).outputFile().store(&fnNodes_)
                               .defaultBasename("nodes-list").description("Donor/acceptor sites of the system"));
        options->addOption(gmx::FileNameOption("edges").filetype(gmx::eftPlot).outputFile().store(&fnEdges_)
                               .defaultBasename("edges-frames").description("Hydrogen bonds of every frame"));
        options->addOption(gmx::SelectionOption("select").store(&solventSel_).required()
                               .defaultSelectionText("Water").description("Solvent (3-site waters)"));
        options->addOption(gmx::SelectionOption("protein").store(&proteinSel_).required()
                               .defaultSelectionText("Protein").description("Solute"));
        options->addOption(gmx::DoubleOption("don").store(&lengthOn_).description("Radial switch starts here (Angstrom)"));
        options->addOption(gmx::DoubleOption("doff").store(&lengthOff_).description("Radial switch ends here (Angstrom)"));
        options->addOption(gmx::DoubleOption("aon").store(&angleOn_).description("Angle switch bound (cosine squared)"));
        options->addOption(gmx::DoubleOption("aoff").store(&angleOff_).description("Angle switch bound (cosine squared)"));
        options->addOption(gmx::BooleanOption("switchopts").store(&useSwitchOptions_)
                               .description("Apply -don/-doff/-aon/-aoff to the energy (the reference parses and ignores them)"));
        options->addOption(gmx::IntegerOption("batch").store(&batchFrames_).description("Frames per GPU batch"));
        options->addOption(gmx::BooleanOption("usepbc").store(&usePbc_).description("Minimum-image search in the frame's box"));
    }

    void optionsFinished(gmx::TrajectoryAnalysisSettings* settings) override
    {
        settings->setFlag(gmx::TrajectoryAnalysisSettings::efRequireTop);
        settings->setFlag(gmx::TrajectoryAnalysisSettings::efUseTopX);
    }

    void initAnalysis(const gmx::TrajectoryAnalysisSettings& settings, const gmx::TopologyInformation& top) override
    {
        // topology view for the GROMACS-free makeSites (names, elements, residues of every atom)
        const t_atoms* atoms = top.atoms();
        TopologyView view;
        for (int a = 0; a < atoms->nr; ++a) {
            view.atomName.push_back(*atoms->atomname[a]);
            view.element.push_back(atoms->atom[a].elem);
            view.resIndex.push_back(atoms->atom[a].resind);
        }
        for (int r = 0; r < atoms->nres; ++r) view.resName.push_back(*atoms->resinfo[r].name);

        const gmx::ArrayRef<const int> pIdx = proteinSel_.atomIndices(), sIdx = solventSel_.atomIndices();
        const std::vector<int> proteinAtoms(pIdx.begin(), pIdx.end()), solventAtoms(sIdx.begin(), sIdx.end());
        makeSites(proteinAtoms, view, proteinSites_);
        makeSites(solventAtoms, view, solventSites_);
        offsetSolventIds(solventSites_, proteinSites_.size());

        // frame layout handed to the batcher: protein selection coordinates, then solvent selection coordinates
        nProteinAtoms_ = (int)proteinAtoms.size();
        natoms_ = nProteinAtoms_ + (int)solventAtoms.size();
        std::vector<SiteInfo> table;
        std::vector<std::vector<int> > hydrogens;
        for (const SiteInfo& s : proteinSites_) {
            table.push_back(s);                                              // proteinPoints.at(info.atmIndex)
            hydrogens.push_back(proteinHydrogenAtoms(s, /*bugCompat*/ true)); // proteinPoints.at(info.id+i+1)
        }
        for (size_t w = 0; w < solventSites_.size(); ++w) {
            SiteInfo s = solventSites_[w];
            s.atmIndex = (unsigned)(nProteinAtoms_ + 3 * (int)w);            // solventPoints.at(3*index)
            std::vector<int> h = waterHydrogenAtoms((int)w, solventSites_[w]);
            for (int& a : h) a += nProteinAtoms_;                            // solventPoints.at(3*index+i)
            table.push_back(s);
            hydrogens.push_back(h);
        }
        for (const SiteInfo& s : table) ids_.push_back(s.id);

        finder_ = std::make_shared<CudaFindPairs>(0);        // the FindPairs drop-in (mp_); the batch path below fuses it
        engine_.reset(new HydrogenBondEdges(0));
        if (useSwitchOptions_) {                              // default: the reference's fixed 5.5 / 6.5 / 0.25 / 0.0
            engine_->setRadiusShift((float)lengthOn_, (float)lengthOff_);
            engine_->setAngleShift((float)angleOn_, (float)angleOff_);
        }
        engine_->setSites(table, hydrogens);

        buriedData_.setColumnCount(0, 2);
        hydroBondData_.setColumnCount(0, 6);
        addPlot(settings, buriedData_, fnBuried_, "Filter Statistics");
        addPlot(settings, hydroBondData_, fnHydroBond_, "Graph Statistics");

        if (!fnNodes_.empty()) {
            std::ofstream nodes(fnNodes_ + ".dat");
            EdgeFileWriter::writeNodesHeader(nodes);
            for (const SiteInfo& s : proteinSites_) EdgeFileWriter::writeNode(nodes, s);
            for (const SiteInfo& s : solventSites_) EdgeFileWriter::writeNode(nodes, s);
        }
        if (!fnEdges_.empty()) {
            edgeStream_.open(fnEdges_ + ".dat");
            writer_.reset(new EdgeFileWriter(edgeStream_, ids_));
            writer_->writeHeader();
        }

        batcher_.reset(new FrameBatcher(*engine_, natoms_, batchFrames_,
            [this](int frnr, const wn_edge* e, size_t n, const wn_frame_stats& st) {
                if (writer_ && !batchTextDone_) writer_->writeFrame(frnr, e, n);
                pending_.push_back(Row{frnr, st.num_sites, st.num_edges, st.ratio});
            }));
        batcher_->setBatchHook([this](const std::vector<int>& frameNumbers, const HydrogenBondBatch&) {
            batchTextDone_ = false;
            if (!writer_) return;
            const char* text = nullptr; uint64_t bytes = 0;
            if (engine_->formatLastBatch(frameNumbers, ids_, &text, &bytes)) {
                edgeStream_.write(text, (std::streamsize)bytes);
                batchTextDone_ = true;
            }
        });
    }

    void analyzeFrame(int frnr, const t_trxframe& fr, t_pbc* /*pbc*/, gmx::TrajectoryAnalysisModuleData* pdata) override
    {
        const gmx::Selection& proteinsel = pdata->parallelSelection(proteinSel_);
        const gmx::Selection& solventsel = pdata->parallelSelection(solventSel_);
        frame_.resize((size_t)natoms_ * 3);
        size_t k = 0;
        for (const gmx::RVec& x : proteinsel.coordinates()) { frame_[k++] = x[XX]; frame_[k++] = x[YY]; frame_[k++] = x[ZZ]; }
        for (const gmx::RVec& x : solventsel.coordinates()) { frame_[k++] = x[XX]; frame_[k++] = x[YY]; frame_[k++] = x[ZZ]; }
        times_.push_back(fr.time);
        if (usePbc_) batcher_->push(frnr, frame_.data(), fr.box); else batcher_->push(frnr, frame_.data());
        // data-set rows are written when their batch completes, in frame order; the frame loop of the
        // trajectory-analysis framework is sequential here (no efUseParallel), so handles can be taken late
        drain(pdata);
    }

    // The last batch completes after the last analyzeFrame call: its data-set rows are added from the
    // frame-local data object's finish(), which the runner calls after the frame loop and before
    // finishAnalysis() while the data handles are still open.
    class ModuleData : public gmx::TrajectoryAnalysisModuleData
    {
    public:
        ModuleData(WaterNetworkB200* owner, const gmx::AnalysisDataParallelOptions& opt,
                   const gmx::SelectionCollection& selections)
            : gmx::TrajectoryAnalysisModuleData(owner, opt, selections), owner_(owner) {}
        void finish() override
        {
            owner_->batcher_->finish();
            owner_->drain(this);
            finishDataHandles();
        }
    private:
        WaterNetworkB200* owner_;
    };

    gmx::TrajectoryAnalysisModuleDataPointer startFrames(const gmx::AnalysisDataParallelOptions& opt,
                                                         const gmx::SelectionCollection& selections) override
    {
        return gmx::TrajectoryAnalysisModuleDataPointer(new ModuleData(this, opt, selections));
    }

    void finishAnalysis(int /*nframes*/) override
    {
        if (edgeStream_.is_open()) edgeStream_.close();
    }

    void writeOutput() override {}

private:
    struct Row { int frnr; double sites, edges, ratio; };

    void addPlot(const gmx::TrajectoryAnalysisSettings& settings, gmx::AnalysisData& data, const std::string& fn,
                 const char* title)
    {
        if (fn.empty()) return;
        gmx::AnalysisDataPlotModulePointer plotm(new gmx::AnalysisDataPlotModule(settings.plotSettings()));
        plotm->setFileName(fn);
        plotm->setTitle(title);
        plotm->setXAxisIsTime();
        plotm->setYLabel("#");
        data.addModule(plotm);
    }

    // rows of completed batches -> the two data sets (reference :654-666)
    void drain(gmx::TrajectoryAnalysisModuleData* pdata)
    {
        if (pending_.empty()) return;
        gmx::AnalysisDataHandle buried = pdata->dataHandle(buriedData_), hbond = pdata->dataHandle(hydroBondData_);
        for (const Row& r : pending_) {
            const real t = (size_t)r.frnr < times_.size() ? times_[(size_t)r.frnr] : (real)r.frnr;
            buried.startFrame(r.frnr, t);
            buried.setPoint(0, (real)solventSites_.size());
            buried.setPoint(1, 0.0);
            buried.finishFrame();
            hbond.startFrame(r.frnr, t);
            hbond.setPoint(0, r.sites); hbond.setPoint(1, r.edges); hbond.setPoint(2, r.ratio);
            hbond.setPoint(3, 0.0); hbond.setPoint(4, 0.0); hbond.setPoint(5, 0.0);
            hbond.finishFrame();
        }
        pending_.clear();
    }

    std::string fnBuried_, fnHydroBond_, fnNodes_, fnEdges_;
    double lengthOn_, lengthOff_, angleOn_, angleOff_;
    int batchFrames_;
    bool usePbc_;
    bool useSwitchOptions_ = false;
    gmx::Selection solventSel_, proteinSel_;
    gmx::AnalysisData buriedData_, hydroBondData_;
    std::shared_ptr<FindPairs> finder_;                   // was std::shared_ptr<DelaunayFindPairs> mp_ (include/waternetwork.hpp:94)
    std::unique_ptr<HydrogenBondEdges> engine_;
    std::unique_ptr<FrameBatcher> batcher_;
    std::unique_ptr<EdgeFileWriter> writer_;
    std::vector<SiteInfo> solventSites_, proteinSites_;
    std::vector<unsigned> ids_;
    std::vector<float> frame_;
    std::vector<real> times_;
    std::vector<Row> pending_;
    std::ofstream edgeStream_;
    int nProteinAtoms_ = 0, natoms_ = 0;
    bool batchTextDone_ = false;
};

int main(int argc, char* argv[])
{
    return gmx::TrajectoryAnalysisCommandLineRunner::runAsMain<WaterNetworkB200>(argc, argv);
}

#endif  // WN_WITH_GROMACS
