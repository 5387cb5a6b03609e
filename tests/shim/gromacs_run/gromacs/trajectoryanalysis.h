// tests/shim/gromacs_run/gromacs/trajectoryanalysis.h -- TEST INFRASTRUCTURE ONLY.
//
// A small WORKING miniature of the part of the GROMACS 2019 trajectory-analysis API that
// gmx-acqueduct_b200/host/gromacs/waternetwork-b200.cpp uses (same declarations as the declaration-only shim in
// tests/shim/gromacs, written from the public API documentation), so that the gated module can be compiled, linked
// and RUN on a .gro trajectory where GROMACS itself is absent -- against the CPU stand-in of the C ABI (tests/mock).
// What it models: option registration and "-name value" parsing (file-name options of type eftPlot get ".xvg"
// appended when they carry no extension, as GROMACS completes them), the default selections "Protein" / "Water"
// (by residue name), topology and frames from a multi-frame .gro file, AnalysisData + handles + the plot module
// (data rows " %11.3f" time, " %8.3f" values; frames must be finished in index order, as in serial GROMACS), the
// frame-local module data object, and the runner's call order: initOptions, optionsFinished, initAnalysis,
// startFrames, analyzeFrame per frame, ModuleData::finish, finishAnalysis, writeOutput.
// It proves the module's own logic (frame layout, site table, batching, late data-set rows) -- nothing about the
// real GROMACS headers or runtime beyond what is restated here.
#ifndef WN_TEST_SHIM_GROMACS_RUN_TRAJECTORYANALYSIS_H
#define WN_TEST_SHIM_GROMACS_RUN_TRAJECTORYANALYSIS_H

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

typedef float real;
typedef real rvec[3];
typedef real matrix[3][3];
enum { XX = 0, YY = 1, ZZ = 2 };

struct t_atom { int resind; char elem[4]; };
struct t_resinfo { char** name; };
struct t_atoms { int nr; t_atom* atom; char*** atomname; int nres; t_resinfo* resinfo; };
struct t_trxframe { real time; matrix box; rvec* x; int natoms; };
struct t_pbc;

namespace gmx
{
enum { eftPlot = 1 };

template <typename T> class ArrayRef
{
public:
    ArrayRef() : p_(nullptr), n_(0) {}
    ArrayRef(const T* p, size_t n) : p_(p), n_(n) {}
    const T* begin() const { return p_; }
    const T* end() const { return p_ + n_; }
    size_t size() const { return n_; }
    const T& operator[](size_t i) const { return p_[i]; }
    const T& at(size_t i) const { if (i >= n_) throw std::out_of_range("ArrayRef::at"); return p_[i]; }
private:
    const T* p_;
    size_t n_;
};

class RVec
{
public:
    RVec() { v_[0] = v_[1] = v_[2] = 0; }
    RVec(real x, real y, real z) { v_[0] = x; v_[1] = y; v_[2] = z; }
    real operator[](int i) const { return v_[i]; }
private:
    real v_[3];
};

// ---- selections
struct SelectionData { std::vector<int> indices; std::vector<RVec> coords; };
class Selection
{
public:
    Selection() {}
    explicit Selection(std::shared_ptr<SelectionData> d) : d_(d) {}
    bool isValid() const { return (bool)d_; }
    int posCount() const { return d_ ? (int)d_->indices.size() : 0; }
    ArrayRef<const int> atomIndices() const { need(); return ArrayRef<const int>(d_->indices.data(), d_->indices.size()); }
    ArrayRef<const RVec> coordinates() const { need(); return ArrayRef<const RVec>(d_->coords.data(), d_->coords.size()); }
    std::shared_ptr<SelectionData> data() const { return d_; }
private:
    void need() const { if (!d_) throw std::logic_error("gmx shim: selection used before it was set"); }
    std::shared_ptr<SelectionData> d_;
};
class SelectionCollection {};

// ---- options
struct OptionRecord {
    std::string name, kind, defaultText, defaultBasename;
    void* store = nullptr;
    bool isRequired = false, isFile = false;
    int fileType = 0;
};
template <typename Self, typename T> class OptionTemplate
{
public:
    Self& store(T* p) { rec_.store = p; return self(); }
    Self& description(const char*) { return self(); }
    Self& required(bool b = true) { rec_.isRequired = b; return self(); }
    const OptionRecord& record() const { return rec_; }
protected:
    OptionTemplate(const char* name, const char* kind) { rec_.name = name; rec_.kind = kind; }
    Self& self() { return static_cast<Self&>(*this); }
    OptionRecord rec_;
};
class FileNameOption : public OptionTemplate<FileNameOption, std::string>
{
public:
    explicit FileNameOption(const char* name) : OptionTemplate(name, "file") { rec_.isFile = true; }
    FileNameOption& filetype(int t) { rec_.fileType = t; return *this; }
    FileNameOption& outputFile() { return *this; }
    FileNameOption& defaultBasename(const char* b) { rec_.defaultBasename = b; return *this; }
};
class SelectionOption : public OptionTemplate<SelectionOption, Selection>
{
public:
    explicit SelectionOption(const char* name) : OptionTemplate(name, "selection") {}
    SelectionOption& defaultSelectionText(const char* t) { rec_.defaultText = t; return *this; }
};
class DoubleOption : public OptionTemplate<DoubleOption, double> { public: explicit DoubleOption(const char* name) : OptionTemplate(name, "double") {} };
class IntegerOption : public OptionTemplate<IntegerOption, int> { public: explicit IntegerOption(const char* name) : OptionTemplate(name, "int") {} };
class BooleanOption : public OptionTemplate<BooleanOption, bool> { public: explicit BooleanOption(const char* name) : OptionTemplate(name, "bool") {} };

class IOptionsContainer
{
public:
    template <typename Opt> void addOption(const Opt& o)
    {
        if (!o.record().store) throw std::logic_error("gmx shim: option -" + o.record().name + " has no store()");
        for (const OptionRecord& r : records) if (r.name == o.record().name) throw std::logic_error("gmx shim: duplicate option -" + r.name);
        records.push_back(o.record());
    }
    std::vector<OptionRecord> records;
};

class AnalysisDataPlotSettings {};
class TrajectoryAnalysisSettings
{
public:
    enum { efRequireTop = 1, efUseTopX = 2 };
    template <size_t N> void setHelpText(const char* const (&)[N]) {}
    void setFlag(int flag, bool on = true) { if (on) flags_ |= flag; else flags_ &= ~flag; }
    bool hasFlag(int flag) const { return (flags_ & flag) != 0; }
    const AnalysisDataPlotSettings& plotSettings() const { return plot_; }
private:
    int flags_ = 0;
    AnalysisDataPlotSettings plot_;
};

// ---- topology of a .gro file in t_atoms form
class TopologyInformation
{
public:
    const t_atoms* atoms() const { return &atoms_; }
    void build(const std::vector<std::string>& atomName, const std::vector<std::string>& resNameAtom, const std::vector<int>& resNr)
    {
        const size_t n = atomName.size();
        names_ = atomName;
        namePtr_.resize(n); namePtrPtr_.resize(n); atom_.resize(n);
        resNames_.clear();
        for (size_t a = 0; a < n; ++a) {
            if (a == 0 || resNr[a] != resNr[a - 1] || resNameAtom[a] != resNameAtom[a - 1]) resNames_.push_back(resNameAtom[a]);
            atom_[a].resind = (int)resNames_.size() - 1;
            std::memset(atom_[a].elem, 0, sizeof atom_[a].elem);
            for (char ch : atomName[a]) if (!(ch >= '0' && ch <= '9')) { atom_[a].elem[0] = ch; break; }    // a .gro carries no elements
        }
        for (size_t a = 0; a < n; ++a) { namePtr_[a] = &names_[a][0]; namePtrPtr_[a] = &namePtr_[a]; }
        resPtr_.resize(resNames_.size()); res_.resize(resNames_.size());
        for (size_t r = 0; r < resNames_.size(); ++r) { resPtr_[r] = &resNames_[r][0]; res_[r].name = &resPtr_[r]; }
        atoms_.nr = (int)n; atoms_.atom = atom_.data(); atoms_.atomname = namePtrPtr_.data();
        atoms_.nres = (int)res_.size(); atoms_.resinfo = res_.data();
    }
private:
    t_atoms atoms_{0, nullptr, nullptr, 0, nullptr};
    std::vector<std::string> names_, resNames_;
    std::vector<char*> namePtr_, resPtr_;
    std::vector<char**> namePtrPtr_;
    std::vector<t_atom> atom_;
    std::vector<t_resinfo> res_;
};

// ---- analysis data
class AnalysisDataPlotModule
{
public:
    explicit AnalysisDataPlotModule(const AnalysisDataPlotSettings&) {}
    void setFileName(const std::string& fn) { fn_ = fn; }
    void setTitle(const char* t) { title_ = t; }
    void setXAxisIsTime() { xIsTime_ = true; }
    void setYLabel(const char* l) { ylabel_ = l; }
    void write(const std::vector<std::pair<real, std::vector<real> > >& rows) const
    {
        std::ofstream os(fn_);
        os << "# written by the test shim of the GROMACS plot module\n@    title \"" << title_ << "\"\n@    xaxis  label \"" << (xIsTime_ ? "Time (ps)" : "")
           << "\"\n@    yaxis  label \"" << ylabel_ << "\"\n@TYPE xy\n";
        char buf[64];
        for (const auto& r : rows) {
            std::snprintf(buf, sizeof buf, " %11.3f", (double)r.first); os << buf;
            for (real v : r.second) { std::snprintf(buf, sizeof buf, " %8.3f", (double)v); os << buf; }
            os << "\n";
        }
    }
private:
    std::string fn_, title_, ylabel_;
    bool xIsTime_ = false;
};
typedef std::shared_ptr<AnalysisDataPlotModule> AnalysisDataPlotModulePointer;

class AnalysisData
{
public:
    void setColumnCount(int dataSet, int columns) { if (dataSet != 0) throw std::logic_error("gmx shim: one data set per AnalysisData"); columns_ = columns; }
    void addModule(AnalysisDataPlotModulePointer m) { modules_.push_back(m); }
    // used by the handle / the runner
    int columns_ = 0;
    std::vector<std::pair<real, std::vector<real> > > rows_;
    int nextFrame_ = 0;
    bool open_ = false;
    std::vector<AnalysisDataPlotModulePointer> modules_;
    void finished() { for (auto& m : modules_) m->write(rows_); }
};
class AnalysisDataHandle
{
public:
    AnalysisDataHandle() : d_(nullptr) {}
    explicit AnalysisDataHandle(AnalysisData* d) : d_(d) {}
    void startFrame(int index, real x, real /*dx*/ = 0.0)
    {
        if (!d_) throw std::logic_error("gmx shim: invalid data handle");
        if (d_->open_) throw std::logic_error("gmx shim: startFrame inside a frame");
        if (index != d_->nextFrame_) throw std::logic_error("gmx shim: data frames must be added in index order (serial mode)");
        d_->open_ = true;
        d_->rows_.push_back(std::make_pair(x, std::vector<real>((size_t)d_->columns_, (real)0)));
    }
    void setPoint(int column, real value, bool /*present*/ = true)
    {
        if (!d_ || !d_->open_) throw std::logic_error("gmx shim: setPoint outside a frame");
        if (column < 0 || column >= d_->columns_) throw std::out_of_range("gmx shim: data column");
        d_->rows_.back().second[(size_t)column] = value;
    }
    void finishFrame()
    {
        if (!d_ || !d_->open_) throw std::logic_error("gmx shim: finishFrame outside a frame");
        d_->open_ = false; ++d_->nextFrame_;
    }
private:
    AnalysisData* d_;
};
class AnalysisDataParallelOptions {};

class TrajectoryAnalysisModule;
class TrajectoryAnalysisModuleData
{
public:
    virtual ~TrajectoryAnalysisModuleData() {}
    virtual void finish() = 0;
    AnalysisDataHandle dataHandle(const AnalysisData& data)
    {
        if (handlesFinished_) throw std::logic_error("gmx shim: dataHandle after finishDataHandles");
        return AnalysisDataHandle(const_cast<AnalysisData*>(&data));
    }
    Selection parallelSelection(const Selection& selection) { return selection; }
    bool handlesFinished() const { return handlesFinished_; }
protected:
    TrajectoryAnalysisModuleData(TrajectoryAnalysisModule*, const AnalysisDataParallelOptions&, const SelectionCollection&) {}
    void finishDataHandles() { handlesFinished_ = true; }
private:
    bool handlesFinished_ = false;
};
typedef std::unique_ptr<TrajectoryAnalysisModuleData> TrajectoryAnalysisModuleDataPointer;

class TrajectoryAnalysisModule
{
public:
    virtual ~TrajectoryAnalysisModule() {}
    virtual void initOptions(IOptionsContainer* options, TrajectoryAnalysisSettings* settings) = 0;
    virtual void optionsFinished(TrajectoryAnalysisSettings*) {}
    virtual void initAnalysis(const TrajectoryAnalysisSettings& settings, const TopologyInformation& top) = 0;
    virtual TrajectoryAnalysisModuleDataPointer startFrames(const AnalysisDataParallelOptions& opt, const SelectionCollection& selections) = 0;
    virtual void analyzeFrame(int frnr, const t_trxframe& fr, t_pbc* pbc, TrajectoryAnalysisModuleData* pdata) = 0;
    virtual void finishAnalysis(int) {}
    virtual void writeOutput() = 0;
    std::vector<AnalysisData*> datasets_;
protected:
    TrajectoryAnalysisModule() {}
    void registerAnalysisDataset(AnalysisData* data, const char*) { datasets_.push_back(data); }
};

// ---- the command-line runner on a multi-frame .gro file
namespace shim_detail
{
inline std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r' || s[b - 1] == '\n')) --b;
    return s.substr(a, b - a);
}
struct GroFrame { std::vector<float> xyz; matrix box; double time; bool hasTime; };
inline bool readGro(std::istream& in, GroFrame& fr, std::vector<std::string>* atomName, std::vector<std::string>* resName, std::vector<int>* resNr)
{
    std::string title, line;
    if (!std::getline(in, title)) return false;
    if (trim(title).empty() && in.peek() == EOF) return false;
    if (!std::getline(in, line)) return false;
    const long natoms = std::strtol(line.c_str(), nullptr, 10);
    fr.hasTime = false; fr.time = 0.0;
    const size_t tp = title.rfind("t=");
    if (tp != std::string::npos) { fr.time = std::strtod(title.c_str() + tp + 2, nullptr); fr.hasTime = true; }
    fr.xyz.assign((size_t)natoms * 3, 0.f);
    for (long a = 0; a < natoms; ++a) {
        if (!std::getline(in, line) || line.size() < 44) throw std::runtime_error("gmx shim: bad .gro atom line");
        if (atomName) {
            resNr->push_back((int)std::strtol(line.substr(0, 5).c_str(), nullptr, 10));
            resName->push_back(trim(line.substr(5, 5)));
            atomName->push_back(trim(line.substr(10, 5)));
        }
        const size_t p1 = line.find('.', 20), p2 = p1 == std::string::npos ? p1 : line.find('.', p1 + 1);
        const size_t w = (p1 != std::string::npos && p2 != std::string::npos) ? p2 - p1 : 8;
        for (int c = 0; c < 3; ++c) fr.xyz[(size_t)a * 3 + c] = (float)std::strtod(line.substr(20 + (size_t)c * w, w).c_str(), nullptr);
    }
    if (!std::getline(in, line)) throw std::runtime_error("gmx shim: missing .gro box line");
    double b[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    std::istringstream bs(line);
    for (int k = 0; k < 9 && (bs >> b[k]); ++k) {}
    fr.box[0][0] = (real)b[0]; fr.box[1][1] = (real)b[1]; fr.box[2][2] = (real)b[2];
    fr.box[0][1] = (real)b[3]; fr.box[0][2] = (real)b[4]; fr.box[1][0] = (real)b[5];
    fr.box[1][2] = (real)b[6]; fr.box[2][0] = (real)b[7]; fr.box[2][1] = (real)b[8];
    return true;
}
}  // namespace shim_detail

class TrajectoryAnalysisCommandLineRunner
{
public:
    template <class Module> static int runAsMain(int argc, char* argv[])
    {
        try {
            Module module;
            IOptionsContainer options;
            TrajectoryAnalysisSettings settings;
            module.initOptions(&options, &settings);
            // "-name value" pairs; booleans are "-name" / "-noname"; -f is the trajectory (and the topology)
            std::string traj;
            std::map<std::string, std::string> given;
            for (int k = 1; k < argc; ++k) {
                std::string a = argv[k];
                if (a.size() < 2 || a[0] != '-') throw std::runtime_error("gmx shim: expected an option, got " + a);
                a = a.substr(1);
                bool isBool = false;
                for (const OptionRecord& r : options.records)
                    if (r.kind == "bool" && (a == r.name || a == "no" + r.name)) { isBool = true; given[r.name] = a == r.name ? "1" : "0"; }
                if (isBool) continue;
                if (k + 1 >= argc) throw std::runtime_error("gmx shim: missing value after -" + a);
                if (a == "f") { traj = argv[++k]; continue; }
                bool known = false;
                for (const OptionRecord& r : options.records) known = known || r.name == a;
                if (!known) throw std::runtime_error("gmx shim: unknown option -" + a);
                given[a] = argv[++k];
            }
            if (traj.empty()) throw std::runtime_error("gmx shim: -f traj.gro is required");
            std::vector<std::shared_ptr<SelectionData> > selections;
            std::vector<std::string> selectionText;
            for (const OptionRecord& r : options.records) {
                const bool has = given.count(r.name) != 0;
                if (r.kind == "file") {
                    std::string v = has ? given[r.name] : std::string();          // optional output files stay empty
                    if (!v.empty() && r.fileType == eftPlot && v.find('.', v.find_last_of('/') == std::string::npos ? 0 : v.find_last_of('/')) == std::string::npos)
                        v += ".xvg";                                               // GROMACS completes the extension
                    *static_cast<std::string*>(r.store) = v;
                } else if (r.kind == "double") { if (has) *static_cast<double*>(r.store) = std::atof(given[r.name].c_str()); }
                else if (r.kind == "int") { if (has) *static_cast<int*>(r.store) = std::atoi(given[r.name].c_str()); }
                else if (r.kind == "bool") { if (has) *static_cast<bool*>(r.store) = given[r.name] == "1"; }
                else if (r.kind == "selection") {
                    const std::string text = has ? given[r.name] : r.defaultText;
                    if (text.empty() && r.isRequired) throw std::runtime_error("gmx shim: selection -" + r.name + " is required");
                    std::shared_ptr<SelectionData> d(new SelectionData());
                    selections.push_back(d); selectionText.push_back(text);
                    *static_cast<Selection*>(r.store) = Selection(d);
                }
            }
            module.optionsFinished(&settings);
            // topology and first frame
            std::ifstream in(traj);
            if (!in) throw std::runtime_error("gmx shim: cannot open " + traj);
            shim_detail::GroFrame fr;
            std::vector<std::string> atomName, resNameAtom;
            std::vector<int> resNr;
            if (!shim_detail::readGro(in, fr, &atomName, &resNameAtom, &resNr)) throw std::runtime_error("gmx shim: empty trajectory");
            TopologyInformation top;
            top.build(atomName, resNameAtom, resNr);
            static const std::set<std::string> water{"SOL", "WAT", "HOH", "TIP3", "SPC"};
            for (size_t s = 0; s < selections.size(); ++s) {
                for (size_t a = 0; a < atomName.size(); ++a) {
                    const bool isWater = water.count(resNameAtom[a]) != 0;
                    if ((selectionText[s] == "Water" && isWater) || (selectionText[s] == "Protein" && !isWater)) selections[s]->indices.push_back((int)a);
                }
                if (selectionText[s] != "Water" && selectionText[s] != "Protein") throw std::runtime_error("gmx shim: only the selections Water and Protein are modelled");
            }
            module.initAnalysis(settings, top);
            AnalysisDataParallelOptions popt;
            SelectionCollection coll;
            TrajectoryAnalysisModuleDataPointer pdata = module.startFrames(popt, coll);
            int frnr = 0;
            std::vector<RVec> x;
            do {
                if (fr.xyz.size() != atomName.size() * 3) throw std::runtime_error("gmx shim: atom count changes between frames");
                for (auto& d : selections) {
                    d->coords.clear();
                    for (int a : d->indices) d->coords.push_back(RVec(fr.xyz[(size_t)a * 3], fr.xyz[(size_t)a * 3 + 1], fr.xyz[(size_t)a * 3 + 2]));
                }
                t_trxframe tf;
                tf.time = (real)(fr.hasTime ? fr.time : (double)frnr);
                std::memcpy(tf.box, fr.box, sizeof(matrix));
                tf.x = reinterpret_cast<rvec*>(fr.xyz.data());
                tf.natoms = (int)atomName.size();
                module.analyzeFrame(frnr, tf, nullptr, pdata.get());
                ++frnr;
            } while (shim_detail::readGro(in, fr, nullptr, nullptr, nullptr));
            pdata->finish();
            if (!pdata->handlesFinished()) throw std::logic_error("gmx shim: ModuleData::finish did not call finishDataHandles");
            module.finishAnalysis(frnr);
            for (AnalysisData* d : module.datasets_) {
                if (d->open_) throw std::logic_error("gmx shim: a data frame was left open");
                if (!d->modules_.empty() && d->nextFrame_ != frnr) throw std::logic_error("gmx shim: a data set is missing frames");
                d->finished();
            }
            module.writeOutput();
            return 0;
        } catch (const std::exception& e) {
            std::cerr << "Program aborted: " << e.what() << std::endl;
            return 1;
        }
    }
};
}  // namespace gmx

#endif
