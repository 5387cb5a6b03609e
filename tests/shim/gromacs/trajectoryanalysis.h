// tests/shim/gromacs/trajectoryanalysis.h -- TEST INFRASTRUCTURE ONLY.
//
// A declaration-only stand-in for the part of the GROMACS 2019 trajectory-analysis API that
// gmx-acqueduct_b200/host/gromacs/waternetwork-b200.cpp uses, written from the public API documentation
// (gromacs/trajectoryanalysis.h and the headers it pulls in).  GROMACS is absent here; this shim lets the
// CPU test suite run the compiler's front end over the gated module (`g++ -fsyntax-only -DWN_WITH_GROMACS`)
// so that typos and type errors against THESE declarations are caught.  It proves nothing about the real
// headers beyond what is restated here, and nothing links against it.
#ifndef WN_TEST_SHIM_GROMACS_TRAJECTORYANALYSIS_H
#define WN_TEST_SHIM_GROMACS_TRAJECTORYANALYSIS_H

#include <cstddef>
#include <memory>
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
    const T* begin() const;
    const T* end() const;
    size_t size() const;
    const T& operator[](size_t i) const;
    const T& at(size_t i) const;
};

class RVec
{
public:
    real operator[](int i) const;
};

template <typename Self, typename T> class OptionTemplate
{
public:
    Self& store(T* p);
    Self& description(const char* d);
    Self& required(bool b = true);
};
class FileNameOption : public OptionTemplate<FileNameOption, std::string>
{
public:
    explicit FileNameOption(const char* name);
    FileNameOption& filetype(int t);
    FileNameOption& outputFile();
    FileNameOption& defaultBasename(const char* b);
};
class Selection;
class SelectionOption : public OptionTemplate<SelectionOption, Selection>
{
public:
    explicit SelectionOption(const char* name);
    SelectionOption& defaultSelectionText(const char* t);
};
class DoubleOption : public OptionTemplate<DoubleOption, double> { public: explicit DoubleOption(const char* name); };
class IntegerOption : public OptionTemplate<IntegerOption, int> { public: explicit IntegerOption(const char* name); };
class BooleanOption : public OptionTemplate<BooleanOption, bool> { public: explicit BooleanOption(const char* name); };

class IOptionsContainer
{
public:
    template <typename Opt> void addOption(const Opt& o);
};

class AnalysisDataPlotSettings {};
class TrajectoryAnalysisSettings
{
public:
    enum { efRequireTop = 1, efUseTopX = 2 };
    template <size_t N> void setHelpText(const char* const (&text)[N]);
    void setFlag(int flag, bool on = true);
    const AnalysisDataPlotSettings& plotSettings() const;
};

class TopologyInformation
{
public:
    const t_atoms* atoms() const;
};

class Selection
{
public:
    bool isValid() const;
    int posCount() const;
    ArrayRef<const int> atomIndices() const;
    ArrayRef<const RVec> coordinates() const;
};
class SelectionCollection {};

class AnalysisDataPlotModule
{
public:
    explicit AnalysisDataPlotModule(const AnalysisDataPlotSettings& s);
    void setFileName(const std::string& fn);
    void setTitle(const char* t);
    void setXAxisIsTime();
    void setYLabel(const char* l);
};
typedef std::shared_ptr<AnalysisDataPlotModule> AnalysisDataPlotModulePointer;

class AnalysisData
{
public:
    void setColumnCount(int dataSet, int columns);
    void addModule(AnalysisDataPlotModulePointer m);
};
class AnalysisDataHandle
{
public:
    void startFrame(int index, real x, real dx = 0.0);
    void setPoint(int column, real value, bool present = true);
    void finishFrame();
};
class AnalysisDataParallelOptions {};

class TrajectoryAnalysisModule;
class TrajectoryAnalysisModuleData
{
public:
    virtual ~TrajectoryAnalysisModuleData();
    virtual void finish() = 0;
    AnalysisDataHandle dataHandle(const AnalysisData& data);
    Selection parallelSelection(const Selection& selection);
protected:
    TrajectoryAnalysisModuleData(TrajectoryAnalysisModule* module, const AnalysisDataParallelOptions& opt,
                                 const SelectionCollection& selections);
    void finishDataHandles();
};
typedef std::unique_ptr<TrajectoryAnalysisModuleData> TrajectoryAnalysisModuleDataPointer;

class TrajectoryAnalysisModule
{
public:
    virtual ~TrajectoryAnalysisModule();
    virtual void initOptions(IOptionsContainer* options, TrajectoryAnalysisSettings* settings) = 0;
    virtual void optionsFinished(TrajectoryAnalysisSettings* settings);
    virtual void initAnalysis(const TrajectoryAnalysisSettings& settings, const TopologyInformation& top) = 0;
    virtual TrajectoryAnalysisModuleDataPointer startFrames(const AnalysisDataParallelOptions& opt,
                                                            const SelectionCollection& selections);
    virtual void analyzeFrame(int frnr, const t_trxframe& fr, t_pbc* pbc, TrajectoryAnalysisModuleData* pdata) = 0;
    virtual void finishAnalysis(int nframes);
    virtual void writeOutput() = 0;
protected:
    TrajectoryAnalysisModule();
    void registerAnalysisDataset(AnalysisData* data, const char* name);
};

class TrajectoryAnalysisCommandLineRunner
{
public:
    template <class Module> static int runAsMain(int argc, char* argv[]);
};
}  // namespace gmx

#endif
