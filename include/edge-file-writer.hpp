// edge-file-writer.hpp -- the text outputs of the path, byte-identical to the reference.
//
//   edges-frames.xvg.dat : header written in initAnalysis (reference src/waternetwork.cpp:485-499),
//                          per frame "#--- frame N ---" and one row per edge (:638-652):
//                          setw(6) id1, setw(6) id2, then energy, length, cosine as
//                          std::fixed << std::setprecision(4) << std::setw(8).
//   nodes-list.xvg.dat   : header (:463-470) and one row per site, operator<<(SiteInfo) (:83-93).
//
// The ids printed are SiteInfo::id (solvent ids are offset by proteinSites_.size()+1, :420-423), so
// the writer takes a site-index -> id table.  Formatting ~16 MB of text per 100k-atom frame through
// iostreams is what limits the reference once the search is on the GPU; formatRows() is a
// fixed-point formatter that produces the same bytes (checked against std::ostream with the
// reference's manipulators in tests/cpp/test_writer.cpp) at several hundred MB/s per core.
#ifndef WN_EDGEFILEWRITER_HPP
#define WN_EDGEFILEWRITER_HPP

#include <cstddef>
#include <cstdint>
#include <ostream>
#include <string>
#include <vector>

#include "wn_b200.h"
#include "wn_types.hpp"

class EdgeFileWriter
{
public:
    // ids[s] = SiteInfo::id of site s (what the reference prints); empty => id = s.
    explicit EdgeFileWriter(std::ostream& os, std::vector<unsigned> ids = std::vector<unsigned>());

    void writeHeader();                                                  // :487-497
    void writeFrame(int frnr, const wn_edge* edges, size_t n_edges);     // :640-651

    // Appends the rows of `edges` to `out` (no frame line).  Public for tests/benchmarks.
    void formatRows(const wn_edge* edges, size_t n_edges, std::string& out) const;

    static void writeNodesHeader(std::ostream& os);                      // :463-470
    static void writeNode(std::ostream& os, const SiteInfo& site);       // :83-93

private:
    std::ostream& os_;
    std::vector<unsigned> ids_;
    std::string buf_;
};

#endif
