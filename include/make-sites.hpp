// make-sites.hpp -- site typing of the waternetwork module without GROMACS types.
//
// Restates makeSites (reference src/waternetwork.cpp:95-180) on plain arrays, so a host that
// has the topology (GROMACS' t_atoms: atom names, element strings, residue indices and
// names) can build the site tables that HydrogenBondEdges::setSites consumes.  Rules kept
// exactly, quirks included:
//   * an atom is a site iff its name is one of 16 names (:99-118; "NE2" listed twice);
//   * nbHydrogen = number of directly following atoms whose element string contains "H",
//     probing stops when  atomIndex + nbHydrogen + 1 > selection size  (:155-161 -- the bound
//     compares a TOPOLOGY index with the SELECTION size, so it only behaves for a selection
//     that starts at atom 0; reproduced as is);
//   * type = map[name] (:120-139), names missing from the map ("OG1") get the
//     default-constructed value DONOR; residue name "WAT" forces WATER (:170-173);
//   * id is 1-based in order of appearance (:164); the caller offsets solvent ids by
//     proteinSites.size()+1 (:420-423) -- offsetSolventIds() does that.
// hydrogenAtoms() gives the hydrogen list of a site in the terms of setSites: for waters the
// reference takes solventPoints[3*w + i], i < nbHydrogen (:564-570: the oxygen itself, then
// H1); for protein sites it takes proteinPoints[id + i + 1] (:554 -- indexed by the 1-based
// site id, a bug; bugCompat = false uses atmIndex + i + 1 instead).
#ifndef WN_MAKESITES_HPP
#define WN_MAKESITES_HPP

#include <string>
#include <vector>

#include "wn_types.hpp"

struct TopologyView {
    std::vector<std::string> atomName;   // per topology atom
    std::vector<std::string> element;    // per topology atom (t_atom::elem)
    std::vector<int> resIndex;           // per topology atom (t_atom::resind)
    std::vector<std::string> resName;    // per residue
};

// selection: topology atom indices of the selection, in order.  Appends to `sites`, returns sites.size().
int makeSites(const std::vector<int>& selection, const TopologyView& top, std::vector<SiteInfo>& sites);

// site.id += nProteinSites + 1 for every solvent site (src/waternetwork.cpp:420-423)
void offsetSolventIds(std::vector<SiteInfo>& solventSites, size_t nProteinSites);

// Atom indices (into the frame the site's coordinates come from) of the hydrogen list.
// water: index w of the water in the solvent selection -> {3w, 3w+1, ...} (nbHydrogen entries).
std::vector<int> waterHydrogenAtoms(int waterIndex, const SiteInfo& site);
std::vector<int> proteinHydrogenAtoms(const SiteInfo& site, bool bugCompat);

#endif
