// make-sites.cpp -- see include/make-sites.hpp.
#include "make-sites.hpp"

#include <map>
#include <set>

int makeSites(const std::vector<int>& selection, const TopologyView& top, std::vector<SiteInfo>& sites)
{
    static const std::set<std::string> atomNames{"OW", "O", "N", "SG", "NE2", "OE1", "OD1", "ND2", "OG",
                                                 "OG1", "OH", "NE1", "NZ", "NH1", "NH2", "ND1"};
    static const std::map<std::string, SiteType> siteType{
        {"NE", DONOR},  {"NH1", DONOR}, {"NH2", DONOR},    {"ND1", DONOR},    {"ND2", DONOR}, {"OD1", DONOR},
        {"OD2", DONOR}, {"NE1", DONOR}, {"NE2", DONOR},    {"OE1", ACCEPTOR}, {"OE2", ACCEPTOR}, {"NZ", DONOR},
        {"OG", DONOR},  {"OH", DONOR},  {"O", ACCEPTOR},   {"N", DONOR},      {"SG", DONOR},  {"OW", WATER},
        {"OCD", ACCEPTOR}};
    const std::string watres("WAT");
    unsigned int count = 0;
    for (size_t i = 0; i < selection.size(); ++i) {
        const int index = selection[i];
        const std::string& name = top.atomName.at((size_t)index);
        if (atomNames.find(name) == atomNames.end()) continue;
        ++count;
        unsigned int nhb = 0;
        while (true) {
            if ((size_t)index + nhb + 1 > selection.size()) break;                 // :157, as written
            if ((size_t)index + nhb + 1 >= top.element.size()) break;              // stay inside the topology
            if (top.element[(size_t)index + nhb + 1].find("H") == std::string::npos) break;
            ++nhb;
        }
        SiteInfo site;
        site.id = count;
        site.name = name;
        site.atmIndex = (unsigned)index;
        site.resIndex = (unsigned)top.resIndex.at((size_t)index);
        site.nbHydrogen = (unsigned short)nhb;
        const std::map<std::string, SiteType>::const_iterator it = siteType.find(name);
        site.type = it == siteType.end() ? SiteType() : it->second;                // operator[] default: DONOR
        site.resName = top.resName.at(site.resIndex);
        if (watres == site.resName) site.type = WATER;
        sites.push_back(site);
    }
    return (int)sites.size();
}

void offsetSolventIds(std::vector<SiteInfo>& solventSites, size_t nProteinSites)
{
    for (SiteInfo& s : solventSites) s.id += (unsigned)(nProteinSites + 1);
}

std::vector<int> waterHydrogenAtoms(int waterIndex, const SiteInfo& site)
{
    std::vector<int> h(site.nbHydrogen);
    for (unsigned i = 0; i < site.nbHydrogen; ++i) h[i] = 3 * waterIndex + (int)i;
    return h;
}

std::vector<int> proteinHydrogenAtoms(const SiteInfo& site, bool bugCompat)
{
    std::vector<int> h(site.nbHydrogen);
    const unsigned base = bugCompat ? site.id : site.atmIndex;
    for (unsigned i = 0; i < site.nbHydrogen; ++i) h[i] = (int)(base + i + 1);
    return h;
}
