// test_sites.cpp -- makeSites / hydrogen-list rules against a hand-made topology (CPU only).
#include <cstdio>

#include "make-sites.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); return 1; } } while (0)

int main()
{
    TopologyView top;
    auto add = [&](const char* name, const char* elem, int res) {
        top.atomName.push_back(name); top.element.push_back(elem); top.resIndex.push_back(res);
    };
    top.resName = {"LYS", "THR", "GLN", "WAT", "SOL"};
    // LYS: N H CA HA C O NZ HZ1 HZ2 HZ3
    add("N", "N", 0); add("H", "H", 0); add("CA", "C", 0); add("HA", "H", 0); add("C", "C", 0); add("O", "O", 0);
    add("NZ", "N", 0); add("HZ1", "H", 0); add("HZ2", "H", 0); add("HZ3", "H", 0);
    // THR: OG1 HG1 ; GLN: OE1 NE2 HE21 HE22
    add("OG1", "O", 1); add("HG1", "H", 1);
    add("OE1", "O", 2); add("NE2", "N", 2); add("HE21", "H", 2); add("HE22", "H", 2);
    // waters: residue WAT (forced WATER) and SOL (WATER through the OW name)
    add("OW", "O", 3); add("HW1", "H", 3); add("HW2", "H", 3);
    add("OW", "O", 4); add("HW1", "H", 4); add("HW2", "H", 4);
    std::vector<int> all;
    for (size_t i = 0; i < top.atomName.size(); ++i) all.push_back((int)i);

    std::vector<SiteInfo> sites;
    CHECK(makeSites(all, top, sites) == 8);
    const char* names[] = {"N", "O", "NZ", "OG1", "OE1", "NE2", "OW", "OW"};
    const int nh[] = {1, 0, 3, 1, 0, 2, 2, 2};
    const SiteType ty[] = {DONOR, ACCEPTOR, DONOR, DONOR /* OG1: not in the map -> default */, ACCEPTOR, DONOR, WATER, WATER};
    const unsigned at[] = {0, 5, 6, 10, 12, 13, 16, 19};
    for (int s = 0; s < 8; ++s) {
        CHECK(sites[s].id == (unsigned)(s + 1));
        CHECK(sites[s].name == names[s]);
        CHECK(sites[s].nbHydrogen == nh[s]);
        CHECK(sites[s].type == ty[s]);
        CHECK(sites[s].atmIndex == at[s]);
    }
    CHECK(sites[6].resName == "WAT" && sites[7].resName == "SOL");
    // a non-site atom named like nothing in the list; an "O" in residue WAT is forced to WATER
    top.atomName[5] = "O"; top.resName[0] = "WAT";
    std::vector<SiteInfo> again;
    makeSites(all, top, again);
    CHECK(again[1].type == WATER && again[0].type == WATER);
    top.resName[0] = "LYS";

    // the probe bound compares the topology index with the selection size (:157): a selection of the
    // two waters only (6 atoms, topology indices 16..21) cuts the probe at once -> 0 hydrogens
    std::vector<int> waters = {16, 17, 18, 19, 20, 21};
    std::vector<SiteInfo> ws;
    CHECK(makeSites(waters, top, ws) == 2);
    CHECK(ws[0].nbHydrogen == 0 && ws[1].nbHydrogen == 0);
    // ... whereas a solvent-only topology (selection starts at atom 0) gives 2 each
    TopologyView wt;
    wt.resName = {"SOL", "SOL"};
    for (int w = 0; w < 2; ++w) { wt.atomName.insert(wt.atomName.end(), {"OW", "HW1", "HW2"}); wt.element.insert(wt.element.end(), {"O", "H", "H"});
                                  wt.resIndex.insert(wt.resIndex.end(), {w, w, w}); }
    std::vector<SiteInfo> w2;
    CHECK(makeSites({0, 1, 2, 3, 4, 5}, wt, w2) == 2);
    CHECK(w2[0].nbHydrogen == 2 && w2[1].nbHydrogen == 2 && w2[1].atmIndex == 3);
    offsetSolventIds(w2, 8);
    CHECK(w2[0].id == 10 && w2[1].id == 11);                       // 1 + 8 + 1, 2 + 8 + 1
    // hydrogen lists: water w -> {3w, 3w+1} (the oxygen itself, then H1; H2 never used)
    CHECK((waterHydrogenAtoms(1, w2[1]) == std::vector<int>{3, 4}));
    // protein: reference indexes by the 1-based site id (bug), fixed variant by atom index
    CHECK((proteinHydrogenAtoms(sites[2], true) == std::vector<int>{4, 5, 6}));     // id 3 -> 4,5,6
    CHECK((proteinHydrogenAtoms(sites[2], false) == std::vector<int>{7, 8, 9}));    // NZ at 6 -> HZ1..3
    std::printf("test_sites OK\n");
    return 0;
}
