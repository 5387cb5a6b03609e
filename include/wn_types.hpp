// wn_types.hpp -- the plain data carriers of the path, with the reference's names
// (reference include/utils.hpp:7-32): SiteType, SiteInfo, Direction, HydrogenBond.
// SiteInfo_ptr / SiteInfoPair keep the reference's aliases so make_edges-style
// consumers (the .dat writer, src/waternetwork.cpp:638-652) compile unchanged.
#ifndef WN_TYPES_HPP
#define WN_TYPES_HPP

#include <memory>
#include <string>
#include <utility>

enum SiteType { DONOR, ACCEPTOR, WATER };   // values 0,1,2 == WN_DONOR, WN_ACCEPTOR, WN_WATER

struct SiteInfo {
    SiteType type;
    std::string name;
    unsigned id;                  // what the edge file prints (solvent ids are offset, :420-423)
    unsigned int atmIndex;
    unsigned int resIndex;
    std::string resName;
    short unsigned int nbHydrogen;
};

using SiteInfo_ptr = std::shared_ptr<SiteInfo>;
using SiteInfoPair = std::pair<SiteInfo_ptr, SiteInfo_ptr>;

enum Direction { FORWARD, BACKWARD };

struct HydrogenBond {
    SiteInfoPair sites;
    Direction direction;          // always FORWARD in the reference (:255)
    double length;                // float-rounded value widened (:195-197)
    double angle;                 // the selected cosine
    double energy;
};

#endif
