// find-pairs.hpp -- the pair-finder strategy interface the reference defines
// (reference include/find-pairs.hpp:8-13): one pure virtual
//     find(std::vector<Point>& points) -> std::vector<std::pair<int,int>>
// called once per frame from WaterNetwork::analyzeFrame (src/waternetwork.cpp:583).
// Same name, same signature, same semantics; the only addition is a virtual
// destructor (the reference lacks one), so that owning a finder through
// std::shared_ptr<FindPairs> -- what include/waternetwork.hpp:94 has to become to
// accept any finder -- destroys GPU state correctly.
#ifndef WN_FINDPAIRS_HPP
#define WN_FINDPAIRS_HPP

#include <utility>
#include <vector>

#include "wn_point.hpp"

class FindPairs
{
public:
    virtual ~FindPairs() {}
    // Neither reference finder mutates `points`; implementations must not reorder it.
    virtual std::vector<std::pair<int, int>> find(std::vector<Point>& points) = 0;
};

#endif
