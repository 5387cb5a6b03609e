/*
 * ref_wrap.cpp -- C-ABI wrapper around the REFERENCE's own hot-path source text.
 * Test infrastructure only (oracle/_ref/libwn_ref.so); never shipped.
 *
 * Built by oracle/Makefile target `ref`, only where /root/reference exists:
 *   - src/brute-find-pairs.cpp is compiled as its own translation unit, in place;
 *   - the bodies of switch_function_radius, switch_function_angle, computeEnergy
 *     (src/waternetwork.cpp:12-66) and make_edges (:182-263) are sliced out of
 *     waternetwork.cpp into a temporary file at build time (the rest of that file
 *     needs GROMACS, which is not installed) and pulled in below through
 *     WN_REF_SLICE.  No reference text is stored in this repository.
 * CGAL arithmetic comes from oracle/shim (restated semantics, see there).
 */
#include "cgal.hpp"              /* reference include/cgal.hpp over the shim   */
#include "utils.hpp"             /* reference include/utils.hpp                 */
#include "brute-find-pairs.hpp"  /* reference include/brute-find-pairs.hpp      */

#include <cmath>
#include <cstdint>
#include <chrono>
#include <cstring>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include WN_REF_SLICE            /* computeEnergy + make_edges, reference text  */

extern "C" {

struct wnref_edge { int32_t i, j; float energy, length, cosine; };

double wnref_switch_function_radius(double r, double on, double off)
{
    return switch_function_radius(r, on, off);
}
double wnref_switch_function_angle(double a, double on, double off)
{
    return switch_function_angle(a, on, off);
}
double wnref_compute_energy(double r, double cosine) { return computeEnergy(r, cosine); }

size_t wnref_brute_find_pairs(const double* xyz, int n, int32_t* pairs, size_t cap)
{
    std::vector<Point> pts;
    pts.reserve(n);
    for (int i = 0; i < n; ++i) pts.push_back(Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
    BruteFindPairs finder;
    std::vector<std::pair<int, int>> res = finder.find(pts);
    for (size_t k = 0; k < res.size() && k < cap; ++k) {
        pairs[2 * k] = res[k].first;
        pairs[2 * k + 1] = res[k].second;
    }
    return res.size();
}

size_t wnref_make_edges(const int32_t* pairs, size_t m, const double* site_xyz, int n,
                        const double* h_xyz, const int32_t* h_off, const uint8_t* site_type,
                        wnref_edge* out)
{
    std::vector<std::pair<int, int>> pv(m);
    for (size_t k = 0; k < m; ++k) pv[k] = std::pair<int, int>(pairs[2 * k], pairs[2 * k + 1]);
    std::vector<Point> sitePoints;
    std::vector<std::vector<Point_ptr>> hydrogens(n);
    std::vector<SiteInfo_ptr> sites(n);
    for (int s = 0; s < n; ++s) {
        sitePoints.push_back(Point(site_xyz[3 * s], site_xyz[3 * s + 1], site_xyz[3 * s + 2]));
        for (int h = h_off[s]; h < h_off[s + 1]; ++h)
            hydrogens[s].push_back(std::make_shared<Point>(h_xyz[3 * h], h_xyz[3 * h + 1], h_xyz[3 * h + 2]));
        SiteInfo info;
        info.type = (SiteType)site_type[s];
        info.id = (unsigned)s;           /* id carries the 0-based site index back out */
        info.atmIndex = (unsigned)s;
        info.resIndex = 0;
        info.nbHydrogen = (unsigned short)(h_off[s + 1] - h_off[s]);
        sites[s] = std::make_shared<SiteInfo>(info);
    }
    std::vector<std::pair<int, int>>::iterator b = pv.begin(), e = pv.end();
    std::vector<HydrogenBond> hb = make_edges(b, e, sitePoints, hydrogens, sites);
    for (size_t k = 0; k < hb.size(); ++k) {
        out[k].i = (int32_t)hb[k].sites.first->id;
        out[k].j = (int32_t)hb[k].sites.second->id;
        out[k].energy = (float)hb[k].energy;   /* exact: value is a widened float */
        out[k].length = (float)hb[k].length;
        out[k].cosine = (float)hb[k].angle;
    }
    return hb.size();
}

/* One frame end to end the way WaterNetwork::analyzeFrame does it for waters
 * (src/waternetwork.cpp:516-523 widening, :562-575 gather with hydrogens
 * solventPoints[3*index+i], i < nbHydrogen = 2, :583 find, :590-633 make_edges),
 * with BruteFindPairs as the finder.  One make_edges call over the whole pair
 * list (the reference splits it over hardware_concurrency() std::async chunks and
 * joins in order, which gives the same list).  Used as the CPU baseline. */
size_t wnref_frame_hbonds_water(const float* frame, int n_waters, wnref_edge* out, size_t cap,
                                uint64_t* n_pairs)
{
    std::vector<Point> solventPoints;
    solventPoints.reserve(3 * (size_t)n_waters);
    for (int i = 0; i < 3 * n_waters; ++i)
        solventPoints.push_back(Point(frame[3 * i], frame[3 * i + 1], frame[3 * i + 2]));
    std::vector<Point> sitePoints;
    std::vector<std::vector<Point_ptr>> hydrogenPoints;
    std::vector<SiteInfo_ptr> siteInfos;
    for (int index = 0; index < n_waters; ++index) {
        std::vector<Point_ptr> hydrogens(2);
        for (unsigned int i = 0; i < 2; i++)
            hydrogens.at(i) = std::make_shared<Point>(solventPoints.at(3 * index + i));
        hydrogenPoints.push_back(hydrogens);
        SiteInfo info;
        info.type = WATER;
        info.id = (unsigned)index;
        info.atmIndex = (unsigned)(3 * index);
        info.resIndex = (unsigned)index;
        info.nbHydrogen = 2;
        siteInfos.push_back(std::make_shared<SiteInfo>(info));
        sitePoints.push_back(solventPoints.at(3 * index));
    }
    BruteFindPairs finder;
    std::vector<std::pair<int, int>> pairs = finder.find(sitePoints);
    if (n_pairs) *n_pairs = pairs.size();
    std::vector<std::pair<int, int>>::iterator b = pairs.begin(), e = pairs.end();
    std::vector<HydrogenBond> hb = make_edges(b, e, sitePoints, hydrogenPoints, siteInfos);
    for (size_t k = 0; k < hb.size() && k < cap; ++k) {
        out[k].i = (int32_t)hb[k].sites.first->id;
        out[k].j = (int32_t)hb[k].sites.second->id;
        out[k].energy = (float)hb[k].energy;
        out[k].length = (float)hb[k].length;
        out[k].cosine = (float)hb[k].angle;
    }
    return hb.size();
}

/* The same frame the way analyzeFrame itself schedules it (src/waternetwork.cpp:583-633): the pair search
 * on the calling thread, then the pair list cut into `n_chunks` contiguous ranges (the reference uses
 * std::thread::hardware_concurrency()), one std::async(std::launch::async, make_edges, ...) per range on the
 * shared read-only site arrays (:608-618), futures joined in order and concatenated (:621-633).  The range
 * arithmetic is restated (even split, remainder spread over the first ranges); the reference's own has
 * undefined behaviour for lists shorter than the thread count and the same result
 * otherwise.  seconds[0] = find, seconds[1] = make_edges fan-out + join.  This is the reference's
 * INTRA-frame parallel variant: one frame at a time on all cores. */
size_t wnref_frame_hbonds_water_chunked(const float* frame, int n_waters, int n_chunks, wnref_edge* out, size_t cap,
                                        uint64_t* n_pairs, double* seconds)
{
    std::vector<Point> solventPoints;
    solventPoints.reserve(3 * (size_t)n_waters);
    for (int i = 0; i < 3 * n_waters; ++i)
        solventPoints.push_back(Point(frame[3 * i], frame[3 * i + 1], frame[3 * i + 2]));
    std::vector<Point> sitePoints;
    std::vector<std::vector<Point_ptr>> hydrogenPoints;
    std::vector<SiteInfo_ptr> siteInfos;
    for (int index = 0; index < n_waters; ++index) {
        std::vector<Point_ptr> hydrogens(2);
        for (unsigned int i = 0; i < 2; i++)
            hydrogens.at(i) = std::make_shared<Point>(solventPoints.at(3 * index + i));
        hydrogenPoints.push_back(hydrogens);
        SiteInfo info;
        info.type = WATER;
        info.id = (unsigned)index;
        info.atmIndex = (unsigned)(3 * index);
        info.resIndex = (unsigned)index;
        info.nbHydrogen = 2;
        siteInfos.push_back(std::make_shared<SiteInfo>(info));
        sitePoints.push_back(solventPoints.at(3 * index));
    }
    const auto t0 = std::chrono::steady_clock::now();
    BruteFindPairs finder;
    std::vector<std::pair<int, int>> pairs = finder.find(sitePoints);
    const auto t1 = std::chrono::steady_clock::now();
    if (n_pairs) *n_pairs = pairs.size();
    if (n_chunks < 1) n_chunks = 1;
    typedef std::vector<std::pair<int, int>>::iterator It;
    std::vector<std::future<std::vector<HydrogenBond>>> futures;
    const size_t m = pairs.size(), base = m / (size_t)n_chunks, extra = m % (size_t)n_chunks;
    size_t start = 0;
    for (int k = 0; k < n_chunks; ++k) {
        const size_t len = base + ((size_t)k < extra ? 1 : 0);
        It b = pairs.begin() + (std::ptrdiff_t)start, e = pairs.begin() + (std::ptrdiff_t)(start + len);
        futures.push_back(std::async(std::launch::async, [&, b, e]() mutable {
            return make_edges(b, e, sitePoints, hydrogenPoints, siteInfos);
        }));
        start += len;
    }
    size_t count = 0;
    for (auto& fut : futures) {
        std::vector<HydrogenBond> hb = fut.get();
        for (size_t k = 0; k < hb.size(); ++k, ++count) {
            if (count >= cap) continue;
            out[count].i = (int32_t)hb[k].sites.first->id;
            out[count].j = (int32_t)hb[k].sites.second->id;
            out[count].energy = (float)hb[k].energy;
            out[count].length = (float)hb[k].length;
            out[count].cosine = (float)hb[k].angle;
        }
    }
    const auto t2 = std::chrono::steady_clock::now();
    if (seconds) {
        seconds[0] = std::chrono::duration<double>(t1 - t0).count();
        seconds[1] = std::chrono::duration<double>(t2 - t1).count();
    }
    return count;
}

} /* extern "C" */
