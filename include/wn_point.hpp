// wn_point.hpp -- the `Point` type of the drop-in boundary.
//
// The reference takes `Point = CGAL::Exact_predicates_inexact_constructions_kernel::Point_3`
// (include/cgal.hpp:21-22): three doubles, 24 bytes, accessors x() y() z().  CGAL is not
// a dependency of this library; by default a layout-compatible plain type is used.  Define
// WN_USE_CGAL before including to get the real CGAL type (for a build inside the
// reference tree); either way std::vector<Point> storage is AoS double[n][3], which is
// what wn_find_pairs() consumes.
#ifndef WN_POINT_HPP
#define WN_POINT_HPP

#if defined(WN_USE_CGAL)
#include <CGAL/Exact_predicates_inexact_constructions_kernel.h>
typedef CGAL::Exact_predicates_inexact_constructions_kernel K;
typedef K::Point_3 Point;
#else
class Point
{
public:
    Point() : c_{0.0, 0.0, 0.0} {}
    Point(double x, double y, double z) : c_{x, y, z} {}
    double x() const { return c_[0]; }
    double y() const { return c_[1]; }
    double z() const { return c_[2]; }
    const double* data() const { return c_; }
private:
    double c_[3];
};
#endif

static_assert(sizeof(Point) == 3 * sizeof(double), "Point must be three packed doubles");

#endif
