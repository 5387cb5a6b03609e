/*
 * wn_cgal_shim.h -- CGAL-free stand-in used ONLY to compile the reference's own
 * hot-path sources in place (oracle/Makefile target `ref`).  Test infrastructure.
 *
 * CGAL is a third-party dependency of the reference that is not vendored in
 * /root/reference and not installed here ("CGAL 4.X", unpinned, README.md:12).
 * This shim restates, from CGAL's published Cartesian-kernel semantics, exactly
 * the operations the hot path calls (reference call sites):
 *   CGAL::squared_distance(Point,Point)      src/brute-find-pairs.cpp:25
 *   Point - Point -> Vector_3                src/waternetwork.cpp:194,219,231
 *   Vector_3 * Vector_3 -> double            src/waternetwork.cpp:195,220,221,232,233
 *   Vector_3 /= scalar                       src/waternetwork.cpp:196,220,232
 *   scalar * Vector_3                        src/waternetwork.cpp:233
 *   CGAL::sqrt(double)                       src/waternetwork.cpp:195,220,232
 * Everything else in include/cgal.hpp (Delaunay, alpha shapes, spatial sort) is
 * declared as an empty template so the typedefs parse; none of it is used.
 */
#ifndef WN_CGAL_SHIM_H
#define WN_CGAL_SHIM_H

#include <cmath>
#include <memory>
#include <utility>
#include <vector>

namespace CGAL {

struct Sequential_tag {};
struct Parallel_tag {};
struct Fast_location {};

class ShimVector_3 {
public:
    ShimVector_3() : x_(0), y_(0), z_(0) {}
    ShimVector_3(double x, double y, double z) : x_(x), y_(y), z_(z) {}
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
    /* Construct_divided_vector_3 (Cartesian): component-wise true division */
    ShimVector_3& operator/=(double c) { x_ = x_ / c; y_ = y_ / c; z_ = z_ / c; return *this; }
private:
    double x_, y_, z_;
};

class ShimPoint_3 {
public:
    ShimPoint_3() : x_(0), y_(0), z_(0) {}
    ShimPoint_3(double x, double y, double z) : x_(x), y_(y), z_(z) {}
    double x() const { return x_; }
    double y() const { return y_; }
    double z() const { return z_; }
private:
    double x_, y_, z_;
};

/* Construct_vector_3(p, q) as used by operator-(Point,Point): p - q */
inline ShimVector_3 operator-(const ShimPoint_3& p, const ShimPoint_3& q)
{
    return ShimVector_3(p.x() - q.x(), p.y() - q.y(), p.z() - q.z());
}
/* Compute_scalar_product_3 (Cartesian): vx*wx + vy*wy + vz*wz */
inline double operator*(const ShimVector_3& v, const ShimVector_3& w)
{
    return v.x() * w.x() + v.y() * w.y() + v.z() * w.z();
}
/* Construct_scaled_vector_3 (Cartesian): (c*vx, c*vy, c*vz) */
inline ShimVector_3 operator*(double c, const ShimVector_3& v)
{
    return ShimVector_3(c * v.x(), c * v.y(), c * v.z());
}
/* squared_distanceC3: square(px-qx) + square(py-qy) + square(pz-qz) */
inline double squared_distance(const ShimPoint_3& p, const ShimPoint_3& q)
{
    const double dx = p.x() - q.x(), dy = p.y() - q.y(), dz = p.z() - q.z();
    return dx * dx + dy * dy + dz * dz;
}
inline double sqrt(double x) { return std::sqrt(x); }

struct Exact_predicates_inexact_constructions_kernel {
    typedef ShimPoint_3 Point_3;
    typedef ShimVector_3 Vector_3;
    typedef double FT;
};

/* Unused by the hot path: only so include/cgal.hpp's typedefs parse. */
template <class...> struct Triangulation_hierarchy_vertex_base_3 {};
template <class...> struct Alpha_shape_vertex_base_3 {};
template <class...> struct Alpha_shape_cell_base_3 {};
template <class...> struct Triangulation_data_structure_3 {};
template <class...> struct Delaunay_triangulation_3 {};
template <class...> struct Alpha_shape_3 {};
template <class...> struct Triangulation_vertex_base_with_info_3 {};
template <class...> struct Delaunay_triangulation_cell_base_3 {};
template <class...> struct Spatial_sort_traits_adapter_3 {};
template <class T> struct Pointer_property_map { typedef T* type; };

} // namespace CGAL
#endif
