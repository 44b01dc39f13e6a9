// ct_bvh.cpp -- host BVH build whose output is identical, node for node and index for index, to the
// reference's InitializeBVHState + BuildBVH (bvh.cpp:16-120), because the GPU traversal reproduces the
// reference's DFS order over exactly that tree (ties and the t=0 reflection rays depend on it).
//
// Same algorithm (midpoint split on the longest axis, leaf when <= 2 triangles or when a side would be
// empty, children allocated adjacently), different shape: iterative pre-order worklist instead of
// recursion (no stack overflow on degenerate 1M-triangle inputs), centroids in a flat SoA array.
//
// Rounding points that decide the partition (all reproduced):
//   centroid = (p1+p2+p3) * 0.3333f   -> fp64 product with the float constant widened   (bvh.cpp:112)
//   axis pick compares extent.z (double) with float(extent[axis])                       (bvh.cpp:61-65)
//   splitPos = float(min[axis]) + float(extent[axis]) * 0.5f   in fp32                  (bvh.cpp:67)
//   partition predicate float(centroid[axis]) < splitPos                                (bvh.cpp:73)
#include <cstring>
#include <stdexcept>
#include <vector>

#include "ct_scene.hpp"

namespace cth {
namespace {

inline double lo(double a, double b) { return (a < b) ? a : b; }   // mymath.h:15 macro semantics
inline double hi(double a, double b) { return (a > b) ? a : b; }   // mymath.h:11

void fit_bounds(const Scene &s, ct_bvh_node &n) {                 // UpdateNodeBounds bvh.cpp:30-49
    for (int a = 0; a < 3; a++) { n.aabb_min[a] = (double)1e30f; n.aabb_max[a] = (double)-1e30f; }
    for (uint32_t i = 0; i < n.triangle_count; i++) {
        const double *v = &s.tris[s.tri_index[n.first_triangle_index + i]].p1.x;   // 9 packed doubles
        for (int p = 0; p < 3; p++)
            for (int a = 0; a < 3; a++) {
                n.aabb_min[a] = lo(n.aabb_min[a], v[3 * p + a]);
                n.aabb_max[a] = hi(n.aabb_max[a], v[3 * p + a]);
            }
    }
}

}  // namespace

void build_bvh(Scene &s) {
    const uint32_t n = (uint32_t)s.tris.size();
    if (n == 0) throw std::runtime_error("cannot build a BVH over zero triangles");
    s.nodes.assign((size_t)2 * n - 1, ct_bvh_node{});              // calloc'd, bvh.cpp:18
    s.tri_index.resize(n);
    std::vector<double> centroid((size_t)3 * n);
    const double third = (double)0.3333f;
    for (uint32_t k = 0; k < n; k++) {
        s.tri_index[k] = k;
        const Triangle &t = s.tris[k];
        centroid[3 * (size_t)k + 0] = third * ((t.p1.x + t.p2.x) + t.p3.x);
        centroid[3 * (size_t)k + 1] = third * ((t.p1.y + t.p2.y) + t.p3.y);
        centroid[3 * (size_t)k + 2] = third * ((t.p1.z + t.p2.z) + t.p3.z);
    }
    uint32_t used = 1;
    s.nodes[0].left_node = 0; s.nodes[0].first_triangle_index = 0; s.nodes[0].triangle_count = n;
    fit_bounds(s, s.nodes[0]);

    std::vector<uint32_t> work;                                    // pre-order: pop, split, push right then left
    work.push_back(0);
    while (!work.empty()) {
        const uint32_t idx = work.back();
        work.pop_back();
        ct_bvh_node &node = s.nodes[idx];
        if (node.triangle_count <= 2) continue;                    // bvh.cpp:54
        double ext[3];
        for (int a = 0; a < 3; a++) ext[a] = node.aabb_max[a] - node.aabb_min[a];
        int axis = 0;
        if (ext[1] > ext[0]) axis = 1;
        if (ext[2] > (double)(float)ext[axis]) axis = 2;
        const float split = (float)node.aabb_min[axis] + (float)ext[axis] * 0.5f;
        int i = (int)node.first_triangle_index;
        int j = i + (int)node.triangle_count - 1;
        while (i <= j) {                                           // in-place partition, bvh.cpp:70-81
            if ((float)centroid[3 * (size_t)s.tri_index[i] + axis] < split) {
                i++;
            } else {
                uint32_t t = s.tri_index[i];
                s.tri_index[i] = s.tri_index[j];
                s.tri_index[j--] = t;
            }
        }
        const uint32_t left_count = (uint32_t)i - node.first_triangle_index;
        if (left_count == 0 || left_count == node.triangle_count) continue;   // bvh.cpp:84-86
        const uint32_t l = used++, r = used++;
        node.left_node = l;
        s.nodes[l].first_triangle_index = node.first_triangle_index;
        s.nodes[l].triangle_count = left_count;
        s.nodes[r].first_triangle_index = (uint32_t)i;
        s.nodes[r].triangle_count = node.triangle_count - left_count;
        node.triangle_count = 0;
        fit_bounds(s, s.nodes[l]);
        fit_bounds(s, s.nodes[r]);
        work.push_back(r);
        work.push_back(l);
    }
    s.nodes.resize(used);
}

}  // namespace cth
