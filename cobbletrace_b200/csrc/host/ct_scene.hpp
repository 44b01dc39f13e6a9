// ct_scene.hpp -- host-side scene model (what crosses into ct_gpu_upload_scene).
// Mirrors the *meaning* of the reference's scene_t (scenefile.h:95-102) for the ray-traced path:
// triangles in GetSceneTriangles order (raythread.cpp:621) with their materials, lights in file
// order, camera, settings.  Spheres are parsed and counted but dropped (raythread.cpp:208).
#pragma once

#include <cstdint>
#include <memory>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "ct_gpu.h"
#include "ct_host.h"

namespace cth {

struct Vec3 { double x = 0, y = 0, z = 0; };

struct Triangle { Vec3 p1, p2, p3; };   // 72 bytes, == 9 packed doubles
static_assert(sizeof(Triangle) == 72, "Triangle must be 9 packed doubles");

// std::allocator whose value-less construct() default-initialises: resize() of a vector of PODs allocates without
// filling, so that big arrays are first touched by the threads that write them.
template <class T>
struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = NoInitAlloc<U>; };
    template <class U, class... A>
    void construct(U *p, A &&...a) {
        if constexpr (sizeof...(A) == 0) ::new ((void *)p) U;
        else ::new ((void *)p) U(std::forward<A>(a)...);
    }
};

struct Scene {
    std::vector<Triangle> tris;
    std::vector<ct_material> mats;       // per triangle
    std::vector<ct_light> lights;
    Vec3 cam_pos{0, 0, -8};              // InitSceneData scenefile.cpp:28
    double cam_rot[9] = {1, 0, 0, 0, 1, 0, -0.0, 0, 1};   // HandleUpdates with yaw=pitch=roll=0 (raythread.cpp:564-572)
    ct_host_settings settings{8, 1, 0, 0};                // scenefile.cpp:23-26
    uint32_t n_spheres = 0;
    // BVH (reference layout)
    std::vector<ct_bvh_node, NoInitAlloc<ct_bvh_node>> nodes;
    std::vector<uint32_t> tri_index;
};

// scene file + imports; throws std::runtime_error
void parse_scene_file(const std::string &path, const std::string &base_dir, Scene &out);
void build_bvh(Scene &s);
void camera_rotation(float yaw, float pitch, float roll, double out[9]);

}  // namespace cth
