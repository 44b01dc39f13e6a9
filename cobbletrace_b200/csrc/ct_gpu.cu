// ct_gpu.cu -- sm_100a wavefront renderer behind the C ABI in include/ct_gpu.h.
//
// Replaces the worker half of the reference's boss/worker (raythread.cpp:437-543
// RayTracePartition -> TraceRay :353 -> ClosestIntersection :197 -> bvh.cpp:198
// IntersectBVHClosest / ComputeLighting :275).  Not a port: the recursive per-pixel CPU loop
// becomes a wavefront of persistent-warp kernels over SoA path state in HBM:
//
//   k_primary   raygen (CanvasToViewport :186) + closest-hit DFS            -> hit records
//   k_emit      the recursion step (:369-373): end the path or append the reflection ray to the next queue
//   k_bounce    the reference's degenerate t=0 reflection rays (:373): first barycentric
//               pass in DFS order                               -> hit records
//   k_shadow    ComputeLighting's shadow rays (:288-306), one work item per (light, path) in
//               light-major order (a warp = 32 neighbouring paths, same light); any-hit walk with deferred
//               leaves                                                       -> occlusion bit masks
//   k_shade     NormalOfSceneObject :329 + ComputeLighting :275 accumulation (lights in file
//               order, fp32) + HsvToColor                                    -> per-depth colour stack / framebuffer
//   k_overflow, k_overflow_huge
//               rays whose walk exceeds a visit budget (e.g. shadow rays cast from a shading point
//               4.3e9 units away after a reflection "miss": fp32 slab quotients all round to the
//               same value and EVERY box passes) are parked by k_shadow / k_bounce and finished by a warp
//               each, or -- the every-box-passes ones -- by the whole grid testing all triangles at once
//   k_resolve   unwinds the per-pixel blend chain (:375-379) and stores 0x00BBGGRR pixels
//
// Arithmetic is the reference's mixed fp64/fp32, reproduced exactly: certified fp32 filters decide what they can,
// the reference's own operations (explicit round-to-nearest intrinsics, ct_exact.cuh) decide the rest.
// No tensor cores: the path has no dense contraction.  No CPU fallback.
#include "../../include/ct_gpu.h"
#include "ct_exact.cuh"

#include <cuda_runtime.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>


namespace {

using namespace ct;

constexpr int kStackMax = 96;        // DFS stack entries per ray (tree depth limit, checked at upload)
constexpr int kMaxDevices = 16;
constexpr int kBlockThreads = 128;   // 4 warps per CTA
#ifndef CT_MIN_BLOCKS
#define CT_MIN_BLOCKS 6
#endif
constexpr int kMinBlocks = CT_MIN_BLOCKS;   // traversal kernels: resident CTAs per SM the register allocation must allow
constexpr int kOvfThreads = 256;     // k_overflow CTA
constexpr int kMaxLaunches = 80;     // launches of one tile (work cursors / stage events)
constexpr uint32_t kNoPos = 0xffffffffu;
// Slots a warp takes from the tile's cursor at a time: 32 = one 8x4 pixel block, one ray per lane.  When the cursor is
// another GPU's memory, 64 would halve the NVLink round trips, but a kernel ends with its slowest warp and a warp's
// chunk is walked one ray per lane at a time: measured on dragon 4K, 2 / 4 GPUs: 4.10 / 2.49 ms with 64 against
// 4.01 / 2.36 ms with 32 (option "shared_chunk_shift" for experiments).
constexpr uint32_t kChunkLocalShift = 5;
constexpr uint32_t kChunkSharedShift = 5;
constexpr uint32_t kChunkMaxShift = 6;
constexpr uint32_t kChunkMax = 1u << kChunkMaxShift;
constexpr uint32_t kDefaultBudget = 384;    // node visits + triangle tests before a ray is parked for k_overflow

// ---- device-side scene layout (SoA arrays in HBM, uploaded once) -----------------------------------
// The BVH is stored per INTERIOR node as the pair of its two children (bvh.cpp:89-97 allocates them adjacently
// and the traversal always needs both).  A child is described by (ref, cnt): cnt > 0 -> leaf holding triangles
// [ref, ref + cnt) of the leaf-ordered triangle array; cnt == 0 -> interior, ref = its own pair index.
struct __align__(16) DevPair32 {     // 64 B = two 32-B sectors: what the certified fp32 filter reads
    float lmin[3], lmax[3], rmin[3], rmax[3];    // float(bounds), round to nearest
    uint32_t l_ref, l_cnt, r_ref, r_cnt;
};
struct __align__(16) DevPair64 {     // 96 B: the reference's fp64 bounds, read only when the filter cannot decide
    double lmin[3], lmax[3], rmin[3], rmax[3];
};
struct __align__(16) DevTri {        // 80 B, stored in LEAF order (position = slot in bvh indexes[])
    double p1[3], e1[3], e2[3];      // e1 = p2-p1, e2 = p3-p1 (bvh.cpp:148-149, raythread.cpp:337-338)
    uint32_t orig, pad;              // original triangle id (= closestIndex of the reference)
};
struct __align__(16) DevTri32 {      // 48 B, same order: what the certified fp32 triangle filter reads
    float p1[3], k3;                 // k3 = max|p1_i| rounded up
    float e1[3], k1;                 // k1 = max|e1_i| rounded up (NaN: magnitudes outside the filter's range)
    float e2[3], k2;
};
struct DevLight { int32_t type; float intensity; double pos[3]; double dir[3]; };
struct DevShadowLight { int32_t type; uint32_t index; double v[3]; };   // non-ambient lights, file order; index = light number
struct __align__(16) OvfRay {        // a parked ray: 64 B
    double o[3], d[3];
    uint32_t target;                 // kAnyHit: word of the occlusion mask; kFirstLine: queue slot of the path
    uint32_t bit;                    // kAnyHit: bit inside that word
};

struct DevSched {                    // zeroed at the start of every tile render
    unsigned long long work[kMaxLaunches];   // dynamic-fetch cursors, one per launch of the tile
    uint32_t queue_count[16];        // paths alive at depth d (d >= 1)
    uint32_t own_count;              // chunks of the tile this device took from the (possibly shared) cursor
    unsigned long long steal_local;  // the cursor of a tile rendered by this device alone
    uint32_t static_next;            // shared frame with a declared partition: next entry of this device's dealt share
    uint32_t ovf_count[40];          // rays parked for the k_overflow launch 2*depth + {0: bounce, 1: shadow}
    uint32_t ovf_cursor[40];         // k_overflow's warp-cooperative pass: next parked ray to take
    uint32_t huge_count[40];         // ... rays it handed on to k_overflow_huge
};
struct DevTotals {                   // running ray / test counters (never reset by a tile)
    unsigned long long rays_primary, rays_shadow, rays_reflection, box_tests, tri_tests;
    unsigned long long rays_overflow;    // rays whose DFS ran past the budget
    unsigned long long rays_in_place;    // ... of which the parking buffer was full: finished by their own thread
    unsigned long long box_exact, tri_exact;   // tests the fp32 filters left to the fp64 arithmetic (CT_FLAG_COUNT_TESTS)
};

struct Params {
    const DevPair32 *pairs32;
    const DevPair64 *pairs64;
    double root_min[3], root_max[3];     // node 0
    float root_min32[3], root_max32[3];  // ... as floats, for the slab filter
    uint32_t root_ref, root_cnt;
    double bound[3];                     // >= |b| for every node bound b per axis (+inf disables the filter), see ray_finish
    const DevTri *tris;
    const DevTri32 *tris32;
    const ct_material *materials;    // by original id
    const DevLight *lights;
    const DevShadowLight *slights;
    uint32_t n_lights, n_slights, n_tri, n_nodes, n_pairs;
    uint32_t occ_words;              // words of occlusion bits per path = ceil(n_lights / 32)
    uint32_t pos_of_tri0;            // leaf position of original triangle 0 (closestIndex default, raythread.cpp:205)
    uint32_t budget;                 // see kDefaultBudget
    uint32_t warp_budget;            // see kWarpBudget
    double cam[3], rot[9];
    float vp_w, vp_h, vp_d;
    int W, H, max_depth;
    uint32_t background;
    // tile
    int x_lo, n_x, y_lo, n_y;        // canvas x in [x_lo, x_lo+n_x), n_y traced rows starting at y_lo
    int subsample, n_rows;           // CT_FLAG_SUBSAMPLING: the tile spans n_rows canvas rows of which every other one is traced
    int supersample;                 // CT_FLAG_SUPERSAMPLING: 16 consecutive slots = the 4x4 jittered samples of one pixel
    uint32_t *final_color;           // the traced pixels' / samples' colours by slot, for k_subsample / k_supersample
    int blocks_x;                    // ceil(n_x / 8): pixel blocks of 8x4 per warp
    uint32_t n_slots;                // blocks_x * ceil(n_y/4) * 32
    uint32_t cap;                    // capacity of every per-slot array
    // per-slot path state
    float *hit0_t; uint32_t *hit0_pos;           // depth-0 hit records, by slot (pos = kNoPos: miss)
    float *hitb_t; uint32_t *hitb_pos;           // depth>=1 hits, by queue slot
    double *ray_buf[2];                          // depth>=1 rays: 6 doubles per queue slot, ping-pong
    uint32_t *path_slot[2];                      // queue slot -> depth-0 slot, ping-pong
    uint32_t *occ;                               // [path][occ_words] shadow-ray verdicts of the current depth, bit i = light i occluded
    uint32_t *stack_color; float *stack_refl;    // [depth][slot]
    uint8_t *term_level;                         // [slot] level at which the chain ended
    uint32_t *fb;                                // W*H, this device's framebuffer
    uint32_t *fb_out;                            // where finished pixels are stored: fb, or the root GPU's fb (peer memory)
    unsigned long long *steal;                   // the tile's chunk cursor: local, or on the root GPU (peer memory)
    uint32_t chunk_shift;                        // log2(slots per chunk)
    uint32_t steal_stride;                       // 1; R > 1 (option "emulate_ranks") takes every R-th chunk only: the share of one of R GPUs
    uint32_t part_index, part_count;             // shared frame: this device is participant part_index of part_count (0: not declared)
    uint32_t static_eighths;                     // ... of every 8 * part_count chunks, static_eighths * part_count are dealt, the rest stolen
    uint32_t *own_chunks;                        // chunk numbers this device took, in the order it took them
    uint32_t *dbg_found, *dbg_index; float *dbg_t;   // optional (CT_FLAG_KEEP_HITS), framebuffer layout
    // parked rays
    OvfRay *ovf; uint32_t ovf_cap;
    uint32_t *ovf_huge;                          // indices (into ovf) of the rays k_overflow left to k_overflow_huge
    const uint32_t *pair_parent;                 // pair -> 2 * parent pair + side of its own box (kNoPos for the root's children pair)
    const uint32_t *tri_parent;                  // leaf position -> 2 * pair + side of the box of the leaf that holds it (kNoPos: root leaf)
    DevSched *sched;
    DevTotals *tot;
};

struct LocalCount { uint32_t box = 0, tri = 0, box_exact = 0, tri_exact = 0; };

CT_DEV V3 ld3(const double *p) { return {p[0], p[1], p[2]}; }

CT_DEV void load_tri(const DevTri *tris, uint32_t pos, V3 &p1, V3 &e1, V3 &e2) {
    const double2 *p = reinterpret_cast<const double2 *>(tris + pos);
    double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
    double e = __ldg(reinterpret_cast<const double *>(p + 4));
    p1 = {a.x, a.y, b.x};
    e1 = {b.y, c.x, c.y};
    e2 = {d.x, d.y, e};
}

// Optional (CT_PREFETCH=1): request the records of a pair's children as soon as the pair has arrived, so that their
// latency overlaps this visit's slab tests.  Measured neutral-to-negative on full frames (the walks are issue
// bound there), kept as a build-time experiment.
CT_DEV void prefetch_children(const Params &P, const DevPair32 &pr) {
#if defined(CT_PREFETCH) && CT_PREFETCH
    const char *l = pr.l_cnt ? reinterpret_cast<const char *>(P.tris32 + pr.l_ref) : reinterpret_cast<const char *>(P.pairs32 + pr.l_ref);
    const char *r = pr.r_cnt ? reinterpret_cast<const char *>(P.tris32 + pr.r_ref) : reinterpret_cast<const char *>(P.pairs32 + pr.r_ref);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(l)); asm volatile("prefetch.global.L2 [%0];" ::"l"(l + 32));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(r)); asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 32));
#endif
}

CT_DEV void load_pair32(const DevPair32 *pairs, uint32_t pid, DevPair32 &p) {
    const float4 *q = reinterpret_cast<const float4 *>(pairs + pid);
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    uint4 m = __ldg(reinterpret_cast<const uint4 *>(q + 3));
    p.lmin[0] = a.x; p.lmin[1] = a.y; p.lmin[2] = a.z; p.lmax[0] = a.w; p.lmax[1] = b.x; p.lmax[2] = b.y;
    p.rmin[0] = b.z; p.rmin[1] = b.w; p.rmin[2] = c.x; p.rmax[0] = c.y; p.rmax[1] = c.z; p.rmax[2] = c.w;
    p.l_ref = m.x; p.l_cnt = m.y; p.r_ref = m.z; p.r_cnt = m.w;
}

// ---- cold paths: the reference's own fp64 arithmetic, out of line so that its operands only occupy registers
// while it runs.  `r64` = the ray's origin (0..2) and direction (3..5) in local memory.
__device__ __noinline__ BoxTimes exact_child(const DevPair64 *pairs, uint32_t pid, uint32_t side, double *r64) {
    const double2 *q = reinterpret_cast<const double2 *>(pairs + pid) + 3u * side;
    double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const double bmin[3] = {a.x, a.y, b.x}, bmax[3] = {b.y, c.x, c.y};
    return box_times(r64, bmin, bmax);
}

CT_DEV bool exact_root(const Params &P, double *r64, float ray_t) {
    return box_accept(box_times(r64, P.root_min, P.root_max), ray_t);
}

// IntersectAABB's verdict for the root box: the fp32 bracket when it is certain, the fp64 arithmetic otherwise.
CT_DEV bool root_accept(const Params &P, const TRay &r) {
    if (r.filt) {
        const BoxBracket b = box_filter(r, P.root_min32, P.root_max32);
        if (bracket_geom_no(b) | bracket_t_no(b, r.t)) return false;
        if (bracket_geom_yes(b) & bracket_t_yes(b, r.t)) return true;
    }
    return exact_root(P, r.r64, r.t);
}

struct TriHit { bool hit; float t; };
__device__ __noinline__ TriHit tri_exact(const DevTri *tris, uint32_t pos, const double *r64) {
    V3 p1, e1, e2;
    load_tri(tris, pos, p1, e1, e2);
    Ray r;
    r.o = {r64[0], r64[1], r64[2]}; r.d = {r64[3], r64[4], r64[5]}; r.t = 0.0f;
    TriHit h;
    h.t = 0.0f;
    h.hit = intersect_triangle(r, p1, e1, e2, &h.t);
    return h;
}

// IntersectTriangle's verdict for the triangle at `pos`: the fp32 filter discards what certainly has no effect,
// the fp64 arithmetic decides the rest.
template <bool ANY_HIT, bool COUNT>
CT_DEV TriHit leaf_triangle(const Params &P, const TRay &r, uint32_t pos, LocalCount &lc) {
    const float4 *q = reinterpret_cast<const float4 *>(P.tris32 + pos);
    const bool miss = tri_filter_miss<ANY_HIT>(r, __ldg(q), __ldg(q + 1), __ldg(q + 2));
    if (r.tfilt & miss) return {false, 0.0f};
    if (COUNT) lc.tri_exact++;
    return tri_exact(P.tris, pos, r.r64);
}

// IntersectAABB's verdicts for the two children of pair `pid`, bit-exact: the fp32 brackets decide when they
// can (no branch on the way), the fp64 arithmetic otherwise.  On return [r_lo, r_hi] brackets the reference's
// tmin of the RIGHT child (collapsed to the exact value when the fp64 path ran), which is what its deferred
// `tmin < ray.t` re-check needs.
// T_FAR: the caller's ray.t is 1e30f for good (shadow rays); a filtered ray has |quotients| < 2^99 < 1e30 (tray_setup),
// so `tmin < ray.t` needs no test.
template <bool COUNT, bool T_FAR = false>
CT_DEV void pair_accept(const Params &P, const TRay &r, uint32_t pid, const DevPair32 &pr, bool &hit_l, bool &hit_r,
                        float &r_lo, float &r_hi, LocalCount &lc) {
    const BoxBracket bl = box_filter(r, pr.lmin, pr.lmax), br = box_filter(r, pr.rmin, pr.rmax);
    const bool no_l = T_FAR ? bracket_geom_no(bl) : (bracket_geom_no(bl) | bracket_t_no(bl, r.t));
    const bool yes_l = T_FAR ? bracket_geom_yes(bl) : (bracket_geom_yes(bl) & bracket_t_yes(bl, r.t));
    const bool no_r = T_FAR ? bracket_geom_no(br) : (bracket_geom_no(br) | bracket_t_no(br, r.t));
    const bool yes_r = T_FAR ? bracket_geom_yes(br) : (bracket_geom_yes(br) & bracket_t_yes(br, r.t));
    hit_l = yes_l; hit_r = yes_r;
    r_lo = br.near_lo; r_hi = br.near_hi;
    const bool open_l = !r.filt | !(no_l | yes_l), open_r = !r.filt | !(no_r | yes_r);
    if (open_l | open_r) {
        if (open_l) {
            if (COUNT) lc.box_exact++;
            hit_l = box_accept(exact_child(P.pairs64, pid, 0u, r.r64), r.t);
        }
        if (open_r) {
            if (COUNT) lc.box_exact++;
            BoxTimes e = exact_child(P.pairs64, pid, 1u, r.r64);
            r_lo = r_hi = e.tmin;
            hit_r = box_accept(e, r.t);
        }
    }
}

enum TraverseMode { kClosest, kAnyHit, kFirstLine };
enum { kTravMiss = 0, kTravHit = 1, kTravOverBudget = -1 };

constexpr uint32_t kFullMask = 0xffffffffu;

// Build-time bounds checks (-DCT_DEBUG_BOUNDS=1; compute-sanitizer is not available on every pool): trap with a message.
#if defined(CT_DEBUG_BOUNDS) && CT_DEBUG_BOUNDS
#define CT_CHECK(cond) do { if (!(cond)) { printf("CT_CHECK failed: %s (line %d)\n", #cond, __LINE__); __trap(); } } while (0)
#else
#define CT_CHECK(cond) do { } while (0)
#endif

// IntersectBVHClosest (bvh.cpp:198-222) as an explicit-stack DFS in the reference's visit order (left subtree,
// then right) -- the closest-hit walk of primary rays and of ct_gpu_debug_closest (any initial ray.t).
// The reference tests a node's box when it VISITS the node; here both children of a passing interior node are
// fetched and tested together (one 64-byte fetch, two independent slab tests in flight):
//   * the left child is visited next, so "now" is its visit time;
//   * the right child's tmin/tmax do not depend on ray.t; of the three accept conditions (bvh.cpp:178) only
//     `tmin < ray.t` does, and ray.t only ever decreases -- so a right child failing now fails at visit time too
//     and is dropped, and one that passes now is pushed WITH (a bracket of) its tmin and re-checked against the
//     then-current ray.t when popped.  Same boxes accepted, same triangles tested in the same order, same counts.
// Leaves are tested on the spot: ray.t must be up to date for the next box (an early-exit walk may defer them,
// traverse_early; this one may not, and for the same reason it cannot be parked and finished out of order).
// WARP-SYNCHRONOUS: all 32 lanes call it (lanes without a ray pass active = false); every iteration = one node
// visit per live lane, and the lanes re-converge at the vote that ends it (left to itself the compiler lets the
// lanes of a warp drift apart for the whole walk: measured 8 of 32 lanes active).
// Returns kTravHit/kTravMiss = ray.t != 1e30f ("found").
template <bool COUNT>
CT_DEV int traverse_closest(const Params &P, TRay &r, bool active, float &tclosest, uint32_t &closest_pos, LocalCount &lc) {
    // stack entry = a pushed right child: (ref, cnt), the bracket of its tmin and its parent pair (to find its fp64
    // bounds again when the bracket cannot decide)
    uint32_t stk_ref[kStackMax], stk_cnt[kStackMax], stk_src[kStackMax];
    float stk_lo[kStackMax], stk_hi[kStackMax];
    int sp = 0;
    tclosest = kFinf;          // raythread.cpp:204
    closest_pos = kNoPos;      // "closestIndex = 0" default, resolved by the caller via pos_of_tri0
    uint32_t cur_ref = P.root_ref, cur_cnt = P.root_cnt;   // current (already accepted) node
    bool live = false;
    if (active) {
        if (COUNT) lc.box++;
        live = root_accept(P, r);
    }
    while (__any_sync(kFullMask, live)) {
        if (live) {
            bool need_pop = true;
            if (cur_cnt > 0) {
                for (uint32_t i = 0; i < cur_cnt; i++) {
                    uint32_t pos = cur_ref + i;
                    if (COUNT) lc.tri++;
                    const TriHit th = leaf_triangle<false, COUNT>(P, r, pos, lc);
                    if (th.hit) {
                        if (th.t > kEps) r.t = macro_min(r.t, th.t);               // bvh.cpp:161
                        if (r.t != kRayTInit && r.t < tclosest) {                  // bvh.cpp:212
                            closest_pos = pos; tclosest = r.t;
                        }
                    }
                }
            } else {
                CT_CHECK(cur_ref < P.n_pairs);
                DevPair32 pr;
                load_pair32(P.pairs32, cur_ref, pr);
                prefetch_children(P, pr);
                if (COUNT) lc.box += 2;
                bool hit_l, hit_r; float r_lo, r_hi;
                pair_accept<COUNT>(P, r, cur_ref, pr, hit_l, hit_r, r_lo, r_hi, lc);
                if (hit_l & hit_r) {
                    CT_CHECK(sp < kStackMax);
                    stk_ref[sp] = pr.r_ref; stk_cnt[sp] = pr.r_cnt;
                    stk_lo[sp] = r_lo; stk_hi[sp] = r_hi; stk_src[sp] = cur_ref;
                    sp++;
                }
                if (hit_l | hit_r) {
                    cur_ref = hit_l ? pr.l_ref : pr.r_ref; cur_cnt = hit_l ? pr.l_cnt : pr.r_cnt;
                    need_pop = false;
                }
            }
            if (need_pop) {
                live = false;
                while (sp > 0) {
                    --sp;
                    if (stk_lo[sp] >= r.t) continue;                          // the deferred `tmin < ray.t` of bvh.cpp:178
                    if (!(stk_hi[sp] < r.t)) {
                        if (COUNT) lc.box_exact++;
                        BoxTimes e = exact_child(P.pairs64, stk_src[sp], 1u, r.r64);
                        if (!(e.tmin < r.t)) continue;
                    }
                    cur_ref = stk_ref[sp]; cur_cnt = stk_cnt[sp]; live = true;
                    break;
                }
            }
        }
    }
    if (!active) return kTravMiss;
    return r.t != kRayTInit ? kTravHit : kTravMiss;
}

// The two early-exit walks.
//   kAnyHit    shadow rays (ray.t = 1e30f): only `found` is used (raythread.cpp:306), i.e. whether SOME triangle
//              reachable through accepted boxes has a barycentric pass with 1e-4 < t < 1e30;
//   kFirstLine reflection rays (ray.t = 0, raythread.cpp:373): the first barycentric pass in DFS order becomes
//              closestIndex with tclosest = 0 = the passing reachable triangle with the LOWEST leaf
//              position (leaf positions increase along the DFS).
// ray.t never changes before the exit, so the set of accepted boxes is fixed and the moment a leaf is tested
// cannot change the answer.  The loop therefore walks interior nodes only and DEFERS accepted leaves to a short
// list, in DFS order; the warp alternates between a walk phase and a leaf phase in which its lanes test their
// triangles together, oldest leaf first, instead of one lane at a time in the middle of the walk (measured: 4 of
// 32 lanes active in an inline leaf path, 16 in the leaf phase).  A leaf phase runs when some lane's list is full
// and after the walk; kFirstLine stops at the first pass of a phase (every leaf before it has been tested).
// WARP-SYNCHRONOUS like traverse_closest().  Returns kTravHit (kAnyHit: occluded; kFirstLine: always -- `found` is
// 0 != 1e30f, raythread.cpp:227 -- with closest_pos = kNoPos when nothing passed), kTravMiss or kTravOverBudget.
#ifndef CT_LEAF_LIST
#define CT_LEAF_LIST 8
#endif
constexpr int kLeafList = CT_LEAF_LIST;       // deferred leaves per lane before a leaf phase is forced
template <TraverseMode MODE, bool COUNT>
CT_DEV int traverse_early(const Params &P, const TRay &r, bool active, const uint32_t budget, float &tclosest, uint32_t &closest_pos, LocalCount &lc) {
    static_assert(MODE != kClosest, "closest-hit rays use traverse_closest");
    uint32_t stk_ref[kStackMax], stk_cnt[kStackMax];     // pushed right children
    uint32_t leaf_ref[kLeafList], leaf_cnt[kLeafList];   // deferred leaves, DFS order
    int sp = 0, nleaf = 0;
    uint32_t spent = 1u;
    uint32_t cur_ref = P.root_ref, cur_cnt = P.root_cnt;
    int state = 0;                                        // 1: nodes left to walk (cur_* pending); 0: walk finished
    int result = MODE == kFirstLine ? kTravHit : kTravMiss;
    tclosest = kFinf;          // raythread.cpp:204
    closest_pos = kNoPos;      // "closestIndex = 0" default, resolved by the caller via pos_of_tri0
    if (active) {
        if (COUNT) lc.box++;
        state = root_accept(P, r) ? 1 : 0;
    } else {
        result = kTravMiss;
    }
    while (true) {
        // ---- walk phase: one interior-node visit per walking lane and iteration; leaves go to the list
        while (__any_sync(kFullMask, (state == 1) & (nleaf < kLeafList))) {
            if ((state == 1) & (nleaf < kLeafList)) {
                bool descend = false;
                if (cur_cnt > 0) {
                    CT_CHECK(nleaf < kLeafList && cur_ref + cur_cnt <= P.n_tri);
                    leaf_ref[nleaf] = cur_ref; leaf_cnt[nleaf] = cur_cnt; nleaf++;
                    spent += cur_cnt;
                } else {
                    CT_CHECK(cur_ref < P.n_pairs);
                    DevPair32 pr;
                    load_pair32(P.pairs32, cur_ref, pr);
                    if (COUNT) lc.box += 2;
                    spent += 2u;
                    bool hit_l, hit_r; float r_lo, r_hi;
                    pair_accept<COUNT, MODE == kAnyHit>(P, r, cur_ref, pr, hit_l, hit_r, r_lo, r_hi, lc);
                    if (hit_l & hit_r) { CT_CHECK(sp < kStackMax); stk_ref[sp] = pr.r_ref; stk_cnt[sp] = pr.r_cnt; sp++; }
                    descend = hit_l | hit_r;
                    cur_ref = hit_l ? pr.l_ref : pr.r_ref; cur_cnt = hit_l ? pr.l_cnt : pr.r_cnt;
                }
                if (!descend) {
                    if (sp == 0) state = 0;
                    else { --sp; cur_ref = stk_ref[sp]; cur_cnt = stk_cnt[sp]; }
                }
                if (spent > budget) { result = kTravOverBudget; state = 0; nleaf = 0; }
            }
        }
        // ---- leaf phase: one triangle per lane and iteration, oldest leaf first
        int li = 0;
        uint32_t tri = 0;                                 // next triangle inside leaf li
        while (__any_sync(kFullMask, li < nleaf)) {
            if (li < nleaf) {
                const uint32_t pos = leaf_ref[li] + tri;
                if (COUNT) lc.tri++;
                const TriHit th = leaf_triangle<MODE == kAnyHit, COUNT>(P, r, pos, lc);
                const bool done = MODE == kAnyHit ? (th.hit & (th.t > kEps) & (th.t < kRayTInit)) : th.hit;
                if (done) {
                    if (MODE == kAnyHit) result = kTravHit;
                    else { closest_pos = pos; tclosest = 0.0f; }
                    state = 0; nleaf = 0;
                } else if (++tri == leaf_cnt[li]) { tri = 0; li++; }
            }
        }
        nleaf = 0;
        if (!__any_sync(kFullMask, state == 1)) break;
    }
    return result;
}

// The counter-based stand-in for rand() in the supersampling jitter (the parity harness patches the same
// function into the compiled reference in place of rand(); DESIGN.md, sampling modes).
CT_DEV uint32_t hash3(uint32_t x, uint32_t y, uint32_t k) {
    uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u) ^ (k * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

// Canvas point of sample k (= 4 xs + ys) of pixel (x, y): the reference's float bookkeeping (raythread.cpp:461-505)
// replayed up to that sample -- sampleX/sampleY are jittered, used, un-jittered and stepped in float, so each
// sample's point depends on the rounding of the ones before it.
CT_DEV void sample_point(int x, int y, int k, float &px, float &py) {
    const float stepsize = 0.25f, jitter = 0.03125f;          // 1/(float)samples, stepsize/8
    float sample_x = (float)x, sample_y = (float)y;
    uint32_t call = 0;
    for (int xs = 0; xs < 4; xs++) {
        sample_y = (float)y;
        for (int ys = 0; ys < 4; ys++) {
            float rx = __fdiv_rn(__int2float_rn((int)(hash3((uint32_t)x, (uint32_t)y, call++) & 0x7fffffffu)), 2147483648.0f);   // (float)RAND_MAX
            float ry = __fdiv_rn(__int2float_rn((int)(hash3((uint32_t)x, (uint32_t)y, call++) & 0x7fffffffu)), 2147483648.0f);
            rx = __fsub_rn(__fmul_rn(rx, 0.0625f), jitter);
            ry = __fsub_rn(__fmul_rn(ry, 0.0625f), jitter);
            sample_x = __fadd_rn(sample_x, rx);
            sample_y = __fadd_rn(sample_y, ry);
            if (xs * 4 + ys == k) { px = sample_x; py = sample_y; return; }
            sample_y = __fadd_rn(__fsub_rn(sample_y, ry), stepsize);
            sample_x = __fsub_rn(sample_x, rx);
        }
        sample_x = __fadd_rn(sample_x, stepsize);
    }
    px = sample_x; py = sample_y;
}

// Primary ray through canvas point (px, py): CanvasToViewport (raythread.cpp:186-194) * camera.rotation (mymath.h:68-75)
CT_DEV Ray primary_ray_at(const Params &P, float px, float py) {
    float hh = (float)P.H;                                  // "Keep it square": both scales use bitmap->height
    double sx = (double)__fdiv_rn(P.vp_w, hh), sy = (double)__fdiv_rn(P.vp_h, hh);
    double vx = __dmul_rn((double)px, sx), vy = __dmul_rn((double)py, sy), vz = (double)P.vp_d;
    Ray r;
    r.o = {P.cam[0], P.cam[1], P.cam[2]};
    r.d.x = __dadd_rn(__dadd_rn(__dmul_rn(vx, P.rot[0]), __dmul_rn(vy, P.rot[3])), __dmul_rn(vz, P.rot[6]));
    r.d.y = __dadd_rn(__dadd_rn(__dmul_rn(vx, P.rot[1]), __dmul_rn(vy, P.rot[4])), __dmul_rn(vz, P.rot[7]));
    r.d.z = __dadd_rn(__dadd_rn(__dmul_rn(vx, P.rot[2]), __dmul_rn(vy, P.rot[5])), __dmul_rn(vz, P.rot[8]));
    r.t = kRayTInit;
    return r;
}

// Primary ray of depth-0 slot `slot`, whose pixel is (x, y): the pixel centre, or one of its 16 jittered samples.
CT_DEV Ray primary_ray(const Params &P, uint32_t slot, int x, int y) {
    if (!P.supersample) return primary_ray_at(P, (float)x, (float)y);
    float px, py;
    sample_point(x, y, (int)(slot & 15u), px, py);
    return primary_ray_at(P, px, py);
}

// slot -> canvas pixel.  A warp owns an 8x4 pixel block (coherent rays); returns false for padding lanes
// and for pixels PutPixel would drop (draw2d.h:11-14), which are not traced at all (with subsampling a dropped
// row is still traced -- fb_index = -1 -- because its colour enters the average stored in the row above it).
CT_DEV void store_pixel(const Params &P, uint32_t slot, int fb_index, uint32_t color) {
    if (fb_index >= 0) P.fb_out[fb_index] = color;
    if (P.final_color) P.final_color[slot] = color;
}

CT_DEV bool slot_pixel(const Params &P, uint32_t slot, int &x, int &y, int &fb_index) {
    if (P.supersample) slot >>= 4;                            // 16 samples per pixel
    uint32_t blk = slot >> 5, lane = slot & 31u;
    int bx = (int)(blk % (uint32_t)P.blocks_x), by = (int)(blk / (uint32_t)P.blocks_x);
    int ix = bx * 8 + (int)(lane & 7u), iy = by * 4 + (int)(lane >> 3);
    if (ix >= P.n_x || iy >= P.n_y) return false;
    x = P.x_lo + ix;
    if (P.subsample)    // raythread.cpp:527-530: y += 2, except that the last row of the partition is always traced
        y = P.y_lo + ((iy == P.n_y - 1 && (P.n_rows & 1) == 0) ? P.n_rows - 1 : 2 * iy);
    else
        y = P.y_lo + iy;
    int col = x + P.W / 2, row = P.H / 2 - y;               // CanvasPutPixel raythread.cpp:181-182
    if (col < 0 || col >= P.W) return false;
    fb_index = row * P.W + col;
    if (row < 0 || row >= P.H) {
        if (!P.subsample) return false;
        fb_index = -1;                                       // traced for the average of the row above it, never stored
    }
    if (P.supersample) fb_index = -1;                        // a sample: k_supersample blends the 16 of a pixel and stores it
    return true;
}

CT_DEV unsigned long long warp_fetch(unsigned long long *cursor) {   // persistent warps pull 32 work items at a time
    unsigned long long base = 0;
    if ((threadIdx.x & 31u) == 0) base = atomicAdd(cursor, 32ull);
    return __shfl_sync(0xffffffffu, base, 0);
}

CT_DEV void warp_add(unsigned long long *dst, uint32_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31u) == 0 && v) atomicAdd(dst, (unsigned long long)v);
}

// Depth-0 paths are numbered in the order this device took their chunks from the cursor.
CT_DEV uint32_t own_slot(const Params &P, uint32_t q) {
    return (P.own_chunks[q >> P.chunk_shift] << P.chunk_shift) + (q & ((1u << P.chunk_shift) - 1u));
}
CT_DEV uint32_t depth0_count(const Params &P) { return P.sched->own_count << P.chunk_shift; }

// The path with queue index q at `depth`: its ray (direction only for depth 0 is regenerated from the pixel),
// its closest-hit record and its depth-0 slot.  False for padding lanes / untraced pixels.
CT_DEV bool load_path(const Params &P, int depth, uint32_t q, uint32_t &slot, int &fbi, Ray &r, float &tc, uint32_t &pos) {
    if (depth == 0) {
        int x, y;
        slot = own_slot(P, q);
        if (!slot_pixel(P, slot, x, y, fbi)) return false;
        r = primary_ray(P, slot, x, y);
        tc = P.hit0_t[slot]; pos = P.hit0_pos[slot];
    } else {
        const int cur = depth & 1;
        slot = P.path_slot[cur][q];
        fbi = 0;
        const double2 *rb = reinterpret_cast<const double2 *>(P.ray_buf[cur] + 6ull * q);
        double2 a = rb[0], b = rb[1], c = rb[2];
        r.o = {a.x, a.y, b.x}; r.d = {b.y, c.x, c.y}; r.t = 0.0f;
        tc = P.hitb_t[q]; pos = P.hitb_pos[q];
        if (pos == kNoPos) pos = P.pos_of_tri0;                 // "closestIndex = 0" (raythread.cpp:205): no barycentric pass at all
    }
    return true;
}

CT_DEV void clear_occ(const Params &P, uint32_t q) {
    for (uint32_t w = 0; w < P.occ_words; w++) P.occ[(size_t)q * P.occ_words + w] = 0u;
}

// Park a ray for k_overflow.  False when the buffer is full (the caller then finishes the ray in place).
CT_DEV bool park_ray(const Params &P, int ovf_idx, const double *r64, uint32_t target, uint32_t bit) {
    uint32_t i = atomicAdd(&P.sched->ovf_count[ovf_idx], 1u);
    if (i >= P.ovf_cap) { atomicAdd(&P.tot->rays_in_place, 1ull); return false; }
    OvfRay &o = P.ovf[i];
    o.o[0] = r64[0]; o.o[1] = r64[1]; o.o[2] = r64[2];
    o.d[0] = r64[3]; o.d[1] = r64[4]; o.d[2] = r64[5];
    o.target = target; o.bit = bit;
    return true;
}

// Primary rays.  Persistent warps take chunks of 32 or 64 slots from the tile's cursor -- one counter for the whole
// tile, which in a multi-GPU frame lives on the root GPU and is shared by all devices over NVLink (dynamic
// stealing at chunk granularity) -- and remember which chunks they took: the later stages of this
// device work on exactly those.
// Which chunk of the tile does this warp trace next?  (called by lane 0)
//   one device, or a shared frame without a declared partition: the next one from the cursor (P.steal: this device's own
//     or, shared, the root GPU's over NVLink) -- pure dynamic stealing;
//   shared frame of R declared participants: the chunks are numbered in groups of 8R; of every group the first E*R are
//     DEALT (participant r owns r, r + R, ...: no atomics on another GPU, and -- being interleaved at 32-pixel grain --
//     an equal share of every later stage's work too), the other (8 - E)*R are STOLEN from the root's cursor (absorbs a
//     slower or busier GPU).  Stealing everything balances only this kernel: the GPU that holds the cursor steals
//     cheaper and ends up with more paths to light (measured at 8 GPUs: 1.75 ms on the root against 1.41 ms elsewhere).
CT_DEV bool next_chunk(const Params &P, uint32_t n_chunks, bool &dealt_left, uint32_t &idx) {
    const uint32_t R = P.part_count, E = P.static_eighths;
    if (R <= 1u) {
        const unsigned long long d = atomicAdd(P.steal, (unsigned long long)P.steal_stride);
        idx = (uint32_t)d;
        return d < n_chunks;
    }
    const uint32_t G = 8u * R, n_groups = (n_chunks + G - 1u) / G;
    while (dealt_left) {
        const uint32_t c = atomicAdd(&P.sched->static_next, 1u);
        const uint32_t g = c / E;
        if (g >= n_groups) { dealt_left = false; break; }
        idx = g * G + (c - g * E) * R + P.part_index;
        if (idx < n_chunks) return true;
    }
    const uint32_t per = (8u - E) * R;
    while (per) {
        const unsigned long long d = atomicAdd(P.steal, 1ull);
        const unsigned long long g = d / per;
        if (g >= n_groups) break;
        idx = (uint32_t)(g * G + E * R + (d - g * per));
        if (idx < n_chunks) return true;
    }
    return false;
}

template <bool COUNT>
__global__ void __launch_bounds__(kBlockThreads, kMinBlocks) k_primary(const __grid_constant__ Params P) {
    LocalCount lc;
    uint32_t n_rays = 0;
    const uint32_t lane = threadIdx.x & 31u, chunk = 1u << P.chunk_shift;
    const uint32_t n_chunks = (P.n_slots + chunk - 1u) >> P.chunk_shift;
    bool dealt_left = P.part_count > 1u && P.static_eighths > 0u;     // lane 0's view of this device's dealt share
    while (true) {
        unsigned long long base = ~0ull;
        uint32_t mine = 0;
        if (lane == 0) {
            uint32_t idx;
            if (next_chunk(P, n_chunks, dealt_left, idx)) {
                base = (unsigned long long)idx << P.chunk_shift;
                mine = atomicAdd(&P.sched->own_count, 1u);
                CT_CHECK(mine <= (P.cap >> kChunkLocalShift));
                P.own_chunks[mine] = idx;
            }
        }
        base = __shfl_sync(kFullMask, base, 0);
        mine = __shfl_sync(kFullMask, mine, 0);
        if (base == ~0ull) break;
        for (uint32_t sub = 0; sub < chunk; sub += 32u) {
            const uint32_t slot = (uint32_t)base + sub + lane;
            const uint32_t q = (mine << P.chunk_shift) + sub + lane;   // this path's depth-0 number on this device
            int x, y, fbi;
            const bool active = slot < P.n_slots && slot_pixel(P, slot, x, y, fbi);
            double r64[kRay64];
            TRay r;
            if (active) {
                Ray ray = primary_ray(P, slot, x, y);
                tray_setup(r, ray, P.bound, r64);
            }
            float tc; uint32_t pos;
            bool found = traverse_closest<COUNT>(P, r, active, tc, pos, lc) == kTravHit;   // warp-synchronous
            if (!active) continue;
            n_rays++;
            CT_CHECK(slot < P.cap && q < P.cap);
            P.hit0_t[slot] = tc;
            P.hit0_pos[slot] = found ? (pos == kNoPos ? P.pos_of_tri0 : pos) : kNoPos;
            if (found) clear_occ(P, q);
            if (P.dbg_found && fbi >= 0) {
                P.dbg_found[fbi] = found ? 1u : 0u;
                P.dbg_index[fbi] = (pos == kNoPos) ? 0u : P.tris[pos].orig;
                P.dbg_t[fbi] = tc;
            }
        }
    }
    warp_add(&P.tot->rays_primary, n_rays);
    if (COUNT) { warp_add(&P.tot->box_tests, lc.box); warp_add(&P.tot->tri_tests, lc.tri); warp_add(&P.tot->box_exact, lc.box_exact); warp_add(&P.tot->tri_exact, lc.tri_exact); }
}

// ComputeLighting's shadow rays (raythread.cpp:288-306) for the paths alive at `depth`.  Work item =
// (shadow light j, path q), j-major, so the 32 lanes of a warp trace 32 neighbouring shading points towards
// the same light.  Verdicts go to the per-path occlusion mask read by k_shade.
template <bool COUNT>
__global__ void __launch_bounds__(kBlockThreads, kMinBlocks) k_shadow(const __grid_constant__ Params P, int depth, int work_idx, int ovf_idx) {
    LocalCount lc;
    uint32_t n_shadow = 0, n_parked = 0;
    const uint32_t n = depth == 0 ? depth0_count(P) : P.sched->queue_count[depth];
    const unsigned long long n_pad = ((unsigned long long)n + 31ull) & ~31ull;
    const unsigned long long total = n_pad * P.n_slights;
    while (true) {
        unsigned long long base = warp_fetch(&P.sched->work[work_idx]);
        if (base >= total) break;
        uint32_t j = (uint32_t)(base / n_pad);
        uint32_t q = (uint32_t)(base - (unsigned long long)j * n_pad) + (threadIdx.x & 31u);
        uint32_t slot, pos = kNoPos; int fbi; Ray r; float tc = 0.0f;
        bool active = q < n && load_path(P, depth, q, slot, fbi, r, tc, pos) && pos != kNoPos;
        double r64[kRay64];
        TRay tr;
        uint32_t word = 0, bit = 0;
        if (active) {
            const DevShadowLight &L = P.slights[j];
            V3 position = vadd(r.o, vscale((double)tc, r.d));                                  // :360
            V3 lray = (L.type == CT_LIGHT_POINT) ? vsub(ld3(L.v), position) : ld3(L.v);        // :288 / :293
            Ray sr; sr.o = position; sr.d = lray; sr.t = kRayTInit;                            // :304 no offset, no t<=1 test
            tray_setup(tr, sr, P.bound, r64);
            n_shadow++;
            word = q * P.occ_words + (L.index >> 5); bit = L.index & 31u;
        }
        uint32_t budget = P.budget;
        while (true) {                                                  // warp-uniform: traverse_early is warp-synchronous
            float stc; uint32_t spos;
            int res = traverse_early<kAnyHit, COUNT>(P, tr, active, budget, stc, spos, lc);
            bool again = false;
            if (active) {
                if (res == kTravOverBudget) {
                    n_parked++;
                    again = !park_ray(P, ovf_idx, r64, word, bit);      // parking buffer full: finish in place, no budget
                    budget = 0xffffffffu;
                } else if (res == kTravHit) {
                    atomicOr(&P.occ[word], 1u << bit);
                }
            }
            active = again;
            if (!__any_sync(0xffffffffu, again)) break;
        }
    }
    warp_add(&P.tot->rays_shadow, n_shadow);
    warp_add(&P.tot->rays_overflow, n_parked);
    if (COUNT) { warp_add(&P.tot->box_tests, lc.box); warp_add(&P.tot->tri_tests, lc.tri); warp_add(&P.tot->box_exact, lc.box_exact); warp_add(&P.tot->tri_exact, lc.tri_exact); }
}

// TraceRay's recursion step (raythread.cpp:369-373) for the paths alive at `depth`: does the path end here, or does
// it continue with the reflection ray {position, ReflectRay(-dir, normal), t = 0}?  Needs only the hit records --
// not the shadow verdicts -- so it runs ahead of k_shadow / k_shade of the same depth (separate streams) and feeds
// k_bounce of the next one.
__global__ void __launch_bounds__(kBlockThreads) k_emit(const __grid_constant__ Params P, int depth, int work_idx) {
    uint32_t n_refl = 0;
    const uint32_t n = depth == 0 ? depth0_count(P) : P.sched->queue_count[depth];
    const int nxt = (depth & 1) ^ 1;
    while (true) {
        unsigned long long base = warp_fetch(&P.sched->work[work_idx]);
        if (base >= n) break;
        uint32_t q = (uint32_t)base + (threadIdx.x & 31u);
        uint32_t slot = q, pos = kNoPos; int fbi = 0;
        Ray r; float tc = 0.0f;
        bool active = q < n && load_path(P, depth, q, slot, fbi, r, tc, pos);
        bool emit = false;
        V3 position = {0, 0, 0}, rdir = {0, 0, 0};
        if (active) {
            float reflection = 0.0f;
            if (pos != kNoPos) reflection = P.materials[P.tris[pos].orig].reflection;
            const int remaining = P.max_depth - depth;                      // recursionDepth of this TraceRay call
            if (pos == kNoPos || remaining <= 0 || !(reflection > 0.0f)) {  // miss :385 / :369 (reflection <= 0, NaN-safe)
                P.term_level[slot] = (uint8_t)depth;
            } else {
                V3 p1, e1, e2;
                load_tri(P.tris, pos, p1, e1, e2);
                position = vadd(r.o, vscale((double)tc, r.d));              // :360
                V3 nn = vcross(e1, e2);                                     // NormalOfSceneObject :337-339
                float dd = vdot(nn, r.d);
                V3 normal = (dd < 0.0f) ? nn : vneg(nn);                    // :341-345
                P.stack_refl[(size_t)depth * P.cap + slot] = reflection;
                rdir = reflect_ray(vneg(r.d), normal);                      // :372
                emit = true;
            }
        }
        // warp-aggregated append of the reflection rays to the next queue
        uint32_t mask = __ballot_sync(0xffffffffu, emit);
        if (mask) {
            uint32_t lane = threadIdx.x & 31u, leader = __ffs(mask) - 1, qbase = 0;
            if (lane == leader) qbase = atomicAdd(&P.sched->queue_count[depth + 1], (uint32_t)__popc(mask));
            qbase = __shfl_sync(0xffffffffu, qbase, leader);
            if (emit) {
                uint32_t nq = qbase + __popc(mask & ((1u << lane) - 1u));
                CT_CHECK(nq < P.cap && slot < P.cap);
                double2 *rb = reinterpret_cast<double2 *>(P.ray_buf[nxt] + 6ull * nq);
                rb[0] = make_double2(position.x, position.y);
                rb[1] = make_double2(position.z, rdir.x);
                rb[2] = make_double2(rdir.y, rdir.z);
                P.path_slot[nxt][nq] = slot;
                n_refl++;
            }
        }
    }
    warp_add(&P.tot->rays_reflection, n_refl);
}

// TraceRay's local colour (raythread.cpp:359-366) for the paths alive at `depth`; the shadow verdicts were computed
// by k_shadow (+ k_overflow).
__global__ void __launch_bounds__(kBlockThreads) k_shade(const __grid_constant__ Params P, int depth, int work_idx) {
    const uint32_t n = depth == 0 ? depth0_count(P) : P.sched->queue_count[depth];
    while (true) {
        unsigned long long base = warp_fetch(&P.sched->work[work_idx]);
        if (base >= n) break;
        uint32_t q = (uint32_t)base + (threadIdx.x & 31u);
        uint32_t slot = q, pos = kNoPos; int fbi = 0;
        Ray r; float tc = 0.0f;
        if (!(q < n && load_path(P, depth, q, slot, fbi, r, tc, pos))) continue;
        uint32_t *sc = P.stack_color + (size_t)depth * P.cap + slot;
        if (pos == kNoPos) {                                           // miss (depth 0 only): raythread.cpp:385
            if (depth == 0) store_pixel(P, slot, fbi, P.background);
            *sc = P.background;
            continue;
        }
        V3 p1, e1, e2;
        load_tri(P.tris, pos, p1, e1, e2);
        const ct_material mat = P.materials[P.tris[pos].orig];
        V3 position = vadd(r.o, vscale((double)tc, r.d));              // :360
        V3 nn = vcross(e1, e2);                                         // NormalOfSceneObject :337-339
        float dd = vdot(nn, r.d);
        V3 normal = (dd < 0.0f) ? nn : vneg(nn);                        // :341-345
        V3 view = vneg(r.d);
        // ---- ComputeLighting :275-327, lights in file order, fp32 accumulator
        float intensity = 0.0f;
        const uint32_t *occ = P.occ + (size_t)q * P.occ_words;
        uint32_t occ_word = 0;
        for (uint32_t i = 0; i < P.n_lights; i++) {
            const DevLight &L = P.lights[i];
            float li = L.intensity;
            if ((i & 31u) == 0) occ_word = occ[i >> 5];
            if (L.type == CT_LIGHT_AMBIENT) { intensity = __fadd_rn(intensity, li); continue; }
            if ((occ_word >> (i & 31u)) & 1u) continue;                  // :306 shadowed
            V3 lray = (L.type == CT_LIGHT_POINT) ? vsub(ld3(L.pos), position) : ld3(L.dir);
            float ndl = vdot(normal, lray);                              // :310
            if (ndl > 0.0f)
                intensity = __fadd_rn(intensity, __fdiv_rn(__fmul_rn(li, ndl), __fmul_rn(vmag(normal), vmag(lray))));
            if (mat.specular != -1) {                                    // :316
                V3 refl = reflect_ray(lray, normal);
                float rdv = vdot(refl, view);
                if (rdv > 0.0f) {                                        // :319-321 double pow, += rounds to float
                    float qv = __fdiv_rn(rdv, __fmul_rn(vmag(refl), vmag(view)));
                    double term = __dmul_rn((double)li, pow((double)qv, (double)mat.specular));
                    intensity = __double2float_rn(__dadd_rn((double)intensity, term));
                }
            }
        }
        uint32_t local = shade_color(mat.color, intensity);
        *sc = local;
        // a depth-0 path that ends here (:369) is the pixel; longer chains are blended by k_resolve
        if (depth == 0 && (P.max_depth <= 0 || !(mat.reflection > 0.0f))) store_pixel(P, slot, fbi, local);
    }
}

// Closest "hit" of the reflection rays {position, reflected, t = 0} (raythread.cpp:373).
template <bool COUNT>
__global__ void __launch_bounds__(kBlockThreads, kMinBlocks) k_bounce(const __grid_constant__ Params P, int depth, int work_idx, int ovf_idx) {
    LocalCount lc;
    uint32_t n_parked = 0;
    const uint32_t n = P.sched->queue_count[depth];
    const int cur = depth & 1;
    while (true) {
        unsigned long long base = warp_fetch(&P.sched->work[work_idx]);
        if (base >= n) break;
        uint32_t q = (uint32_t)base + (threadIdx.x & 31u);
        bool active = q < n;
        double r64[kRay64];
        TRay r;
        if (active) {
            const double2 *rb = reinterpret_cast<const double2 *>(P.ray_buf[cur] + 6ull * q);
            double2 a = rb[0], b = rb[1], c = rb[2];
            Ray ray;
            ray.o = {a.x, a.y, b.x}; ray.d = {b.y, c.x, c.y}; ray.t = 0.0f;
            tray_setup(r, ray, P.bound, r64);
            clear_occ(P, q);
        }
        uint32_t budget = P.budget;
        while (true) {                                                  // warp-uniform: traverse is warp-synchronous
            float tc; uint32_t pos;
            int res = traverse_early<kFirstLine, COUNT>(P, r, active, budget, tc, pos, lc);
            bool again = false;
            if (active) {
                if (res == kTravOverBudget) {
                    n_parked++;
                    again = !park_ray(P, ovf_idx, r64, q, 0u);          // parking buffer full: finish in place, no budget
                    budget = 0xffffffffu;
                } else {                                                // found is always true: 0 != 1e30f (:227)
                    P.hitb_t[q] = tc;
                    P.hitb_pos[q] = (pos == kNoPos) ? P.pos_of_tri0 : pos;
                }
            }
            active = again;
            if (!__any_sync(kFullMask, again)) break;
        }
    }
    warp_add(&P.tot->rays_overflow, n_parked);
    if (COUNT) { warp_add(&P.tot->box_tests, lc.box); warp_add(&P.tot->tri_tests, lc.tri); warp_add(&P.tot->box_exact, lc.box_exact); warp_add(&P.tot->tri_exact, lc.tri_exact); }
}

// Parked rays (see the file header).  Both early-exit modes have an answer that does not depend on the visit
// order, because ray.t never changes before the exit, so the set of boxes that pass is fixed:
//   kAnyHit     occluded  <=>  SOME triangle reachable through passing boxes has a bary pass with 1e-4 < t < 1e30;
//   kFirstLine  closestIndex = the bary-passing reachable triangle that the DFS meets first = the one with the
//               smallest leaf position (BuildBVH hands the left child the lower part of the parent's index range,
//               bvh.cpp:70-97, so leaf positions increase along the DFS).
// Tests the triangles of an accepted leaf for a parked ray.  kAnyHit: 1 if one of them occludes, else 0;
// kFirstLine: the first (lowest) leaf position with a barycentric pass, else kNoPos.
template <TraverseMode MODE, bool COUNT>
CT_DEV uint32_t overflow_leaf(const Params &P, const TRay &r, uint32_t first, uint32_t cnt, LocalCount &lc) {
    for (uint32_t k = 0; k < cnt; k++) {
        uint32_t pos = first + k;
        if (COUNT) lc.tri++;
        const TriHit th = leaf_triangle<MODE == kAnyHit, COUNT>(P, r, pos, lc);
        if (!th.hit) continue;
        if (MODE == kAnyHit) { if (th.t > kEps && th.t < kRayTInit) return 1u; }
        else return pos;                                      // later positions of this leaf are larger
    }
    return MODE == kAnyHit ? 0u : kNoPos;
}

template <TraverseMode MODE>
CT_DEV uint32_t overflow_merge(uint32_t a, uint32_t b) { return MODE == kAnyHit ? (a | b) : min(a, b); }

template <TraverseMode MODE>
CT_DEV void overflow_store(const Params &P, const OvfRay &o, uint32_t res) {
    if (MODE == kAnyHit) {
        if (res) atomicOr(&P.occ[o.target], 1u << o.bit);
    } else {
        P.hitb_t[o.target] = (res == kNoPos) ? kFinf : 0.0f;          // raythread.cpp:204 / first line pass
        P.hitb_pos[o.target] = (res == kNoPos) ? P.pos_of_tri0 : res;
    }
}

constexpr int kWarpStack = 1024;         // pending interior nodes of one ray in the warp-cooperative pass
constexpr uint32_t kWarpBudget = 4096;       // default node visits before a ray is handed to the grid-wide pass

// Parked rays, pass 1: one WARP per ray -- the 32 lanes pop up to 32 pending interior nodes from a shared-memory
// stack, test their child pairs, test accepted leaves on the spot and push accepted interior children back.
// A ray whose stack outgrows kWarpStack, that needs more than P.warp_budget node visits or whose origin is so far
// outside the scene that every box passes (the rays described in the header: ~1M visits) goes to k_overflow_huge.
template <TraverseMode MODE, bool COUNT>
__global__ void __launch_bounds__(kOvfThreads) k_overflow(const __grid_constant__ Params P, int ovf_idx) {
    const uint32_t n = min(P.sched->ovf_count[ovf_idx], P.ovf_cap);
    if (n == 0) return;
    LocalCount lc;
    const uint32_t lane = threadIdx.x & 31u;
    __shared__ uint32_t wstack[kOvfThreads / 32][kWarpStack];
    uint32_t *stk = wstack[threadIdx.x >> 5];
    while (true) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(&P.sched->ovf_cursor[ovf_idx], 1u);
        idx = __shfl_sync(kFullMask, idx, 0);
        if (idx >= n) break;
        const OvfRay &o = P.ovf[idx];
        Ray ray; ray.o = ld3(o.o); ray.d = ld3(o.d); ray.t = (MODE == kAnyHit) ? kRayTInit : 0.0f;
        double r64[kRay64];
        TRay r;
        tray_setup(r, ray, P.bound, r64);             // every lane holds the same ray
        uint32_t res = MODE == kAnyHit ? 0u : kNoPos;
        // An origin this far outside the scene (a shading point 2^32 ray lengths away) makes all slab
        // quotients of an axis round to the same float: every box passes and no filter can help.
        bool too_big = !r.filt || (double)r.om > 0x1p20 * fmax(fmax(P.bound[0], P.bound[1]), P.bound[2]);
        if (COUNT && lane == 0 && !too_big) lc.box++;
        if (!too_big && exact_root(P, r64, r.t)) {
            if (P.root_cnt > 0) {
                if (lane == 0) res = overflow_leaf<MODE, COUNT>(P, r, P.root_ref, P.root_cnt, lc);
            } else {
                if (lane == 0) stk[0] = P.root_ref;
                __syncwarp();
                uint32_t sp = 1, visits = 0;
                while (sp > 0) {
                    const uint32_t take = min(sp, 32u);
                    sp -= take;
                    uint32_t n_out = 0, out_a = 0, out_b = 0;
                    if (lane < take) {
                        const uint32_t pid = stk[sp + lane];
                        DevPair32 pr;
                        load_pair32(P.pairs32, pid, pr);
                        if (COUNT) lc.box += 2;
                        bool hit_l, hit_r; float lo, hi;
                        pair_accept<COUNT>(P, r, pid, pr, hit_l, hit_r, lo, hi, lc);
                        if (hit_l) {
                            if (pr.l_cnt > 0) res = overflow_merge<MODE>(res, overflow_leaf<MODE, COUNT>(P, r, pr.l_ref, pr.l_cnt, lc));
                            else { out_a = pr.l_ref; n_out = 1; }
                        }
                        if (hit_r) {
                            if (pr.r_cnt > 0) res = overflow_merge<MODE>(res, overflow_leaf<MODE, COUNT>(P, r, pr.r_ref, pr.r_cnt, lc));
                            else { if (n_out) out_b = pr.r_ref; else out_a = pr.r_ref; n_out++; }
                        }
                    }
                    __syncwarp();                     // every lane has read its entry before the pushes below
                    if (MODE == kAnyHit && __any_sync(kFullMask, res != 0u)) break;
                    uint32_t incl = n_out;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(kFullMask, incl, d); if ((int)lane >= d) incl += v; }
                    const uint32_t total = __shfl_sync(kFullMask, incl, 31);
                    visits += take;
                    if (sp + total > (uint32_t)kWarpStack || visits > P.warp_budget) { too_big = true; break; }
                    const uint32_t at = sp + incl - n_out;
                    if (n_out > 0) stk[at] = out_a;
                    if (n_out > 1) stk[at + 1u] = out_b;
                    sp += total;
                    __syncwarp();
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) res = overflow_merge<MODE>(res, __shfl_xor_sync(kFullMask, res, d));
        if (lane == 0) {
            if (too_big) {
                P.ovf_huge[atomicAdd(&P.sched->huge_count[ovf_idx], 1u)] = idx;
                if (MODE == kFirstLine) { P.hitb_t[o.target] = kFinf; P.hitb_pos[o.target] = kNoPos; }   // until k_overflow_huge finds a pass
            } else {
                overflow_store<MODE>(P, o, res);
            }
        }
    }
    if (COUNT) { warp_add(&P.tot->box_tests, lc.box); warp_add(&P.tot->tri_tests, lc.tri); warp_add(&P.tot->box_exact, lc.box_exact); warp_add(&P.tot->tri_exact, lc.tri_exact); }
}

// Is the leaf that holds a triangle REACHED by the reference's walk?  Every box on the way down must accept the ray:
// the leaf's own box, its ancestors' boxes, the root's.  `code` = 2 * pair + side of the leaf's box.
CT_DEV bool chain_accepts(const Params &P, double *r64, float ray_t, uint32_t code) {
    while (code != kNoPos) {
        const uint32_t pid = code >> 1;
        if (!box_accept(exact_child(P.pairs64, pid, code & 1u, r64), ray_t)) return false;
        code = P.pair_parent[pid];
    }
    return exact_root(P, r64, ray_t);
}

// Parked rays, pass 2 (what pass 1 gave up on).  Walking a tree in which every box passes level by level costs a
// grid-wide barrier per level; instead the whole grid tests ALL triangles against the ray at once and, for the few
// that pass, checks whether the reference's walk would have reached them at all (chain_accepts).  No barrier, and
// the answers are merged with idempotent atomics (OR into the occlusion mask / MIN of the leaf position).
template <TraverseMode MODE, bool COUNT>
__global__ void __launch_bounds__(256) k_overflow_huge(const __grid_constant__ Params P, int ovf_idx) {
    const uint32_t nh = P.sched->huge_count[ovf_idx];
    if (nh == 0) return;
    LocalCount lc;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
    for (uint32_t h = 0; h < nh; h++) {
        const OvfRay &o = P.ovf[P.ovf_huge[h]];
        Ray ray; ray.o = ld3(o.o); ray.d = ld3(o.d); ray.t = (MODE == kAnyHit) ? kRayTInit : 0.0f;
        double r64[kRay64];
        TRay r;
        tray_setup(r, ray, P.bound, r64);
        for (uint32_t pos = tid; pos < P.n_tri; pos += n_threads) {
            if (MODE == kAnyHit && (*(volatile uint32_t *)&P.occ[o.target] >> o.bit) & 1u) break;      // already occluded
            if (MODE == kFirstLine && *(volatile uint32_t *)&P.hitb_pos[o.target] < pos) break;        // a lower position already passed
            if (COUNT) lc.tri++;
            const TriHit th = leaf_triangle<MODE == kAnyHit, COUNT>(P, r, pos, lc);
            if (!th.hit) continue;
            if (MODE == kAnyHit && !(th.t > kEps && th.t < kRayTInit)) continue;
            if (!chain_accepts(P, r64, r.t, P.tri_parent[pos])) continue;
            if (MODE == kAnyHit) atomicOr(&P.occ[o.target], 1u << o.bit);
            else { atomicMin(&P.hitb_pos[o.target], pos); P.hitb_t[o.target] = 0.0f; }
        }
    }
    if (COUNT) { warp_add(&P.tot->box_tests, lc.box); warp_add(&P.tot->tri_tests, lc.tri); warp_add(&P.tot->box_exact, lc.box_exact); warp_add(&P.tot->tri_exact, lc.tri_exact); }
}

// Unwind TraceRay's recursion (raythread.cpp:375-379) for pixels whose chain went past depth 0.
__global__ void __launch_bounds__(256) k_resolve(const __grid_constant__ Params P) {
    const uint32_t n = depth0_count(P);
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t slot = own_slot(P, q);
        int x, y, fbi;
        if (!slot_pixel(P, slot, x, y, fbi)) continue;
        int lvl = P.term_level[slot];
        if (lvl == 0) continue;                              // already stored by k_shade
        uint32_t color = P.stack_color[(size_t)lvl * P.cap + slot];
        for (int d = lvl - 1; d >= 0; d--)
            color = blend_color(P.stack_color[(size_t)d * P.cap + slot], color, P.stack_refl[(size_t)d * P.cap + slot]);
        store_pixel(P, slot, fbi, color);
    }
}

// settings.subsampling (raythread.cpp:512-531): after a traced pixel (x, y) the reference stores the average of its
// colour and the previously traced colour of the column (its own for the first row of the partition) one row
// below, at (x, y - 1).  Runs after every traced pixel of the tile has its final colour; a later store wins where
// the reference's sequential loop would overwrite (the always-traced last row of an even partition).
__global__ void __launch_bounds__(256) k_subsample(const __grid_constant__ Params P) {
    const uint32_t n = depth0_count(P);
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint32_t slot = own_slot(P, q);
        int x, y, fbi;
        if (!slot_pixel(P, slot, x, y, fbi)) continue;
        const uint32_t blk = slot >> 5, lane = slot & 31u;
        const uint32_t iy = (blk / (uint32_t)P.blocks_x) * 4u + (lane >> 3);
        const uint32_t color = P.final_color[slot];
        uint32_t last = color;                                                 // :513-514
        if (iy > 0) {
            const uint32_t py = iy - 1u, ix = (blk % (uint32_t)P.blocks_x) * 8u + (lane & 7u);
            last = P.final_color[(((py >> 2) * (uint32_t)P.blocks_x + (ix >> 3)) << 5) + ((py & 3u) << 3) + (ix & 7u)];
        }
        uint32_t avg = 0;                                                      // :517-523: float (a + b) / 2, min 0xff, truncated
        for (int sh = 0; sh <= 16; sh += 8) avg |= ((((last >> sh) & 0xffu) + ((color >> sh) & 0xffu)) >> 1) << sh;
        const int col = x + P.W / 2, row = P.H / 2 - (y - 1);                  // CanvasPutPixel(bitmap, {x, y-1}, avgColor) :524
        if (row >= 0 && row < P.H && col >= 0 && col < P.W) P.fb_out[row * P.W + col] = avg;
    }
}

// settings.supersampling (raythread.cpp:460-505): the 16 samples of a pixel are folded into its colour one after the
// other -- colour -= colour/8; colour += sample/8 per channel in float, truncated to uint8 after every sample (:486-497).
__global__ void __launch_bounds__(256) k_supersample(const __grid_constant__ Params P) {
    const uint32_t n = depth0_count(P) >> 4;                                  // chunks hold whole pixels (32 or 64 slots)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = own_slot(P, i << 4);
        int x, y, fbi;
        if (!slot_pixel(P, slot, x, y, fbi)) continue;
        uint32_t color = P.final_color[slot];
        for (uint32_t k = 1; k < 16u; k++) {
            const uint32_t temp = P.final_color[slot + k];
            uint32_t out = 0;
            for (int sh = 0; sh <= 16; sh += 8) {
                float c = (float)((color >> sh) & 0xffu), t = (float)((temp >> sh) & 0xffu);
                c = __fsub_rn(c, __fdiv_rn(c, 8.0f));
                c = __fadd_rn(c, __fdiv_rn(t, 8.0f));
                out |= to_u8(c) << sh;
            }
            color = out;
        }
        const int col = x + P.W / 2, row = P.H / 2 - y;
        P.fb_out[row * P.W + col] = color;
    }
}

// ---- KAT kernels -----------------------------------------------------------------------------------------
__global__ void k_debug_closest(const __grid_constant__ Params P, uint32_t n, const double *org, const double *dir,
                                const float *t0, uint32_t *found, uint32_t *index, float *tclosest) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;                           // the traversals are warp-synchronous: every lane calls both
    Ray r;
    r.t = 1.0f;
    double r64[kRay64];
    TRay tr;
    if (active) {
        r.o = ld3(org + 3ull * i); r.d = ld3(dir + 3ull * i); r.t = t0[i];
        tray_setup(tr, r, P.bound, r64);
    }
    LocalCount lc; float tc, tc2; uint32_t pos, pos2;
    const bool first_line = r.t == 0.0f;
    bool f = traverse_early<kFirstLine, false>(P, tr, active && first_line, 0xffffffffu, tc, pos, lc) == kTravHit;
    bool f2 = traverse_closest<false>(P, tr, active && !first_line, tc2, pos2, lc) == kTravHit;
    if (!active) return;
    if (!first_line) { f = f2; tc = tc2; pos = pos2; }
    if (found) found[i] = f ? 1u : 0u;
    if (index) index[i] = (pos == kNoPos) ? 0u : P.tris[pos].orig;
    if (tclosest) tclosest[i] = tc;
}

__global__ void k_debug_primitives(uint32_t n, const double *org, const double *dir, float *ray_t, const double *tri,
                                   const double *bmin, const double *bmax, uint32_t *tri_hit, uint32_t *box_hit,
                                   uint32_t *filter_out, double bound_scale, bool tri_filter) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r;
    r.o = ld3(org + 3ull * i); r.d = ld3(dir + 3ull * i); r.t = ray_t[i];
    double mn[3] = {bmin[3ull * i], bmin[3ull * i + 1], bmin[3ull * i + 2]};
    double mx[3] = {bmax[3ull * i], bmax[3ull * i + 1], bmax[3ull * i + 2]};
    const bool exact = intersect_aabb(r, mn, mx);
    box_hit[i] = exact ? 1u : 0u;
    if (filter_out) {
        // the certified filter on the same box: bit 0 exact verdict, bits 1-2 filter (0 undecided, 1 accept, 2 reject),
        // bit 3 = the filter was usable for this ray.  A certain verdict that contradicts bit 0 is a soundness bug.
        double bnd[3];
        bool ordered = true;
        for (int k = 0; k < 3; k++) {
            bnd[k] = fmax(fabs(mn[k]), fabs(mx[k])) * bound_scale;
            ordered = ordered && (mn[k] <= mx[k]) && isfinite(mn[k]) && isfinite(mx[k]);
        }
        if (!ordered) bnd[0] = bnd[1] = bnd[2] = INFINITY;
        double r64[kRay64];
        TRay tr;
        tray_setup(tr, r, bnd, r64);
        uint32_t f = 0;
        if (tr.filt) {
            const float fmn[3] = {(float)mn[0], (float)mn[1], (float)mn[2]}, fmx[3] = {(float)mx[0], (float)mx[1], (float)mx[2]};
            BoxBracket b = box_filter(tr, fmn, fmx);
            if (bracket_geom_no(b) || bracket_t_no(b, r.t)) f = 2;
            else if (bracket_geom_yes(b) && bracket_t_yes(b, r.t)) f = 1;
            BoxTimes e = box_times(r, mn, mx);
            bool inside = b.near_lo <= e.tmin && e.tmin <= b.near_hi && b.far_lo <= e.tmax && e.tmax <= b.far_hi;
            if (!inside) f |= 8u;                                  // bracket does not contain the reference's floats: bug
        }
        // the division-free evaluation used for undecided tests must give the literal arithmetic's verdict and floats
        {
            const BoxTimes lit = box_times(r, mn, mx), rec = box_times(r64, mn, mx);
            const bool same = (lit.tmin == rec.tmin || (lit.tmin != lit.tmin && rec.tmin != rec.tmin)) &&
                              (lit.tmax == rec.tmax || (lit.tmax != lit.tmax && rec.tmax != rec.tmax));
            if (!same || box_accept(rec, r.t) != exact) f |= 16u;
        }
        uint32_t out = (exact ? 1u : 0u) | ((f & 3u) << 1) | (tr.filt ? 8u : 0u) | ((f & 8u) ? 16u : 0u) | ((f & 16u) ? 32u : 0u);
        if (tri_filter) {
            // the certified triangle filter on the same (ray, triangle), with the magnitudes the upload would store:
            // bit 8 = reference returns true, bit 9 = ... and the hit would occlude a shadow ray (1e-4 < t < 1e30),
            // bit 10 / 11 = tri_filter_miss<false> / <true> say "certainly no effect", bit 12 = filter usable
            const V3 q1 = ld3(tri + 9ull * i), q2 = ld3(tri + 9ull * i + 3), q3 = ld3(tri + 9ull * i + 6);
            const V3 e1 = vsub(q2, q1), e2 = vsub(q3, q1);
            float tt = 0.0f;
            const bool th = intersect_triangle(r, q1, e1, e2, &tt);
            const double k1 = fmax(fmax(fabs(e1.x), fabs(e1.y)), fabs(e1.z)), k2 = fmax(fmax(fabs(e2.x), fabs(e2.y)), fabs(e2.z));
            const double k3 = fmax(fmax(fabs(q1.x), fabs(q1.y)), fabs(q1.z));
            const bool in_range = k1 >= 0x1p-30 && k1 <= 0x1p30 && k2 >= 0x1p-30 && k2 <= 0x1p30 && k3 <= 0x1p40;
            const float4 t0 = make_float4((float)q1.x, (float)q1.y, (float)q1.z, __double2float_ru(k3));
            const float4 t1 = make_float4((float)e1.x, (float)e1.y, (float)e1.z, in_range ? __double2float_ru(k1) : NAN);
            const float4 t2 = make_float4((float)e2.x, (float)e2.y, (float)e2.z, __double2float_ru(k2));
            const bool m0 = tr.tfilt && tri_filter_miss<false>(tr, t0, t1, t2), m1 = tr.tfilt && tri_filter_miss<true>(tr, t0, t1, t2);
            out |= (th ? 1u << 8 : 0u) | ((th && tt > kEps && tt < kRayTInit) ? 1u << 9 : 0u) | (m0 ? 1u << 10 : 0u) | (m1 ? 1u << 11 : 0u) | (tr.tfilt ? 1u << 12 : 0u);
        }
        filter_out[i] = out;
    }
    V3 p1 = ld3(tri + 9ull * i), p2 = ld3(tri + 9ull * i + 3), p3 = ld3(tri + 9ull * i + 6);
    float t;
    bool hit = intersect_triangle(r, p1, vsub(p2, p1), vsub(p3, p1), &t);
    if (hit && t > kEps) r.t = macro_min(r.t, t);
    tri_hit[i] = hit ? 1u : 0u;
    ray_t[i] = r.t;
}

// ==== host side of the ABI ===================================================================================
thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? CT_ERR_OOM : CT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

struct DeviceState {
    bool loaded = false;
    Params p{};
    uint32_t flags = 0;
    bool any_reflective = false;
    int n_sm = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t tile_done[8] = {};   // ring: completion of the last 8 submitted tiles (ct_gpu_throttle)
    unsigned long long tiles_submitted = 0;
    int row_lo = 0, row_hi = 0;      // hull of framebuffer rows rendered since upload (readback clips to it)
    unsigned long long launches = 0; // kernels launched since upload / reset
    cudaEvent_t stage_ev[kMaxLaunches + 1] = {};   // CT_FLAG_STAGE_TIMING: boundaries between the launches of the last tile
    const char *stage_name[kMaxLaunches] = {};
    int stage_depth[kMaxLaunches] = {};
    bool can_overflow = false;       // some traversal could exceed the visit budget: k_overflow launches are needed
    // per-depth path state (render_impl hands each launch a Params view of its depth, so that the shadow / shade
    // chain of one depth can run on its own stream next to the bounce chain of the next)
    float *hitb_t_all = nullptr; uint32_t *hitb_pos_all = nullptr;      // [depth][cap]
    double *rays_all = nullptr; uint32_t *path_slot_all = nullptr;      // [depth][cap * 6], [depth][cap]
    uint32_t *occ_all = nullptr;                                        // [depth][cap * occ_words]
    OvfRay *ovf_all = nullptr; uint32_t *ovf_huge_all = nullptr;        // [2 * depth + kind][ovf_cap]
    int levels = 1;
    cudaStream_t aux[4] = {};
    cudaEvent_t ev_hit[16] = {}, ev_done[16] = {};
    // multi-GPU frame sharing (ct_gpu_share_*): the cursor and framebuffer a shared render uses (own or the root's)
    unsigned long long *cursor_own = nullptr, *share_cursor = nullptr;
    int part_index = 0, part_count = 0;      // ct_gpu_share_partition
    uint32_t *share_fb = nullptr;
    void *ipc_opened[2] = {nullptr, nullptr};
    int n_stages = 0;
    bool timed = false;
    std::vector<void *> allocs;
    ct_ray_counters snapshot{};      // totals at the end of the previous counted tile
    unsigned long long rays_overflow = 0, rays_in_place = 0;   // DevTotals' overflow counters as of the last read_totals
    unsigned long long box_exact = 0, tri_exact = 0;
    // rows rendered so far (framebuffer rows), for readback clipping
    int col_lo = 0, col_hi = 0;
};

DeviceState g_dev[kMaxDevices];
std::mutex g_mutex;
long long g_budget_option = 0;       // ct_gpu_set_option("traversal_budget"); 0 = default
long long g_warp_budget_option = 0;  // ct_gpu_set_option("overflow_warp_budget"); 0 = default
long long g_static_eighths = 7;      // ct_gpu_set_option("shared_static_eighths"), see next_chunk
long long g_shared_chunk_shift = 0;  // ct_gpu_set_option("shared_chunk_shift"): 0 = kChunkSharedShift
long long g_emulate_ranks = 0;       // ct_gpu_set_option("emulate_ranks"): profiling aid, see ct_gpu.h

int check_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail(CT_ERR_NO_DEVICE, "no CUDA device available (%s): this library has no CPU fallback",
                                                 e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (device < 0 || device >= n || device >= kMaxDevices) return fail(CT_ERR_NO_DEVICE, "device %d out of range (0..%d)", device, n - 1);
    CU(cudaSetDevice(device));
    return CT_OK;
}

void free_device(DeviceState &s) {
    for (void *q : s.ipc_opened) if (q) cudaIpcCloseMemHandle(q);
    for (void *p : s.allocs) cudaFree(p);
    s.allocs.clear();
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    for (cudaEvent_t e : s.tile_done) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s.stage_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s.ev_hit) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s.ev_done) if (e) cudaEventDestroy(e);
    for (cudaStream_t a : s.aux) if (a) cudaStreamDestroy(a);
    if (s.own_stream) cudaStreamDestroy(s.own_stream);
    s = DeviceState{};
}

template <typename T>
int dev_alloc(DeviceState &s, T **out, size_t count, bool zero = false) {
    void *p = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(CT_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    s.allocs.push_back(p);
    if (zero) CU(cudaMemset(p, 0, bytes));
    *out = static_cast<T *>(p);
    return CT_OK;
}

#define TRY(expr) do { int rc_ = (expr); if (rc_ != CT_OK) return rc_; } while (0)


// depth of the reference's DFS (stack entries needed) -- iterative to survive degenerate trees
// ---- scene build on the device (ct_gpu_upload_scene) -------------------------------------------------------------
// The reference's arrays go to the device as they are (nodes, triangles at their stride, the leaf permutation) and
// two kernels turn them into the layout above -- a gather through the permutation plus conversions is bandwidth work
// the host does an order of magnitude slower (868k triangles: 130 ms on 16 host threads, < 1 ms here).  Every value is
// produced by the same IEEE operations as before (fp64 subtractions, round-to-nearest / round-up conversions).
struct BuildReport {
    unsigned long long bound_bits[3];    // max |bound| over all nodes, per axis (bit pattern of a non-negative double)
    uint32_t boxes_bad;                  // some box is unordered or not finite
    uint32_t bad_pos;                    // a tri_indexes entry out of range (kNoPos: none)
    uint32_t pos0;                       // leaf position of triangle 0
    uint32_t any_reflective;
};

__global__ void __launch_bounds__(256) k_build_pairs(const ct_bvh_node *__restrict__ nodes, const uint32_t *__restrict__ pid_of, uint32_t n_nodes,
                                                     DevPair32 *__restrict__ pairs32, DevPair64 *__restrict__ pairs64,
                                                     uint32_t *__restrict__ pair_parent, uint32_t *__restrict__ tri_parent, uint32_t n_tri, BuildReport *rep) {
    double bound[3] = {0.0, 0.0, 0.0};
    bool bad = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        const ct_bvh_node n = nodes[i];
        for (int a = 0; a < 3; a++) {
            // the filter needs finite, ordered boxes (box_filter picks near/far by the ray's sign)
            if (!(n.aabb_min[a] <= n.aabb_max[a]) || isinf(n.aabb_min[a]) || isinf(n.aabb_max[a])) bad = true;
            else bound[a] = fmax(bound[a], fmax(fabs(n.aabb_min[a]), fabs(n.aabb_max[a])));
        }
        if (n.triangle_count != 0 || n.left_node == 0u || (uint64_t)n.left_node + 1u >= n_nodes) continue;      // leaf, or an unreachable slot holding zeros or garbage (the root is nobody's child)
        const uint32_t pid = pid_of[i];
        const ct_bvh_node L = nodes[n.left_node], R = nodes[n.left_node + 1u];
        DevPair32 p32;
        DevPair64 p64;
        for (int a = 0; a < 3; a++) {
            p64.lmin[a] = L.aabb_min[a]; p64.lmax[a] = L.aabb_max[a]; p64.rmin[a] = R.aabb_min[a]; p64.rmax[a] = R.aabb_max[a];
            p32.lmin[a] = __double2float_rn(L.aabb_min[a]); p32.lmax[a] = __double2float_rn(L.aabb_max[a]);
            p32.rmin[a] = __double2float_rn(R.aabb_min[a]); p32.rmax[a] = __double2float_rn(R.aabb_max[a]);
        }
        p32.l_cnt = L.triangle_count; p32.l_ref = L.triangle_count ? L.first_triangle_index : pid_of[n.left_node];
        p32.r_cnt = R.triangle_count; p32.r_ref = R.triangle_count ? R.first_triangle_index : pid_of[n.left_node + 1u];
        pairs32[pid] = p32;
        pairs64[pid] = p64;
        if (pair_parent) {
            // who holds whose box: lets k_overflow_huge check a triangle's ancestor chain without walking down
            for (uint32_t side = 0; side < 2u; side++) {
                const ct_bvh_node &ch = side ? R : L;
                const uint32_t code = 2u * pid + side;
                if (ch.triangle_count == 0) pair_parent[pid_of[n.left_node + side]] = code;
                else for (uint32_t k = 0; k < ch.triangle_count && (uint64_t)ch.first_triangle_index + k < n_tri; k++) tri_parent[ch.first_triangle_index + k] = code;
            }
        }
    }
    for (int a = 0; a < 3; a++) {
        for (int off = 16; off > 0; off >>= 1) bound[a] = fmax(bound[a], __shfl_xor_sync(0xffffffffu, bound[a], off));
        if ((threadIdx.x & 31u) == 0 && bound[a] > 0.0) atomicMax(&rep->bound_bits[a], (unsigned long long)__double_as_longlong(bound[a]));
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31u) == 0) atomicOr(&rep->boxes_bad, 1u);
}

__global__ void __launch_bounds__(256) k_build_tris(const unsigned char *__restrict__ raw, uint32_t stride, const uint32_t *__restrict__ tri_indexes,
                                                    const ct_material *__restrict__ materials, uint32_t n_tri,
                                                    DevTri *__restrict__ tris, DevTri32 *__restrict__ tris32, BuildReport *rep) {
    bool refl = false;
    auto dmax = [](double a, double b) { return (a < b) ? b : a; };      // std::max: a NaN component is skipped
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < n_tri; pos += gridDim.x * blockDim.x) {
        const uint32_t k = tri_indexes[pos];
        if (k >= n_tri) { atomicMin(&rep->bad_pos, pos); continue; }
        if (k == 0) rep->pos0 = pos;                     // a permutation holds it once (checked by the host afterwards)
        const double *v = reinterpret_cast<const double *>(raw + (size_t)k * stride);
        DevTri t;
        DevTri32 t32;
        double k1 = 0.0, k2 = 0.0, k3 = 0.0;
        for (int a = 0; a < 3; a++) {
            t.p1[a] = v[a];
            t.e1[a] = __dsub_rn(v[3 + a], v[a]);
            t.e2[a] = __dsub_rn(v[6 + a], v[a]);
            // fp32 copy + magnitudes for tri_filter_miss (rounded up; NaN k1 = "never certify")
            t32.p1[a] = __double2float_rn(t.p1[a]); t32.e1[a] = __double2float_rn(t.e1[a]); t32.e2[a] = __double2float_rn(t.e2[a]);
            k3 = dmax(k3, fabs(t.p1[a])); k1 = dmax(k1, fabs(t.e1[a])); k2 = dmax(k2, fabs(t.e2[a]));
        }
        t.orig = k; t.pad = 0;
        const bool in_range = k1 >= 0x1p-30 && k1 <= 0x1p30 && k2 >= 0x1p-30 && k2 <= 0x1p30 && k3 <= 0x1p40;   // NaNs fail
        t32.k1 = in_range ? __double2float_ru(k1) : __int_as_float(0x7fc00000);
        t32.k2 = __double2float_ru(k2); t32.k3 = __double2float_ru(k3);
        tris[pos] = t;
        tris32[pos] = t32;
        if (materials[k].reflection > 0.0f) refl = true;
    }
    if (__any_sync(0xffffffffu, refl) && (threadIdx.x & 31u) == 0) atomicOr(&rep->any_reflective, 1u);
}

int bvh_depth(const ct_bvh_node *nodes, uint32_t n_nodes, uint32_t n_tri, bool *ok) {
    std::vector<std::pair<uint32_t, int>> st;
    st.push_back({0u, 1});
    int best = 0;
    size_t visited = 0;
    *ok = true;
    while (!st.empty()) {
        auto [i, d] = st.back();
        st.pop_back();
        if (i >= n_nodes || ++visited > (size_t)n_nodes) { *ok = false; return 0; }
        best = std::max(best, d);
        const ct_bvh_node &nd = nodes[i];
        if (nd.triangle_count == 0) {
            st.push_back({nd.left_node, d + 1});
            st.push_back({nd.left_node + 1, d + 1});
        } else if ((uint64_t)nd.first_triangle_index + nd.triangle_count > n_tri) {
            *ok = false; return 0;
        }
    }
    return best;
}

int launch_grid(const DeviceState &s, int blocks_per_sm) { return s.n_sm * blocks_per_sm; }

int read_totals(DeviceState &s, ct_ray_counters *out) {   // synchronises the stream
    DevTotals h;
    CU(cudaMemcpyAsync(&h, s.p.tot, sizeof h, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    out->rays_primary = h.rays_primary; out->rays_shadow = h.rays_shadow; out->rays_reflection = h.rays_reflection;
    out->box_tests = h.box_tests; out->tri_tests = h.tri_tests;
    s.rays_overflow = h.rays_overflow; s.rays_in_place = h.rays_in_place;
    s.box_exact = h.box_exact; s.tri_exact = h.tri_exact;
    return CT_OK;
}

}  // namespace

// ==== exported C ABI ==========================================================================================
extern "C" {

int ct_gpu_abi_version(void) { return CT_GPU_ABI_VERSION; }

const char *ct_gpu_last_error(void) { return g_err; }

int ct_gpu_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(CT_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int ct_gpu_upload_scene(int device, const ct_scene_desc *d) {
    if (!d || d->struct_size != sizeof(ct_scene_desc)) return fail(CT_ERR_INVALID, "ct_scene_desc missing or struct_size != %zu", sizeof(ct_scene_desc));
    if (d->n_triangles == 0 || !d->triangles || !d->materials) return fail(CT_ERR_INVALID, "scene has no triangles");
    if (d->triangle_stride < 72 || d->triangle_stride % 8) return fail(CT_ERR_INVALID, "triangle_stride must be >= 72 and a multiple of 8");
    if (d->n_nodes == 0 || !d->nodes || !d->tri_indexes) return fail(CT_ERR_INVALID, "scene has no BVH (build it with the reference's BuildBVH or ct_host_build_bvh)");
    if (d->n_lights && !d->lights) return fail(CT_ERR_INVALID, "n_lights > 0 but lights == NULL");
    if (d->width <= 0 || d->height <= 0 || (int64_t)d->width * d->height > (1ll << 30)) return fail(CT_ERR_INVALID, "bad frame size %dx%d", d->width, d->height);
    if (d->max_depth < 0 || d->max_depth > 15) return fail(CT_ERR_LIMIT, "max_depth %d outside 0..15", d->max_depth);
    if ((d->flags & CT_FLAG_SUPERSAMPLING) && (d->flags & (CT_FLAG_SUBSAMPLING | CT_FLAG_KEEP_HITS)))
        return fail(CT_ERR_INVALID, "CT_FLAG_SUPERSAMPLING cannot be combined with CT_FLAG_SUBSAMPLING or CT_FLAG_KEEP_HITS");
    bool ok = true;
    int depth = bvh_depth(d->nodes, d->n_nodes, d->n_triangles, &ok);
    if (!ok) return fail(CT_ERR_INVALID, "BVH is malformed (child or triangle range out of bounds, or a cycle)");
    if (depth > kStackMax) return fail(CT_ERR_LIMIT, "BVH depth %d exceeds the device traversal stack (%d)", depth, kStackMax);
    TRY(check_device(device));
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState &s = g_dev[device];
    free_device(s);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    s.n_sm = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&s.own_stream, cudaStreamNonBlocking));
    s.stream = s.own_stream;
    CU(cudaEventCreate(&s.ev0));
    CU(cudaEventCreate(&s.ev1));
    for (cudaEvent_t &e : s.tile_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t &e : s.ev_hit) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaEvent_t &e : s.ev_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaStream_t &a : s.aux) CU(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    if (d->flags & CT_FLAG_STAGE_TIMING) for (cudaEvent_t &e : s.stage_ev) CU(cudaEventCreate(&e));
    s.p.budget = g_budget_option > 0 ? (uint32_t)std::min<long long>(g_budget_option, 1ll << 30) : kDefaultBudget;
    s.p.warp_budget = g_warp_budget_option > 0 ? (uint32_t)std::min<long long>(g_warp_budget_option, 1ll << 30) : kWarpBudget;
    s.flags = d->flags;

    const bool timing = getenv("CT_GPU_TIMING") != nullptr;
    auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now_ms();
    double t_mark = t_begin;
    auto lap = [&](const char *what) { if (timing) { double t = now_ms(); fprintf(stderr, "ct_gpu_upload_scene: %-28s %7.1f ms\n", what, t - t_mark); t_mark = t; } };
    Params &p = s.p;
    p.n_tri = d->n_triangles; p.n_lights = d->n_lights;
    // nodes: reference layout -> per-interior-node child pairs (fp32 for the filter, fp64 for the exact path), built on
    // the device from the raw arrays (k_build_pairs / k_build_tris); the host only numbers the interior nodes
    std::vector<uint32_t> pid_of(d->n_nodes, kNoPos);
    uint32_t n_pairs = 0;
    for (uint32_t i = 0; i < d->n_nodes; i++)
        if (d->nodes[i].triangle_count == 0) pid_of[i] = n_pairs++;
    auto child_ref = [&](uint32_t c, uint32_t &ref, uint32_t &cnt) {
        const ct_bvh_node &n = d->nodes[c];
        cnt = n.triangle_count;
        ref = cnt ? n.first_triangle_index : pid_of[c];
    };
    child_ref(0, p.root_ref, p.root_cnt);
    p.n_pairs = n_pairs;
    p.n_nodes = d->n_nodes;
    s.can_overflow = (uint64_t)d->n_nodes + d->n_triangles > p.budget;      // parked rays: only when a DFS can run past the budget at all
    lap("interior-node numbering (host)");

    DevPair32 *dp32 = nullptr; DevPair64 *dp64 = nullptr; DevTri *dt = nullptr; DevTri32 *dt32 = nullptr; ct_material *dm = nullptr;
    uint32_t *dpp = nullptr, *dtp = nullptr;
    TRY(dev_alloc(s, &dp32, std::max<uint32_t>(n_pairs, 1))); TRY(dev_alloc(s, &dp64, std::max<uint32_t>(n_pairs, 1)));
    TRY(dev_alloc(s, &dt, d->n_triangles)); TRY(dev_alloc(s, &dt32, d->n_triangles)); TRY(dev_alloc(s, &dm, d->n_triangles));
    if (s.can_overflow) { TRY(dev_alloc(s, &dpp, std::max<uint32_t>(n_pairs, 1))); TRY(dev_alloc(s, &dtp, d->n_triangles)); }
    BuildReport rep{};
    {
        // staging copies of the caller's arrays, freed again below
        ct_bvh_node *raw_nodes = nullptr; uint32_t *raw_pid = nullptr, *raw_idx = nullptr; unsigned char *raw_tris = nullptr; BuildReport *drep = nullptr;
        const size_t tri_bytes = (size_t)(d->n_triangles - 1) * d->triangle_stride + 72;      // the last triangle may end at its third vertex
        auto release = [&] { cudaFree(raw_nodes); cudaFree(raw_pid); cudaFree(raw_idx); cudaFree(raw_tris); cudaFree(drep); };
        auto staged = [&](cudaError_t e) { if (e != cudaSuccess) { release(); free_device(s); } return e; };
        CU(staged(cudaMalloc(&raw_nodes, (size_t)d->n_nodes * sizeof(ct_bvh_node))));
        CU(staged(cudaMalloc(&raw_pid, (size_t)d->n_nodes * 4)));
        CU(staged(cudaMalloc(&raw_idx, (size_t)d->n_triangles * 4)));
        CU(staged(cudaMalloc(&raw_tris, tri_bytes)));
        CU(staged(cudaMalloc(&drep, sizeof rep)));
        rep.bad_pos = kNoPos; rep.pos0 = kNoPos;
        CU(staged(cudaMemcpyAsync(drep, &rep, sizeof rep, cudaMemcpyHostToDevice, s.stream)));
        CU(staged(cudaMemcpyAsync(raw_nodes, d->nodes, (size_t)d->n_nodes * sizeof(ct_bvh_node), cudaMemcpyHostToDevice, s.stream)));
        CU(staged(cudaMemcpyAsync(raw_pid, pid_of.data(), (size_t)d->n_nodes * 4, cudaMemcpyHostToDevice, s.stream)));
        if (dpp) {
            CU(staged(cudaMemsetAsync(dpp, 0xff, (size_t)std::max<uint32_t>(n_pairs, 1) * 4, s.stream)));
            CU(staged(cudaMemsetAsync(dtp, 0xff, (size_t)d->n_triangles * 4, s.stream)));
        }
        const int build_blocks = s.n_sm * 8;
        k_build_pairs<<<build_blocks, 256, 0, s.stream>>>(raw_nodes, raw_pid, d->n_nodes, dp32, dp64, dpp, dtp, d->n_triangles, drep);
        CU(staged(cudaMemcpyAsync(raw_idx, d->tri_indexes, (size_t)d->n_triangles * 4, cudaMemcpyHostToDevice, s.stream)));
        CU(staged(cudaMemcpyAsync(raw_tris, d->triangles, tri_bytes, cudaMemcpyHostToDevice, s.stream)));
        CU(staged(cudaMemcpyAsync(dm, d->materials, (size_t)d->n_triangles * sizeof(ct_material), cudaMemcpyHostToDevice, s.stream)));
        k_build_tris<<<build_blocks, 256, 0, s.stream>>>(raw_tris, (uint32_t)d->triangle_stride, raw_idx, dm, d->n_triangles, dt, dt32, drep);
        CU(staged(cudaGetLastError()));
        CU(staged(cudaMemcpyAsync(&rep, drep, sizeof rep, cudaMemcpyDeviceToHost, s.stream)));
        CU(staged(cudaStreamSynchronize(s.stream)));
        release();
    }
    if (rep.bad_pos != kNoPos) { free_device(s); return fail(CT_ERR_INVALID, "tri_indexes[%u] = %u out of range", rep.bad_pos, d->tri_indexes[rep.bad_pos]); }
    if (rep.pos0 == kNoPos) { free_device(s); return fail(CT_ERR_INVALID, "tri_indexes is not a permutation (triangle 0 missing)"); }
    p.pos_of_tri0 = rep.pos0;
    s.any_reflective = rep.any_reflective != 0;
    for (int a = 0; a < 3; a++) {
        double bound;
        memcpy(&bound, &rep.bound_bits[a], sizeof bound);
        p.bound[a] = (!rep.boxes_bad && bound < 1e30) ? bound : INFINITY;
        p.root_min[a] = d->nodes[0].aabb_min[a]; p.root_max[a] = d->nodes[0].aabb_max[a];
        p.root_min32[a] = (float)p.root_min[a]; p.root_max32[a] = (float)p.root_max[a];
    }
    p.pairs32 = dp32; p.pairs64 = dp64; p.tris = dt; p.tris32 = dt32; p.materials = dm;
    p.pair_parent = dpp; p.tri_parent = dtp;
    lap("scene arrays built on the device");
    std::vector<DevLight> lights(std::max<uint32_t>(d->n_lights, 1));
    for (uint32_t i = 0; i < d->n_lights; i++) {
        lights[i].type = d->lights[i].type; lights[i].intensity = d->lights[i].intensity;
        for (int a = 0; a < 3; a++) { lights[i].pos[a] = d->lights[i].position[a]; lights[i].dir[a] = d->lights[i].direction[a]; }
    }
    std::vector<DevShadowLight> slights;
    for (uint32_t i = 0; i < d->n_lights; i++) {
        if (d->lights[i].type == CT_LIGHT_AMBIENT) continue;            // raythread.cpp:284-286: no shadow ray
        DevShadowLight sl{};
        sl.type = d->lights[i].type; sl.index = i;
        const double *v = (sl.type == CT_LIGHT_POINT) ? d->lights[i].position : d->lights[i].direction;   // :288 / :293 (every non-point light is directional)
        for (int a = 0; a < 3; a++) sl.v[a] = v[a];
        slights.push_back(sl);
    }
    p.n_slights = (uint32_t)slights.size();
    p.occ_words = std::max<uint32_t>((d->n_lights + 31u) / 32u, 1u);
    DevLight *dl = nullptr; DevShadowLight *dsl = nullptr;
    TRY(dev_alloc(s, &dl, lights.size()));
    TRY(dev_alloc(s, &dsl, slights.size()));
    if (!slights.empty()) CU(cudaMemcpy(dsl, slights.data(), slights.size() * sizeof(DevShadowLight), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dl, lights.data(), lights.size() * sizeof(DevLight), cudaMemcpyHostToDevice));
    p.slights = dsl; p.lights = dl;

    memcpy(p.cam, d->camera_position, sizeof p.cam);
    memcpy(p.rot, d->camera_rotation, sizeof p.rot);
    p.vp_w = d->viewport[0]; p.vp_h = d->viewport[1]; p.vp_d = d->viewport[2];
    p.W = d->width; p.H = d->height; p.max_depth = d->max_depth; p.background = d->background;

    // x range: the reference's centred square (raythread.cpp:454-455) or the whole width
    int half = d->height / 2;
    int x_lo = -half, n_x = 2 * half;
    if (d->flags & CT_FLAG_WIDE) { x_lo = -(d->width / 2); n_x = d->width; }
    p.x_lo = x_lo; p.n_x = n_x;
    p.blocks_x = (n_x + 7) / 8;
    s.col_lo = std::max(0, x_lo + d->width / 2);
    s.col_hi = std::min(d->width, x_lo + n_x + d->width / 2);
    int rows_max = 2 * half + 1;
    {
        const uint64_t pixel_slots = (uint64_t)p.blocks_x * (uint64_t)((rows_max + 3) / 4) * 32u;
        const uint64_t slots = pixel_slots * ((d->flags & CT_FLAG_SUPERSAMPLING) ? 16u : 1u);
        if (slots > (1ull << 31)) { free_device(s); return fail(CT_ERR_LIMIT, "frame needs %llu path slots (limit 2^31)", (unsigned long long)slots); }
        p.cap = (uint32_t)((slots + kChunkMax - 1u) / kChunkMax * kChunkMax);
    }

    const int levels = s.any_reflective ? d->max_depth + 1 : 1;
    TRY(dev_alloc(s, &p.hit0_t, p.cap)); TRY(dev_alloc(s, &p.hit0_pos, p.cap));
    TRY(dev_alloc(s, &p.stack_color, (size_t)p.cap * levels));
    TRY(dev_alloc(s, &p.term_level, p.cap, true));
    s.levels = levels;
    TRY(dev_alloc(s, &s.occ_all, (size_t)p.cap * p.occ_words * levels, true));
    p.occ = s.occ_all;
    if (s.can_overflow) {
        p.ovf_cap = 1u << 18;          // parked rays per launch (16 MB); a full buffer means finishing rays in place, which must stay hypothetical
        TRY(dev_alloc(s, &s.ovf_all, (size_t)p.ovf_cap * 2 * levels));
        TRY(dev_alloc(s, &s.ovf_huge_all, (size_t)p.ovf_cap * 2 * levels));
        p.ovf = s.ovf_all; p.ovf_huge = s.ovf_huge_all;
    }
    if (levels > 1) {
        TRY(dev_alloc(s, &s.hitb_t_all, (size_t)p.cap * levels)); TRY(dev_alloc(s, &s.hitb_pos_all, (size_t)p.cap * levels));
        TRY(dev_alloc(s, &p.stack_refl, (size_t)p.cap * levels));
        TRY(dev_alloc(s, &s.rays_all, (size_t)p.cap * 6 * (levels + 1))); TRY(dev_alloc(s, &s.path_slot_all, (size_t)p.cap * (levels + 1)));
    }
    TRY(dev_alloc(s, &p.fb, (size_t)d->width * d->height, true));       // calloc'd like cobbletrace.cpp:57
    p.subsample = (d->flags & CT_FLAG_SUBSAMPLING) ? 1 : 0;
    p.supersample = (d->flags & CT_FLAG_SUPERSAMPLING) ? 1 : 0;
    if (p.subsample || p.supersample) TRY(dev_alloc(s, &p.final_color, p.cap, true));
    TRY(dev_alloc(s, &p.own_chunks, (p.cap >> kChunkLocalShift) + 1u));
    TRY(dev_alloc(s, &s.cursor_own, 1, true));                           // its own allocation: exported over CUDA IPC
    s.share_cursor = s.cursor_own; s.share_fb = p.fb;
    if (d->flags & CT_FLAG_KEEP_HITS) {
        size_t npx = (size_t)d->width * d->height;
        TRY(dev_alloc(s, &p.dbg_found, npx)); TRY(dev_alloc(s, &p.dbg_index, npx, true)); TRY(dev_alloc(s, &p.dbg_t, npx, true));
        CU(cudaMemset(p.dbg_found, 0xff, npx * sizeof(uint32_t)));
    }
    TRY(dev_alloc(s, &p.sched, 1, true));
    TRY(dev_alloc(s, &p.tot, 1, true));
    lap("path state allocation");
    if (timing) fprintf(stderr, "ct_gpu_upload_scene: total %.1f ms\n", now_ms() - t_begin);
    s.loaded = true;
    return CT_OK;
}

int ct_gpu_set_camera(int device, const double position[3], const double rotation[9]) {
    if (!position || !rotation) return fail(CT_ERR_INVALID, "NULL camera");
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    memcpy(s.p.cam, position, sizeof s.p.cam);
    memcpy(s.p.rot, rotation, sizeof s.p.rot);
    return CT_OK;
}

int ct_gpu_set_stream(int device, void *cuda_stream) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    CU(cudaStreamSynchronize(s.stream));
    s.stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s.own_stream;
    return CT_OK;
}

static int render_impl(int device, int y_start, int y_end, ct_ray_counters *counters, bool shared) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    int half = s.p.H / 2;
    // rows outside [-half+?, half] can never be stored (draw2d.h:11): clip so no ray is wasted on them
    int y0 = std::max(y_start, half - (s.p.H - 1)), y1 = std::min(y_end, half + 1);
    if (s.p.subsample) {
        // the partition's own bounds define which rows are traced and averaged (raythread.cpp:457,513,527): no clipping
        if (shared) return fail(CT_ERR_INVALID, "CT_FLAG_SUBSAMPLING needs the whole partition on one device (neighbouring rows are averaged)");
        y0 = y_start; y1 = y_end;
        if (y0 < -(1 << 20) || y1 > (1 << 20)) return fail(CT_ERR_INVALID, "tile [%d,%d) out of range", y_start, y_end);
    }
    if (y1 <= y0) { if (counters) memset(counters, 0, sizeof *counters); return CT_OK; }
    Params p = s.p;
    p.y_lo = y0; p.n_rows = y1 - y0;
    p.n_y = s.p.subsample ? (p.n_rows + 1) / 2 + ((p.n_rows & 1) == 0 ? 1 : 0) : p.n_rows;
    p.n_slots = (uint32_t)p.blocks_x * (uint32_t)((p.n_y + 3) / 4) * 32u * (s.p.supersample ? 16u : 1u);
    if (p.n_slots > p.cap) return fail(CT_ERR_INVALID, "tile [%d,%d) larger than the frame", y_start, y_end);
    const bool count = (s.flags & CT_FLAG_COUNT_TESTS) != 0;
    const int depth_max = s.any_reflective ? p.max_depth : 0;
    // one device: its own cursor (zeroed with the rest of DevSched below) and framebuffer; shared frame: the root's
    p.steal = shared ? s.share_cursor : &p.sched->steal_local;
    const uint32_t shared_shift = g_shared_chunk_shift ? (uint32_t)g_shared_chunk_shift : kChunkSharedShift;
    p.chunk_shift = shared ? shared_shift : kChunkLocalShift;
    p.steal_stride = g_emulate_ranks > 1 ? (uint32_t)g_emulate_ranks : 1u;
    if (p.steal_stride > 1) p.chunk_shift = shared_shift;
    const bool dealt = shared && p.steal_stride == 1 && s.part_count > 1;
    p.part_index = dealt ? (uint32_t)s.part_index : 0u;
    p.part_count = dealt ? (uint32_t)s.part_count : 0u;
    p.static_eighths = (uint32_t)g_static_eighths;
    p.fb_out = shared ? s.share_fb : p.fb;
    Params pk = p; pk.max_depth = depth_max;   // with no reflective material the recursion never goes past depth 0 (:369)
    cudaStream_t st = s.stream;
    CU(cudaMemsetAsync(p.sched, 0, sizeof(DevSched), st));
    CU(cudaEventRecord(s.ev0, st));
    const int grid = launch_grid(s, 8);
    int work = 0;
    const bool stages = (s.flags & CT_FLAG_STAGE_TIMING) != 0;
    s.n_stages = 0;
    if (stages) CU(cudaEventRecord(s.stage_ev[0], st));
    auto mark = [&](const char *name, int depth) -> int {          // called after each launch
        s.launches++;
        if (!stages || s.n_stages >= kMaxLaunches) return CT_OK;
        s.stage_name[s.n_stages] = name; s.stage_depth[s.n_stages] = depth;
        s.n_stages++;
        CU(cudaEventRecord(s.stage_ev[s.n_stages], st));
        return CT_OK;
    };
    // Params as the launches of depth d see them: their own hit records, occlusion masks, ray queue in / out
    auto view = [&](int d) {
        Params v = pk;
        const size_t cap = pk.cap;
        if (s.hitb_t_all) { v.hitb_t = s.hitb_t_all + cap * d; v.hitb_pos = s.hitb_pos_all + cap * d; }
        if (s.rays_all) {
            v.ray_buf[d & 1] = s.rays_all + cap * 6 * d;            v.path_slot[d & 1] = s.path_slot_all + cap * d;
            v.ray_buf[(d & 1) ^ 1] = s.rays_all + cap * 6 * (d + 1); v.path_slot[(d & 1) ^ 1] = s.path_slot_all + cap * (d + 1);
        }
        v.occ = s.occ_all + cap * pk.occ_words * d;
        return v;
    };
    auto overflow = [&](const Params &v0, cudaStream_t q, int mode_anyhit, int ovf_idx, int depth) -> int {   // parked rays of the launch just made
        if (!s.can_overflow) return CT_OK;
        Params v = v0;
        v.ovf = s.ovf_all + (size_t)pk.ovf_cap * ovf_idx; v.ovf_huge = s.ovf_huge_all + (size_t)pk.ovf_cap * ovf_idx;
        const int g1 = s.n_sm * 2, g2 = s.n_sm * 4;
        if (mode_anyhit) {
            if (count) { k_overflow<kAnyHit, true><<<g1, kOvfThreads, 0, q>>>(v, ovf_idx); k_overflow_huge<kAnyHit, true><<<g2, 256, 0, q>>>(v, ovf_idx); }
            else { k_overflow<kAnyHit, false><<<g1, kOvfThreads, 0, q>>>(v, ovf_idx); k_overflow_huge<kAnyHit, false><<<g2, 256, 0, q>>>(v, ovf_idx); }
        } else {
            if (count) { k_overflow<kFirstLine, true><<<g1, kOvfThreads, 0, q>>>(v, ovf_idx); k_overflow_huge<kFirstLine, true><<<g2, 256, 0, q>>>(v, ovf_idx); }
            else { k_overflow<kFirstLine, false><<<g1, kOvfThreads, 0, q>>>(v, ovf_idx); k_overflow_huge<kFirstLine, false><<<g2, 256, 0, q>>>(v, ovf_idx); }
        }
        s.launches++;                                       // two kernels, one stage
        return mark(mode_anyhit ? "overflow_shadow" : "overflow_bounce", depth);
    };
    // Two dependency chains per frame:
    //   hits      primary -> emit(0) -> bounce(1) -> emit(1) -> bounce(2) -> ...          (main stream)
    //   lighting  shadow(d) -> shade(d), needs only the hits of depth d                   (one side stream per depth)
    // so the lighting of depth d overlaps the hit chain of depth d+1 and the lighting of the other depths; the
    // persistent CTAs of a later kernel move into an SM as those of an earlier one run out of work, which fills
    // the kernels' tails.  With CT_FLAG_STAGE_TIMING everything is serialised on the main stream instead.
    if (count) k_primary<true><<<grid, kBlockThreads, 0, st>>>(pk);
    else k_primary<false><<<grid, kBlockThreads, 0, st>>>(pk);
    TRY(mark("primary", 0));
    for (int d = 0; d <= depth_max; d++) {
        const Params v = view(d);
        const int ovf_b = 2 * d, ovf_s = 2 * d + 1;
        if (d > 0) {
            Params vb = v;
            if (s.can_overflow) { vb.ovf = s.ovf_all + (size_t)pk.ovf_cap * ovf_b; vb.ovf_huge = s.ovf_huge_all + (size_t)pk.ovf_cap * ovf_b; }
            if (count) k_bounce<true><<<grid, kBlockThreads, 0, st>>>(vb, d, work++, ovf_b);
            else k_bounce<false><<<grid, kBlockThreads, 0, st>>>(vb, d, work++, ovf_b);
            TRY(mark("bounce", d));
            TRY(overflow(v, st, 0, ovf_b, d));
        }
        cudaStream_t side = stages ? st : s.aux[d % 4];
        if (!stages) { CU(cudaEventRecord(s.ev_hit[d], st)); CU(cudaStreamWaitEvent(side, s.ev_hit[d], 0)); }
        if (depth_max > 0) { k_emit<<<grid, kBlockThreads, 0, st>>>(v, d, work++); TRY(mark("emit", d)); }
        if (pk.n_slights > 0) {
            Params vs = v;
            if (s.can_overflow) { vs.ovf = s.ovf_all + (size_t)pk.ovf_cap * ovf_s; vs.ovf_huge = s.ovf_huge_all + (size_t)pk.ovf_cap * ovf_s; }
            if (count) k_shadow<true><<<grid, kBlockThreads, 0, side>>>(vs, d, work++, ovf_s);
            else k_shadow<false><<<grid, kBlockThreads, 0, side>>>(vs, d, work++, ovf_s);
            if (stages) TRY(mark("shadow", d)); else s.launches++;
            if (stages) { TRY(overflow(v, side, 1, ovf_s, d)); }
            else if (s.can_overflow) {
                Params vo = v;
                vo.ovf = s.ovf_all + (size_t)pk.ovf_cap * ovf_s; vo.ovf_huge = s.ovf_huge_all + (size_t)pk.ovf_cap * ovf_s;
                if (count) { k_overflow<kAnyHit, true><<<s.n_sm * 2, kOvfThreads, 0, side>>>(vo, ovf_s); k_overflow_huge<kAnyHit, true><<<s.n_sm * 4, 256, 0, side>>>(vo, ovf_s); }
                else { k_overflow<kAnyHit, false><<<s.n_sm * 2, kOvfThreads, 0, side>>>(vo, ovf_s); k_overflow_huge<kAnyHit, false><<<s.n_sm * 4, 256, 0, side>>>(vo, ovf_s); }
                s.launches += 2;
            }
        }
        k_shade<<<grid, kBlockThreads, 0, side>>>(v, d, work++);
        if (stages) TRY(mark("shade", d)); else s.launches++;
        if (!stages) CU(cudaEventRecord(s.ev_done[d], side));
    }
    if (!stages) for (int d = 0; d <= depth_max; d++) CU(cudaStreamWaitEvent(st, s.ev_done[d], 0));
    if (depth_max > 0) { k_resolve<<<s.n_sm * 4, 256, 0, st>>>(pk); TRY(mark("resolve", 0)); }
    if (pk.subsample) { k_subsample<<<s.n_sm * 4, 256, 0, st>>>(pk); TRY(mark("subsample", 0)); }
    if (pk.supersample) { k_supersample<<<s.n_sm * 4, 256, 0, st>>>(pk); TRY(mark("supersample", 0)); }
    CU(cudaEventRecord(s.ev1, st));
    CU(cudaEventRecord(s.tile_done[s.tiles_submitted % 8], st));
    s.tiles_submitted++;
    CU(cudaGetLastError());
    s.timed = true;
    {   // framebuffer rows this tile writes: row = H/2 - y
        int lo = half - (y1 - 1), hi = half - y0 + 1 + (s.p.subsample ? 1 : 0);   // subsampling also stores row y0 - 1
        lo = std::max(lo, 0); hi = std::min(hi, s.p.H);
        if (s.row_hi <= s.row_lo) { s.row_lo = lo; s.row_hi = hi; }
        else { s.row_lo = std::min(s.row_lo, lo); s.row_hi = std::max(s.row_hi, hi); }
    }
    if (counters) {
        ct_ray_counters now;
        TRY(read_totals(s, &now));
        counters->rays_primary = now.rays_primary - s.snapshot.rays_primary;
        counters->rays_shadow = now.rays_shadow - s.snapshot.rays_shadow;
        counters->rays_reflection = now.rays_reflection - s.snapshot.rays_reflection;
        counters->box_tests = now.box_tests - s.snapshot.box_tests;
        counters->tri_tests = now.tri_tests - s.snapshot.tri_tests;
        s.snapshot = now;
    }
    return CT_OK;
}

int ct_gpu_render_tile(int device, int y_start, int y_end, ct_ray_counters *counters) {
    return render_impl(device, y_start, y_end, counters, false);
}

int ct_gpu_render_shared(int device, int y_start, int y_end, ct_ray_counters *counters) {
    return render_impl(device, y_start, y_end, counters, true);
}

int ct_gpu_share_export(int device, ct_gpu_share *out) {
    if (!out || out->struct_size != sizeof(ct_gpu_share)) return fail(CT_ERR_INVALID, "ct_gpu_share missing or struct_size != %zu", sizeof(ct_gpu_share));
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ct_gpu_share reserves 64 bytes per IPC handle");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s.p.fb));
    memcpy(out->fb_ipc, &h, sizeof h);
    CU(cudaIpcGetMemHandle(&h, s.cursor_own));
    memcpy(out->cursor_ipc, &h, sizeof h);
    out->device = device;
    out->pid = (int64_t)getpid();
    out->fb_ptr = (uint64_t)(uintptr_t)s.p.fb;
    out->cursor_ptr = (uint64_t)(uintptr_t)s.cursor_own;
    out->width = s.p.W; out->height = s.p.H;
    return CT_OK;
}

int ct_gpu_share_attach(int device, const ct_gpu_share *root) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    for (void *&q : s.ipc_opened) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    if (!root) {                                             // detach: back to this device's own cursor and framebuffer
        s.share_cursor = s.cursor_own; s.share_fb = s.p.fb;
        return CT_OK;
    }
    if (root->struct_size != sizeof(ct_gpu_share)) return fail(CT_ERR_INVALID, "ct_gpu_share struct_size != %zu", sizeof(ct_gpu_share));
    if (root->width != s.p.W || root->height != s.p.H) return fail(CT_ERR_INVALID, "shared frame is %dx%d, this device renders %dx%d", root->width, root->height, s.p.W, s.p.H);
    if (root->pid == (int64_t)getpid()) {                    // same process: raw pointers (+ peer access between the two devices)
        if (root->device != device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, device, root->device));
            if (!can) return fail(CT_ERR_CUDA, "device %d cannot access device %d's memory", device, root->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(root->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(CT_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        s.share_fb = (uint32_t *)(uintptr_t)root->fb_ptr;
        s.share_cursor = (unsigned long long *)(uintptr_t)root->cursor_ptr;
        return CT_OK;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, root->fb_ipc, sizeof h);
    CU(cudaIpcOpenMemHandle(&s.ipc_opened[0], h, cudaIpcMemLazyEnablePeerAccess));
    memcpy(&h, root->cursor_ipc, sizeof h);
    CU(cudaIpcOpenMemHandle(&s.ipc_opened[1], h, cudaIpcMemLazyEnablePeerAccess));
    s.share_fb = (uint32_t *)s.ipc_opened[0];
    s.share_cursor = (unsigned long long *)s.ipc_opened[1];
    return CT_OK;
}

int ct_gpu_share_partition(int device, int index, int count) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    if (count < 0 || count > 64 || (count > 0 && (index < 0 || index >= count))) return fail(CT_ERR_INVALID, "bad partition %d of %d", index, count);
    s.part_index = count > 1 ? index : 0;
    s.part_count = count > 1 ? count : 0;
    return CT_OK;
}

int ct_gpu_share_reset(int device) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    CU(cudaMemsetAsync(s.cursor_own, 0, sizeof(unsigned long long), s.stream));
    CU(cudaStreamSynchronize(s.stream));
    return CT_OK;
}

int ct_gpu_set_option(const char *name, long long value) {
    if (!name) return fail(CT_ERR_INVALID, "NULL option name");
    if (!strcmp(name, "traversal_budget")) {
        if (value < 0) return fail(CT_ERR_INVALID, "traversal_budget must be >= 0 (0 = default)");
        g_budget_option = value;
        return CT_OK;
    }
    if (!strcmp(name, "emulate_ranks")) {
        if (value < 0 || value > 64) return fail(CT_ERR_INVALID, "emulate_ranks must be 0..64");
        g_emulate_ranks = value;
        return CT_OK;
    }
    if (!strcmp(name, "shared_static_eighths")) {
        if (value < 0 || value > 8) return fail(CT_ERR_INVALID, "shared_static_eighths must be 0..8");
        g_static_eighths = value;
        return CT_OK;
    }
    if (!strcmp(name, "shared_chunk_shift")) {
        if (value != 0 && (value < kChunkLocalShift || value > kChunkMaxShift)) return fail(CT_ERR_INVALID, "shared_chunk_shift must be 0 (default), 5 or 6");
        g_shared_chunk_shift = value;
        return CT_OK;
    }
    if (!strcmp(name, "overflow_warp_budget")) {
        if (value < 0) return fail(CT_ERR_INVALID, "overflow_warp_budget must be >= 0 (0 = default)");
        g_warp_budget_option = value;
        return CT_OK;
    }
    return fail(CT_ERR_INVALID, "unknown option '%s'", name);
}

int ct_gpu_overflow_stats(int device, uint64_t *parked, uint64_t *finished_in_place) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    ct_ray_counters tmp;
    TRY(read_totals(s, &tmp));
    if (parked) *parked = s.rays_overflow;
    if (finished_in_place) *finished_in_place = s.rays_in_place;
    return CT_OK;
}

int ct_gpu_filter_stats(int device, uint64_t *box_exact, uint64_t *tri_exact) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    ct_ray_counters tmp;
    TRY(read_totals(s, &tmp));
    if (box_exact) *box_exact = s.box_exact;
    if (tri_exact) *tri_exact = s.tri_exact;
    return CT_OK;
}

int ct_gpu_sync(int device) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    CU(cudaStreamSynchronize(s.stream));
    return CT_OK;
}

int ct_gpu_kernel_launches(int device, uint64_t *out, int reset) {
    if (!out) return fail(CT_ERR_INVALID, "NULL out");
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    *out = s.launches;
    if (reset) s.launches = 0;
    return CT_OK;
}

int ct_gpu_last_tile_stages(int device, int max, float *ms, const char **names, int *depth) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    if (!(s.flags & CT_FLAG_STAGE_TIMING)) return fail(CT_ERR_INVALID, "scene was uploaded without CT_FLAG_STAGE_TIMING");
    CU(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < s.n_stages && i < max; i++) {
        if (ms) CU(cudaEventElapsedTime(&ms[i], s.stage_ev[i], s.stage_ev[i + 1]));
        if (names) names[i] = s.stage_name[i];
        if (depth) depth[i] = s.stage_depth[i];
    }
    return s.n_stages;
}

int ct_gpu_throttle(int device, int max_in_flight) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    if (max_in_flight <= 0) { CU(cudaStreamSynchronize(s.stream)); return CT_OK; }
    unsigned long long keep = (unsigned long long)std::min(max_in_flight, 7);
    if (s.tiles_submitted > keep) CU(cudaEventSynchronize(s.tile_done[(s.tiles_submitted - keep - 1) % 8]));
    return CT_OK;
}

int ct_gpu_last_tile_ms(int device, float *ms) {
    if (!ms) return fail(CT_ERR_INVALID, "NULL ms");
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded || !s.timed) return fail(CT_ERR_NO_SCENE, "no tile rendered on device %d", device);
    CU(cudaEventSynchronize(s.ev1));
    CU(cudaEventElapsedTime(ms, s.ev0, s.ev1));
    return CT_OK;
}

int ct_gpu_get_counters(int device, ct_ray_counters *out, int reset) {
    if (!out) return fail(CT_ERR_INVALID, "NULL out");
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    TRY(read_totals(s, out));
    if (reset) {
        CU(cudaMemset(s.p.tot, 0, sizeof(DevTotals)));
        s.snapshot = ct_ray_counters{};
    }
    return CT_OK;
}

int ct_gpu_readback(int device, uint32_t *dst, int dst_stride_pixels, int row_start, int row_end) {
    if (!dst) return fail(CT_ERR_INVALID, "NULL dst");
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    const Params &p = s.p;
    // Only rows some tile has rendered are copied: e.g. row 0 (y = H/2) is never reached by the reference's
    // loops (y < yEnd <= H/2) and so stays untouched in dst here as well.
    int r0 = std::max(row_start, s.row_lo), r1 = std::min(row_end, s.row_hi);
    int cols = s.col_hi - s.col_lo;
    if (dst_stride_pixels < p.W) return fail(CT_ERR_INVALID, "dst stride %d < width %d", dst_stride_pixels, p.W);
    CU(cudaStreamSynchronize(s.stream));
    if (r1 <= r0 || cols <= 0) return CT_OK;
    CU(cudaMemcpy2D(dst + (size_t)r0 * dst_stride_pixels + s.col_lo, (size_t)dst_stride_pixels * 4,
                    p.fb + (size_t)r0 * p.W + s.col_lo, (size_t)p.W * 4, (size_t)cols * 4, (size_t)(r1 - r0),
                    cudaMemcpyDeviceToHost));
    return CT_OK;
}

int ct_gpu_readback_hits(int device, uint32_t *found, uint32_t *index, float *t, int stride_pixels, int row_start, int row_end) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    const Params &p = s.p;
    if (!p.dbg_found) return fail(CT_ERR_INVALID, "scene was uploaded without CT_FLAG_KEEP_HITS");
    if (stride_pixels < p.W) return fail(CT_ERR_INVALID, "stride %d < width %d", stride_pixels, p.W);
    int r0 = std::max(row_start, 0), r1 = std::min(row_end, p.H);
    CU(cudaStreamSynchronize(s.stream));
    if (r1 <= r0) return CT_OK;
    size_t rows = (size_t)(r1 - r0);
    if (found) CU(cudaMemcpy2D(found + (size_t)r0 * stride_pixels, (size_t)stride_pixels * 4, p.dbg_found + (size_t)r0 * p.W, (size_t)p.W * 4, (size_t)p.W * 4, rows, cudaMemcpyDeviceToHost));
    if (index) CU(cudaMemcpy2D(index + (size_t)r0 * stride_pixels, (size_t)stride_pixels * 4, p.dbg_index + (size_t)r0 * p.W, (size_t)p.W * 4, (size_t)p.W * 4, rows, cudaMemcpyDeviceToHost));
    if (t) CU(cudaMemcpy2D(t + (size_t)r0 * stride_pixels, (size_t)stride_pixels * 4, p.dbg_t + (size_t)r0 * p.W, (size_t)p.W * 4, (size_t)p.W * 4, rows, cudaMemcpyDeviceToHost));
    return CT_OK;
}

int ct_gpu_framebuffer(int device, void **device_ptr, int *width, int *height) {
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    if (device_ptr) *device_ptr = s.p.fb;
    if (width) *width = s.p.W;
    if (height) *height = s.p.H;
    return CT_OK;
}

int ct_gpu_gather_rows(int src_device, int dst_device, int row_start, int row_end) {
    TRY(check_device(dst_device));
    TRY(check_device(src_device));
    DeviceState &a = g_dev[src_device], &b = g_dev[dst_device];
    if (!a.loaded || !b.loaded) return fail(CT_ERR_NO_SCENE, "both devices need an uploaded scene");
    if (a.p.W != b.p.W || a.p.H != b.p.H) return fail(CT_ERR_INVALID, "frame sizes differ between devices");
    int r0 = std::max(row_start, 0), r1 = std::min(row_end, a.p.H);
    if (r1 <= r0 || src_device == dst_device) return CT_OK;
    size_t off = (size_t)r0 * a.p.W, bytes = (size_t)(r1 - r0) * a.p.W * 4;
    CU(cudaMemcpyPeerAsync(b.p.fb + off, dst_device, a.p.fb + off, src_device, bytes, a.stream));
    return CT_OK;
}

int ct_gpu_debug_closest(int device, uint32_t n, const double *origins, const double *directions, const float *t0,
                         uint32_t *found, uint32_t *index, float *tclosest) {
    if (!origins || !directions || !t0) return fail(CT_ERR_INVALID, "NULL ray arrays");
    TRY(check_device(device));
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    if (n == 0) return CT_OK;
    double *d_o = nullptr, *d_d = nullptr; float *d_t0 = nullptr, *d_tc = nullptr; uint32_t *d_f = nullptr, *d_i = nullptr;
    int rc = CT_OK;
    auto cleanup = [&]() { cudaFree(d_o); cudaFree(d_d); cudaFree(d_t0); cudaFree(d_tc); cudaFree(d_f); cudaFree(d_i); };
#define CUX(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(CT_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); cleanup(); return rc; } } while (0)
    CUX(cudaMalloc(&d_o, 24ull * n)); CUX(cudaMalloc(&d_d, 24ull * n)); CUX(cudaMalloc(&d_t0, 4ull * n));
    CUX(cudaMalloc(&d_tc, 4ull * n)); CUX(cudaMalloc(&d_f, 4ull * n)); CUX(cudaMalloc(&d_i, 4ull * n));
    CUX(cudaMemcpy(d_o, origins, 24ull * n, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(d_d, directions, 24ull * n, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(d_t0, t0, 4ull * n, cudaMemcpyHostToDevice));
    CUX(cudaStreamSynchronize(s.stream));
    k_debug_closest<<<(n + 127) / 128, 128, 0, s.stream>>>(s.p, n, d_o, d_d, d_t0, d_f, d_i, d_tc);
    CUX(cudaGetLastError());
    CUX(cudaStreamSynchronize(s.stream));
    if (found) CUX(cudaMemcpy(found, d_f, 4ull * n, cudaMemcpyDeviceToHost));
    if (index) CUX(cudaMemcpy(index, d_i, 4ull * n, cudaMemcpyDeviceToHost));
    if (tclosest) CUX(cudaMemcpy(tclosest, d_tc, 4ull * n, cudaMemcpyDeviceToHost));
    cleanup();
    return CT_OK;
}

static int debug_primitives_impl(int device, uint32_t n, const double *origins, const double *directions, float *ray_t,
                                 const double *tri, const double *bmin, const double *bmax, uint32_t *tri_hit, uint32_t *box_hit,
                                 uint32_t *filter_out, double bound_scale, bool keep_ray_t = false) {
    TRY(check_device(device));
    if (n == 0) return CT_OK;
    double *d_o = nullptr, *d_d = nullptr, *d_tri = nullptr, *d_mn = nullptr, *d_mx = nullptr; float *d_t = nullptr;
    uint32_t *d_th = nullptr, *d_bh = nullptr, *d_f = nullptr;
    int rc = CT_OK;
    auto cleanup = [&]() { cudaFree(d_o); cudaFree(d_d); cudaFree(d_tri); cudaFree(d_mn); cudaFree(d_mx); cudaFree(d_t); cudaFree(d_th); cudaFree(d_bh); cudaFree(d_f); };
    CUX(cudaMalloc(&d_o, 24ull * n)); CUX(cudaMalloc(&d_d, 24ull * n)); CUX(cudaMalloc(&d_tri, 72ull * n));
    CUX(cudaMalloc(&d_mn, 24ull * n)); CUX(cudaMalloc(&d_mx, 24ull * n)); CUX(cudaMalloc(&d_t, 4ull * n));
    CUX(cudaMalloc(&d_th, 4ull * n)); CUX(cudaMalloc(&d_bh, 4ull * n));
    if (filter_out) CUX(cudaMalloc(&d_f, 4ull * n));
    CUX(cudaMemcpy(d_o, origins, 24ull * n, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(d_d, directions, 24ull * n, cudaMemcpyHostToDevice));
    if (tri) CUX(cudaMemcpy(d_tri, tri, 72ull * n, cudaMemcpyHostToDevice));
    else CUX(cudaMemset(d_tri, 0, 72ull * n));
    CUX(cudaMemcpy(d_mn, bmin, 24ull * n, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(d_mx, bmax, 24ull * n, cudaMemcpyHostToDevice));
    CUX(cudaMemcpy(d_t, ray_t, 4ull * n, cudaMemcpyHostToDevice));
    k_debug_primitives<<<(n + 127) / 128, 128>>>(n, d_o, d_d, d_t, d_tri, d_mn, d_mx, d_th, d_bh, d_f, bound_scale, filter_out != nullptr && tri != nullptr);
    CUX(cudaGetLastError());
    CUX(cudaDeviceSynchronize());
    if (tri && !keep_ray_t) CUX(cudaMemcpy(ray_t, d_t, 4ull * n, cudaMemcpyDeviceToHost));
    if (tri_hit) CUX(cudaMemcpy(tri_hit, d_th, 4ull * n, cudaMemcpyDeviceToHost));
    if (box_hit) CUX(cudaMemcpy(box_hit, d_bh, 4ull * n, cudaMemcpyDeviceToHost));
    if (filter_out) CUX(cudaMemcpy(filter_out, d_f, 4ull * n, cudaMemcpyDeviceToHost));
    cleanup();
#undef CUX
    return CT_OK;
}

int ct_gpu_debug_primitives(int device, uint32_t n, const double *origins, const double *directions, float *ray_t,
                            const double *tri, const double *bmin, const double *bmax, uint32_t *tri_hit, uint32_t *box_hit) {
    if (!origins || !directions || !ray_t || !tri || !bmin || !bmax || !tri_hit || !box_hit) return fail(CT_ERR_INVALID, "NULL array");
    return debug_primitives_impl(device, n, origins, directions, ray_t, tri, bmin, bmax, tri_hit, box_hit, nullptr, 1.0);
}

int ct_gpu_debug_filter(int device, uint32_t n, const double *origins, const double *directions, const float *ray_t,
                        const double *tri, const double *bmin, const double *bmax, double bound_scale, uint32_t *verdict) {
    if (!origins || !directions || !ray_t || !bmin || !bmax || !verdict) return fail(CT_ERR_INVALID, "NULL array");
    if (!(bound_scale >= 1.0)) return fail(CT_ERR_INVALID, "bound_scale must be >= 1");
    return debug_primitives_impl(device, n, origins, directions, const_cast<float *>(ray_t), tri, bmin, bmax, nullptr, nullptr, verdict, bound_scale, true);
}

int ct_gpu_shutdown(int device) {
    if (device < 0 || device >= kMaxDevices) return fail(CT_ERR_NO_DEVICE, "device %d out of range", device);
    std::lock_guard<std::mutex> lock(g_mutex);
    DeviceState &s = g_dev[device];
    if (s.loaded || !s.allocs.empty()) {
        cudaSetDevice(device);
        cudaDeviceSynchronize();
        free_device(s);
    }
    return CT_OK;
}

}  // extern "C"
