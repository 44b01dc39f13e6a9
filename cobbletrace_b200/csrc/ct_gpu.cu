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


#include "ct_layout.cuh"     // constants, device structs, Params
#include "ct_traverse.cuh"   // the BVH walks
#include "ct_kernels.cuh"    // the kernels of a tile
#include "ct_build.cuh"      // upload-time scene build kernels

namespace {

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
    uint32_t *parent_q_all = nullptr;                                   // [depth][cap]: parent path of a reflection path (k_emit)
    uint32_t *far_all = nullptr;                                        // [depth][cap]: k_bounce's far lists
    cudaEvent_t ev_lit[16] = {};                                        // shadow verdicts of depth d complete (k_shade of deeper levels may read them)
    unsigned long long rays_shadow_reused = 0;
    OvfRay *ovf_all = nullptr; uint32_t *ovf_huge_all = nullptr;        // [2 * depth + kind][ovf_cap]
    int levels = 1;
    cudaStream_t aux[4] = {};
    cudaEvent_t ev_hit[16] = {}, ev_done[16] = {};
    // multi-GPU frame sharing (ct_gpu_share_*): the cursor and framebuffer a shared render uses (own or the root's)
    unsigned long long *cursor_own = nullptr, *share_cursor = nullptr;
    unsigned long long shared_frames = 0;    // shared renders so far: frame f uses cursor f & 1 (the root zeroes the other one meanwhile)
    int part_index = 0, part_count = 0;      // ct_gpu_share_partition
    bool remote_cursor_ok = true;            // atomics on the attached root's cursor are native (same device, or P2P native atomics)
    uint32_t *share_fb = nullptr;
    void *ipc_opened[2] = {nullptr, nullptr};
    int n_stages = 0;
    bool timed = false;
    // ct_gpu_readback_async: a device-side snapshot of the framebuffer and the stream that copies it to the host
    uint32_t *fb_snapshot = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;
    bool copy_pending = false;
    void *hot_base = nullptr; size_t hot_bytes = 0;      // pairs32 + tris32 (one allocation): the L2 access-policy window
    size_t l2_window = 0, l2_set_aside = 0;              // what apply_l2_policy obtained (0: none)
    std::vector<void *> allocs;
    ct_ray_counters snapshot{};      // totals at the end of the previous counted tile
    unsigned long long rays_overflow = 0, rays_in_place = 0;   // DevTotals' overflow counters as of the last read_totals
    unsigned long long box_exact = 0, tri_exact = 0;
    // rows rendered so far (framebuffer rows), for readback clipping
    int col_lo = 0, col_hi = 0;
};

DeviceState g_dev[kMaxDevices];
std::mutex g_mutex;
// One host thread drives a device at a time (the boss uses one thread per GPU): every entry point that touches a device's
// state holds that device's lock, so a second thread calling in for the same device waits instead of corrupting the state.
std::recursive_mutex g_dev_mutex[kMaxDevices];
#define DEVICE_GUARD(device) std::lock_guard<std::recursive_mutex> dev_lock_(g_dev_mutex[((device) >= 0 && (device) < kMaxDevices) ? (device) : 0])
long long g_budget_option = 0;       // ct_gpu_set_option("traversal_budget"); 0 = default
long long g_warp_budget_option = 0;  // ct_gpu_set_option("overflow_warp_budget"); 0 = default
long long g_any_leaves = 0;               // ct_gpu_set_option("any_leaves"); 0 = by the number of lights
long long g_primary_split = 0;            // ct_gpu_set_option("primary_split"): parked primary rays are finished one ray per warp (k_primary_split)
long long g_primary_budget_option = 0;   // ct_gpu_set_option("primary_budget"); 0 = default (kDefaultPrimaryBudget: never), < 0 = never
long long g_static_eighths = 7;      // ct_gpu_set_option("shared_static_eighths"), see next_chunk
long long g_shared_chunk_shift = 0;  // ct_gpu_set_option("shared_chunk_shift"): 0 = kChunkSharedShift
long long g_run_shift = 0;           // ct_gpu_set_option("shared_run_shift")
long long g_emulate_ranks = 0;       // ct_gpu_set_option("emulate_ranks"): profiling aid, see ct_gpu.h
long long g_shadow_reuse = 1;        // ct_gpu_set_option("shadow_reuse"): a shading point that repeats its parent's takes the parent's shadow verdicts
long long g_hold_frame = 0;          // ct_gpu_set_option("shared_hold_frame"): test aid, see ct_gpu.h
long long g_l2_persist = 1;          // ct_gpu_set_option("l2_persist"): pin the walk's fp32 records in L2 (apply_l2_policy)

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
    if (s.l2_set_aside) {                                  // hand the persisting lines back
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.num_bytes = 0;
        if (s.own_stream) cudaStreamSetAttribute(s.own_stream, cudaStreamAttributeAccessPolicyWindow, &attr);   // (a caller's stream may be gone by now: ct_gpu_set_stream clears it when it is replaced)
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
    }
    for (void *q : s.ipc_opened) if (q) cudaIpcCloseMemHandle(q);
    for (void *p : s.allocs) cudaFree(p);
    s.allocs.clear();
    if (s.ev0) cudaEventDestroy(s.ev0);
    if (s.ev1) cudaEventDestroy(s.ev1);
    for (cudaEvent_t e : s.tile_done) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s.stage_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s.ev_hit) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s.ev_done) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s.ev_lit) if (e) cudaEventDestroy(e);
    for (cudaStream_t a : s.aux) if (a) cudaStreamDestroy(a);
    if (s.copy_stream) { cudaStreamSynchronize(s.copy_stream); cudaStreamDestroy(s.copy_stream); }
    if (s.ev_snap) cudaEventDestroy(s.ev_snap);
    if (s.ev_copied) cudaEventDestroy(s.ev_copied);
    if (s.own_stream) cudaStreamDestroy(s.own_stream);
    s = DeviceState{};
}

// The hull of rows holding rendered pixels grows from two sides: tiles this device renders (its own host thread) and rows
// gathered into it from other devices (their host threads, ct_gpu_gather_rows).
std::mutex g_rows_mutex;
void extend_rows(DeviceState &s, int lo, int hi) {
    std::lock_guard<std::mutex> lock(g_rows_mutex);
    if (s.row_hi <= s.row_lo) { s.row_lo = lo; s.row_hi = hi; }
    else { s.row_lo = std::min(s.row_lo, lo); s.row_hi = std::max(s.row_hi, hi); }
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


// depth of the reference's DFS (stack entries needed) -- iterative to survive degenerate trees.  Also marks the nodes the
// DFS reaches (array slots nobody points at may hold anything) and rejects leaves whose triangle ranges overlap.
int bvh_depth(const ct_bvh_node *nodes, uint32_t n_nodes, uint32_t n_tri, bool *ok, std::vector<uint8_t> &reachable) {
    std::vector<std::pair<uint32_t, int>> st;
    std::vector<uint8_t> covered(n_tri, 0);
    reachable.assign(n_nodes, 0);
    st.push_back({0u, 1});
    int best = 0;
    size_t visited = 0;
    *ok = true;
    while (!st.empty()) {
        auto [i, d] = st.back();
        st.pop_back();
        if (i >= n_nodes || ++visited > (size_t)n_nodes || reachable[i]) { *ok = false; return 0; }      // out of range, a cycle, or a node with two parents
        reachable[i] = 1;
        best = std::max(best, d);
        const ct_bvh_node &nd = nodes[i];
        if (nd.triangle_count == 0) {
            if ((uint64_t)nd.left_node + 1u >= n_nodes || nd.left_node == 0u) { *ok = false; return 0; }
            st.push_back({nd.left_node, d + 1});
            st.push_back({nd.left_node + 1, d + 1});
        } else {
            if ((uint64_t)nd.first_triangle_index + nd.triangle_count > n_tri) { *ok = false; return 0; }
            for (uint32_t k = 0; k < nd.triangle_count; k++) {
                uint8_t &c = covered[nd.first_triangle_index + k];
                if (c) { *ok = false; return 0; }                                  // two leaves hold the same position
                c = 1;
            }
        }
    }
    return best;
}

int launch_grid(const DeviceState &s, int blocks_per_sm) { return s.n_sm * blocks_per_sm; }

// L2 residency of the walk's hot set.  The fp32 child pairs and fp32 triangles (76 MB for the dragon-class scene) are
// read by every walk step, at random; everything else a frame touches (ray queues, hit records, occlusion masks, colour
// stacks, the fp64 records behind the filters) streams through once.  Left to the default policy the streams evict the
// tree (measured: 70 % L2 hit rate on a hot set that fits, every other warp step waiting for DRAM).  So: set aside as
// much L2 as the device allows for persisting lines and open an access-policy window over the hot allocation on every
// stream the tile's kernels run on; hitRatio scales the window down to the set-aside size when the scene is larger.
int apply_l2_policy(DeviceState &s, int device, cudaStream_t st) {
    if (!g_l2_persist || !s.hot_base || s.hot_bytes == 0) return CT_OK;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return CT_OK;
    if (s.l2_set_aside == 0) {
        const size_t want = std::min<size_t>(s.hot_bytes, (size_t)prop.persistingL2CacheMaxSize);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); return CT_OK; }
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize);
        s.l2_set_aside = got;
        s.l2_window = std::min<size_t>(s.hot_bytes, (size_t)prop.accessPolicyMaxWindowSize);
    }
    if (s.l2_set_aside == 0) return CT_OK;
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = s.hot_base;
    attr.accessPolicyWindow.num_bytes = s.l2_window;
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)s.l2_set_aside / (double)s.l2_window);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    return CT_OK;
}

int read_totals(DeviceState &s, ct_ray_counters *out) {   // synchronises the stream
    DevTotals h;
    CU(cudaMemcpyAsync(&h, s.p.tot, sizeof h, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    out->rays_primary = h.rays_primary; out->rays_shadow = h.rays_shadow; out->rays_reflection = h.rays_reflection;
    out->box_tests = h.box_tests; out->tri_tests = h.tri_tests;
    s.rays_overflow = h.rays_overflow; s.rays_in_place = h.rays_in_place;
    s.box_exact = h.box_exact; s.tri_exact = h.tri_exact;
    s.rays_shadow_reused = h.rays_shadow_reused;
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
    if ((d->flags & CT_FLAG_SUPERSAMPLING) && (d->flags & CT_FLAG_KEEP_HITS))
        return fail(CT_ERR_INVALID, "CT_FLAG_SUPERSAMPLING cannot be combined with CT_FLAG_KEEP_HITS (a pixel has 16 primary rays)");
    bool ok = true;
    std::vector<uint8_t> reachable;
    int depth = bvh_depth(d->nodes, d->n_nodes, d->n_triangles, &ok, reachable);
    if (!ok) return fail(CT_ERR_INVALID, "BVH is malformed (child or triangle range out of bounds, overlapping leaves, a cycle or a shared child)");
    if (depth > kStackMax) return fail(CT_ERR_LIMIT, "BVH depth %d exceeds the device traversal stack (%d)", depth, kStackMax);
    TRY(check_device(device));
    DEVICE_GUARD(device);
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
    for (cudaEvent_t &e : s.ev_lit) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaStream_t &a : s.aux) CU(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
    if (d->flags & CT_FLAG_STAGE_TIMING) for (cudaEvent_t &e : s.stage_ev) CU(cudaEventCreate(&e));
    s.p.budget = g_budget_option > 0 ? (uint32_t)std::min<long long>(g_budget_option, 1ll << 30) : kDefaultBudget;
    s.p.warp_budget = g_warp_budget_option > 0 ? (uint32_t)std::min<long long>(g_warp_budget_option, 1ll << 30) : kWarpBudget;
    s.p.primary_budget = g_primary_budget_option < 0 ? 0u : g_primary_budget_option > 0 ? (uint32_t)std::min<long long>(g_primary_budget_option, 1ll << 30) : kDefaultPrimaryBudget;
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
        if (reachable[i]) pid_of[i] = d->nodes[i].triangle_count == 0 ? n_pairs++ : kLeafMark;      // unreachable slots get no record and write nothing
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
    // the early-exit walks' wide tree: kWideLevels levels of the reference's tree per node, children in DFS order
    std::vector<DevWide> wide;
    if (d->nodes[0].triangle_count == 0) {
        std::vector<uint32_t> root_of;                  // wide node -> the interior node whose descendants it holds
        std::vector<uint32_t> wid_of(d->n_nodes, kNoPos);
        root_of.reserve(n_pairs / 2 + 1);
        root_of.push_back(0u); wid_of[0] = 0u;
        wide.reserve(n_pairs / 2 + 1);
        for (size_t w = 0; w < root_of.size(); w++) {
            uint32_t cur[kWide], nxt[kWide];
            int n_cur = 2;
            cur[0] = d->nodes[root_of[w]].left_node; cur[1] = cur[0] + 1u;
            for (int lvl = 1; lvl < kWideLevels; lvl++) {
                int n_nxt = 0;
                for (int i = 0; i < n_cur; i++) {
                    const ct_bvh_node &c = d->nodes[cur[i]];
                    if (c.triangle_count != 0) nxt[n_nxt++] = cur[i];
                    else { nxt[n_nxt++] = c.left_node; nxt[n_nxt++] = c.left_node + 1u; }
                }
                n_cur = n_nxt;
                memcpy(cur, nxt, sizeof cur);
            }
            DevWide rec;
            for (int i = 0; i < kWide; i++) {
                DevWideChild &e = rec.c[i];
                if (i >= n_cur) {
                    for (int a = 0; a < 3; a++) { e.bmin[a] = 1e30f; e.bmax[a] = -1e30f; }
                    e.ref = kNoPos; e.cnt = 0u;
                    continue;
                }
                const ct_bvh_node &c = d->nodes[cur[i]];
                for (int a = 0; a < 3; a++) { e.bmin[a] = (float)c.aabb_min[a]; e.bmax[a] = (float)c.aabb_max[a]; }
                e.cnt = c.triangle_count;
                if (c.triangle_count != 0) e.ref = c.first_triangle_index;
                else {
                    if (wid_of[cur[i]] == kNoPos) { wid_of[cur[i]] = (uint32_t)root_of.size(); root_of.push_back(cur[i]); }
                    e.ref = wid_of[cur[i]];
                }
            }
            wide.push_back(rec);
        }
    }
    lap("wide tree (host)");

    DevPair32 *dp32 = nullptr; DevPair64 *dp64 = nullptr; DevTri *dt = nullptr; DevTri32 *dt32 = nullptr; ct_material *dm = nullptr;
    uint32_t *dpp = nullptr, *dtp = nullptr;
    {
        // the records every walk step reads -- fp32 child pairs and fp32 triangles -- share ONE allocation, so that a single
        // L2 access-policy window can cover them (apply_l2_policy)
        const size_t pair_bytes = ((size_t)std::max<uint32_t>(n_pairs, 1) * sizeof(DevPair32) + 255u) & ~(size_t)255u;
        const size_t wide_bytes = (wide.size() * sizeof(DevWide) + 255u) & ~(size_t)255u;
        unsigned char *hot = nullptr;
        s.hot_bytes = wide_bytes + pair_bytes + (size_t)d->n_triangles * sizeof(DevTri32);
        TRY(dev_alloc(s, &hot, s.hot_bytes));
        s.hot_base = hot;
        if (!wide.empty()) {
            CU(cudaMemcpyAsync(hot, wide.data(), wide.size() * sizeof(DevWide), cudaMemcpyHostToDevice, s.stream));
            CU(cudaStreamSynchronize(s.stream));
            p.wide = reinterpret_cast<const DevWide *>(hot);
            p.n_wide = (uint32_t)wide.size();
        }
        dp32 = reinterpret_cast<DevPair32 *>(hot + wide_bytes); dt32 = reinterpret_cast<DevTri32 *>(hot + wide_bytes + pair_bytes);
    }
    TRY(dev_alloc(s, &dp64, std::max<uint32_t>(n_pairs, 1)));
    TRY(dev_alloc(s, &dt, d->n_triangles)); TRY(dev_alloc(s, &dm, d->n_triangles));
    TRY(dev_alloc(s, &dtp, d->n_triangles));
    if (s.can_overflow) TRY(dev_alloc(s, &dpp, std::max<uint32_t>(n_pairs, 1)));
    BuildReport rep{};
    {
        // staging copies of the caller's arrays, freed again below
        ct_bvh_node *raw_nodes = nullptr; uint32_t *raw_pid = nullptr, *raw_idx = nullptr, *seen = nullptr; unsigned char *raw_tris = nullptr; BuildReport *drep = nullptr;
        const size_t tri_bytes = (size_t)(d->n_triangles - 1) * d->triangle_stride + 72;      // the last triangle may end at its third vertex
        auto release = [&] { cudaFree(raw_nodes); cudaFree(raw_pid); cudaFree(raw_idx); cudaFree(raw_tris); cudaFree(drep); cudaFree(seen); };
        auto staged = [&](cudaError_t e) { if (e != cudaSuccess) { release(); free_device(s); } return e; };
        CU(staged(cudaMalloc(&raw_nodes, (size_t)d->n_nodes * sizeof(ct_bvh_node))));
        CU(staged(cudaMalloc(&raw_pid, (size_t)d->n_nodes * 4)));
        CU(staged(cudaMalloc(&raw_idx, (size_t)d->n_triangles * 4)));
        CU(staged(cudaMalloc(&raw_tris, tri_bytes)));
        CU(staged(cudaMalloc(&drep, sizeof rep)));
        const size_t seen_words = ((size_t)d->n_triangles + 31u) / 32u;
        CU(staged(cudaMalloc(&seen, seen_words * 4)));
        CU(staged(cudaMemsetAsync(seen, 0, seen_words * 4, s.stream)));
        rep.bad_pos = kNoPos; rep.pos0 = kNoPos;
        CU(staged(cudaMemcpyAsync(drep, &rep, sizeof rep, cudaMemcpyHostToDevice, s.stream)));
        CU(staged(cudaMemcpyAsync(raw_nodes, d->nodes, (size_t)d->n_nodes * sizeof(ct_bvh_node), cudaMemcpyHostToDevice, s.stream)));
        CU(staged(cudaMemcpyAsync(raw_pid, pid_of.data(), (size_t)d->n_nodes * 4, cudaMemcpyHostToDevice, s.stream)));
        if (dpp) CU(staged(cudaMemsetAsync(dpp, 0xff, (size_t)std::max<uint32_t>(n_pairs, 1) * 4, s.stream)));
        CU(staged(cudaMemsetAsync(dtp, 0xff, (size_t)d->n_triangles * 4, s.stream)));
        const int build_blocks = s.n_sm * 8;
        k_build_pairs<<<build_blocks, 256, 0, s.stream>>>(raw_nodes, raw_pid, d->n_nodes, dp32, dp64, dpp, dtp, d->n_triangles, drep);
        CU(staged(cudaMemcpyAsync(raw_idx, d->tri_indexes, (size_t)d->n_triangles * 4, cudaMemcpyHostToDevice, s.stream)));
        CU(staged(cudaMemcpyAsync(raw_tris, d->triangles, tri_bytes, cudaMemcpyHostToDevice, s.stream)));
        CU(staged(cudaMemcpyAsync(dm, d->materials, (size_t)d->n_triangles * sizeof(ct_material), cudaMemcpyHostToDevice, s.stream)));
        k_build_tris<<<build_blocks, 256, 0, s.stream>>>(raw_tris, (uint32_t)d->triangle_stride, raw_idx, dm, d->n_triangles, dt, dt32, seen, drep);
        CU(staged(cudaGetLastError()));
        CU(staged(cudaMemcpyAsync(&rep, drep, sizeof rep, cudaMemcpyDeviceToHost, s.stream)));
        CU(staged(cudaStreamSynchronize(s.stream)));
        release();
    }
    if (rep.bad_pos != kNoPos) { free_device(s); return fail(CT_ERR_INVALID, "tri_indexes[%u] = %u out of range", rep.bad_pos, d->tri_indexes[rep.bad_pos]); }
    if (rep.duplicate || rep.pos0 == kNoPos) { free_device(s); return fail(CT_ERR_INVALID, "tri_indexes is not a permutation of 0..n_triangles-1 (an index occurs twice)"); }
    p.pos_of_tri0 = rep.pos0;
    p.nested = (!rep.not_nested && !rep.boxes_bad) ? 1u : 0u;
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
    // the any-hit walk's leaf-list length trades the lit rays' phase overhead against the occluded rays' early exit: measured 12 best for
    // the 3-light dragon frame (k_shadow 1.11 ms against 1.15 at 8), 8 for the 66-light scene (6.75 ms against 7.14 at 12) -- a shading point
    // inside a sphere of lights has half of them behind its own surface.  Option "any_leaves" overrides (0 = this rule).
    p.any_leaves = g_any_leaves > 0 ? (uint32_t)g_any_leaves : (p.n_slights >= 16u ? 8u : 12u);
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
        // parked rays per launch: 2^18 (16 MB) -- with primary walks parked as well (option "primary_budget": 2-3 % of a dragon-class
        // frame's primary rays) 1/16 of the frame's paths where that is more; a full buffer means finishing rays in place, which
        // must stay hypothetical
        p.ovf_cap = 1u << 18;
        if (p.primary_budget > 0u) p.ovf_cap = std::max<uint32_t>(p.ovf_cap, p.cap / 16u);
        TRY(dev_alloc(s, &s.ovf_all, (size_t)p.ovf_cap * 2 * levels));
        TRY(dev_alloc(s, &s.ovf_huge_all, (size_t)p.ovf_cap * 2 * levels));
        p.ovf = s.ovf_all; p.ovf_huge = s.ovf_huge_all;
    }
    if (levels > 1) {
        TRY(dev_alloc(s, &s.hitb_t_all, (size_t)p.cap * levels)); TRY(dev_alloc(s, &s.hitb_pos_all, (size_t)p.cap * levels));
        TRY(dev_alloc(s, &p.stack_refl, (size_t)p.cap * levels));
        TRY(dev_alloc(s, &s.rays_all, (size_t)p.cap * 6 * (levels + 1))); TRY(dev_alloc(s, &s.path_slot_all, (size_t)p.cap * (levels + 1)));
        TRY(dev_alloc(s, &s.parent_q_all, (size_t)p.cap * (levels + 1)));
        TRY(dev_alloc(s, &s.far_all, (size_t)p.cap * levels));
    }
    TRY(dev_alloc(s, &p.fb, (size_t)d->width * d->height, true));       // calloc'd like cobbletrace.cpp:57
    p.subsample = (d->flags & CT_FLAG_SUBSAMPLING) ? 1 : 0;
    p.supersample = (d->flags & CT_FLAG_SUPERSAMPLING) ? 1 : 0;
    if (p.subsample || p.supersample) TRY(dev_alloc(s, &p.final_color, p.cap, true));
    TRY(dev_alloc(s, &p.own_chunks, (p.cap >> kChunkLocalShift) + 1u));
    TRY(dev_alloc(s, &s.cursor_own, 2 * kCursorStride, true));           // its own allocation: exported over CUDA IPC; two cursors, 128 bytes apart
    s.share_cursor = s.cursor_own; s.share_fb = p.fb;
    if (d->flags & CT_FLAG_KEEP_HITS) {
        size_t npx = (size_t)d->width * d->height;
        TRY(dev_alloc(s, &p.dbg_found, npx)); TRY(dev_alloc(s, &p.dbg_index, npx, true)); TRY(dev_alloc(s, &p.dbg_t, npx, true));
        CU(cudaMemset(p.dbg_found, 0xff, npx * sizeof(uint32_t)));
    }
    TRY(dev_alloc(s, &p.sched, 1, true));
    TRY(dev_alloc(s, &p.tot, 1, true));
    lap("path state allocation");
    TRY(apply_l2_policy(s, device, s.stream));
    for (cudaStream_t a : s.aux) TRY(apply_l2_policy(s, device, a));
    if (timing) fprintf(stderr, "ct_gpu_upload_scene: L2 window %zu bytes, set-aside %zu bytes; total %.1f ms\n", s.l2_window, s.l2_set_aside, now_ms() - t_begin);
    s.loaded = true;
    return CT_OK;
}

int ct_gpu_set_camera(int device, const double position[3], const double rotation[9]) {
    if (!position || !rotation) return fail(CT_ERR_INVALID, "NULL camera");
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    memcpy(s.p.cam, position, sizeof s.p.cam);
    memcpy(s.p.rot, rotation, sizeof s.p.rot);
    return CT_OK;
}

int ct_gpu_set_stream(int device, void *cuda_stream) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    CU(cudaStreamSynchronize(s.stream));
    if (s.stream != s.own_stream && s.l2_set_aside) {        // take the access-policy window off the caller's previous stream
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.num_bytes = 0;
        if (cudaStreamSetAttribute(s.stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
    }
    s.stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s.own_stream;
    TRY(apply_l2_policy(s, device, s.stream));
    return CT_OK;
}

static int render_impl(int device, int y_start, int y_end, ct_ray_counters *counters, bool shared) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
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
    // shared frames alternate between two cursors: while frame f steals from cursor f & 1, the root zeroes the other one on
    // its stream for frame f + 1 -- nobody touches that one before the frame-end rendezvous, so no reset call and no
    // host synchronisation stands between two frames
    const bool same_frame = shared && g_hold_frame && s.shared_frames > 0;      // (tests: one device plays several participants of ONE frame)
    const unsigned long long frame_no = same_frame ? s.shared_frames - 1ull : s.shared_frames;
    p.steal = shared ? s.share_cursor + kCursorStride * (frame_no & 1ull) : &p.sched->steal_local;
    const uint32_t shared_shift = g_shared_chunk_shift ? (uint32_t)g_shared_chunk_shift : kChunkSharedShift;
    p.chunk_shift = shared ? shared_shift : kChunkLocalShift;
    p.steal_stride = g_emulate_ranks > 1 ? (uint32_t)g_emulate_ranks : 1u;
    if (p.steal_stride > 1) p.chunk_shift = shared_shift;
    const bool dealt = shared && p.steal_stride == 1 && s.part_count > 1;
    p.part_index = dealt ? (uint32_t)s.part_index : 0u;
    p.part_count = dealt ? (uint32_t)s.part_count : 0u;
    p.static_eighths = (uint32_t)g_static_eighths;
    p.run_shift = (shared || p.steal_stride > 1) ? (uint32_t)g_run_shift : 0u;
    if (shared && !s.remote_cursor_ok) {
        // no native atomics to the root's cursor: stealing from it would hand chunks out twice or not at all
        if (!dealt) return fail(CT_ERR_CUDA, "device %d has no native atomics to the shared frame's cursor: declare the participants (ct_gpu_share_partition), then every chunk is dealt", device);
        p.static_eighths = 8u;
    }
    p.fb_out = shared ? s.share_fb : p.fb;
    Params pk = p; pk.max_depth = depth_max;   // with no reflective material the recursion never goes past depth 0 (:369)
    pk.parent_q_all = s.parent_q_all; pk.hitb_t_all = s.hitb_t_all; pk.occ_all = s.occ_all;
    pk.reuse_shadow = (g_shadow_reuse && s.parent_q_all) ? 1u : 0u;
    cudaStream_t st = s.stream;
    if (shared && !same_frame) {
        if (s.share_cursor == s.cursor_own) CU(cudaMemsetAsync(s.cursor_own + kCursorStride * ((s.shared_frames + 1ull) & 1ull), 0, sizeof(unsigned long long), st));
        s.shared_frames++;
    }
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
            v.parent_q[d & 1] = s.parent_q_all + cap * d; v.parent_q[(d & 1) ^ 1] = s.parent_q_all + cap * (d + 1);
        }
        v.occ = s.occ_all + cap * pk.occ_words * d;
        if (s.far_all) v.far_list = s.far_all + cap * d;
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
    {
        Params v0 = view(0);              // (k_primary also runs the recursion step of depth 0: it needs the depth's queues)
        // walks longer than primary_budget pair visits are parked (slot 0 of the parking buffers: depth 0 has no bounce launch)
        // and walked again by k_primary_long, in warps made of long walks only
        const bool park_primary = s.can_overflow && pk.primary_budget > 0u;
        v0.ovf = park_primary ? s.ovf_all : nullptr;
        if (park_primary) {
            if (count) k_primary<true, true><<<grid, kBlockThreads, 0, st>>>(v0);
            else k_primary<false, true><<<grid, kBlockThreads, 0, st>>>(v0);
        } else {
            if (count) k_primary<true, false><<<grid, kBlockThreads, 0, st>>>(v0);
            else k_primary<false, false><<<grid, kBlockThreads, 0, st>>>(v0);
        }
        TRY(mark("primary", 0));
        if (park_primary) {
            if (g_primary_split) {
                if (count) k_primary_split<true><<<s.n_sm * 4, kOvfThreads, 0, st>>>(v0);
                else k_primary_split<false><<<s.n_sm * 4, kOvfThreads, 0, st>>>(v0);
                TRY(mark("primary_split", 0));
            } else {
                if (count) k_primary_long<true><<<grid, kBlockThreads, 0, st>>>(v0, work++);
                else k_primary_long<false><<<grid, kBlockThreads, 0, st>>>(v0, work++);
                TRY(mark("primary_long", 0));
            }
        }
    }
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
        if (depth_max > 0) {
            // TraceRay's recursion step runs inside the hit kernels (emit_paths); what is left for a kernel of its own are the
            // reflection paths k_bounce had parked, whose hit k_overflow has just delivered
            if (!kFusedEmit) { k_emit<<<grid, kBlockThreads, 0, st>>>(v, d, work++, nullptr, 0); TRY(mark("emit", d)); }
            else if (d > 0 && s.can_overflow) {
                k_emit<<<s.n_sm * 2, kBlockThreads, 0, st>>>(v, d, work++, s.ovf_all + (size_t)pk.ovf_cap * ovf_b, ovf_b);
                TRY(mark("emit_late", d));
            }
        }
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
        if (!stages) {
            // k_shade(d) may read the shadow verdicts of shallower depths (occlusion_source): they are complete at ev_lit
            CU(cudaEventRecord(s.ev_lit[d], side));
            for (int dd = 0; dd < d; dd++) CU(cudaStreamWaitEvent(side, s.ev_lit[dd], 0));
        }
        k_shade<<<grid, kBlockThreads, 0, side>>>(v, d, work++);
        if (stages) TRY(mark("shade", d)); else s.launches++;
        if (!stages) CU(cudaEventRecord(s.ev_done[d], side));
    }
    if (!stages) for (int d = 0; d <= depth_max; d++) CU(cudaStreamWaitEvent(st, s.ev_done[d], 0));
    if (depth_max > 0) { k_resolve<<<s.n_sm * 4, 256, 0, st>>>(pk); TRY(mark("resolve", 0)); }
    if (pk.supersample) { k_supersample<<<s.n_sm * 4, 256, 0, st>>>(pk); TRY(mark("supersample", 0)); }
    if (pk.subsample) { k_subsample<<<s.n_sm * 4, 256, 0, st>>>(pk); TRY(mark("subsample", 0)); }
    CU(cudaEventRecord(s.ev1, st));
    CU(cudaEventRecord(s.tile_done[s.tiles_submitted % 8], st));
    s.tiles_submitted++;
    CU(cudaGetLastError());
    s.timed = true;
    {   // framebuffer rows this tile writes: row = H/2 - y
        int lo = half - (y1 - 1), hi = half - y0 + 1 + (s.p.subsample ? 1 : 0);   // subsampling also stores row y0 - 1
        lo = std::max(lo, 0); hi = std::min(hi, s.p.H);
        extend_rows(s, lo, hi);
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
    DEVICE_GUARD(device);
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
    out->frames = s.shared_frames;
    return CT_OK;
}

int ct_gpu_share_attach(int device, const ct_gpu_share *root) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    for (void *&q : s.ipc_opened) if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
    s.remote_cursor_ok = true;
    if (root) s.shared_frames = root->frames;                // which of the two cursors the next shared frame uses
    if (!root) {                                             // detach: back to this device's own cursor and framebuffer
        s.share_cursor = s.cursor_own; s.share_fb = s.p.fb;
        return CT_OK;
    }
    if (root->struct_size != sizeof(ct_gpu_share)) return fail(CT_ERR_INVALID, "ct_gpu_share struct_size != %zu", sizeof(ct_gpu_share));
    if (root->width != s.p.W || root->height != s.p.H) return fail(CT_ERR_INVALID, "shared frame is %dx%d, this device renders %dx%d", root->width, root->height, s.p.W, s.p.H);
    if (root->device != device) {
        // next_chunk steals with atomicAdd on the root's cursor: across devices that is only atomic where the link has
        // native atomics (NVLink; not PCIe peer access).  Without them this device takes dealt chunks only (render_impl).
        int native = 0;
        if (cudaDeviceGetP2PAttribute(&native, cudaDevP2PAttrNativeAtomicSupported, device, root->device) != cudaSuccess) { cudaGetLastError(); native = 0; }
        s.remote_cursor_ok = native != 0;
    }
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
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    if (count < 0 || count > 64 || (count > 0 && (index < 0 || index >= count))) return fail(CT_ERR_INVALID, "bad partition %d of %d", index, count);
    s.part_index = count > 1 ? index : 0;
    s.part_count = count > 1 ? count : 0;
    return CT_OK;
}

int ct_gpu_share_reset(int device) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    CU(cudaMemsetAsync(s.cursor_own, 0, 2 * kCursorStride * sizeof(unsigned long long), s.stream));
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
    if (!strcmp(name, "shared_run_shift")) {
        if (value < 0 || value > 10) return fail(CT_ERR_INVALID, "shared_run_shift must be 0..10");
        g_run_shift = value;
        return CT_OK;
    }
    if (!strcmp(name, "shadow_reuse")) {
        if (value != 0 && value != 1) return fail(CT_ERR_INVALID, "shadow_reuse must be 0 or 1");
        g_shadow_reuse = value;
        return CT_OK;
    }
    if (!strcmp(name, "shared_hold_frame")) {
        if (value != 0 && value != 1) return fail(CT_ERR_INVALID, "shared_hold_frame must be 0 or 1");
        g_hold_frame = value;
        return CT_OK;
    }
    if (!strcmp(name, "l2_persist")) {
        if (value != 0 && value != 1) return fail(CT_ERR_INVALID, "l2_persist must be 0 or 1");
        g_l2_persist = value;
        return CT_OK;
    }
    if (!strcmp(name, "any_leaves")) {
        if (value != 0 && (value < 5 || value > 64)) return fail(CT_ERR_INVALID, "any_leaves must be 0 (default) or 5..64");
        g_any_leaves = value;
        return CT_OK;
    }
    if (!strcmp(name, "primary_split")) {
        if (value != 0 && value != 1) return fail(CT_ERR_INVALID, "primary_split must be 0 or 1");
        g_primary_split = value;
        return CT_OK;
    }
    if (!strcmp(name, "primary_budget")) {
        g_primary_budget_option = value;
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
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    ct_ray_counters tmp;
    TRY(read_totals(s, &tmp));
    if (parked) *parked = s.rays_overflow;
    if (finished_in_place) *finished_in_place = s.rays_in_place;
    return CT_OK;
}

int ct_gpu_reuse_stats(int device, uint64_t *shadow_rays_reused) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    ct_ray_counters tmp;
    TRY(read_totals(s, &tmp));
    if (shadow_rays_reused) *shadow_rays_reused = s.rays_shadow_reused;
    return CT_OK;
}

int ct_gpu_filter_stats(int device, uint64_t *box_exact, uint64_t *tri_exact) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
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
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    CU(cudaStreamSynchronize(s.stream));
    return CT_OK;
}

int ct_gpu_kernel_launches(int device, uint64_t *out, int reset) {
    if (!out) return fail(CT_ERR_INVALID, "NULL out");
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    *out = s.launches;
    if (reset) s.launches = 0;
    return CT_OK;
}

int ct_gpu_last_tile_stages(int device, int max, float *ms, const char **names, int *depth) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
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
    DEVICE_GUARD(device);
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
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded || !s.timed) return fail(CT_ERR_NO_SCENE, "no tile rendered on device %d", device);
    CU(cudaEventSynchronize(s.ev1));
    CU(cudaEventElapsedTime(ms, s.ev0, s.ev1));
    return CT_OK;
}

int ct_gpu_get_counters(int device, ct_ray_counters *out, int reset) {
    if (!out) return fail(CT_ERR_INVALID, "NULL out");
    TRY(check_device(device));
    DEVICE_GUARD(device);
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
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    const Params &p = s.p;
    // Only rows some tile has rendered are copied: e.g. row 0 (y = H/2) is never reached by the reference's
    // loops (y < yEnd <= H/2) and so stays untouched in dst here as well.
    int row_lo, row_hi;
    { std::lock_guard<std::mutex> lock(g_rows_mutex); row_lo = s.row_lo; row_hi = s.row_hi; }
    int r0 = std::max(row_start, row_lo), r1 = std::min(row_end, row_hi);
    int cols = s.col_hi - s.col_lo;
    if (dst_stride_pixels < p.W) return fail(CT_ERR_INVALID, "dst stride %d < width %d", dst_stride_pixels, p.W);
    CU(cudaStreamSynchronize(s.stream));
    if (r1 <= r0 || cols <= 0) return CT_OK;
    CU(cudaMemcpy2D(dst + (size_t)r0 * dst_stride_pixels + s.col_lo, (size_t)dst_stride_pixels * 4,
                    p.fb + (size_t)r0 * p.W + s.col_lo, (size_t)p.W * 4, (size_t)cols * 4, (size_t)(r1 - r0),
                    cudaMemcpyDeviceToHost));
    return CT_OK;
}

int ct_gpu_readback_async(int device, uint32_t *dst, int dst_stride_pixels, int row_start, int row_end) {
    if (!dst) return fail(CT_ERR_INVALID, "NULL dst");
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    const Params &p = s.p;
    if (dst_stride_pixels < p.W) return fail(CT_ERR_INVALID, "dst stride %d < width %d", dst_stride_pixels, p.W);
    if (!s.copy_stream) {
        TRY(dev_alloc(s, &s.fb_snapshot, (size_t)p.W * p.H));
        CU(cudaStreamCreateWithFlags(&s.copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s.ev_snap, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_copied, cudaEventDisableTiming));
    }
    int row_lo, row_hi;
    { std::lock_guard<std::mutex> lock(g_rows_mutex); row_lo = s.row_lo; row_hi = s.row_hi; }
    const int r0 = std::max(row_start, row_lo), r1 = std::min(row_end, row_hi), cols = s.col_hi - s.col_lo;
    // in stream order behind the tiles submitted so far: snapshot the rows on the device (microseconds), so that the next
    // frame may overwrite the framebuffer while the copy stream moves the snapshot to the host (PCIe, ~0.4 ms for a 4K frame)
    if (s.copy_pending) CU(cudaStreamWaitEvent(s.stream, s.ev_copied, 0));       // the previous snapshot has left the device
    if (r1 > r0 && cols > 0)
        CU(cudaMemcpy2DAsync(s.fb_snapshot + (size_t)r0 * p.W + s.col_lo, (size_t)p.W * 4, p.fb + (size_t)r0 * p.W + s.col_lo, (size_t)p.W * 4,
                             (size_t)cols * 4, (size_t)(r1 - r0), cudaMemcpyDeviceToDevice, s.stream));
    CU(cudaEventRecord(s.ev_snap, s.stream));
    CU(cudaStreamWaitEvent(s.copy_stream, s.ev_snap, 0));
    if (r1 > r0 && cols > 0)
        CU(cudaMemcpy2DAsync(dst + (size_t)r0 * dst_stride_pixels + s.col_lo, (size_t)dst_stride_pixels * 4,
                             s.fb_snapshot + (size_t)r0 * p.W + s.col_lo, (size_t)p.W * 4, (size_t)cols * 4, (size_t)(r1 - r0),
                             cudaMemcpyDeviceToHost, s.copy_stream));
    CU(cudaEventRecord(s.ev_copied, s.copy_stream));
    s.copy_pending = true;
    return CT_OK;
}

int ct_gpu_readback_wait(int device) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    if (s.copy_pending) CU(cudaEventSynchronize(s.ev_copied));
    s.copy_pending = false;
    return CT_OK;
}

int ct_gpu_readback_hits(int device, uint32_t *found, uint32_t *index, float *t, int stride_pixels, int row_start, int row_end) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
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
    DEVICE_GUARD(device);
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
    extend_rows(b, r0, r1);                 // the destination's readback must cover rows it did not render itself
    return CT_OK;
}

int ct_gpu_mark_rows(int device, int row_start, int row_end) {
    TRY(check_device(device));
    DEVICE_GUARD(device);
    DeviceState &s = g_dev[device];
    if (!s.loaded) return fail(CT_ERR_NO_SCENE, "no scene uploaded on device %d", device);
    int r0 = std::max(row_start, 0), r1 = std::min(row_end, s.p.H);
    if (r1 > r0) extend_rows(s, r0, r1);
    return CT_OK;
}

int ct_gpu_debug_closest(int device, uint32_t n, const double *origins, const double *directions, const float *t0,
                         uint32_t *found, uint32_t *index, float *tclosest) {
    if (!origins || !directions || !t0) return fail(CT_ERR_INVALID, "NULL ray arrays");
    TRY(check_device(device));
    DEVICE_GUARD(device);
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
    k_debug_closest<<<(n + kBlockThreads - 1) / kBlockThreads, kBlockThreads, 0, s.stream>>>(s.p, n, d_o, d_d, d_t0, d_f, d_i, d_tc);
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
    DEVICE_GUARD(device);
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
