// ct_bvh.cpp -- host BVH build whose output is identical, node for node and index for index, to the
// reference's InitializeBVHState + BuildBVH (bvh.cpp:16-120), because the GPU traversal reproduces the
// reference's DFS order over exactly that tree (ties and the t=0 reflection rays depend on it).
//
// Same algorithm (midpoint split on the longest axis, leaf when <= 2 triangles or when a side would be
// empty, children allocated adjacently), different shape:
//   * iterative pre-order worklist instead of recursion (no stack overflow on degenerate 1M-triangle inputs),
//     centroids in a flat SoA array;
//   * multi-threaded (once the tracer takes milliseconds, the 0.5 s single-threaded build is what
//     a user waits for).  The reference numbers nodes in the order its recursion allocates them (nodesUsed++ at each
//     split, left subtree before right), so a subtree occupies one contiguous block of indices and its internal
//     numbering does not depend on anything outside it.  Nodes are tasks in one pool: a big node is split (the
//     in-place partition of the index array, bvh.cpp:70-81, is order-dependent and stays sequential per node;
//     disjoint ranges run concurrently) and its children are queued; a node at or below the grain is built with its
//     whole subtree as a block with block-local numbering; one pre-order pass over the top then assigns every block
//     its place and the blocks are copied there concurrently.
//
// Rounding points that decide the partition (all reproduced):
//   centroid = (p1+p2+p3) * 0.3333f   -> fp64 product with the float constant widened   (bvh.cpp:112)
//   axis pick compares extent.z (double) with float(extent[axis])                       (bvh.cpp:61-65)
//   splitPos = float(min[axis]) + float(extent[axis]) * 0.5f   in fp32                  (bvh.cpp:67)
//   partition predicate float(centroid[axis]) < splitPos                                (bvh.cpp:73)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

#include "ct_scene.hpp"

namespace cth {
namespace {

inline double lo(double a, double b) { return (a < b) ? a : b; }   // mymath.h:15 macro semantics
inline double hi(double a, double b) { return (a > b) ? a : b; }   // mymath.h:11

// The builder works on copies of the triangles and centroids kept IN INDEX ORDER (entry i = triangle tri_index[i]):
// the partition swaps them together with the indices, so that the bounds of a range are a streaming read instead
// of a gather through tri_index (4x faster single-threaded on the 868k-triangle scene).  Same comparisons, same
// swaps, same result.
struct Tri9 { double v[9]; };        // a Triangle without constructors (the copies are written by worker threads)
struct Work {
    std::vector<Tri9, NoInitAlloc<Tri9>> tri;          // [n]
    std::vector<double, NoInitAlloc<double>> centroid; // [3 n]
    std::vector<uint32_t> *index;   // -> Scene::tri_index
    // scratch of the same shape for split_node_parallel (a node only ever uses its own index range of it)
    std::vector<Tri9, NoInitAlloc<Tri9>> tri2;
    std::vector<double, NoInitAlloc<double>> centroid2;
    std::vector<uint32_t, NoInitAlloc<uint32_t>> index2;
    std::vector<unsigned char, NoInitAlloc<unsigned char>> small;   // the partition predicate per entry
    uint32_t par_min = 0xffffffffu;  // nodes with at least this many triangles are partitioned by all threads
    int threads = 1;
};

// UpdateNodeBounds bvh.cpp:30-49 over index positions [first, first + count)
void fit_range(const Work &s, uint32_t first, uint32_t count, double mn[3], double mx[3]) {
    for (int a = 0; a < 3; a++) { mn[a] = (double)1e30f; mx[a] = (double)-1e30f; }
    for (uint32_t i = 0; i < count; i++) {
        const double *v = s.tri[first + i].v;                     // 9 packed doubles
        for (int p = 0; p < 3; p++)
            for (int a = 0; a < 3; a++) {
                mn[a] = lo(mn[a], v[3 * p + a]);
                mx[a] = hi(mx[a], v[3 * p + a]);
            }
    }
}

// The same bounds computed by `threads` workers over slices of the range: min / max of finite coordinates do not
// depend on the order they are taken in (a NaN coordinate would; then the slices are merged in index order, which
// is what the sequential loop does as well up to which NaN-free prefix wins -- scenes with NaN vertices are not
// promised to match).
void fit_range_parallel(const Work &s, uint32_t first, uint32_t count, double mn[3], double mx[3], int threads) {
    if (threads <= 1 || count < 200000) { fit_range(s, first, count, mn, mx); return; }
    std::vector<double> part((size_t)threads * 6);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++)
        th.emplace_back([&, t] {
            uint32_t a = first + (uint32_t)((uint64_t)count * t / threads), b = first + (uint32_t)((uint64_t)count * (t + 1) / threads);
            fit_range(s, a, b - a, &part[(size_t)t * 6], &part[(size_t)t * 6 + 3]);
        });
    for (auto &t : th) t.join();
    for (int a = 0; a < 3; a++) { mn[a] = (double)1e30f; mx[a] = (double)-1e30f; }
    for (int t = 0; t < threads; t++)
        for (int a = 0; a < 3; a++) { mn[a] = lo(mn[a], part[(size_t)t * 6 + a]); mx[a] = hi(mx[a], part[(size_t)t * 6 + 3 + a]); }
}

// Subdivide's decision for one node (bvh.cpp:51-99): partitions the node's index range in place and returns the size of
// the left part, or 0 when the node stays a leaf.
uint32_t split_node(Work &s, const ct_bvh_node &node) {
    uint32_t *index = s.index->data();
    double *centroid = s.centroid.data();
    if (node.triangle_count <= 2) return 0;                        // bvh.cpp:54
    double ext[3];
    for (int a = 0; a < 3; a++) ext[a] = node.aabb_max[a] - node.aabb_min[a];
    int axis = 0;
    if (ext[1] > ext[0]) axis = 1;
    if (ext[2] > (double)(float)ext[axis]) axis = 2;
    const float split = (float)node.aabb_min[axis] + (float)ext[axis] * 0.5f;
    int i = (int)node.first_triangle_index;
    int j = i + (int)node.triangle_count - 1;
    while (i <= j) {                                               // in-place partition, bvh.cpp:70-81
        if ((float)centroid[3 * (size_t)i + axis] < split) {
            i++;
        } else {
            std::swap(index[i], index[j]);
            std::swap(s.tri[i], s.tri[j]);
            for (int a = 0; a < 3; a++) std::swap(centroid[3 * (size_t)i + a], centroid[3 * (size_t)j + a]);
            j--;
        }
    }
    const uint32_t left_count = (uint32_t)i - node.first_triangle_index;
    if (left_count == 0 || left_count == node.triangle_count) return 0;   // bvh.cpp:84-86
    return left_count;
}

// fn(slice, begin, end) over [0, n) in `nt` contiguous slices, one thread each
template <typename F>
void for_slices(uint32_t n, int nt, F fn) {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back([=] { fn(t, (uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt)); });
    fn(0, 0u, (uint32_t)((uint64_t)n / nt));
    for (auto &t : th) t.join();
}

// split_node for a big node, by all threads, with the SAME resulting order.  The reference's loop
// (`if small(a[i]) i++ else swap(a[i], a[j--])`, bvh.cpp:70-81) sends every entry to a place that prefix counts
// determine (derivation and an exhaustive check against the loop: tools/partition_closed_form.py).  With k smalls:
//   left part [0, k): a small stays; the m-th large found there is replaced by the m-th small of the right part counted
//     from the end;
//   right part, written from the end backwards: for m = 0, 1, ...: the m-th left-part large, then the right-part larges
//     between the (m-1)-th and the m-th of those smalls;
//   the untouched middle [k, hi] (all large) ends up rotated left by one.
// Note that the loop permutes the range even when it then declines to split (left part empty or everything): the order
// inside the resulting leaf is the order its triangles are tested in, so that case is reproduced as well.
uint32_t split_node_parallel(Work &s, const ct_bvh_node &node) {
    const uint32_t first = node.first_triangle_index, n = node.triangle_count;
    double ext[3];
    for (int a = 0; a < 3; a++) ext[a] = node.aabb_max[a] - node.aabb_min[a];
    int axis = 0;
    if (ext[1] > ext[0]) axis = 1;
    if (ext[2] > (double)(float)ext[axis]) axis = 2;
    const float split = (float)node.aabb_min[axis] + (float)ext[axis] * 0.5f;
    uint32_t *index = s.index->data() + first, *index2 = s.index2.data() + first;
    Tri9 *tri = s.tri.data() + first, *tri2 = s.tri2.data() + first;
    double *cen = s.centroid.data() + 3 * (size_t)first, *cen2 = s.centroid2.data() + 3 * (size_t)first;
    unsigned char *small = s.small.data() + first;
    const int nt = s.threads;
    std::vector<uint32_t> smalls_in(nt + 1, 0), left_larges_in(nt, 0);
    for_slices(n, nt, [&](int t, uint32_t a, uint32_t b) {
        uint32_t c = 0;
        for (uint32_t p = a; p < b; p++) { small[p] = (float)cen[3 * (size_t)p + axis] < split; c += small[p]; }
        smalls_in[t + 1] = c;
    });
    for (int t = 0; t < nt; t++) smalls_in[t + 1] += smalls_in[t];            // -> smalls before slice t
    const uint32_t k = smalls_in[nt], n_large = n - k;
    const uint32_t pairs_max = std::min(k, n_large);
    std::vector<uint32_t> ll_pos(pairs_max + 1), rs_pos(pairs_max + 1), before(pairs_max + 2);
    before[0] = 0;
    for_slices(n, nt, [&](int t, uint32_t a, uint32_t b) {                      // who pairs with whom
        uint32_t sm = smalls_in[t], ll = 0;                                     // smalls in [0, p)
        for (uint32_t p = a; p < b; p++) {
            if (small[p]) {
                if (p >= k) {
                    const uint32_t r = k - (sm + 1);                            // smalls above p = its rank from the end
                    rs_pos[r] = p;
                    before[r + 1] = (r + 1) + (n_large - (p + 1 - (sm + 1)));   // entries written to the right part before left large r + 1
                }
                sm++;
            } else if (p < k) {
                ll_pos[p - sm] = p;                                             // larges in [0, p) = its rank among the left larges
                ll++;
            }
        }
        left_larges_in[t] = ll;
    });
    uint32_t n_pairs = 0;
    for (int t = 0; t < nt; t++) n_pairs += left_larges_in[t];
    const uint32_t hi = n_pairs ? rs_pos[n_pairs - 1] - 1 : n - 1;              // last entry of the untouched middle
    for_slices(n, nt, [&](int t, uint32_t a, uint32_t b) {                      // scatter
        uint32_t sm = smalls_in[t];
        for (uint32_t p = a; p < b; p++) {
            uint32_t dest;
            if (small[p]) {
                dest = p < k ? p : ll_pos[k - (sm + 1)];
                sm++;
            } else if (p < k) {
                dest = n - 1 - before[p - sm];
            } else {
                const uint32_t above = k - sm;                                  // smalls above p
                if (above < n_pairs) dest = n - 1 - (above + 1) - (n_large - (p + 1 - sm));
                else dest = p > k ? p - 1 : hi;                                 // the middle, rotated left by one
            }
            index2[dest] = index[p];
            tri2[dest] = tri[p];
            for (int c = 0; c < 3; c++) cen2[3 * (size_t)dest + c] = cen[3 * (size_t)p + c];
        }
    });
    for_slices(n, nt, [&](int, uint32_t a, uint32_t b) {
        memcpy(index + a, index2 + a, (size_t)(b - a) * sizeof(uint32_t));
        memcpy(tri + a, tri2 + a, (size_t)(b - a) * sizeof(Tri9));
        memcpy(cen + 3 * (size_t)a, cen2 + 3 * (size_t)a, (size_t)(b - a) * 3 * sizeof(double));
    });
    if (k == 0 || k == n) return 0;                                             // bvh.cpp:84-86
    return k;
}

// Builds the subtree below nodes[0] (bounds, range and count already set) with block-local numbering: the children
// of the first split are nodes 1 and 2, and so on in the reference's allocation order.
void build_block(Work &s, std::vector<ct_bvh_node> &nodes) {
    std::vector<uint32_t> work;                                    // pre-order: pop, split, push right then left
    work.push_back(0);
    while (!work.empty()) {
        const uint32_t idx = work.back();
        work.pop_back();
        const uint32_t left_count = split_node(s, nodes[idx]);
        if (left_count == 0) continue;
        const uint32_t l = (uint32_t)nodes.size();
        nodes.push_back(ct_bvh_node{});
        nodes.push_back(ct_bvh_node{});
        ct_bvh_node &node = nodes[idx];
        node.left_node = l;
        nodes[l].first_triangle_index = node.first_triangle_index;
        nodes[l].triangle_count = left_count;
        nodes[l + 1].first_triangle_index = node.first_triangle_index + left_count;
        nodes[l + 1].triangle_count = node.triangle_count - left_count;
        node.triangle_count = 0;
        fit_range(s, nodes[l].first_triangle_index, nodes[l].triangle_count, nodes[l].aabb_min, nodes[l].aabb_max);
        fit_range(s, nodes[l + 1].first_triangle_index, nodes[l + 1].triangle_count, nodes[l + 1].aabb_min, nodes[l + 1].aabb_max);
        work.push_back(l + 1);
        work.push_back(l);
    }
}

// fn(begin, end) over [0, n) on up to `threads` workers (big, independent slices only)
template <typename F>
void run_slices(uint32_t n, int threads, F fn) {
    const int nt = (int)std::min<uint64_t>((uint64_t)std::max(threads, 1), (n + 65535u) / 65536u);
    if (nt <= 1) { fn(0u, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([=] { fn((uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt)); });
    for (auto &t : th) t.join();
}

int build_threads() {
    if (const char *e = std::getenv("CT_HOST_THREADS")) { int v = std::atoi(e); if (v >= 1) return std::min(v, 64); }
    unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(std::max(hw, 1u), 32u);
}

}  // namespace

void build_bvh(Scene &s) {
    const uint32_t n = (uint32_t)s.tris.size();
    if (n == 0) throw std::runtime_error("cannot build a BVH over zero triangles");
    const int threads = build_threads();
    const bool timing = std::getenv("CT_HOST_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    s.tri_index.resize(n);
    Work w;
    w.tri.resize(n);
    w.centroid.resize((size_t)3 * n);
    w.index = &s.tri_index;
    const double third = (double)0.3333f;
    run_slices(n, threads, [&](uint32_t k0, uint32_t k1) {
        for (uint32_t k = k0; k < k1; k++) {
            s.tri_index[k] = k;
            const Triangle &t = s.tris[k];
            memcpy(w.tri[k].v, &t.p1.x, sizeof(Tri9));
            w.centroid[3 * (size_t)k + 0] = third * ((t.p1.x + t.p2.x) + t.p3.x);
            w.centroid[3 * (size_t)k + 1] = third * ((t.p1.y + t.p2.y) + t.p3.y);
            w.centroid[3 * (size_t)k + 2] = third * ((t.p1.z + t.p2.z) + t.p3.z);
        }
    });
    w.threads = threads;
    if (threads > 1) {
        // Opt-in (CT_HOST_PAR_PARTITION_MIN = smallest node partitioned by all threads): measured SLOWER than the one-thread
        // loop on the 16-core box (tree 62 ms against 50 ms for 868k triangles) -- the loop touches every entry once, the
        // scatter needs the predicate pass, the pairing pass, the out-of-place write and the copy back, and the pool
        // already runs the other subtrees meanwhile.  Kept, and tested, because it is the form a device-side build needs.
        if (const char *e = std::getenv("CT_HOST_PAR_PARTITION_MIN")) { long v = std::atol(e); if (v >= 3) w.par_min = (uint32_t)v; }
        if (n >= w.par_min) { w.tri2.resize(n); w.centroid2.resize((size_t)3 * n); w.index2.resize(n); w.small.resize(n); }
        else w.par_min = 0xffffffffu;
    }
    const double t_prologue = now();

    // ---- top of the tree and the blocks below it, as one pool of tasks.  A task = one node: at or below `grain`
    // triangles it becomes a block (its whole subtree is built on the spot, block-local numbering); above, it is split
    // (the order-dependent partition runs on this thread, disjoint ranges run concurrently) and its children become
    // tasks.  The critical path is the root's partition plus half of it, a quarter, ...: about twice the root instead
    // of once per level.  top[i].left_node indexes `top` (temporary numbering); is_block marks frontier nodes.
    const uint32_t grain = (threads <= 1 || n < 16384) ? n : std::max<uint32_t>(4096, n / (uint32_t)(threads * 3));
    std::deque<ct_bvh_node> top(1);                                // deques: references stay valid while other tasks append
    std::deque<char> is_block(1, 0);
    std::deque<uint32_t> block_of(1, 0xffffffffu);
    std::deque<std::vector<ct_bvh_node>> blocks;
    top[0].first_triangle_index = 0; top[0].triangle_count = n;
    fit_range_parallel(w, 0, n, top[0].aabb_min, top[0].aabb_max, threads);
    {
        std::mutex mu;
        std::condition_variable cv;
        std::vector<uint32_t> ready(1, 0u);
        int running = 0;
        auto process = [&](uint32_t idx) {
            ct_bvh_node node;
            { std::lock_guard<std::mutex> g(mu); node = top[idx]; }
            if (node.triangle_count <= grain) {
                std::vector<ct_bvh_node> *blk;
                {
                    std::lock_guard<std::mutex> g(mu);
                    is_block[idx] = 1; block_of[idx] = (uint32_t)blocks.size();
                    blocks.emplace_back(1, node);
                    blk = &blocks.back();
                }
                blk->reserve((size_t)2 * node.triangle_count);
                build_block(w, *blk);
                return;
            }
            const uint32_t left_count = node.triangle_count >= w.par_min ? split_node_parallel(w, node) : split_node(w, node);
            if (left_count == 0) return;                            // a big leaf (unsplittable): stays in the top part
            ct_bvh_node l{}, r{};
            l.first_triangle_index = node.first_triangle_index; l.triangle_count = left_count;
            r.first_triangle_index = node.first_triangle_index + left_count; r.triangle_count = node.triangle_count - left_count;
            // big children: all hands on the bounds while there are not yet enough tasks to go round
            const int fit_threads = node.triangle_count >= n / 4 ? threads : 1;
            fit_range_parallel(w, l.first_triangle_index, l.triangle_count, l.aabb_min, l.aabb_max, fit_threads);
            fit_range_parallel(w, r.first_triangle_index, r.triangle_count, r.aabb_min, r.aabb_max, fit_threads);
            std::lock_guard<std::mutex> g(mu);
            const uint32_t li = (uint32_t)top.size();
            top.push_back(l); top.push_back(r);
            is_block.push_back(0); is_block.push_back(0);
            block_of.push_back(0xffffffffu); block_of.push_back(0xffffffffu);
            top[idx].left_node = li;
            top[idx].triangle_count = 0;
            ready.push_back(li + 1);
            ready.push_back(li);
            cv.notify_all();
        };
        auto worker = [&] {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                cv.wait(lk, [&] { return !ready.empty() || running == 0; });
                if (ready.empty()) { cv.notify_all(); return; }    // nothing queued and nobody left to queue more
                const uint32_t idx = ready.back();
                ready.pop_back();
                running++;
                lk.unlock();
                process(idx);
                lk.lock();
                running--;
                if (ready.empty() && running == 0) cv.notify_all();
            }
        };
        if (threads <= 1) worker();
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < threads; t++) th.emplace_back(worker);
            for (auto &t : th) t.join();
        }
    }
    const double t_top = now();

    // ---- final numbering: the reference's allocation order (pre-order over splits, left before right)
    size_t total = top.size();
    for (uint32_t i = 0; i < top.size(); i++) if (is_block[i]) total += blocks[block_of[i]].size() - 1;
    s.nodes.resize(total);                                         // not filled: every entry is written below
    uint32_t used = 1;
    struct Item { uint32_t top_idx, final_idx; };
    struct Placed { uint32_t block, root_idx, base; };
    std::vector<Placed> placed;                                    // where every block goes; copied concurrently below
    std::vector<Item> work(1, Item{0u, 0u});
    while (!work.empty()) {
        const Item it = work.back();
        work.pop_back();
        if (is_block[it.top_idx]) {
            // the block's root sits at final_idx (allocated by its parent); its other nodes take the next indices in
            // block order, which IS the reference's order because the recursion finishes a subtree before leaving it
            const uint32_t b = block_of[it.top_idx];
            placed.push_back(Placed{b, it.final_idx, used - 1});   // block-local index k >= 1 -> final index base + k
            used += (uint32_t)blocks[b].size() - 1;
            continue;
        }
        s.nodes[it.final_idx] = top[it.top_idx];
        if (top[it.top_idx].triangle_count != 0) continue;          // leaf inside the top part
        const uint32_t l = used; used += 2;
        s.nodes[it.final_idx].left_node = l;
        work.push_back(Item{top[it.top_idx].left_node + 1, l + 1});
        work.push_back(Item{top[it.top_idx].left_node, l});
    }
    if (used != total) throw std::runtime_error("BVH numbering is inconsistent (internal error)");
    {
        std::atomic<uint32_t> next{0};
        auto copier = [&] {
            for (uint32_t i; (i = next.fetch_add(1)) < placed.size();) {
                const auto &blk = blocks[placed[i].block];
                const uint32_t base = placed[i].base;
                s.nodes[placed[i].root_idx] = blk[0];
                if (blk[0].triangle_count == 0) s.nodes[placed[i].root_idx].left_node = base + blk[0].left_node;
                for (uint32_t k = 1; k < blk.size(); k++) {
                    s.nodes[base + k] = blk[k];
                    if (blk[k].triangle_count == 0) s.nodes[base + k].left_node = base + blk[k].left_node;
                }
            }
        };
        const int nt = (int)std::min<size_t>((size_t)threads, placed.size());
        if (nt <= 1) copier();
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) th.emplace_back(copier);
            for (auto &t : th) t.join();
        }
    }
    if (timing)
        fprintf(stderr, "build_bvh: %d threads, %zu top nodes, %zu blocks: working copies %.1f ms, tree %.1f ms, numbering %.1f ms\n", threads,
                top.size(), blocks.size(), t_prologue - t_start, t_top - t_prologue, now() - t_top);
}

}  // namespace cth
