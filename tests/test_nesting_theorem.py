"""The argument the early-exit walks rest on (box_maybe in ct_traverse.cuh), checked on the CPU with the reference's
own slab arithmetic: in a tree built by the reference's BuildBVH every child's box lies inside its parent's, the floats
IntersectAABB compares are monotone in the box, and therefore -- for a ray without NaN quotients and a fixed ray.t --

    the reference's walk reaches a node   <=>   the node's OWN box passes IntersectAABB.

The slab test is restated in numpy (fp64 quotients rounded to fp32, the reference's min / max macros; bvh.cpp:165-179),
pinned against oracle/ct_oracle.c on a sample, and then evaluated for EVERY node of the bundled scenes' reference-built trees
and a few hundred rays of each kind the walks serve: shadow rays leaving a surface (ray.t = 1e30) and the degenerate
reflection rays (ray.t = 0)."""
import ctypes as C

import numpy as np
import pytest

from oracle import ct_oracle_py as O
from conftest import load_golden_scene


def slab_accept(o, d, ray_t, bmin, bmax):
    """IntersectAABB for one ray against many boxes ([n, 3] each), in the reference's arithmetic."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t1 = ((bmin - o) / d).astype(np.float32)
        t2 = ((bmax - o) / d).astype(np.float32)
    mn = lambda a, b: np.where(a < b, a, b)          # mymath.h:11-17: a NaN operand yields the SECOND operand
    mx = lambda a, b: np.where(a > b, a, b)
    tmin, tmax = mn(t1[:, 0], t2[:, 0]), mx(t1[:, 0], t2[:, 0])
    for k in (1, 2):
        tmin = mx(tmin, mn(t1[:, k], t2[:, k]))
        tmax = mn(tmax, mx(t1[:, k], t2[:, k]))
    return (tmax >= tmin) & (tmin < np.float32(ray_t)) & (tmax > np.float32(0))


def tree_arrays(fs):
    n = fs.n_nodes
    parent = np.full(n, -1, np.int64)
    interior = np.flatnonzero(fs.node_count == 0)
    parent[fs.node_left[interior]] = interior
    parent[fs.node_left[interior] + 1] = interior
    depth = np.zeros(n, np.int64)
    order = [np.array([0])]
    while True:                                     # nodes level by level (children of the previous level's interior nodes)
        cur = order[-1]
        cur = cur[fs.node_count[cur] == 0]
        if cur.size == 0:
            break
        nxt = np.concatenate([fs.node_left[cur], fs.node_left[cur] + 1]).astype(np.int64)
        depth[nxt] = depth[cur[0]] + 1
        order.append(nxt)
    return parent, order


@pytest.mark.parametrize("name", ["scene_import_bunny", "scene_import", "pc_big", "scene_file_cube"])
def test_a_node_is_reached_exactly_when_its_own_box_passes(name, golden):
    fs = load_golden_scene(name, golden)
    parent, levels = tree_arrays(fs)
    # 1. nesting, bit for bit
    kids = np.flatnonzero(parent >= 0)
    assert (fs.node_min[kids] >= fs.node_min[parent[kids]]).all() and (fs.node_max[kids] <= fs.node_max[parent[kids]]).all()
    # 2. the numpy restatement of the slab test == oracle/ct_oracle.c on a sample
    rng = np.random.default_rng(7)
    L = O.lib()
    p = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    tri = fs.tri.reshape(-1, 3, 3)
    lo, hi = fs.node_min[0], fs.node_max[0]
    for _ in range(200):
        o = rng.uniform(lo - 1, hi + 1); d = rng.normal(size=3)
        k = rng.integers(0, fs.n_nodes, 8)
        t = float(rng.choice([1e30, 0.0, 3.0]))
        want = [L.ct_oracle_intersect_aabb(p(o), p(d), C.c_float(t), p(fs.node_min[i]), p(fs.node_max[i])) for i in k]
        assert slab_accept(o, d, t, fs.node_min[k], fs.node_max[k]).astype(int).tolist() == want
    # 3. the theorem, on every node, for the rays the early-exit walks serve
    n_rays = 120
    for i in range(n_rays):
        a = tri[rng.integers(0, tri.shape[0])]
        w = rng.dirichlet([1, 1, 1])
        origin = w @ a                                                   # a point on a triangle (up to rounding): a shading point
        if i % 3 == 0:
            d, ray_t = np.array([20.0, 110.0, -300.0]) - origin, 1e30     # shadow ray towards a point light
        elif i % 3 == 1:
            d, ray_t = rng.normal(size=3), 0.0                            # a degenerate reflection ray (t = 0, raythread.cpp:373)
        else:
            d, ray_t = rng.normal(size=3) * rng.choice([1e-3, 1.0, 50.0]), 1e30
        assert (d != 0).all()
        own = slab_accept(origin, d, ray_t, fs.node_min, fs.node_max)
        reached = np.zeros(fs.n_nodes, bool)
        reached[0] = own[0]
        for lvl in levels[1:]:                                           # the walk: a node is visited iff its parent was accepted
            reached[lvl] = reached[parent[lvl]] & own[lvl]
        assert np.array_equal(reached, own), (name, i, int((reached != own).sum()))
        assert own[kids].sum() == 0 or (own[parent[kids[own[kids]]]]).all()   # (same statement: an accepted node's parent is accepted)
