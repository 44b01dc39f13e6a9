"""Parity tests proper: the CUDA path, called through the C ABI (include/ct_gpu.h), against
  * golden vectors produced by the unmodified reference (tests/golden/),
  * the oracle on the same inputs at sizes it finishes in seconds,
  * size-independent properties at BASELINE's full sizes (tile-split invariance, determinism).
BASELINE.json's bar is: primary hit index equal on >= 99.99 % of rays; RGB within 1 LSB on >= 99.9 % of pixels and
max abs error <= 4/255 elsewhere.  This implementation reproduces the reference's arithmetic exactly, so the
tests assert the stronger property: bit-identical frames, indices and distances."""
import dataclasses
import os

import numpy as np
import pytest

import cobbletrace_b200 as ct
from cobbletrace_b200 import host, procedural
from oracle import ct_oracle_py as O
from conftest import GOLD, load_frames, case_scene

pytestmark = pytest.mark.gpu

CASES = ["cube_160", "cube_rot_160", "cube_wide_200x120", "cube_tall_90x150", "import_160", "pc_big_96", "bunny_160",
         "bunny_refl_d2_160", "bunny_refl_d10_128", "bunny_odd_161x161"]
DBG = ct.CT_FLAG_KEEP_HITS | ct.CT_FLAG_COUNT_TESTS


def north_star_metrics(frame, ref_frame, found, index, ref_hits):
    """The BASELINE.json acceptance numbers (reported in assertion messages)."""
    traced = ref_hits["found"] != 0xFFFFFFFF
    idx_ok = ((found == ref_hits["found"]) & ((index == ref_hits["index"]) | (ref_hits["found"] != 1)))[traced].mean()
    ch = lambda f, s: ((f >> s) & 0xFF).astype(np.int32)
    err = np.maximum.reduce([np.abs(ch(frame, s) - ch(ref_frame, s)) for s in (0, 8, 16)])
    return float(idx_ok), float((err <= 1).mean()), int(err.max())


def assert_same(r, oframe, ohits, what):
    frame = r.readback()
    found, index, t = r.readback_hits()
    idx_ok, lsb_ok, max_err = north_star_metrics(frame, oframe, found, index, ohits)
    msg = f"{what}: hit-index match {idx_ok:.6f}, <=1LSB {lsb_ok:.6f}, max err {max_err}"
    assert idx_ok >= 0.9999 and lsb_ok >= 0.999 and max_err <= 4, msg          # BASELINE.json bar
    traced = ohits["found"] != 0xFFFFFFFF
    assert np.array_equal(frame, oframe), msg                                    # ... and the bar this repo holds itself to
    assert np.array_equal(found, ohits["found"]), msg
    assert np.array_equal(index[traced], ohits["index"][traced]), msg
    assert np.array_equal(t[traced].view(np.uint32), ohits["t"][traced].view(np.uint32)), msg
    return frame


@pytest.fixture
def gpu():
    r = ct.GpuRenderer(0)
    yield r
    r.shutdown()


def test_loaded_library_is_the_in_tree_cuda_build():
    L = ct.load_library()
    assert os.path.samefile(ct.api.GPU_LIB, os.path.join(os.path.dirname(ct.__file__), "libct_gpu.so"))
    assert L.ct_gpu_device_count() >= 1


def test_primitive_kats_against_reference(gpu):
    zf = np.load(os.path.join(GOLD, "kat_primitives.npz"))
    z = {k: zf[k] for k in zf.files}
    th, bh, t_out = gpu.debug_primitives(z["org"], z["dir"], z["t0"], z["tri"], z["bmin"], z["bmax"])
    assert np.array_equal(bh, z["box_hit"]), f"{int((bh != z['box_hit']).sum())} box verdicts differ"
    assert np.array_equal(th, z["tri_hit"]), f"{int((th != z['tri_hit']).sum())} triangle verdicts differ"
    assert np.array_equal(t_out.view(np.uint32), z["t_out"].view(np.uint32))


def _filter_cases(rng, n):
    """(origins, directions, ray_t, bmin, bmax): generic boxes, then adversarial ones whose slab quotients tie."""
    o = rng.normal(size=(n, 3)) * rng.choice([0.1, 5.0, 60.0, 4e9], (n, 1))
    d = rng.normal(size=(n, 3)) * rng.choice([1e-3, 1.0, 300.0, 4e9], (n, 1))
    d[rng.random((n, 3)) < 0.03] = 0.0                               # the x = 0 pixel column has dir.x == 0
    c = o + d * rng.uniform(-1, 3, (n, 1)) + rng.normal(size=(n, 3)) * rng.choice([0.0, 1e-9, 0.05, 10.0], (n, 1))
    half = np.abs(rng.normal(size=(n, 3))) * rng.choice([1e-6, 0.02, 1.0, 30.0], (n, 1))
    half[rng.random((n, 3)) < 0.15] = 0.0                            # zero-thickness boxes of axis-aligned triangles
    bmin, bmax = c - half, c + half
    # ties: every axis enters at exactly t1 and leaves at exactly t2 (corner to corner), then nudged by a few ulps
    k = n // 3
    t1 = rng.uniform(-2, 5, (k, 1)); t2 = t1 + np.abs(rng.normal(size=(k, 1))) * rng.choice([0.0, 1e-7, 1.0], (k, 1))
    pa, pb = o[:k] + t1 * d[:k], o[:k] + t2 * d[:k]
    lo, hi = np.minimum(pa, pb), np.maximum(pa, pb)
    for arr in (lo, hi):
        steps = rng.integers(-3, 4, arr.shape)
        for _ in range(3):
            arr[:] = np.where(steps > 0, np.nextafter(arr, np.inf), np.where(steps < 0, np.nextafter(arr, -np.inf), arr))
            steps = steps - np.sign(steps)
    hi = np.maximum(lo, hi)
    bmin[:k], bmax[:k] = lo, hi
    rt = rng.choice(np.array([1e30, 1e30, 0.0, 1.0, 3.5], np.float32), n)
    rt[:k:2] = (t1[::2, 0] * rng.choice([1.0, 1.0 + 1e-7, 1.0 - 1e-7], t1[::2, 0].shape)).astype(np.float32)   # ray.t next to tmin
    return o, d, rt, bmin, bmax


@pytest.mark.parametrize("bound_scale", [1.0, 64.0])
def test_slab_filter_is_sound_and_mostly_decides(gpu, bound_scale):
    """The certified fp32 filter in front of IntersectAABB may only ever answer what the reference's
    arithmetic answers; undecided cases fall back to that arithmetic.  1M generic + adversarial (ray, box) cases."""
    rng = np.random.default_rng(11)
    o, d, rt, bmin, bmax = _filter_cases(rng, 1 << 20)
    v = gpu.debug_filter(o, d, rt, bmin, bmax, bound_scale)
    exact, filt, usable, escaped, recip_bad = v & 1, (v >> 1) & 3, (v >> 3) & 1, (v >> 4) & 1, (v >> 5) & 1
    assert not escaped.any(), f"{int(escaped.sum())} brackets do not contain the reference's tmin/tmax"
    assert not recip_bad.any(), f"{int(recip_bad.sum())} division-free slab evaluations differ from the literal arithmetic"
    assert not ((filt == 1) & (exact == 0)).any() and not ((filt == 2) & (exact == 1)).any()
    assert (filt[usable == 0] == 0).all()
    zero_dir = (d == 0).any(1)
    assert (usable[zero_dir] == 0).all()                              # +-inf / NaN quotients: reference arithmetic only
    generic = (np.arange(len(v)) >= len(v) // 3) & ((bmax - bmin) > 0).all(1)     # ordinary boxes with some thickness
    decided = (filt[generic & (usable == 1)] != 0).mean()
    assert decided > 0.5, decided                                     # (real scenes: > 99.8 %, see tools/perf_stages.py)
    assert ((filt == 1).sum() > 1000) and ((filt == 2).sum() > 1000)


def test_triangle_filter_is_sound_and_mostly_decides(gpu):
    """The certified fp32 filter in front of IntersectTriangle may only discard a triangle whose reference test has no
    effect: 1M (ray, triangle) cases -- rays aimed at / just past triangle edges and vertices, rays starting ON the
    triangle (what a shadow or reflection ray does), slivers, tiny and huge triangles, far-away origins."""
    rng = np.random.default_rng(5)
    n = 1 << 20
    scale = rng.choice([1e-3, 0.03, 1.0, 40.0], (n, 1))
    p1 = rng.normal(size=(n, 3)) * rng.choice([0.0, 1.0, 30.0], (n, 1))
    e1, e2 = rng.normal(size=(n, 3)) * scale, rng.normal(size=(n, 3)) * scale
    sliver = rng.random(n) < 0.1
    e2[sliver] = e1[sliver] * rng.uniform(0.5, 2, (int(sliver.sum()), 1)) + rng.normal(size=(int(sliver.sum()), 3)) * scale[sliver] * 1e-6
    tri = np.concatenate([p1, p1 + e1, p1 + e2], 1)
    # target point: barycentric (u, v) mostly near the boundary of the triangle
    u = rng.choice([0.0, 1.0, 0.3, -1e-7, 1 + 1e-7, 0.5], n) + rng.normal(size=n) * rng.choice([0.0, 1e-9, 1e-4, 0.5], n)
    v = rng.choice([0.0, 0.7, 0.3, -1e-7, 0.5], n) + rng.normal(size=n) * rng.choice([0.0, 1e-9, 1e-4, 0.5], n)
    target = p1 + u[:, None] * e1 + v[:, None] * e2
    d = rng.normal(size=(n, 3)) * rng.choice([1e-2, 1.0, 300.0], (n, 1))
    tdist = rng.choice([0.0, 1e-5, 1e-4, 1.0, 50.0, -3.0], n) * rng.choice([1.0, 1 + 1e-7], n)
    o = target - tdist[:, None] * d                               # tdist == 0: the ray starts on the triangle's plane
    far = rng.random(n) < 0.02
    o[far] *= 1e9
    rt = np.full(n, 1e30, np.float32)
    box = np.concatenate([np.minimum.reduce([p1, p1 + e1, p1 + e2]), np.maximum.reduce([p1, p1 + e1, p1 + e2])], 1)
    vv = gpu.debug_filter(o, d, rt, box[:, :3], box[:, 3:], 1.0, tri=tri)
    hit, occl, m_closest, m_shadow, usable = (vv >> 8) & 1, (vv >> 9) & 1, (vv >> 10) & 1, (vv >> 11) & 1, (vv >> 12) & 1
    assert not (m_closest & hit).any(), f"{int((m_closest & hit).sum())} triangles discarded although the reference's test returns true"
    assert not (m_shadow & occl).any(), f"{int((m_shadow & occl).sum())} occluders discarded"
    assert (m_closest <= m_shadow).all()                          # the shadow variant only ever discards more
    assert usable.mean() > 0.9 and hit.sum() > 10000 and occl.sum() > 1000
    miss = (hit == 0) & (usable == 1)
    assert m_closest[miss].mean() > 0.5, m_closest[miss].mean()   # (real scenes: ~90 %, tools/perf_stages.py)
    assert (m_shadow & hit & (1 - occl)).sum() > 1000             # t <= 1e-4 hits (rays leaving the triangle) are discarded for shadow rays


@pytest.mark.parametrize("case", CASES)
def test_golden_frames_from_reference(case, golden, scene_loader, gpu):
    fs, meta = case_scene(case, golden, scene_loader)
    g = load_frames(case)
    gpu.upload(fs, meta["width"], meta["height"], max_depth=meta["depth"], flags=DBG)
    gpu.render_tile()
    hits = np.zeros(g["frame"].shape, O.HIT_DT)
    hits["found"], hits["index"], hits["t"] = g["found"], g["index"], g["t"]
    assert_same(gpu, g["frame"], hits, case)


@pytest.mark.parametrize("name,W,H,depth,refl", [
    ("scene_file_cube", 640, 640, 10, None),          # BASELINE config 1a: mirror triangle, 348k degenerate reflection rays
    ("scene_import", 640, 640, 10, None),             # config 1b: leaves of up to 54 triangles
    ("scene_import_bunny", 1920, 1080, 10, None),     # config 2
    ("pc_big", 640, 640, 10, None),                   # config 4 scene (66 lights, any-hit bound)
    ("pc_big", 1920, 1080, 10, None),                 # config 4 at BASELINE's size
    ("scene_import_bunny", 1000, 700, 2, 0.5),        # config-3-like: forced reflection, depth 2
    ("scene_file_cube", 333, 517, 10, 0.9),           # W < H, odd sizes, everything reflective to full depth
    ("scene_file_cube", 401, 301, 10, 0.9),           # odd H, W > H: no pixel is dropped, all counters comparable
])
def test_frames_against_oracle(name, W, H, depth, refl, golden, scene_loader, gpu):
    fs = scene_loader(name)
    if refl is not None:
        fs = fs.with_reflection(refl)
    oframe, ohits, octr = O.OracleScene(fs).render(W, H, max_depth=depth)
    gpu.upload(fs, W, H, max_depth=depth, flags=DBG)
    ctr = gpu.render_tile(counters=True)
    assert_same(gpu, oframe, ohits, f"{name} {W}x{H}")
    if name == "scene_file_cube" and refl is None and W == 640:
        assert ct.frame_fnv1a(gpu.readback()) == golden["scenes"][name]["frame640_fnv1a"] == "a52ca3e236c27312"
    # Ray accounting: the GPU does not trace pixels PutPixel would drop (draw2d.h:11); the oracle does.
    stored = int((ohits["found"] != 0xFFFFFFFF).sum())
    assert ctr["rays_primary"] == stored
    if H % 2 == 1 and W >= H:      # odd H and wide enough: nothing is dropped, every counter must agree
        assert ctr["rays_shadow"] == octr["rays_shadow"] and ctr["rays_reflection"] == octr["rays_reflection"]
    # shadow rays are traced any-hit (fewer triangle tests than the reference's closest-hit walk) with their leaves
    # deferred (no early exit from the box walk), so only the order of magnitude is comparable
    assert 0 < ctr["box_tests"] <= 2 * octr["box_tests"] and 0 < ctr["tri_tests"] <= octr["tri_tests"]


@pytest.mark.parametrize("budget,warp_budget,primary_budget,split", [(16, 0, 0, 0), (16, 4, 1, 0), (200, 64, 6, 0), (0, 0, 1, 0), (0, 0, -1, 0),
                                                                     (0, 0, 1, 1), (200, 64, 6, 1), (0, 0, 40, 1)])
def test_parked_rays_give_the_same_frames(budget, warp_budget, primary_budget, split, golden, scene_loader):
    """Shadow / reflection rays whose walk exceeds the budget are parked and finished by a warp (pass 1) or by the
    whole grid (pass 2) in k_overflow; with tiny budgets nearly every ray takes those paths (and the parking
    buffer overflows, so some are finished in place).  Primary rays whose walk exceeds "primary_budget" pair visits are
    parked and walked again by k_primary_long (1: every ray that gets past the root; -1: none) -- or, with "primary_split", finished
    one ray per warp by the order-free frontier search of k_primary_split.  Frames (and hit records) must not change."""
    ct.api.set_option("primary_split", split)
    ct.api.set_option("traversal_budget", budget)
    ct.api.set_option("overflow_warp_budget", warp_budget)
    ct.api.set_option("primary_budget", primary_budget)
    try:
        for case in ("bunny_refl_d2_160", "cube_160", "pc_big_96", "import_160"):
            fs, meta = case_scene(case, golden, scene_loader)
            r = ct.GpuRenderer(0).upload(fs, meta["width"], meta["height"], max_depth=meta["depth"], flags=ct.CT_FLAG_KEEP_HITS)
            r.render_tile()
            frame = r.readback()
            found, index, t = r.readback_hits()
            parked, in_place = r.overflow_stats()
            r.shutdown()
            gold = load_frames(case)
            assert np.array_equal(frame, gold["frame"]), case
            if "found" in gold:                                             # the primary hit records of the compiled reference
                traced = gold["found"] != 0xFFFFFFFF
                assert np.array_equal(found[traced], gold["found"][traced]) and np.array_equal(index[traced], gold["index"][traced]), case
                assert np.array_equal(t[traced].view(np.uint32), gold["t"][traced].view(np.uint32)), case
            if budget == 16 and case != "cube_160":
                assert parked > 1000, (case, parked)
            if primary_budget == 1 and budget == 0 and case != "cube_160":
                assert parked > 1000, (case, parked)
    finally:
        ct.api.set_option("traversal_budget", 0)
        ct.api.set_option("overflow_warp_budget", 0)
        ct.api.set_option("primary_budget", 0)
        ct.api.set_option("primary_split", 0)


def test_shared_frame_mode_on_one_gpu(golden, scene_loader, gpu):
    """ct_gpu_render_shared (the multi-GPU path: cursor and output framebuffer reached through the share handle)
    with this GPU as its own root: same frame as ct_gpu_render_tile; chunks are handed out exactly once per reset."""
    fs, meta = case_scene("bunny_refl_d2_160", golden, scene_loader)
    want = load_frames("bunny_refl_d2_160")["frame"]
    gpu.upload(fs, meta["width"], meta["height"], max_depth=meta["depth"])
    gpu.share_attach(gpu.share_export())                        # self-attach: same process, same device
    c1 = gpu.render_shared(counters=True)
    assert np.array_equal(gpu.readback(), want)
    ct.api.set_option("shared_hold_frame", 1)
    try:
        c2 = gpu.render_shared(counters=True)                   # the same frame again: nothing left to steal
    finally:
        ct.api.set_option("shared_hold_frame", 0)
    assert c2["rays_primary"] == 0 and c2["rays_shadow"] == 0 and np.array_equal(gpu.readback(), want)
    for _ in range(3):                                          # new frames alternate between the two cursors; no reset call
        c3 = gpu.render_shared(counters=True)
        assert c3 == c1 and np.array_equal(gpu.readback(), want)
    gpu.share_reset()                                           # still allowed: zeroes both cursors
    c4 = gpu.render_shared(counters=True)
    assert c4 == c1 and np.array_equal(gpu.readback(), want)
    gpu.share_attach(None)
    gpu.render_tile()
    assert np.array_equal(gpu.readback(), want)


@pytest.mark.parametrize("count,eighths,run_shift", [(2, 7, 0), (3, 7, 0), (8, 7, 0), (4, 8, 0), (4, 0, 0), (5, 3, 0), (8, 7, 2), (3, 7, 4), (4, 8, 1), (2, 0, 3)])
def test_shared_frame_partition_covers_the_frame_once(count, eighths, run_shift, golden, scene_loader, gpu):
    """ct_gpu_share_partition: of every 8*count chunks (or runs of 2^run_shift chunks, option "shared_run_shift") `eighths`*count
    are dealt round-robin, the rest stolen from the cursor.  One GPU plays all `count` participants in turn (the first one to run
    steals every stolen chunk): together they must trace every pixel exactly once -- the ray counts add up to the one-GPU frame's
    and the frame is identical."""
    fs, meta = case_scene("bunny_refl_d2_160", golden, scene_loader)
    want = load_frames("bunny_refl_d2_160")["frame"]
    gpu.upload(fs, meta["width"], meta["height"], max_depth=meta["depth"])
    gpu.render_tile()
    whole = gpu.counters(reset=True)
    gpu.upload(fs, meta["width"], meta["height"], max_depth=meta["depth"])      # fresh framebuffer
    gpu.share_attach(gpu.share_export())
    ct.api.set_option("shared_static_eighths", eighths)
    ct.api.set_option("shared_run_shift", run_shift)
    try:
        parts = []
        for k in range(count):
            gpu.share_partition(k, count)
            ct.api.set_option("shared_hold_frame", 1 if k > 0 else 0)      # participants 1.. render the SAME frame as participant 0
            parts.append(gpu.render_shared(counters=True))
            if k == 0 and count > 1 and eighths > 0:
                assert not np.array_equal(gpu.readback(), want)      # one participant alone does not finish the frame
        assert np.array_equal(gpu.readback(), want)
        for key in ("rays_primary", "rays_shadow", "rays_reflection"):
            assert sum(p[key] for p in parts) == whole[key], key
        if eighths == 8:
            assert max(p["rays_primary"] for p in parts) <= min(p["rays_primary"] for p in parts) + (64 * 2 << run_shift)
    finally:
        ct.api.set_option("shared_static_eighths", 7)
        ct.api.set_option("shared_run_shift", 0)
        ct.api.set_option("shared_hold_frame", 0)
        gpu.share_partition(0, 0)
        gpu.share_attach(None)


def test_subsampling_against_reference_and_oracle(golden, scene_loader, gpu):
    """CT_FLAG_SUBSAMPLING = settings.subsampling (raythread.cpp:512-531): one tile per frame against frames of the
    compiled reference (one worker thread), then two tiles in sequence against the oracle run the same way."""
    for case, m in golden["frames_subsampling"].items():
        fs = scene_loader(m["scene"])
        if m["force_reflection"] is not None:
            fs = fs.with_reflection(m["force_reflection"])
        gpu.upload(fs, m["width"], m["height"], max_depth=m["depth"], flags=ct.CT_FLAG_SUBSAMPLING)
        gpu.render_tile()
        want = np.load(os.path.join(GOLD, f"frames_sub_{case}.npz"))["frame"]
        got = np.zeros_like(want)
        gpu.readback(got)
        assert np.array_equal(got, want), f"{case}: {int((got != want).sum())} pixels differ"
    fs = scene_loader("scene_import_bunny").with_reflection(0.4)
    W, H = 211, 158
    y0, y1 = -(H // 2), -(H // 2) + H
    for cut in (y0 + 37, y0 + 80):                      # odd and even first partition
        osc = O.OracleScene(fs)
        want = np.zeros((H, W), np.uint32)
        for a, b in ((y0, cut), (cut, y1)):             # the second partition overwrites the seam row, as in submission order
            f, _, _ = osc.render(W, H, y_start=a, y_end=b, max_depth=2, flags=O.SUBSAMPLE, want_hits=False, n_threads=1)
            want = np.where(f != 0, f, want)
        gpu.upload(fs, W, H, max_depth=2, flags=ct.CT_FLAG_SUBSAMPLING)
        gpu.render_tile(y0, cut); gpu.render_tile(cut, y1)
        got = np.zeros_like(want)
        gpu.readback(got)
        assert np.array_equal(got, want), f"cut {cut}: {int((got != want).sum())} pixels differ"
    with pytest.raises(ct.CtError):
        gpu.render_shared()                             # neighbouring rows may not live on different GPUs


def test_supersampling_against_patched_reference_and_oracle(golden, scene_loader, gpu):
    """CT_FLAG_SUPERSAMPLING = settings.supersampling (raythread.cpp:460-505) with the counter-based jitter: frames of
    the compiled reference (--supersampling-hash), then a larger frame, cut into ragged tiles, against the oracle."""
    for case, m in golden["frames_supersampling"].items():
        fs = scene_loader(m["scene"])
        if m["force_reflection"] is not None:
            fs = fs.with_reflection(m["force_reflection"])
        gpu.upload(fs, m["width"], m["height"], max_depth=m["depth"], flags=ct.CT_FLAG_SUPERSAMPLING)
        c = gpu.render_tile(counters=True)
        want = np.load(os.path.join(GOLD, f"frames_ss_{case}.npz"))["frame"]
        got = np.zeros_like(want)
        gpu.readback(got)
        assert np.array_equal(got, want), f"{case}: {int((got != want).sum())} pixels differ"
        assert c["rays_primary"] == 16 * int((want != 0).sum())
    fs = scene_loader("scene_import_bunny").with_reflection(0.3)
    W, H = 333, 250
    want, _, octr = O.OracleScene(fs).render(W, H, max_depth=3, flags=O.SUPERSAMPLE, want_hits=False)
    gpu.upload(fs, W, H, max_depth=3, flags=ct.CT_FLAG_SUPERSAMPLING)
    y0, y1 = gpu.full_range()
    cuts = [y0, y0 + 3, y0 + 70, 1, y1]
    for a, b in zip(cuts[:-1], cuts[1:]):
        gpu.render_tile(a, b)
    got = np.zeros_like(want)
    gpu.readback(got)
    assert np.array_equal(got, want), f"{int((got != want).sum())} pixels differ"
    gpu.share_attach(gpu.share_export())      # whole pixels per chunk: the shared-frame path works too
    gpu.render_shared()
    got2 = np.zeros_like(want)
    gpu.readback(got2)
    assert np.array_equal(got2, want)


@pytest.mark.parametrize("stride", [96, 120])
def test_reference_triangle_stride(stride, golden, scene_loader, gpu):
    """The reference's triangle_t is 96 bytes (p1, p2, p3, centroid: scenefile.h:36-41) and sits inside a 120-byte
    scene_object_t; both strides give the frame of the packed 72-byte layout (the bytes between triangles are NaNs here)."""
    fs, meta = case_scene("bunny_refl_d2_160", golden, scene_loader)
    g = load_frames("bunny_refl_d2_160")
    gpu.upload(fs, meta["width"], meta["height"], max_depth=meta["depth"], triangle_stride=stride)
    gpu.render_tile()
    assert np.array_equal(gpu.readback(), g["frame"])


def test_640_golden_hashes(golden, scene_loader, gpu):
    for name in ("scene_file_cube", "scene_import", "scene_import_bunny", "pc_big"):
        gpu.upload(scene_loader(name), 640, 640)
        gpu.render_tile()
        f = gpu.readback()
        assert ct.frame_fnv1a(f) == golden["scenes"][name]["frame640_fnv1a"], name
        assert int((f == 0).sum()) == 640                       # row 0 never written


def test_ragged_tiles_equal_one_tile(scene_loader, gpu):
    fs = scene_loader("scene_import_bunny").with_reflection(0.3)
    W = H = 301
    gpu.upload(fs, W, H, max_depth=3)
    gpu.render_tile()
    full = gpu.readback()
    gpu.upload(fs, W, H, max_depth=3)
    y0, y1 = gpu.full_range()
    cuts = [y0, y0 + 1, y0 + 4, y0 + 21, y0 + 22, 0, 7, y1 - 1, y1]
    for a, b in reversed(list(zip(cuts[:-1], cuts[1:]))):       # out of order on purpose
        gpu.render_tile(a, b)
    gpu.render_tile(y1 + 5, y1 + 50)                             # completely outside: a no-op
    gpu.render_tile(10, 10)                                      # empty
    assert np.array_equal(gpu.readback(), full)


def test_wide_flag_against_oracle(scene_loader, gpu):
    fs = scene_loader("scene_import")
    W, H = 400, 180
    oframe, ohits, _ = O.OracleScene(fs).render(W, H, max_depth=10, flags=1)
    gpu.upload(fs, W, H, flags=DBG | ct.CT_FLAG_WIDE)
    gpu.render_tile()
    assert_same(gpu, oframe, ohits, "wide")
    assert (oframe[1:, 0] != 0).all() and (oframe[1:, -1] != 0).all()   # the side bars are traced in this mode


def test_readback_leaves_untraced_pixels_untouched(scene_loader, gpu):
    fs = scene_loader("scene_file_cube")
    W, H = 200, 120                                              # square mode: columns 40..159 only
    gpu.upload(fs, W, H)
    gpu.render_tile()
    out = np.full((H, W), 0xDEADBEEF, np.uint32)
    gpu.readback(out)
    assert (out[0] == 0xDEADBEEF).all()                          # row 0 (y = H/2) is never reached
    assert (out[:, :40] == 0xDEADBEEF).all() and (out[:, 160:] == 0xDEADBEEF).all()
    assert (out[1:, 40:160] != 0xDEADBEEF).all()
    ref = load_frames("cube_wide_200x120")["frame"]
    assert np.array_equal(out[1:, 40:160], ref[1:, 40:160])


def test_arbitrary_rays_including_t0_zero(scene_loader, gpu):
    fs = scene_loader("scene_import_bunny")
    osc = O.OracleScene(fs)
    gpu.upload(fs, 64, 64)
    rng = np.random.default_rng(7)
    n = 4000
    centre = fs.tri.reshape(-1, 3).mean(0)
    org = centre + rng.normal(size=(n, 3)) * 15
    target = fs.tri.reshape(-1, 3)[rng.integers(0, fs.n_tri * 3, n)] + rng.normal(size=(n, 3)) * 0.5
    d = (target - org) * rng.uniform(0.2, 3, (n, 1))
    t0 = rng.choice(np.array([1e30, 0.0, 5.0, 0.5, 1e-4], np.float32), n)
    # the reference's reflection rays start ON a surface (raythread.cpp:373): second half of the rays does too
    k = rng.integers(0, fs.n_tri, n // 2)
    w = rng.dirichlet([1, 1, 1], n // 2)
    T = fs.tri[k].reshape(-1, 3, 3)
    org[n // 2:] = (T * w[:, :, None]).sum(1)
    d[n // 2:] = rng.normal(size=(n // 2, 3))
    t0[n // 2:] = rng.choice(np.array([0.0, 0.0, 1e30], np.float32), n // 2)
    found, index, t = gpu.debug_closest(org, d, t0)
    for i in range(n):
        f, k, tt = osc.closest(org[i], d[i], t0[i])
        assert (bool(found[i]), int(index[i])) == (f, k) and np.float32(tt).view(np.uint32) == t[i].view(np.uint32), i
    assert 100 < int(((t0 == 0) & (t == 0)).sum())               # the t=0 "first line hit" path was exercised


def test_set_camera_equals_reupload(golden, scene_loader, gpu):
    fs, meta = case_scene("cube_rot_160", golden, scene_loader)
    base = scene_loader("scene_file_cube")
    gpu.upload(base, 160, 160)
    gpu.render_tile()
    gpu.set_camera(meta["cam_pos"], meta["cam_rot"])
    gpu.render_tile()
    assert np.array_equal(gpu.readback(), load_frames("cube_rot_160")["frame"])


def test_render_is_deterministic_and_idempotent(scene_loader, gpu):
    fs = scene_loader("pc_big")
    gpu.upload(fs, 256, 256)
    gpu.render_tile(); a = gpu.readback()
    gpu.render_tile(); gpu.render_tile(); b = gpu.readback()
    assert np.array_equal(a, b)


def test_error_paths(scene_loader, gpu):
    with pytest.raises(ct.CtError) as e:
        gpu.render_tile(-4, 4)
    assert e.value.code == -4                                    # CT_ERR_NO_SCENE
    fs = scene_loader("scene_file_cube")
    with pytest.raises(ct.CtError) as e:
        gpu.upload(fs, 64, 64, max_depth=99)
    assert e.value.code == -5                                    # CT_ERR_LIMIT
    bad = dataclasses.replace(fs, tri_index=np.zeros_like(fs.tri_index) + 1)
    with pytest.raises(ct.CtError) as e:
        gpu.upload(bad, 64, 64)
    assert e.value.code == -1
    loop = dataclasses.replace(fs, node_left=np.zeros_like(fs.node_left))      # children point back at the root
    with pytest.raises(ct.CtError):
        gpu.upload(loop, 64, 64)
    with pytest.raises(ct.CtError):
        gpu.upload(fs, 0, 64)
    gpu.upload(fs, 64, 64)
    with pytest.raises(ct.CtError):
        gpu.readback_hits()                                      # uploaded without CT_FLAG_KEEP_HITS


def test_boss_matches_direct_calls(scene_loader):
    fs = scene_loader("scene_import_bunny").with_reflection(0.5)
    W, H = 320, 240
    oframe, _, octr = O.OracleScene(fs).render(W, H, max_depth=2, want_hits=False)
    hs = host.HostScene.from_flat(fs.without_bvh())               # the boss builds the BVH itself (raythread.cpp:650-651)
    for tile_rows in (0, 16, 7):
        b = host.Boss(hs, W, H, devices=(0,), max_depth=2, tile_rows=tile_rows)
        bitmap = np.full((H, W), 0xABCDEF01, np.uint32)
        bitmap, stats = b.render(bitmap)
        ref = np.where(oframe == 0, 0xABCDEF01, oframe)
        assert np.array_equal(bitmap[1:, 40:280], oframe[1:, 40:280]) and (bitmap[0] == 0xABCDEF01).all(), tile_rows
        assert stats["tiles_mine"] == stats["tiles_total"] == len(b.tiles())
        assert stats["rays_shadow"] + stats["rays_reflection"] > 0
        b.close()


def _soup_scene(rng, n_tri, spread, size, reflection, same_centroid=False):
    """Random triangle soup in front of a camera at the origin looking down +z (cobbletrace.cpp's default frame)."""
    from cobbletrace_b200.sceneio import FlatScene, LT_AMBIENT, LT_POINT, LT_DIRECTIONAL
    c = rng.normal(size=(n_tri, 3)) * spread + np.array([0.0, 0.0, 6.0])
    if same_centroid:
        c[:] = c[0]                                              # BuildBVH cannot split: one big leaf at the root
    a, b = rng.normal(size=(n_tri, 3)) * size, rng.normal(size=(n_tri, 3)) * size
    tri = np.concatenate([c - (a + b) / 3, c - (a + b) / 3 + a, c - (a + b) / 3 + b], 1)
    return FlatScene(tri=tri, mat_color=rng.integers(0x202020, 0xffffff, n_tri).astype(np.uint32),
                     mat_specular=rng.choice(np.array([-1, 0, 10, 500], np.int32), n_tri),
                     mat_reflection=np.where(rng.random(n_tri) < 0.5, reflection, 0.0).astype(np.float32),
                     light_type=np.array([LT_AMBIENT, LT_POINT, LT_DIRECTIONAL], np.int32), light_intensity=np.array([0.2, 0.6, 0.3], np.float32),
                     light_pos=np.array([[0, 0, 0], [3.0, 5.0, -2.0], [0, 0, 0]]), light_dir=np.array([[0, 0, 0], [0, 0, 0], [1.0, 4.0, -4.0]]),
                     cam_pos=np.zeros(3), cam_rot=np.eye(3).reshape(9))


@pytest.mark.parametrize("n_tri,same_centroid", [(1, False), (2, False), (3, False), (40, True), (700, False)])
def test_small_and_degenerate_trees_against_oracle(n_tri, same_centroid, gpu):
    """Trees the bundled scenes never produce: a root that is a leaf (1-2 triangles, or 40 triangles with one common
    centroid that BuildBVH cannot split), a 3-triangle tree, and a soup with heavy overlap -- each with mirrors."""
    rng = np.random.default_rng(100 + n_tri)
    fs = _soup_scene(rng, n_tri, spread=1.2, size=1.5, reflection=0.6, same_centroid=same_centroid)
    fs = host.HostScene.from_flat(fs).to_flat(with_bvh=True)
    if n_tri <= 2 or same_centroid:
        assert fs.n_nodes == 1 and int(fs.node_count[0]) == n_tri
    W, H = 96, 80
    oframe, ohits, _ = O.OracleScene(fs).render(W, H, max_depth=4)
    for budget in (0, 8):                                        # 8: nearly every early-exit ray goes through k_overflow
        ct.api.set_option("traversal_budget", budget)
        try:
            gpu.upload(fs, W, H, max_depth=4, flags=DBG)
            gpu.render_tile()
            assert_same(gpu, oframe, ohits, f"soup n={n_tri} budget={budget}")
        finally:
            ct.api.set_option("traversal_budget", 0)


def test_boss_two_gpus_in_one_process(scene_loader):
    """ct_host_boss over two devices of this process: one shared frame (device-side chunk stealing, peer stores into
    devices[0]) and the row-tile variant (host-side counter, peer copies).  Needs a box with >= 2 GPUs."""
    if ct.api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fs = scene_loader("scene_import_bunny").with_reflection(0.5)
    W, H = 640, 480
    oframe, _, _ = O.OracleScene(fs).render(W, H, max_depth=2, want_hits=False)
    hs = host.HostScene.from_flat(fs.without_bvh())
    for tile_rows in (0, 32, 7, 0):
        b = host.Boss(hs, W, H, devices=(0, 1), max_depth=2, tile_rows=tile_rows)
        try:
            for _ in range(3):
                # a fresh bitmap every frame: rows that one device rendered and the other has to hand over (gathered
                # rows must count as "hold pixels" for the root's readback, whichever device took the outermost tiles)
                bitmap, stats = b.render(np.zeros((H, W), np.uint32))
                assert np.array_equal(bitmap, oframe), tile_rows
        finally:
            b.close()


@pytest.mark.parametrize("W,H", [(3840, 2160), (7680, 4320)])
def test_dragon_class_full_size(tmp_path_factory, W, H, gpu):
    """BASELINE configs 3 (4K) and 5 (8K, here on one GPU, cut into row tiles) at full size: procedural 868k-triangle
    stand-in, forced reflection, depth 2."""
    d = os.environ.get("CT_SCENE_CACHE") or str(tmp_path_factory.mktemp("dragon"))
    scene, n = procedural.write_dragon_standin(d)
    hs = host.HostScene.load(scene, base_dir=d)
    hs.set_reflection(0.5)
    fs = hs.to_flat(with_bvh=True)
    assert fs.n_tri == n == 868352
    gpu.upload(fs, W, H, max_depth=2, flags=DBG)
    ctr = gpu.render_tile(counters=True)
    full = gpu.readback()
    # property 1: oracle agreement at full size (bit-exact)
    oframe, ohits, octr = O.OracleScene(fs).render(W, H, max_depth=2, n_threads=os.cpu_count() or 8)
    assert_same(gpu, oframe, ohits, f"dragon stand-in {W}x{H}")
    # ... and the frame the compiled, unmodified reference rendered from the same scene file (tests/golden/make_golden_bench.py)
    import json
    pins = json.load(open(os.path.join(GOLD, "bench_frames.json")))
    pin = pins["dragon4k" if W == 3840 else "dragon8k"]
    assert (pin["width"], pin["height"], pin["triangles"]) == (W, H, n) and ct.frame_fnv1a(full) == pin["fnv1a"]
    # property 2: tile-split invariance (what multi-GPU row tiles rely on)
    gpu.upload(fs, W, H, max_depth=2)
    y0, y1 = gpu.full_range()
    step = 135
    for a in range(y0, y1, step):
        gpu.render_tile(a, min(a + step, y1))
    assert np.array_equal(gpu.readback(), full)
    # property 3: every reflection ray "hits" (SURVEY 0.4): per primary hit 3 shading points -> 6 shadow rays
    hits = int((ohits["found"] == 1).sum())
    assert ctr["rays_reflection"] == 2 * hits and ctr["rays_shadow"] == 6 * hits


@pytest.mark.gpu
def test_viewer_sessions_match_reference(golden, scene_loader):
    """cobbletrace.cpp's main loop without the window.  Key presses go through ct_host_controls
    (AddEvent / HandleKeyboard / HandleUpdates), frames through the boss; every tick's bitmap must hash to what the
    compiled reference showed after the same keys (tests/golden/make_golden_sessions.py), and no frame is rendered
    when nothing changed."""
    from cobbletrace_b200.sceneio import frame_fnv1a
    for case, m in golden["sessions"].items():
        hs = host.HostScene.from_flat(scene_loader(m["scene"]).without_bvh())
        boss = host.Boss(hs, m["width"], m["height"], devices=(0,), max_depth=m["depth"])
        shown = []
        v = host.Viewer(boss, present=lambda bitmap, is_new: shown.append((frame_fnv1a(bitmap), is_new)))
        assert v.tick()                                           # the first frame
        assert frame_fnv1a(v.bitmap) == m["ticks"][0]["fnv1a"], (case, 0)
        assert not v.tick() and shown[-1] == (m["ticks"][0]["fnv1a"], False)      # idle tick: Blit only
        for k, batch in enumerate(m["session"].split("|"), 1):
            v.keys(batch + "m")                                   # the reference harness appends 'm' to every batch
            assert v.tick()
            assert frame_fnv1a(v.bitmap) == m["ticks"][k]["fnv1a"], (case, k)
            pos, _, rot = v.controls.camera()
            assert np.array_equal(pos, np.array(m["ticks"][k]["pos"])) and np.array_equal(rot, np.array(m["ticks"][k]["rot"]))
            assert v.last_stats["kernel_launches"] > 0
        assert len(shown) == len(m["ticks"]) + 1 and not v.tick()
        boss.close()


@pytest.mark.gpu
def test_tiny_and_narrow_frames_match_reference(golden, scene_loader, gpu):
    """Degenerate bitmap sizes against frames of the compiled reference (tests/golden/make_golden_tiny.py): H = 1 (nothing
    is traced), odd H (row 0 is reached after all), W < H (columns outside the bitmap wrap into neighbouring rows)."""
    for m in golden["tiny_frames"]:
        W, H = m["width"], m["height"]
        gpu.upload(scene_loader(m["scene"]), W, H, max_depth=m["depth"])
        gpu.render_tile()
        assert np.array_equal(gpu.readback(), np.array(m["frame"], np.uint32)), (m["scene"], W, H, m["depth"])


@pytest.mark.gpu
def test_light_sets_and_depth_limits_against_oracle(scene_loader, gpu):
    """Scenes without lights, with the ambient light only, one light of each kind; recursion depth 0 and 1 on the
    mirror scene (raythread.cpp:366: `recursionDepth <= 0 || reflection <= 0`)."""
    from cobbletrace_b200.sceneio import LT_AMBIENT
    fs = scene_loader("scene_file_cube")
    W, H = 72, 64
    variants = [("depth0", fs, 0), ("depth1", fs, 1)]
    for keep in ([], [int(np.flatnonzero(fs.light_type == LT_AMBIENT)[0])], [1], [2], [0, 1, 2, 1, 2]):
        sel = np.array(keep, np.int64)
        variants.append((f"lights{keep}", dataclasses.replace(fs, light_type=fs.light_type[sel], light_intensity=fs.light_intensity[sel],
                                                             light_pos=fs.light_pos[sel], light_dir=fs.light_dir[sel]), 3))
    for what, scene, depth in variants:
        oframe, ohits, _ = O.OracleScene(scene).render(W, H, max_depth=depth)
        gpu.upload(scene, W, H, max_depth=depth, flags=DBG)
        gpu.render_tile()
        assert_same(gpu, oframe, ohits, what)


@pytest.mark.gpu
def test_random_scene_files_match_reference(golden, gpu):
    """The 20 generated scene files of tests/golden/make_golden_fuzz.py: frames and primary hit records of the compiled
    reference, bit for bit, with the default walk budget and with one that parks nearly every early-exit ray."""
    from conftest import load_fuzz_case
    for k, m in golden["fuzz"].items():
        fs, g = load_fuzz_case(k)
        for budget in (0, 6):
            ct.api.set_option("traversal_budget", budget)
            try:
                gpu.upload(fs, m["width"], m["height"], max_depth=m["depth"], flags=DBG)
                gpu.render_tile()
                assert np.array_equal(gpu.readback(), g["frame"]), (k, budget)
                found, index, t = gpu.readback_hits()
                traced = g["found"] != 0xFFFFFFFF
                assert np.array_equal(found, g["found"]) and np.array_equal(index[traced], g["index"][traced]), (k, budget)
                assert np.array_equal(t[traced].view(np.uint32), g["t"][traced].view(np.uint32)), (k, budget)
            finally:
                ct.api.set_option("traversal_budget", 0)


@pytest.mark.gpu
def test_both_sampling_modes_at_once(golden, scene_loader, gpu):
    """CT_FLAG_SUBSAMPLING | CT_FLAG_SUPERSAMPLING: frames of the compiled reference with both settings on, then a frame cut
    into partitions rendered in sequence against the oracle doing the same."""
    both = ct.CT_FLAG_SUBSAMPLING | ct.CT_FLAG_SUPERSAMPLING
    for case, m in golden["frames_both_sampling"].items():
        fs = scene_loader(m["scene"])
        if m["force_reflection"] is not None:
            fs = fs.with_reflection(m["force_reflection"])
        gpu.upload(fs, m["width"], m["height"], max_depth=m["depth"], flags=both)
        gpu.render_tile()
        assert np.array_equal(gpu.readback(), np.load(os.path.join(GOLD, f"frames_both_{case}.npz"))["frame"]), case
    fs = scene_loader("scene_import_bunny").with_reflection(0.4)
    W, H = 150, 101
    half = H // 2
    cuts = [-half, -31, -30, 2, 17, -half + H]
    want = np.zeros((H, W), np.uint32)
    osc = O.OracleScene(fs)
    for a, b in zip(cuts[:-1], cuts[1:]):                     # the oracle's partitions write into one frame, in order
        part, _, _ = osc.render(W, H, y_start=a, y_end=b, max_depth=2, flags=O.SUBSAMPLE | O.SUPERSAMPLE, want_hits=False, n_threads=1)
        want = np.where(part != 0, part, want)
    gpu.upload(fs, W, H, max_depth=2, flags=both)
    for a, b in zip(cuts[:-1], cuts[1:]):
        gpu.render_tile(a, b)
    got = gpu.readback()
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_upload_validates_the_tree_and_ignores_unreachable_slots(golden, scene_loader, gpu):
    """What ct_gpu_upload_scene guarantees about a caller's BVH: tri_indexes must be a permutation (a duplicate is rejected, not
    silently rendered), leaves may not overlap, a node has one parent -- and array slots the tree never reaches may hold
    anything (NaN boxes, wild child indices) without changing a pixel or switching the filters off."""
    fs, meta = case_scene("bunny_refl_d2_160", golden, scene_loader)
    want = load_frames("bunny_refl_d2_160")["frame"]
    W, H, depth = meta["width"], meta["height"], meta["depth"]
    # garbage in unreachable slots
    k = 5
    junk = dataclasses.replace(
        fs, node_min=np.concatenate([fs.node_min, np.full((k, 3), np.nan)]), node_max=np.concatenate([fs.node_max, np.full((k, 3), -np.inf)]),
        node_left=np.concatenate([fs.node_left, np.array([3, 0xFFFFFFF0, 1, 7, 2], np.uint32)]),
        node_first=np.concatenate([fs.node_first, np.array([0, 9, 0xFFFFFF00, 1, 2], np.uint32)]),
        node_count=np.concatenate([fs.node_count, np.array([0, 0, 2, 0, 1], np.uint32)]))
    gpu.upload(junk, W, H, max_depth=depth, flags=ct.CT_FLAG_COUNT_TESTS)
    gpu.render_tile()
    assert np.array_equal(gpu.readback(), want)
    bx, tx = gpu.filter_stats()
    c = gpu.counters()
    assert bx < 0.02 * c["box_tests"]                       # the certified slab filter still decides nearly everything
    # a duplicated triangle index
    dup = fs.tri_index.copy(); dup[7] = dup[8]
    with pytest.raises(ct.CtError, match="not a permutation"):
        gpu.upload(dataclasses.replace(fs, tri_index=dup), W, H, max_depth=depth)
    # two leaves that share a position
    leaves = np.flatnonzero(fs.node_count > 0)
    a = leaves[np.argmax(fs.node_first[leaves] > 0)]
    first = fs.node_first.copy(); first[a] -= 1
    with pytest.raises(ct.CtError, match="malformed"):
        gpu.upload(dataclasses.replace(fs, node_first=first), W, H, max_depth=depth)
    # a node with two parents
    interior = np.flatnonzero(fs.node_count == 0)
    left = fs.node_left.copy(); left[interior[3]] = left[interior[2]]
    with pytest.raises(ct.CtError, match="malformed"):
        gpu.upload(dataclasses.replace(fs, node_left=left), W, H, max_depth=depth)


@pytest.mark.gpu
def test_tree_that_is_not_nested_takes_the_exact_walks(golden, scene_loader, gpu):
    """The early-exit walks' conservative box tests rest on every child's box lying inside its parent's (box_maybe in
    ct_traverse.cuh).  A caller's tree without that property -- here: boxes of some inner nodes SHRUNK, so that rays the
    parent rejects would have hit the children -- must be walked with the reference's verdict at every box: the frame then
    equals the oracle's on the same (odd) tree, and differs from the frame of the proper tree."""
    fs, meta = case_scene("bunny_refl_d2_160", golden, scene_loader)
    W, H, depth = meta["width"], meta["height"], meta["depth"]
    nmin, nmax = fs.node_min.copy(), fs.node_max.copy()
    interior = np.flatnonzero(fs.node_count == 0)
    rng = np.random.default_rng(3)
    pick = interior[rng.choice(len(interior), len(interior) // 6, replace=False)]
    pick = pick[pick != 0]
    mid = 0.5 * (nmin[pick] + nmax[pick])
    nmin[pick] = mid + 0.6 * (nmin[pick] - mid); nmax[pick] = mid + 0.6 * (nmax[pick] - mid)
    odd = dataclasses.replace(fs, node_min=nmin, node_max=nmax)
    oframe, ohits, _ = O.OracleScene(odd).render(W, H, max_depth=depth)
    gpu.upload(odd, W, H, max_depth=depth, flags=DBG)
    gpu.render_tile()
    assert_same(gpu, oframe, ohits, "shrunk inner boxes")
    assert not np.array_equal(oframe, load_frames("bunny_refl_d2_160")["frame"])


@pytest.mark.gpu
def test_identical_shadow_rays_are_traced_once(golden, scene_loader, gpu):
    """The reference's reflection rays have t = 0, so a reflection "hit" puts the next shading point exactly on the previous
    one and ComputeLighting casts the same shadow rays again (raythread.cpp:360,373,288-304).  Option shadow_reuse (default on)
    answers them from the parent's verdicts: same frames, same ray counts (rays_shadow keeps the reference's definition),
    fewer box tests -- on the mirror scene of configs[0], the bunny with forced mirrors at depth 2 and 10, and a generated scene."""
    # (in scene_file_cube.json no reflection ray finds a pass -- the mirror shows triangle 0's colour, SURVEY 0.4 -- so every
    # reflection "hit" is 2^32 ray lengths away and nothing repeats: the option must not change anything there either)
    cases = [("cube_160", False), ("bunny_refl_d2_160", True), ("bunny_refl_d10_128", True)]
    for case, repeats in cases:
        fs, meta = case_scene(case, golden, scene_loader)
        want = load_frames(case)["frame"]
        res = {}
        for reuse in (0, 1):
            ct.api.set_option("shadow_reuse", reuse)
            try:
                gpu.upload(fs, meta["width"], meta["height"], max_depth=meta["depth"], flags=ct.CT_FLAG_COUNT_TESTS)
                c = gpu.render_tile(counters=True)
                assert np.array_equal(gpu.readback(), want), (case, reuse)
                res[reuse] = (c, gpu.reuse_stats())
            finally:
                ct.api.set_option("shadow_reuse", 1)
        (c0, r0), (c1, r1) = res[0], res[1]
        assert r0 == 0 and (r1 > 0) == repeats, (case, r0, r1)
        for k in ("rays_primary", "rays_shadow", "rays_reflection"):
            assert c0[k] == c1[k], (case, k)
        assert (c1["box_tests"] < c0["box_tests"]) == repeats, case
        assert r1 <= c1["rays_shadow"]
