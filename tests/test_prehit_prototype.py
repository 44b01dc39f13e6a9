"""DESIGN.md "what comes next" 1, checked on the CPU before any kernel is written: the part of a closest-hit walk before
the first effective hit is order-free.  oracle/ct_oracle.c holds a prototype walk (phase 1 in a different visit order,
the reference's stack rebuilt at the first hit, the ordered walk continued from there) that must return exactly what the
reference-order walk returns -- found, triangle index and the bits of tclosest -- on every ray."""
import numpy as np
import pytest

from oracle import ct_oracle_py as O
from conftest import load_fuzz_case


def camera_rays(fs, W, H, rng, jitter):
    """Pixel-grid rays from the scene's camera (raythread.cpp:186: direction (x / H, y / H, 1) . rotation), optionally jittered."""
    half = H // 2
    x, y = np.meshgrid(np.arange(-half, half), np.arange(-half, half))
    d = np.stack([x.ravel() / H, y.ravel() / H, np.ones(x.size)], 1)
    if jitter:
        d[:, :2] += rng.normal(size=(d.shape[0], 2)) * jitter
    rot = np.asarray(fs.cam_rot).reshape(3, 3)
    d = d @ rot
    return np.tile(np.asarray(fs.cam_pos, np.float64), (d.shape[0], 1)), d


@pytest.mark.parametrize("name,H", [("scene_file_cube", 96), ("scene_import", 96), ("scene_import_bunny", 128), ("pc_big", 96)])
def test_prehit_walk_equals_reference_walk_on_bundled_scenes(name, H, scene_loader):
    fs = scene_loader(name)
    sc = O.OracleScene(fs)
    rng = np.random.default_rng(5)
    for jitter in (0.0, 0.003):
        org, d = camera_rays(fs, H, H, rng, jitter)
        bad, found = O.prehit_check(sc, org, d)
        assert bad == 0 and found > 0, (name, jitter, bad, found)
    # rays from inside and around the geometry, any direction (boxes behind the origin, grazing, many misses)
    c = fs.tri.reshape(-1, 3).mean(0)
    ext = np.ptp(fs.tri.reshape(-1, 3), axis=0).max()
    org = c + rng.normal(size=(4000, 3)) * ext * 0.7
    bad, found = O.prehit_check(sc, org, rng.normal(size=(4000, 3)))
    assert bad == 0 and found > 0, (name, "random", bad, found)


def test_prehit_walk_equals_reference_walk_on_generated_scenes(golden):
    rng = np.random.default_rng(6)
    for k in golden["fuzz"]:
        fs, _ = load_fuzz_case(k)
        org, d = camera_rays(fs, 48, 48, rng, 0.002)
        bad, _ = O.prehit_check(O.OracleScene(fs), org, d)
        assert bad == 0, (k, bad)
