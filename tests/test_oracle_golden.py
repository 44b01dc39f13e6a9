"""The oracle (oracle/ct_oracle.c, a plain-C restatement) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py -> oracle/_ref/ct_ref).  This is what pins the oracle."""
import numpy as np
import pytest

from cobbletrace_b200.sceneio import frame_fnv1a
from oracle import ct_oracle_py as O
from conftest import load_frames, case_scene

SCENES = ["scene_file_cube", "scene_import", "pc_big", "scene_import_bunny"]
CASES = ["cube_160", "cube_rot_160", "cube_wide_200x120", "cube_tall_90x150", "import_160", "pc_big_96", "bunny_160",
         "bunny_refl_d2_160", "bunny_refl_d10_128", "bunny_odd_161x161"]


@pytest.mark.parametrize("name", SCENES)
def test_bvh_restatement_matches_reference_build(name, golden, scene_loader):
    fs = scene_loader(name)
    b = O.build_bvh(fs.tri)
    for k, v in b.items():
        assert np.array_equal(v, getattr(fs, k)), k
    assert len(b["node_left"]) == golden["scenes"][name]["n_nodes"]


@pytest.mark.parametrize("name", ["scene_file_cube", "scene_import", "scene_import_bunny"])
def test_frame_640_hash_and_counters(name, golden, scene_loader):
    """The four 640x640 hashes are the ones SURVEY 8(c) recorded from the reference independently."""
    meta = golden["scenes"][name]
    frame, hits, ctr = O.OracleScene(scene_loader(name)).render(640, 640, max_depth=10, want_hits=False)
    assert frame_fnv1a(frame) == meta["frame640_fnv1a"]
    assert int((frame == 0).sum()) == meta["frame640_zero_px"] == 640      # row 0 is never written
    assert ctr == meta["counters640"]


def test_survey_hashes_are_the_recorded_ones(golden):
    expect = {"scene_file_cube": "a52ca3e236c27312", "scene_import": "1920dbc59aefa6ea",
              "scene_import_bunny": "09fe037eb4efa918", "pc_big": "2efe7bf0c10ddd93"}
    for k, v in expect.items():
        assert golden["scenes"][k]["frame640_fnv1a"] == v


@pytest.mark.parametrize("case", CASES)
def test_low_res_frames_and_hits(case, golden, scene_loader):
    fs, meta = case_scene(case, golden, scene_loader)
    g = load_frames(case)
    frame, hits, _ = O.OracleScene(fs).render(meta["width"], meta["height"], max_depth=meta["depth"])
    assert np.array_equal(frame, g["frame"])
    assert np.array_equal(hits["found"], g["found"])
    traced = g["found"] != 0xFFFFFFFF
    assert np.array_equal(hits["index"][traced], g["index"][traced])
    assert np.array_equal(hits["t"][traced].view(np.uint32), g["t"][traced].view(np.uint32))      # bit-exact distances


def test_primitive_kats():
    import os
    from conftest import GOLD
    zf = np.load(os.path.join(GOLD, "kat_primitives.npz"))
    z = {k: zf[k] for k in zf.files}        # NpzFile decompresses on every access
    L = O.lib()
    import ctypes as C
    n = len(z["t0"])
    bad = 0
    for i in range(n):
        t = C.c_float(float(z["t0"][i]))
        o = np.ascontiguousarray(z["org"][i]); d = np.ascontiguousarray(z["dir"][i]); tr = np.ascontiguousarray(z["tri"][i])
        mn = np.ascontiguousarray(z["bmin"][i]); mx = np.ascontiguousarray(z["bmax"][i])
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        bh = L.ct_oracle_intersect_aabb(p(o), p(d), t, p(mn), p(mx))
        th = L.ct_oracle_intersect_triangle(p(o), p(d), C.byref(t), p(tr))
        ok = bh == z["box_hit"][i] and th == z["tri_hit"][i] and np.float32(t.value).view(np.uint32) == z["t_out"][i].view(np.uint32)
        bad += not ok
    assert bad == 0
    assert int(z["tri_hit"].sum()) > 500 and int(z["box_hit"].sum()) > 200      # the vectors exercise both outcomes


def test_camera_matrix_matches_reference_keyboard_path(golden):
    """HandleUpdates evaluates cos/sin in float (raythread.cpp:564-572 with float angles)."""
    import math
    for keys, g in golden["camera"].items():
        yaw = pitch = roll = np.float32(0)
        step = math.pi / 4 / 4
        for k in keys:
            if k == "y": yaw = np.float32(np.float64(yaw) + step)
            if k == "p": pitch = np.float32(np.float64(pitch) + step)
            if k == "r": roll = np.float32(np.float64(roll) + step)
        assert np.array_equal(O.camera_rotation(yaw, pitch, roll), np.array(g["rot"])), keys


def test_color_helpers_edge_cases():
    L = O.lib()
    assert L.ct_oracle_shade_color(0xFFFFFF, 2.0) == 0xFFFFFF          # clamp min(r*255, 255)
    assert L.ct_oracle_shade_color(0x000000, 0.7) == 0xB2B2B2          # black material: s == 0 -> grey of value v
    assert L.ct_oracle_blend(0x00FF00, 0x0000FF, 1.0) == 0x0000FF
    assert L.ct_oracle_blend(0x102030, 0x405060, 0.0) == 0x102030


def test_subsampling_frames_match_reference(golden, scene_loader):
    """settings.subsampling (raythread.cpp:512-531): the restatement against frames of the compiled
    reference run with one worker thread (tests/golden/make_golden_subsampling.py)."""
    import os
    from oracle import ct_oracle_py as O
    from conftest import GOLD
    assert len(golden["frames_subsampling"]) >= 4
    for case, m in golden["frames_subsampling"].items():
        fs = scene_loader(m["scene"])
        if m["force_reflection"] is not None:
            fs = fs.with_reflection(m["force_reflection"])
        frame, _, _ = O.OracleScene(fs).render(m["width"], m["height"], max_depth=m["depth"], flags=O.SUBSAMPLE, want_hits=False, n_threads=1)
        want = np.load(os.path.join(GOLD, f"frames_sub_{case}.npz"))["frame"]
        assert np.array_equal(frame, want), case


def test_supersampling_frames_match_patched_reference(golden, scene_loader):
    """settings.supersampling (raythread.cpp:460-505) with the jitter from the counter-based generator: the restatement
    against frames of the compiled reference run with --supersampling-hash on 1-3 worker threads."""
    import os
    from oracle import ct_oracle_py as O
    from conftest import GOLD
    assert len(golden["frames_supersampling"]) >= 3
    for case, m in golden["frames_supersampling"].items():
        fs = scene_loader(m["scene"])
        if m["force_reflection"] is not None:
            fs = fs.with_reflection(m["force_reflection"])
        frame, _, _ = O.OracleScene(fs).render(m["width"], m["height"], max_depth=m["depth"], flags=O.SUPERSAMPLE, want_hits=False)
        want = np.load(os.path.join(GOLD, f"frames_ss_{case}.npz"))["frame"]
        assert np.array_equal(frame, want), case


def test_session_frames_match_reference(golden, scene_loader):
    """the frames of the reference's interactive sessions (ct_ref --session: key presses through its
    own AddEvent / HandleKeyboard / HandleUpdates) are what the restatement renders from the recorded cameras."""
    import dataclasses
    from oracle import ct_oracle_py as O
    from cobbletrace_b200.sceneio import frame_fnv1a
    for case, m in golden["sessions"].items():
        fs = scene_loader(m["scene"])
        for k, t in enumerate(m["ticks"]):
            cam = dataclasses.replace(fs, cam_pos=np.array(t["pos"]), cam_rot=np.array(t["rot"]))
            frame, _, _ = O.OracleScene(cam).render(m["width"], m["height"], max_depth=m["depth"], want_hits=False)
            assert frame_fnv1a(frame) == t["fnv1a"], (case, k)


def test_tiny_and_narrow_frames_match_reference(golden, scene_loader):
    """Bitmap sizes the reference treats oddly (H = 1 traces nothing, odd H reaches row 0, W < H wraps columns into the
    neighbouring rows through PutPixel, draw2d.h:8-20): the restatement against frames of the compiled reference."""
    from oracle import ct_oracle_py as O
    assert len(golden["tiny_frames"]) >= 30
    for m in golden["tiny_frames"]:
        frame, _, _ = O.OracleScene(scene_loader(m["scene"])).render(m["width"], m["height"], max_depth=m["depth"], want_hits=False)
        assert np.array_equal(frame, np.array(m["frame"], np.uint32)), (m["scene"], m["width"], m["height"], m["depth"])


def test_random_scene_files_match_reference(golden):
    """20 generated scene files rendered by the compiled reference (tests/golden/make_golden_fuzz.py: triangle soups with
    mirrors, every specular kind, 1-6 lights of every type, odd and non-square sizes, depths 0-4, moved cameras): the
    restatement reproduces frames and primary hit records bit for bit."""
    from oracle import ct_oracle_py as O
    from conftest import load_fuzz_case
    assert len(golden["fuzz"]) >= 20
    for k, m in golden["fuzz"].items():
        fs, g = load_fuzz_case(k)
        frame, hits, _ = O.OracleScene(fs).render(m["width"], m["height"], max_depth=m["depth"])
        assert np.array_equal(frame, g["frame"]), k
        traced = g["found"] != 0xFFFFFFFF
        assert np.array_equal(hits["found"], g["found"]) and np.array_equal(hits["index"][traced], g["index"][traced]), k
        assert np.array_equal(hits["t"][traced].view(np.uint32), g["t"][traced].view(np.uint32)), k


def test_both_sampling_modes_at_once_match_patched_reference(golden, scene_loader):
    """settings.subsampling and settings.supersampling together (raythread.cpp:460-531): frames of the compiled reference
    (one thread, --subsampling --supersampling-hash) against the restatement."""
    import os
    from oracle import ct_oracle_py as O
    from conftest import GOLD
    assert len(golden["frames_both_sampling"]) >= 3
    for case, m in golden["frames_both_sampling"].items():
        fs = scene_loader(m["scene"])
        if m["force_reflection"] is not None:
            fs = fs.with_reflection(m["force_reflection"])
        frame, _, _ = O.OracleScene(fs).render(m["width"], m["height"], max_depth=m["depth"], flags=O.SUBSAMPLE | O.SUPERSAMPLE, want_hits=False, n_threads=1)
        assert np.array_equal(frame, np.load(os.path.join(GOLD, f"frames_both_{case}.npz"))["frame"]), case
