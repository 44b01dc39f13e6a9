#!/usr/bin/env python
"""Regenerates tests/golden/* from the UNMODIFIED reference compiled under oracle/_ref.

Run here (needs /root/reference):   make -C oracle ref && python tests/golden/make_golden.py

Outputs (all derived by executing the reference's own code, none hand-written):
  scenes/<name>.ctscene.{gz,xz}  flattened scene (+ BVH for the small ones) dumped by ct_ref --dump-scene
  frames_<case>.npz              frame + primary hit records of ct_ref --frame/--hits at low resolution
  kat_primitives.npz             IntersectTriangle / IntersectAABB answers of ct_ref --kat on edge-case vectors
  golden.json                    digests, 640x640 frame hashes (the SURVEY 8c hashes), counters, camera matrices
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cobbletrace_b200.sceneio import load_ctscene, save_ctscene, frame_fnv1a  # noqa: E402
from oracle import ct_oracle_py as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(O.REF_DIR, "scenes")
HIT_DT = O.HIT_DT

# name -> scene file, stored with BVH?
SCENE_FILES = {
    "scene_file_cube": ("scene_file_cube.json", True, ".gz"),
    "scene_import": ("scene_import.json", True, ".gz"),
    "pc_big": ("pc_big.json", True, ".gz"),
    "scene_import_bunny": ("scene_import_bunny.json", False, ".xz"),
}

# low-resolution frame cases: (case, scene, W, H, depth, force_reflection, keys)
FRAME_CASES = [
    ("cube_160", "scene_file_cube", 160, 160, 10, None, None),
    ("cube_rot_160", "scene_file_cube", 160, 160, 10, None, "yprwd"),
    ("cube_wide_200x120", "scene_file_cube", 200, 120, 10, None, None),
    ("cube_tall_90x150", "scene_file_cube", 90, 150, 3, None, None),
    ("import_160", "scene_import", 160, 160, 10, None, None),
    ("pc_big_96", "pc_big", 96, 96, 10, None, None),
    ("bunny_160", "scene_import_bunny", 160, 160, 10, None, None),
    ("bunny_refl_d2_160", "scene_import_bunny", 160, 160, 2, 0.5, None),
    ("bunny_refl_d10_128", "scene_import_bunny", 128, 128, 10, 0.25, None),
    ("bunny_odd_161x161", "scene_import_bunny", 161, 161, 10, None, None),
]


def run_ref(scene, **kw):
    return O.run_ref(scene, chdir=SCENES, **kw)


def ref_extra(args):
    subprocess.run([os.path.join(O.REF_DIR, "ct_ref")] + args, check=True, capture_output=True)


def kat_vectors(seed=1234):
    """Edge-case-heavy (ray, triangle, box) cases: NaN/inf slabs, |a| near 1e-4, u==0, u+v==1,
    t near 1e-4, t=0 rays, flat boxes, -0.0 directions, huge/tiny directions, float-midpoint quotients."""
    rng = np.random.default_rng(seed)
    org, dr, tri, mn, mx, t0 = [], [], [], [], [], []

    def add(o, d, tr, a, b, t):
        org.append(o); dr.append(d); tri.append(tr); mn.append(a); mx.append(b); t0.append(t)

    def rand_tri(scale=1.0):
        p = rng.uniform(-5, 5, 3)
        return np.concatenate([p, p + rng.uniform(-1, 1, 3) * scale, p + rng.uniform(-1, 1, 3) * scale])

    def rand_box():
        a = rng.uniform(-5, 5, 3); b = a + rng.uniform(0, 3, 3)
        return a, b

    ts = [1e30, 0.0, 1e-4, 0.5, 3.0, 4294967296.0]
    # 1. generic random
    for _ in range(2000):
        o = rng.uniform(-8, 8, 3); d = rng.normal(size=3)
        a, b = rand_box()
        add(o, d, rand_tri(rng.choice([1.0, 0.05, 3.0])), a, b, rng.choice(ts))
    # 2. rays aimed at exact barycentric boundary points of the triangle (u=0, v=0, u+v=1, vertices)
    for _ in range(1500):
        tr = rand_tri(rng.choice([1.0, 0.2]))
        p1, e1, e2 = tr[:3], tr[3:6] - tr[:3], tr[6:] - tr[:3]
        u, v = rng.choice([0.0, 1.0, 0.5, 0.25, 1e-9, 1 - 1e-9]), rng.choice([0.0, 0.5, 0.75, 1.0, 1e-9])
        if rng.random() < 0.5: v = 1.0 - u
        target = p1 + u * e1 + v * e2
        o = rng.uniform(-8, 8, 3)
        s = rng.choice([1.0, 1.0, 0.5, 2.0, -1.0, 1e4, 1e-4])      # hit distance t = 1/s scale, incl. behind and t ~ 1e-4
        a, b = rand_box()
        add(o, (target - o) * s, tr, a, b, rng.choice(ts))
    # 3. axis-aligned / zero / negative-zero direction components, origins on box planes (0/0 NaN slabs), flat boxes
    for _ in range(1500):
        a, b = rand_box()
        flat = rng.integers(0, 4)
        if flat < 3: b[flat] = a[flat]
        o = rng.uniform(-8, 8, 3); d = rng.normal(size=3)
        for ax in range(3):
            r = rng.random()
            if r < 0.35: d[ax] = 0.0
            elif r < 0.5: d[ax] = -0.0
            r = rng.random()
            if r < 0.3: o[ax] = a[ax]
            elif r < 0.5: o[ax] = b[ax]
            elif r < 0.6: o[ax] = 0.5 * (a[ax] + b[ax])
        tr = rand_tri()
        if rng.random() < 0.5:                                       # axis-aligned triangle in the box's flat plane
            ax = int(rng.integers(0, 3)); tr[ax] = tr[3 + ax] = tr[6 + ax] = a[ax]
        add(o, d, tr, a, b, rng.choice(ts))
    # 4. |a| around the 1e-4 parallel threshold: scale the triangle so that a = e1.(d x e2) lands near +-1e-4
    for _ in range(1000):
        tr = rand_tri(); o = rng.uniform(-8, 8, 3); d = rng.normal(size=3)
        e1, e2 = tr[3:6] - tr[:3], tr[6:] - tr[:3]
        a0 = float(np.dot(e1, np.cross(d, e2)))
        if abs(a0) < 1e-12: continue
        k = np.sqrt(abs(rng.choice([1e-4, 1.0000001e-4, 0.9999999e-4, 2e-4, 5e-5]) / a0))
        tr2 = np.concatenate([tr[:3], tr[:3] + e1 * k, tr[:3] + e2 * k])
        a, b = rand_box()
        add(o, d, tr2, a, b, rng.choice(ts))
    # 5. extreme direction magnitudes (reciprocal overflows / underflows: exact-division path)
    for _ in range(500):
        o = rng.uniform(-8, 8, 3); d = rng.normal(size=3) * 10.0 ** rng.choice([-320, -310, -300, -30, 30, 300, 305])
        a, b = rand_box()
        add(o, d, rand_tri(), a, b, rng.choice(ts))
    # 6. slab quotients sitting on float rounding midpoints: (b - 0)/d with b = fl(q*d), q = float midpoint (+- few ulps)
    for _ in range(1500):
        f = np.float32(rng.uniform(0.1, 50.0)) * np.float32(rng.choice([1, -1]))
        q = 0.5 * (np.float64(f) + np.float64(np.nextafter(f, np.float32(np.inf))))
        q = q + rng.integers(-3, 4) * np.spacing(q)
        d = rng.normal(size=3); d[d == 0] = 1.0
        o = np.zeros(3)
        b = q * d
        a = b - np.abs(rng.uniform(0, 2, 3))
        if rng.random() < 0.5: a, b = b - 0.0, b + np.abs(rng.uniform(0, 2, 3))
        add(o, d, rand_tri(), np.minimum(a, b), np.maximum(a, b), rng.choice(ts))
    A = lambda x, dt=np.float64: np.ascontiguousarray(np.array(x), dt)
    return A(org), A(dr), A(tri), A(mn), A(mx), A(t0, np.float32)


def main():
    if not O.ref_available():
        sys.exit("oracle/_ref/ct_ref missing: run `make -C oracle ref` first (needs /root/reference)")
    os.makedirs(os.path.join(GOLD, "scenes"), exist_ok=True)
    gold = {"scenes": {}, "frames": {}, "camera": {}}
    tmp = tempfile.mkdtemp(prefix="ctgold")
    flat = {}
    for name, (jf, with_bvh, ext) in SCENE_FILES.items():
        dump = os.path.join(tmp, name + ".ctscene")
        fr = os.path.join(tmp, name + ".frame")
        info = run_ref(jf, width=640, height=640, depth=10, threads=8, frame=fr, dump_scene=dump, counters=True)
        fs = load_ctscene(dump)
        flat[name] = fs
        frame = np.fromfile(fr, np.uint32).reshape(640, 640)
        gold["scenes"][name] = {
            "file": f"scenes/{name}.ctscene{ext}", "stored_with_bvh": with_bvh,
            "n_tri": fs.n_tri, "n_nodes": fs.n_nodes, "n_lights": fs.n_lights,
            "geometry_sha256": fs.geometry_digest(), "bvh_sha256": fs.bvh_digest(),
            "frame640_fnv1a": frame_fnv1a(frame), "frame640_background_px": int((frame == 0x333333).sum()),
            "frame640_zero_px": int((frame == 0).sum()),
            "counters640": {k: info[k] for k in ("rays_primary", "rays_shadow", "rays_reflection", "box_tests", "tri_tests")},
        }
        save_ctscene(os.path.join(GOLD, "scenes", f"{name}.ctscene{ext}"), fs if with_bvh else fs.without_bvh())
        print(name, gold["scenes"][name]["frame640_fnv1a"], flush=True)

    for case, scene, W, H, depth, refl, keys in FRAME_CASES:
        fr, hi, du = (os.path.join(tmp, case + e) for e in (".frame", ".hits", ".ctscene"))
        args = ["--scene", SCENE_FILES[scene][0], "--chdir", SCENES, "--width", str(W), "--height", str(H), "--depth", str(depth),
                "--threads", "1" if H % 2 else "2", "--frame", fr, "--hits", hi, "--dump-scene", du]
        if refl is not None: args += ["--force-reflection", repr(refl)]
        if keys: args += ["--keys", keys]
        ref_extra(args)
        frame = np.fromfile(fr, np.uint32).reshape(H, W)
        hits = np.fromfile(hi, HIT_DT).reshape(H, W)
        cam = load_ctscene(du)
        np.savez_compressed(os.path.join(GOLD, f"frames_{case}.npz"), frame=frame, found=hits["found"], index=hits["index"], t=hits["t"])
        gold["frames"][case] = {"scene": scene, "width": W, "height": H, "depth": depth, "force_reflection": refl, "keys": keys,
                                "cam_pos": cam.cam_pos.tolist(), "cam_rot": cam.cam_rot.tolist(), "fnv1a": frame_fnv1a(frame)}
        print(case, gold["frames"][case]["fnv1a"], flush=True)

    # camera matrices from the reference's HandleUpdates for a few key sequences (float cos/sin!)
    for keys in ("y", "p", "r", "yyp", "yprwd", "rrrrpppyyyyy"):
        du = os.path.join(tmp, "cam.ctscene")
        ref_extra(["--scene", "scene_file_cube.json", "--chdir", SCENES, "--width", "16", "--height", "16", "--threads", "1", "--keys", keys, "--dump-scene", du])
        c = load_ctscene(du)
        gold["camera"][keys] = {"pos": c.cam_pos.tolist(), "rot": c.cam_rot.tolist()}

    # primitive KATs
    org, dr, tri, mn, mx, t0 = kat_vectors()
    n = org.shape[0]
    fin, fout = os.path.join(tmp, "kat.in"), os.path.join(tmp, "kat.out")
    with open(fin, "wb") as f:
        f.write(np.uint32(n).tobytes())
        for a in (org, dr, tri, mn, mx, t0): f.write(a.tobytes())
    subprocess.run([os.path.join(O.REF_DIR, "ct_ref"), "--kat", fin, fout], check=True)
    out = np.fromfile(fout, np.uint32)
    tri_hit, box_hit, t_out = out[:n], out[n:2 * n], out[2 * n:].view(np.float32)
    np.savez_compressed(os.path.join(GOLD, "kat_primitives.npz"), org=org, dir=dr, tri=tri, bmin=mn, bmax=mx, t0=t0,
                        tri_hit=tri_hit.astype(np.uint8), box_hit=box_hit.astype(np.uint8), t_out=t_out)
    gold["kat"] = {"n": int(n), "tri_hits": int(tri_hit.sum()), "box_hits": int(box_hit.sum())}
    print("kat", gold["kat"])
    with open(os.path.join(GOLD, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
