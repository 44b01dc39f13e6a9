"""Procedural scenes for measurement (BASELINE.json: "procedurally generated and bundled meshes only").

The headline config needs the Stanford dragon (~871k triangles), which is not shipped with the
reference.  The stand-in follows SURVEY 8(d): octahedron-subdivision spheres (the rule of
the reference's utils/genRasterSphere.cpp: split every triangle at its edge midpoints, push the
midpoints onto the unit sphere; 8*4^d triangles at depth d) composed to ~868k triangles, written as an
ASCII PLY + scene JSON with the dragon scene's transform, material, lights and camera
(scene_import_dragon.json), so that BOTH the reference's parser and ours ingest the same bytes.

Numbers are printed in fixed notation without '+' or exponents: the reference's number parser
(fileBuffer.cpp:160-207) understands digits, '.', 'e' and '-' only.
"""
from __future__ import annotations

import json
import os
from typing import List, Sequence, Tuple

import numpy as np


def octasphere(depth: int) -> Tuple[np.ndarray, np.ndarray]:
    """Unit sphere by octahedron subdivision: returns (vertices float64 [V,3], faces int32 [8*4^depth,3])."""
    v = np.array([[1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float64)
    f = np.array([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4], [0, 1, 5], [1, 2, 5], [2, 3, 5], [3, 0, 5]], np.int64)
    for _ in range(depth):
        nv = v.shape[0]
        # unique undirected edges -> one shared midpoint vertex each
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
        e.sort(axis=1)
        key = e[:, 0] * nv + e[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a, b = uniq // nv, uniq % nv
        mid = (v[a] + v[b]) / 2.0
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], 0)
        nf = f.shape[0]
        m01, m12, m20 = nv + inv[:nf], nv + inv[nf:2 * nf], nv + inv[2 * nf:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([m01, f[:, 1], m12], 1),
                            np.stack([m20, m01, m12], 1), np.stack([m20, m12, f[:, 2]], 1)], 0)
    return v, f.astype(np.int32)


# (depth, radius, centre) in MODEL units; the scene's import scale (90) and rotation are applied by the loader.
# 8*(4^8 + 2*4^7 + 2*4^6 + 2*4^5) = 868 352 triangles (the dragon has 871 414).
DRAGON_STANDIN_PARTS: List[Tuple[int, float, Tuple[float, float, float]]] = [
    (8, 0.070, (0.000, 0.125, 0.000)),
    (7, 0.040, (-0.085, 0.105, -0.010)),
    (7, 0.040, (0.085, 0.105, -0.010)),
    (6, 0.025, (-0.045, 0.190, -0.030)),
    (6, 0.025, (0.045, 0.190, -0.030)),
    (5, 0.015, (-0.030, 0.070, -0.075)),
    (5, 0.015, (0.030, 0.070, -0.075)),
]


def compose(parts: Sequence[Tuple[int, float, Tuple[float, float, float]]]) -> Tuple[np.ndarray, np.ndarray]:
    vs, fs, base = [], [], 0
    cache = {}
    for depth, radius, centre in parts:
        if depth not in cache:
            cache[depth] = octasphere(depth)
        v, f = cache[depth]
        vs.append(v * radius + np.asarray(centre, np.float64))
        fs.append(f + base)
        base += v.shape[0]
    return np.concatenate(vs, 0), np.concatenate(fs, 0)


def write_ascii_ply(path: str, verts: np.ndarray, faces: np.ndarray) -> None:
    """ASCII PLY the reference's ImportPlyObject (objectLoader.cpp:142-202) can read."""
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment cobbletrace_b200 procedural mesh\n")
        f.write(f"element vertex {verts.shape[0]}\nproperty float32 x\nproperty float32 y\nproperty float32 z\n")
        f.write(f"element face {faces.shape[0]}\nproperty list uint8 int32 vertex_indices\nend_header\n")
        # fixed notation, 8 decimals, no exponent / '+' (see module docstring)
        np.savetxt(f, verts, fmt="%.8f")
        np.savetxt(f, np.concatenate([np.full((faces.shape[0], 1), 3, np.int32), faces], 1), fmt="%d")


def dragon_scene_json(model_relpath: str, reflection: float = 0.0, threads: int = 32) -> str:
    """scene_import_dragon.json's parameters for a stand-in mesh."""
    scene = {
        "objects": [{
            "type": "import", "filename": model_relpath, "format": "ply",
            "position": [0, 0, 0], "rotation": [0.392, 3.14, 0], "scale": [90, 90, 90],
            "color": [255, 155, 255], "specular": 100, "reflection": reflection,
        }],
        "lights": [
            {"type": "ambient", "intensity": 0.2},
            {"type": "point", "intensity": 0.6, "position": [2, 11, -30]},
            {"type": "directional", "intensity": 0.2, "direction": [-10, 4, -14]},
        ],
        "camera": {"position": [0.50, 8.5, -26.4]},
        "settings": {"numberOfThreads": threads, "subsampling": False, "wireframe": False, "supersampling": False},
    }
    # the reference's parser wants exactly this key order inside "camera" and lower-case booleans: json.dumps gives both
    return json.dumps(scene, indent=2)


def write_dragon_standin(out_dir: str, parts: Sequence = DRAGON_STANDIN_PARTS, name: str = "dragon_standin",
                         reflection: float = 0.0) -> Tuple[str, int]:
    """Writes <out_dir>/models/<name>.ply and <out_dir>/scene_<name>.json; returns (scene path, triangle count)."""
    os.makedirs(os.path.join(out_dir, "models"), exist_ok=True)
    verts, faces = compose(parts)
    ply_rel = f"models/{name}.ply"
    ply = os.path.join(out_dir, ply_rel)
    scene = os.path.join(out_dir, f"scene_{name}.json")
    if not (os.path.exists(ply) and os.path.exists(scene)):
        write_ascii_ply(ply + ".tmp", verts, faces)
        os.replace(ply + ".tmp", ply)
        with open(scene, "w") as f:
            f.write(dragon_scene_json(ply_rel, reflection))
    return scene, int(faces.shape[0])


def small_standin_parts(depth_main: int = 4) -> List[Tuple[int, float, Tuple[float, float, float]]]:
    """Same composition at low subdivision depth (for tests)."""
    d = depth_main
    return [(max(d - (8 - p[0]), 1), p[1], p[2]) for p in DRAGON_STANDIN_PARTS]
