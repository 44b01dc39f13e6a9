"""Flattened scene container (`.ctscene`) used at the upload boundary.

A flattened scene is exactly what crosses the drop-in boundary behind
``RayThread`` (reference raythread.cpp:641-654): the triangle soup
in ``GetSceneTriangles`` order (raythread.cpp:621), the per-triangle material
(``objects[triangleLookup[k]].material``, raythread.cpp:211), the lights in file
order, the camera, and -- optionally -- the BVH in the reference's node layout
(``bvh_node_t`` bvh.h:5-11 + ``indexes`` bvh.h:16).

File layout (little endian)::

    "CTSCENE1"  u32 nTri  u32 nLights  u32 nNodes  u32 0
    f64[3] camera position, f64[9] camera rotation (row-major data[i][j])
    f64[nTri*9] p1,p2,p3
    {u32 color(0x00BBGGRR), i32 specular, f32 reflection}[nTri]
    {i32 type(0 point,1 directional,2 ambient), f32 intensity, f64[3] pos, f64[3] dir}[nLights]
    {f64[3] min, f64[3] max, u32 left, u32 first, u32 count, u32 0}[nNodes]
    u32[nTri] indexes            (only when nNodes > 0)
"""
from __future__ import annotations

import dataclasses
import gzip
import hashlib
import lzma
from typing import Optional

import numpy as np

LT_POINT, LT_DIRECTIONAL, LT_AMBIENT = 0, 1, 2

_MAT_DT = np.dtype([("color", "<u4"), ("specular", "<i4"), ("reflection", "<f4")])
_LIGHT_DT = np.dtype([("type", "<i4"), ("intensity", "<f4"), ("pos", "<f8", 3), ("dir", "<f8", 3)])
_NODE_DT = np.dtype([("min", "<f8", 3), ("max", "<f8", 3), ("left", "<u4"), ("first", "<u4"), ("count", "<u4"), ("pad", "<u4")])


@dataclasses.dataclass
class FlatScene:
    tri: np.ndarray                 # (nTri, 9) float64
    mat_color: np.ndarray           # (nTri,) uint32
    mat_specular: np.ndarray        # (nTri,) int32
    mat_reflection: np.ndarray      # (nTri,) float32
    light_type: np.ndarray          # (nLights,) int32
    light_intensity: np.ndarray     # (nLights,) float32
    light_pos: np.ndarray           # (nLights, 3) float64
    light_dir: np.ndarray           # (nLights, 3) float64
    cam_pos: np.ndarray             # (3,) float64
    cam_rot: np.ndarray             # (9,) float64
    node_min: Optional[np.ndarray] = None    # (nNodes, 3) float64
    node_max: Optional[np.ndarray] = None
    node_left: Optional[np.ndarray] = None   # (nNodes,) uint32
    node_first: Optional[np.ndarray] = None
    node_count: Optional[np.ndarray] = None
    tri_index: Optional[np.ndarray] = None   # (nTri,) uint32

    @property
    def n_tri(self) -> int:
        return int(self.tri.shape[0])

    @property
    def n_lights(self) -> int:
        return int(self.light_type.shape[0])

    @property
    def n_nodes(self) -> int:
        return 0 if self.node_left is None else int(self.node_left.shape[0])

    def has_bvh(self) -> bool:
        return self.node_left is not None

    def without_bvh(self) -> "FlatScene":
        return dataclasses.replace(self, node_min=None, node_max=None, node_left=None, node_first=None,
                                   node_count=None, tri_index=None)

    def with_reflection(self, reflection: float) -> "FlatScene":
        """Harness-level material override used by BASELINE config 3."""
        return dataclasses.replace(self, mat_reflection=np.full(self.n_tri, reflection, dtype=np.float32))

    def geometry_digest(self) -> str:
        h = hashlib.sha256()
        for a in (self.tri, self.mat_color, self.mat_specular, self.mat_reflection, self.light_type,
                  self.light_intensity, self.light_pos, self.light_dir, self.cam_pos, self.cam_rot):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()

    def bvh_digest(self) -> str:
        h = hashlib.sha256()
        for a in (self.node_min, self.node_max, self.node_left, self.node_first, self.node_count, self.tri_index):
            h.update(np.ascontiguousarray(a).tobytes())
        return h.hexdigest()


def _open(path: str, mode: str):
    if str(path).endswith(".gz"):
        return gzip.open(path, mode)
    if str(path).endswith(".xz"):
        return lzma.open(path, mode)
    return open(path, mode)


def load_ctscene(path: str) -> FlatScene:
    with _open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != b"CTSCENE1":
        raise ValueError(f"{path}: not a CTSCENE1 file")
    n_tri, n_lights, n_nodes, _ = np.frombuffer(buf, "<u4", 4, 8)
    n_tri, n_lights, n_nodes = int(n_tri), int(n_lights), int(n_nodes)
    off = 24
    cam = np.frombuffer(buf, "<f8", 12, off).copy(); off += 96
    tri = np.frombuffer(buf, "<f8", n_tri * 9, off).reshape(n_tri, 9).copy(); off += n_tri * 72
    mat = np.frombuffer(buf, _MAT_DT, n_tri, off); off += n_tri * _MAT_DT.itemsize
    lights = np.frombuffer(buf, _LIGHT_DT, n_lights, off); off += n_lights * _LIGHT_DT.itemsize
    fs = FlatScene(
        tri=tri, mat_color=mat["color"].copy(), mat_specular=mat["specular"].copy(), mat_reflection=mat["reflection"].copy(),
        light_type=lights["type"].copy(), light_intensity=lights["intensity"].copy(),
        light_pos=lights["pos"].copy().reshape(n_lights, 3), light_dir=lights["dir"].copy().reshape(n_lights, 3),
        cam_pos=cam[:3].copy(), cam_rot=cam[3:].copy())
    if n_nodes:
        nodes = np.frombuffer(buf, _NODE_DT, n_nodes, off); off += n_nodes * _NODE_DT.itemsize
        fs.node_min = nodes["min"].copy(); fs.node_max = nodes["max"].copy()
        fs.node_left = nodes["left"].copy(); fs.node_first = nodes["first"].copy(); fs.node_count = nodes["count"].copy()
        fs.tri_index = np.frombuffer(buf, "<u4", n_tri, off).copy(); off += n_tri * 4
    if off != len(buf):
        raise ValueError(f"{path}: trailing bytes ({len(buf) - off})")
    return fs


def save_ctscene(path: str, fs: FlatScene) -> None:
    with _open(path, "wb") as f:
        f.write(b"CTSCENE1")
        f.write(np.array([fs.n_tri, fs.n_lights, fs.n_nodes, 0], "<u4").tobytes())
        f.write(np.asarray(fs.cam_pos, "<f8").tobytes())
        f.write(np.asarray(fs.cam_rot, "<f8").tobytes())
        f.write(np.ascontiguousarray(fs.tri, "<f8").tobytes())
        mat = np.zeros(fs.n_tri, _MAT_DT)
        mat["color"], mat["specular"], mat["reflection"] = fs.mat_color, fs.mat_specular, fs.mat_reflection
        f.write(mat.tobytes())
        lights = np.zeros(fs.n_lights, _LIGHT_DT)
        lights["type"], lights["intensity"] = fs.light_type, fs.light_intensity
        lights["pos"], lights["dir"] = fs.light_pos, fs.light_dir
        f.write(lights.tobytes())
        if fs.has_bvh():
            nodes = np.zeros(fs.n_nodes, _NODE_DT)
            nodes["min"], nodes["max"] = fs.node_min, fs.node_max
            nodes["left"], nodes["first"], nodes["count"] = fs.node_left, fs.node_first, fs.node_count
            f.write(nodes.tobytes())
            f.write(np.ascontiguousarray(fs.tri_index, "<u4").tobytes())


def frame_fnv1a(frame: np.ndarray) -> str:
    """64-bit FNV-1a over uint32 pixels in row-major order (the hash SURVEY 8c records goldens with)."""
    h = 1469598103934665603
    mask = (1 << 64) - 1
    for p in np.ascontiguousarray(frame, dtype=np.uint32).ravel().tolist():
        h = ((h ^ p) * 1099511628211) & mask
    return "%016x" % h
