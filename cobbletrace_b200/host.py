"""ctypes binding of include/ct_host.h (libct_host.so): scene ingest, BVH build and the boss loop.

Host code is C++ like the reference's; Python only marshals arrays.  The boss renders through
libct_gpu.so (dlopen'ed by the C++ side) -- no CPU rendering path exists.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import api
from .sceneio import FlatScene

PKG = os.path.dirname(os.path.abspath(__file__))
HOST_LIB = os.path.join(PKG, "libct_host.so")

ABI_SYMBOLS = [
    "ct_host_last_error", "ct_host_scene_load", "ct_host_scene_from_arrays", "ct_host_scene_free",
    "ct_host_scene_share", "ct_host_scene_attach", "ct_host_scene_unshare",
    "ct_host_scene_triangle_count", "ct_host_scene_sphere_count", "ct_host_scene_triangles", "ct_host_scene_materials",
    "ct_host_scene_light_count", "ct_host_scene_lights", "ct_host_scene_camera", "ct_host_scene_set_camera",
    "ct_host_scene_settings", "ct_host_scene_set_reflection", "ct_host_build_bvh", "ct_host_scene_nodes",
    "ct_host_scene_tri_indexes", "ct_host_scene_set_bvh", "ct_host_camera_rotation", "ct_host_fill_desc",
    "ct_host_boss_create", "ct_host_boss_set_stream", "ct_host_boss_set_camera", "ct_host_boss_render", "ct_host_boss_reset_shared_counter",
    "ct_host_boss_tiles", "ct_host_boss_destroy",
    "ct_host_tile_counter_open", "ct_host_tile_counter_next", "ct_host_tile_counter_reset", "ct_host_tile_counter_close",
    "ct_host_controls_create", "ct_host_controls_add_event", "ct_host_controls_update", "ct_host_controls_pending",
    "ct_host_controls_frames", "ct_host_controls_camera", "ct_host_controls_destroy", "ct_host_viewer_tick", "ct_host_write_ppm", "ct_host_scene_render_flags",
]


class HostSettings(C.Structure):
    _fields_ = [("number_of_threads", C.c_int32), ("subsampling", C.c_int32), ("wireframe", C.c_int32), ("supersampling", C.c_int32)]


class BossConfig(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("width", C.c_int32), ("height", C.c_int32), ("max_depth", C.c_int32),
                ("flags", C.c_uint32), ("n_devices", C.c_int32), ("devices", C.c_int32 * 16), ("tile_rows", C.c_int32),
                ("shared_counter_name", C.c_char_p), ("rank", C.c_int32), ("world_size", C.c_int32), ("gpu_library", C.c_char_p)]


class FrameStats(C.Structure):
    _fields_ = [("rays", api.RayCounters), ("device_ms_max", C.c_float), ("wall_ms", C.c_double),
                ("tiles_total", C.c_int32), ("tiles_mine", C.c_int32), ("kernel_launches", C.c_uint64)]


PRESENT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int)   # ct_host_present_fn
EVENT_KEY_UP, EVENT_KEY_DOWN = 2, 3                                         # eventType_t, eventQueue.h:7-12

_lib = None


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_LIB):
            raise RuntimeError(f"{HOST_LIB} not found: build it with `python -m cobbletrace_b200.build`")
        L = C.CDLL(HOST_LIB)
        vp = C.c_void_p
        L.ct_host_last_error.restype = C.c_char_p
        L.ct_host_scene_load.restype = vp; L.ct_host_scene_load.argtypes = [C.c_char_p, C.c_char_p]
        L.ct_host_scene_from_arrays.restype = vp
        L.ct_host_scene_from_arrays.argtypes = [C.c_uint32, vp, vp, C.c_uint32, vp, vp, vp]
        L.ct_host_scene_free.argtypes = [vp]
        L.ct_host_scene_share.argtypes = [vp, C.c_char_p]
        L.ct_host_scene_attach.restype = vp; L.ct_host_scene_attach.argtypes = [C.c_char_p]
        L.ct_host_scene_unshare.argtypes = [C.c_char_p]
        for f in ("ct_host_scene_triangle_count", "ct_host_scene_sphere_count", "ct_host_scene_light_count"):
            getattr(L, f).restype = C.c_uint32; getattr(L, f).argtypes = [vp]
        for f in ("ct_host_scene_triangles", "ct_host_scene_materials", "ct_host_scene_lights", "ct_host_scene_tri_indexes"):
            getattr(L, f).restype = vp; getattr(L, f).argtypes = [vp]
        L.ct_host_scene_camera.argtypes = [vp, vp, vp]
        L.ct_host_scene_set_camera.argtypes = [vp, vp, vp]
        L.ct_host_scene_settings.argtypes = [vp, C.POINTER(HostSettings)]
        L.ct_host_scene_set_reflection.argtypes = [vp, C.c_float]
        L.ct_host_build_bvh.argtypes = [vp]
        L.ct_host_scene_nodes.restype = vp; L.ct_host_scene_nodes.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.ct_host_scene_set_bvh.argtypes = [vp, C.c_uint32, vp, vp]
        L.ct_host_camera_rotation.argtypes = [C.c_float, C.c_float, C.c_float, vp]
        L.ct_host_fill_desc.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(api.SceneDesc)]
        L.ct_host_boss_create.restype = vp; L.ct_host_boss_create.argtypes = [vp, C.POINTER(BossConfig)]
        L.ct_host_boss_set_camera.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float]
        L.ct_host_boss_set_stream.argtypes = [vp, C.c_int, vp]
        L.ct_host_boss_render.argtypes = [vp, vp, C.c_int, C.POINTER(FrameStats)]
        L.ct_host_boss_reset_shared_counter.argtypes = [vp]
        L.ct_host_boss_tiles.argtypes = [vp, vp, C.c_int]
        L.ct_host_boss_destroy.argtypes = [vp]
        L.ct_host_controls_create.restype = vp; L.ct_host_controls_create.argtypes = [vp, C.c_uint32]
        L.ct_host_controls_add_event.argtypes = [vp, C.c_uint32, C.c_uint32]
        L.ct_host_controls_update.argtypes = [vp]
        L.ct_host_controls_pending.restype = C.c_uint32; L.ct_host_controls_pending.argtypes = [vp]
        L.ct_host_controls_frames.restype = C.c_uint64; L.ct_host_controls_frames.argtypes = [vp]
        L.ct_host_controls_camera.argtypes = [vp, vp, vp, vp]
        L.ct_host_controls_destroy.argtypes = [vp]
        L.ct_host_viewer_tick.argtypes = [vp, vp, vp, C.c_int, PRESENT_FN, vp, C.POINTER(C.c_int), C.POINTER(FrameStats)]
        L.ct_host_scene_render_flags.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.ct_host_write_ppm.argtypes = [C.c_char_p, vp, C.c_int, C.c_int, C.c_int]
        L.ct_host_tile_counter_open.restype = vp; L.ct_host_tile_counter_open.argtypes = [C.c_char_p]
        L.ct_host_tile_counter_next.restype = C.c_int32; L.ct_host_tile_counter_next.argtypes = [vp]
        L.ct_host_tile_counter_reset.argtypes = [vp]
        L.ct_host_tile_counter_close.argtypes = [vp, C.c_int]
        _lib = L
    return _lib


def _err(L) -> str:
    return L.ct_host_last_error().decode(errors="replace")


def _np_from(ptr, dtype, count):
    if count == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (np.dtype(dtype).itemsize * count)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=count).copy()


def camera_rotation(yaw: float = 0.0, pitch: float = 0.0, roll: float = 0.0) -> np.ndarray:
    out = np.zeros(9)
    load_library().ct_host_camera_rotation(yaw, pitch, roll, out.ctypes.data_as(C.c_void_p))
    return out


class HostScene:
    """Owns a ct_host_scene (C++ cth::Scene)."""

    def __init__(self, handle):
        self.L = load_library()
        self.h = handle

    @classmethod
    def load(cls, scene_file: str, base_dir: Optional[str] = None) -> "HostScene":
        L = load_library()
        h = L.ct_host_scene_load(os.fsencode(scene_file), os.fsencode(base_dir) if base_dir else None)
        if not h:
            raise RuntimeError("ct_host_scene_load: " + _err(L))
        return cls(h)

    @classmethod
    def from_flat(cls, fs: FlatScene) -> "HostScene":
        L = load_library()
        tri = np.ascontiguousarray(fs.tri, np.float64)
        mats = np.zeros(fs.n_tri, api.MAT_DT)
        mats["color"], mats["specular"], mats["reflection"] = fs.mat_color, fs.mat_specular, fs.mat_reflection
        lights = np.zeros(max(fs.n_lights, 1), api.LIGHT_DT)
        if fs.n_lights:
            lights["type"][:fs.n_lights], lights["intensity"][:fs.n_lights] = fs.light_type, fs.light_intensity
            lights["pos"][:fs.n_lights], lights["dir"][:fs.n_lights] = fs.light_pos, fs.light_dir
        pos = np.ascontiguousarray(fs.cam_pos, np.float64); rot = np.ascontiguousarray(fs.cam_rot, np.float64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        h = L.ct_host_scene_from_arrays(fs.n_tri, p(tri), p(mats), fs.n_lights, p(lights), p(pos), p(rot))
        if not h:
            raise RuntimeError("ct_host_scene_from_arrays: " + _err(L))
        s = cls(h)
        if fs.has_bvh():
            nodes = api.pack_nodes(fs); idx = np.ascontiguousarray(fs.tri_index, np.uint32)
            if L.ct_host_scene_set_bvh(h, fs.n_nodes, p(nodes), p(idx)) < 0:
                raise RuntimeError("ct_host_scene_set_bvh: " + _err(L))
        return s

    def share(self, name: str) -> None:
        """Scene + BVH as built into the POSIX shared-memory object `name` (ct_host_scene_share): the other processes of a
        multi-GPU job attach() instead of parsing and building again."""
        if self.L.ct_host_scene_share(self.h, name.encode()) < 0:
            raise RuntimeError("ct_host_scene_share: " + _err(self.L))

    @classmethod
    def attach(cls, name: str) -> "HostScene":
        L = load_library()
        h = L.ct_host_scene_attach(name.encode())
        if not h:
            raise RuntimeError("ct_host_scene_attach: " + _err(L))
        return cls(h)

    @staticmethod
    def unshare(name: str) -> None:
        load_library().ct_host_scene_unshare(name.encode())

    def __del__(self):
        try:
            if self.h:
                self.L.ct_host_scene_free(self.h); self.h = None
        except Exception:
            pass

    @property
    def n_tri(self) -> int:
        return int(self.L.ct_host_scene_triangle_count(self.h))

    @property
    def n_spheres(self) -> int:
        return int(self.L.ct_host_scene_sphere_count(self.h))

    def settings(self) -> dict:
        s = HostSettings(); self.L.ct_host_scene_settings(self.h, C.byref(s))
        return dict(numberOfThreads=s.number_of_threads, subsampling=bool(s.subsampling), wireframe=bool(s.wireframe),
                    supersampling=bool(s.supersampling))

    def render_flags(self) -> int:
        """CT_FLAG_* implied by the scene file's settings (ct_host_scene_render_flags)."""
        f = C.c_uint32(0)
        if self.L.ct_host_scene_render_flags(self.h, C.byref(f)) < 0:
            raise RuntimeError("ct_host_scene_render_flags: " + _err(self.L))
        return int(f.value)

    @property
    def n_lights(self) -> int:
        return int(self.L.ct_host_scene_light_count(self.h))

    def camera(self):
        """(position[3], rotation[9]) as float64 arrays (ct_host_scene_camera)."""
        pos, rot = np.zeros(3), np.zeros(9)
        self.L.ct_host_scene_camera(self.h, pos.ctypes.data_as(C.c_void_p), rot.ctypes.data_as(C.c_void_p))
        return pos, rot

    def set_reflection(self, reflection: float):
        self.L.ct_host_scene_set_reflection(self.h, reflection)

    def set_camera(self, pos=None, rot=None):
        p = None if pos is None else np.ascontiguousarray(pos, np.float64)
        r = None if rot is None else np.ascontiguousarray(rot, np.float64)
        self.L.ct_host_scene_set_camera(self.h, None if p is None else p.ctypes.data_as(C.c_void_p),
                                        None if r is None else r.ctypes.data_as(C.c_void_p))

    def build_bvh(self) -> int:
        n = self.L.ct_host_build_bvh(self.h)
        if n < 0:
            raise RuntimeError("ct_host_build_bvh: " + _err(self.L))
        return n

    def to_flat(self, with_bvh: bool = True) -> FlatScene:
        L, h = self.L, self.h
        n, nl = self.n_tri, int(L.ct_host_scene_light_count(h))
        tri = _np_from(L.ct_host_scene_triangles(h), np.float64, n * 9).reshape(n, 9)
        mats = _np_from(L.ct_host_scene_materials(h), api.MAT_DT, n)
        lights = _np_from(L.ct_host_scene_lights(h), api.LIGHT_DT, nl) if nl else np.zeros(0, api.LIGHT_DT)
        pos = np.zeros(3); rot = np.zeros(9)
        L.ct_host_scene_camera(h, pos.ctypes.data_as(C.c_void_p), rot.ctypes.data_as(C.c_void_p))
        fs = FlatScene(tri=tri, mat_color=mats["color"].copy(), mat_specular=mats["specular"].copy(),
                       mat_reflection=mats["reflection"].copy(), light_type=lights["type"].copy(),
                       light_intensity=lights["intensity"].copy(), light_pos=lights["pos"].copy().reshape(nl, 3),
                       light_dir=lights["dir"].copy().reshape(nl, 3), cam_pos=pos, cam_rot=rot)
        if with_bvh:
            self.build_bvh()
            nn = C.c_uint32()
            nodes = _np_from(L.ct_host_scene_nodes(h, C.byref(nn)), api.NODE_DT, 0) if False else None
            ptr = L.ct_host_scene_nodes(h, C.byref(nn))
            nodes = _np_from(ptr, api.NODE_DT, int(nn.value))
            fs.node_min, fs.node_max = nodes["min"].copy(), nodes["max"].copy()
            fs.node_left, fs.node_first, fs.node_count = nodes["left"].copy(), nodes["first"].copy(), nodes["count"].copy()
            fs.tri_index = _np_from(L.ct_host_scene_tri_indexes(h), np.uint32, n)
        return fs


class Boss:
    """RayThread's boss half over one or more GPUs of this process (ct_host_boss_*)."""

    def __init__(self, scene: HostScene, width: int, height: int, devices: Sequence[int] = (0,), max_depth: int = api.REFERENCE_MAX_DEPTH,
                 flags: int = 0, tile_rows: int = 0, shared_counter: Optional[str] = None, rank: int = 0, world_size: int = 1):
        self.L = load_library()
        self.scene = scene
        self.width, self.height = width, height
        cfg = BossConfig()
        cfg.struct_size = C.sizeof(BossConfig)
        cfg.width, cfg.height, cfg.max_depth, cfg.flags = width, height, max_depth, flags
        cfg.n_devices = len(devices)
        for i, d in enumerate(devices):
            cfg.devices[i] = d
        cfg.tile_rows = tile_rows
        self._name = shared_counter.encode() if shared_counter else None
        cfg.shared_counter_name = self._name
        cfg.rank, cfg.world_size = rank, world_size
        self._gpu = os.fsencode(api.GPU_LIB)
        cfg.gpu_library = self._gpu
        self.h = self.L.ct_host_boss_create(scene.h, C.byref(cfg))
        if not self.h:
            raise RuntimeError("ct_host_boss_create: " + _err(self.L))

    def set_camera(self, pos, yaw=0.0, pitch=0.0, roll=0.0):
        p = np.ascontiguousarray(pos, np.float64)
        if self.L.ct_host_boss_set_camera(self.h, p.ctypes.data_as(C.c_void_p), yaw, pitch, roll) < 0:
            raise RuntimeError("ct_host_boss_set_camera: " + _err(self.L))

    def set_stream(self, cuda_stream_ptr, slot: int = 0):
        if self.L.ct_host_boss_set_stream(self.h, slot, C.c_void_p(cuda_stream_ptr or 0)) < 0:
            raise RuntimeError("ct_host_boss_set_stream: " + _err(self.L))

    def reset_shared_counter(self):
        self.L.ct_host_boss_reset_shared_counter(self.h)

    def render(self, bitmap: Optional[np.ndarray] = None, want_bitmap: bool = True):
        if bitmap is None and want_bitmap:
            bitmap = np.zeros((self.height, self.width), np.uint32)
        st = FrameStats()
        rc = self.L.ct_host_boss_render(self.h, bitmap.ctypes.data_as(C.c_void_p) if bitmap is not None else None,
                                        bitmap.shape[1] if bitmap is not None else 0, C.byref(st))
        if rc < 0:
            raise RuntimeError("ct_host_boss_render: " + _err(self.L))
        stats = dict(st.rays.as_dict(), wall_ms=float(st.wall_ms), tiles_total=int(st.tiles_total), tiles_mine=int(st.tiles_mine),
                     kernel_launches=int(st.kernel_launches), device_ms_max=float(st.device_ms_max))
        return bitmap, stats

    def tiles(self):
        n = self.L.ct_host_boss_tiles(self.h, None, 0)
        out = np.zeros((max(n, 1), 2), np.int32)
        self.L.ct_host_boss_tiles(self.h, out.ctypes.data_as(C.c_void_p), n)
        return [tuple(map(int, r)) for r in out[:n]]

    def close(self):
        if self.h:
            self.L.ct_host_boss_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_ppm(path: str, bitmap: np.ndarray):
    """ct_host_write_ppm: the bitmap (uint32 0x00BBGGRR, [H, W]) as a binary PPM."""
    L = load_library()
    b = np.ascontiguousarray(bitmap, np.uint32)
    if L.ct_host_write_ppm(os.fsencode(path), b.ctypes.data_as(C.c_void_p), b.shape[1], b.shape[0], b.shape[1]) < 0:
        raise RuntimeError("ct_host_write_ppm: " + _err(L))


class Controls:
    """The event queue + HandleKeyboard + HandleUpdates' camera state (ct_host_controls_*); host state only."""

    def __init__(self, scene: HostScene, event_capacity: int = 0):
        self.L = load_library()
        self.scene = scene
        self.h = self.L.ct_host_controls_create(scene.h, event_capacity)
        if not self.h:
            raise RuntimeError("ct_host_controls_create: " + _err(self.L))

    def key(self, ch: str, event_type: int = EVENT_KEY_DOWN):
        """One SDL_KEYDOWN (cobbletrace.cpp:103-105)."""
        if self.L.ct_host_controls_add_event(self.h, event_type, ord(ch)) < 0:
            raise RuntimeError("ct_host_controls_add_event: " + _err(self.L))

    def keys(self, s: str):
        for ch in s:
            self.key(ch)

    def update(self) -> bool:
        rc = self.L.ct_host_controls_update(self.h)
        if rc < 0:
            raise RuntimeError("ct_host_controls_update: " + _err(self.L))
        return bool(rc)

    @property
    def pending(self) -> int:
        return int(self.L.ct_host_controls_pending(self.h))

    @property
    def frames(self) -> int:
        return int(self.L.ct_host_controls_frames(self.h))

    def camera(self):
        """(position[3] float64, (yaw, pitch, roll) float32, rotation[9] float64)."""
        pos, ypr, rot = np.zeros(3), np.zeros(3, np.float32), np.zeros(9)
        vp = C.c_void_p
        self.L.ct_host_controls_camera(self.h, pos.ctypes.data_as(vp), ypr.ctypes.data_as(vp), rot.ctypes.data_as(vp))
        return pos, ypr, rot

    def close(self):
        if self.h:
            self.L.ct_host_controls_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Viewer:
    """cobbletrace.cpp's main loop without the window: feed key presses, call tick() once per iteration.
    `present(bitmap, frame_is_new)` stands where Blit (draw2d.h:22) is."""

    def __init__(self, boss: "Boss", controls: Optional[Controls] = None, present=None):
        self.boss = boss
        self.controls = controls or Controls(boss.scene)
        self.bitmap = np.zeros((boss.height, boss.width), np.uint32)      # env->bitmap, calloc'ed (cobbletrace.cpp:56)
        self._present_py = present
        bm = self.bitmap

        def _thunk(_user, _ptr, _stride, is_new):
            if self._present_py is not None:
                self._present_py(bm, bool(is_new))
        self._thunk = PRESENT_FN(_thunk)
        self.last_stats = None

    def key(self, ch: str):
        self.controls.key(ch)

    def keys(self, s: str):
        self.controls.keys(s)

    def tick(self) -> bool:
        """One main-loop iteration; True when a new frame was rendered into self.bitmap."""
        L = self.boss.L
        rendered, st = C.c_int(0), FrameStats()
        rc = L.ct_host_viewer_tick(self.boss.h, self.controls.h, self.bitmap.ctypes.data_as(C.c_void_p), self.bitmap.shape[1],
                                   self._thunk, None, C.byref(rendered), C.byref(st))
        if rc < 0:
            raise RuntimeError("ct_host_viewer_tick: " + _err(L))
        if rendered.value:
            self.last_stats = dict(st.rays.as_dict(), wall_ms=float(st.wall_ms), kernel_launches=int(st.kernel_launches))
        return bool(rendered.value)


class TileCounter:
    """The dispenser the boss steals tiles from (process-local, or POSIX shm shared by all ranks)."""

    def __init__(self, shared_name: Optional[str] = None):
        self.L = load_library()
        self.h = self.L.ct_host_tile_counter_open(shared_name.encode() if shared_name else None)
        if not self.h:
            raise RuntimeError("ct_host_tile_counter_open: " + _err(self.L))

    def next(self) -> int:
        return int(self.L.ct_host_tile_counter_next(self.h))

    def reset(self):
        self.L.ct_host_tile_counter_reset(self.h)

    def close(self, unlink: bool = False):
        if self.h:
            self.L.ct_host_tile_counter_close(self.h, int(unlink)); self.h = None
