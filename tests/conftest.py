import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Native libraries are built in-tree once per session (no JIT cache)."""
    from cobbletrace_b200 import build
    build.build_host()
    try:
        build.find_nvcc()
    except Exception:
        # a machine without the CUDA toolkit: the CPU-side tests (oracle, host parser / BVH builder, controls) still run;
        # whatever needs libct_gpu.so reports that it is missing
        if not os.path.exists(build.GPU_LIB):
            print("conftest: nvcc not found and libct_gpu.so not built -- tests that load it will fail", file=sys.stderr)
    else:
        build.build_gpu()
    from oracle import ct_oracle_py
    ct_oracle_py.build()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLD, "golden.json")) as f:
        return json.load(f)


_scene_cache = {}


def load_golden_scene(name, golden):
    """Flattened golden scene, BVH included (rebuilt with the host builder where it is not stored and
    checked against the recorded digest of the reference's own BVH)."""
    if name in _scene_cache:
        return _scene_cache[name]
    import dataclasses
    from cobbletrace_b200 import host
    from cobbletrace_b200.sceneio import load_ctscene
    meta = golden["scenes"][name]
    fs = load_ctscene(os.path.join(GOLD, meta["file"]))
    assert fs.geometry_digest() == meta["geometry_sha256"]
    if not fs.has_bvh():
        fs = host.HostScene.from_flat(fs).to_flat(with_bvh=True)
    assert fs.bvh_digest() == meta["bvh_sha256"], "host BVH differs from the reference's BuildBVH"
    _scene_cache[name] = fs
    return fs


@pytest.fixture(scope="session")
def scene_loader(golden):
    return lambda name: load_golden_scene(name, golden)


def load_frames(case):
    z = np.load(os.path.join(GOLD, f"frames_{case}.npz"))
    return {k: z[k] for k in z.files}


def case_scene(case, golden, loader):
    """FlatScene for a golden frame case with its camera / material override applied."""
    import dataclasses
    meta = golden["frames"][case]
    fs = loader(meta["scene"])
    fs = dataclasses.replace(fs, cam_pos=np.array(meta["cam_pos"]), cam_rot=np.array(meta["cam_rot"]))
    if meta["force_reflection"] is not None:
        fs = fs.with_reflection(meta["force_reflection"])
    return fs, meta


def load_fuzz_case(k):
    """tests/golden/fuzz_<k>.npz (make_golden_fuzz.py) -> (FlatScene as the reference flattened + built it, dict of the rest)."""
    from cobbletrace_b200.sceneio import FlatScene
    z = np.load(os.path.join(GOLD, f"fuzz_{k}.npz"))
    fs = FlatScene(**{f: z[f] for f in ("tri", "mat_color", "mat_specular", "mat_reflection", "light_type", "light_intensity", "light_pos",
                                        "light_dir", "cam_pos", "cam_rot", "node_min", "node_max", "node_left", "node_first", "node_count", "tri_index")})
    rest = {"json": bytes(z["json"]).decode(), "keys": bytes(z["keys"]).decode(), "depth": int(z["depth"]), "frame": z["frame"],
            "found": z["found"], "index": z["index"], "t": z["t"]}
    return fs, rest
