"""CPU-side checks of the drop-in boundary: the C-ABI libraries load without a GPU, export every symbol
the headers declare, use the reference's struct layouts, and fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from cobbletrace_b200 import api, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(%s\w+)\s*\(" % prefix, src)))


def test_gpu_library_exports_every_declared_symbol():
    L = api.load_library()
    names = declared("ct_gpu.h", "ct_gpu_")
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"libct_gpu.so lacks {n}"
    assert sorted(names) == sorted(api.ABI_SYMBOLS)
    assert L.ct_gpu_abi_version() == 1


def test_host_library_exports_every_declared_symbol():
    L = host.load_library()
    names = declared("ct_host.h", "ct_host_")
    for n in names:
        assert hasattr(L, n), f"libct_host.so lacks {n}"
    assert sorted(names) == sorted(host.ABI_SYMBOLS)


def test_struct_layouts_match_reference_sizes():
    # sizeof(bvh_node_t)=64, light_t=56, material_t=12 on x86-64
    assert C.sizeof(api.BvhNode) == 64 and api.BvhNode.left_node.offset == 48
    assert C.sizeof(api.Light) == 56 and api.Light.position.offset == 8
    assert C.sizeof(api.Material) == 12
    assert api.SceneDesc.triangles.offset == 16


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = api.load_library()
    assert L.ct_gpu_device_count() == -2          # CT_ERR_NO_DEVICE
    r = api.GpuRenderer(0)
    with pytest.raises(api.CtError) as e:
        r.render_tile(-4, 4)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under cobbletrace_b200/ or include/ may reference it."""
    bad = []
    for base in ("cobbletrace_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    txt = open(os.path.join(d, f), errors="replace").read()
                    if re.search(r"ct_oracle|from oracle|import oracle|oracle/", txt):
                        bad.append(os.path.join(d, f))
    assert not bad, bad


def test_headers_are_plain_c_and_the_example_host_builds(tmp_path):
    """include/*.h must be usable from C (the reference is C-style C++; other hosts bind the same header): the example
    host compiles as C99 with -Werror, links against libct_host.so only, and -- without a CUDA device -- fails loudly
    instead of rendering anything on the CPU."""
    import shutil, subprocess
    import torch
    host.load_library()                                        # make sure libct_host.so is built
    exe = str(tmp_path / "viewer")
    pkg = os.path.join(ROOT, "cobbletrace_b200")
    subprocess.run([shutil.which("gcc") or "gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", "headless_viewer.c"),
                    "-I", os.path.join(ROOT, "include"), "-L", pkg, "-lct_host", f"-Wl,-rpath,{pkg}", "-o", exe], check=True)
    scene = tmp_path / "s.json"
    scene.write_text('{"objects":[{"type": "triangle", "p1": [-1, -1, 4], "p2": [1, -1, 4], "p3": [0, 1, 4], "color": [255, 0, 0], "specular": 10, "reflection": 0}],'
                     ' "lights":[{"type": "ambient", "intensity": 0.5}], "camera":{"position": [0, 0, -3]},'
                     ' "settings":{"numberOfThreads": 2, "subsampling": false, "wireframe": false, "supersampling": false}}')
    r = subprocess.run([exe, str(scene), "w|", str(tmp_path / "f")], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "tick 1: new frame" in r.stdout and "tick 2: nothing changed" in r.stdout, r.stdout + r.stderr
        assert os.path.exists(str(tmp_path / "f_001.ppm"))
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr, r.stdout + r.stderr
