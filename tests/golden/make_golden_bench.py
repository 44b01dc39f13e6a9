#!/usr/bin/env python
"""Frame hashes of the bench workloads from the compiled, unmodified reference (oracle/_ref/ct_ref) -> bench_frames.json.

    python tests/golden/make_golden_bench.py [workload ...]

bench.py compares the bitmap it reads back at every GPU count with these (and, at N = 1, with a live oracle render).
Needs /root/reference (oracle/_ref built by `make -C oracle ref`); the dragon-class scene is generated on the fly
(cobbletrace_b200.procedural, deterministic)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cobbletrace_b200.sceneio import frame_fnv1a  # noqa: E402
from oracle import ct_oracle_py as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "bench_frames.json")
names = sys.argv[1:] or ["dragon4k", "dragon8k", "dragon1080", "cube640", "import640", "bunny1080", "pcbig1080"]
res = json.load(open(OUT)) if os.path.exists(OUT) else {}
for name in names:
    desc, kind, W, H, depth, refl = bench.WORKLOADS[name][:6]
    scene_path, n_tri = bench.ensure_scene(kind)
    threads = bench.ref_threads(H, min(bench.host_cores(), 16))
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "frame.bin")
        cmd = [os.path.join(O.REF_DIR, "ct_ref"), "--scene", os.path.basename(scene_path), "--chdir", bench.scene_dir(kind), "--width", str(W),
               "--height", str(H), "--depth", str(depth), "--threads", str(threads), "--frame", out]
        if refl is not None:
            cmd += ["--force-reflection", repr(float(refl))]
        subprocess.run(cmd, check=True, capture_output=True, timeout=3600)
        frame = np.fromfile(out, np.uint32).reshape(H, W)
    res[name] = {"fnv1a": frame_fnv1a(frame), "width": W, "height": H, "max_depth": depth, "forced_reflection": refl, "triangles": n_tri,
                 "background_px": int((frame == 0x333333).sum()), "zero_px": int((frame == 0).sum()),
                 "source": f"oracle/_ref/ct_ref (the unmodified reference), numberOfThreads={threads}"}
    print(name, res[name], flush=True)
    json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
