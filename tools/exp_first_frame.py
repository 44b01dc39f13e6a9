import os, sys, time
sys.path.insert(0, "/root/repo")
t0=time.time()
import numpy as np
from cobbletrace_b200 import host, api, procedural
import tempfile
d = os.environ.get("CT_SCENE_CACHE") or tempfile.mkdtemp()
scene, n = procedural.write_dragon_standin(d)
t=time.time(); hs = host.HostScene.load(scene, base_dir=d); print("load", (time.time()-t)*1e3)
hs.set_reflection(0.5)
t=time.time(); hs.build_bvh(); print("bvh", (time.time()-t)*1e3)
import ctypes
t=time.time(); api.device_count(); print("device_count (context?)", (time.time()-t)*1e3)
for i in range(2):
    t=time.time(); b = host.Boss(hs, 3840, 2160, devices=(0,), max_depth=2); print("boss create", (time.time()-t)*1e3)
    t=time.time(); b.render(None, want_bitmap=False); print("first frame", (time.time()-t)*1e3)
    t=time.time(); b.render(None, want_bitmap=False); print("second frame", (time.time()-t)*1e3)
    b.close()
