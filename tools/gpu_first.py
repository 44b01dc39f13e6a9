import sys, time, json, numpy as np
sys.path.insert(0, ".")
import cobbletrace_b200 as ct
from oracle import ct_oracle_py as O

def run(name, W, H, depth=10, refl=None, flags=0):
    fs = ct.load_ctscene(f"oracle/_ref/dumps/{name}.ctscene")
    if refl is not None: fs = fs.with_reflection(refl)
    osc = O.OracleScene(fs)
    t0 = time.time(); oframe, ohits, octr = osc.render(W, H, max_depth=depth); t_or = time.time() - t0
    r = ct.GpuRenderer(0).upload(fs, W, H, max_depth=depth, flags=ct.CT_FLAG_KEEP_HITS | ct.CT_FLAG_COUNT_TESTS | flags)
    ctr = r.render_tile(counters=True)
    frame = r.readback()
    found, index, t = r.readback_hits()
    traced = ohits["found"] != 0xFFFFFFFF
    hit_ok = (found[traced] == ohits["found"][traced]) & (index[traced] == ohits["index"][traced]) & (t[traced].view(np.uint32) == ohits["t"][traced].view(np.uint32))
    nd = int((frame != oframe).sum())
    print(f"{name} {W}x{H} d{depth} refl={refl}: frame diff px {nd}/{W*H}, hit mismatch {int((~hit_ok).sum())}/{int(traced.sum())}, oracle {t_or:.2f}s")
    print("   gpu ctr", ctr); print("   ora ctr", octr)
    # timing without counters
    r.upload(fs, W, H, max_depth=depth, flags=flags)
    for _ in range(3): r.render_tile(); 
    r.sync()
    ms = []
    for _ in range(5):
        r.render_tile(); ms.append(r.last_tile_ms())
    rays = ctr["rays_primary"] + ctr["rays_shadow"] + ctr["rays_reflection"]
    print(f"   gpu ms/frame {min(ms):.3f} (median {sorted(ms)[2]:.3f})  -> {rays/min(ms)/1e3:.1f} Mrays/s; rays {rays}", flush=True)
    r.shutdown()

run("scene_file_cube", 640, 640)
run("scene_import", 640, 640)
run("scene_import_bunny", 640, 640)
run("pc_big", 640, 640)
run("scene_import_bunny", 1920, 1080)
run("scene_import_bunny", 1920, 1080, depth=2, refl=0.5)
run("scene_import_bunny", 3840, 2160, depth=2, refl=0.5)
