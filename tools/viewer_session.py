#!/usr/bin/env python
"""Headless stand-in for the reference's window: replays key presses through
ct_host_controls / ct_host_viewer_tick on a GPU and writes every NEW frame as a PPM -- the `present` hook is where
Blit (draw2d.h:22) would hand the bitmap to SDL.

    python tools/viewer_session.py --scene oracle/_ref/scenes/scene_file_cube.json --keys "yp|wd|oooo" --out gpurun_out/session
Batches are separated by '|', one main-loop tick each; an empty batch is an idle tick (no frame is rendered).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cobbletrace_b200 import host  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", required=True)
    ap.add_argument("--base-dir", default=None)
    ap.add_argument("--width", type=int, default=640)      # cobbletrace.cpp:14-15
    ap.add_argument("--height", type=int, default=640)
    ap.add_argument("--depth", type=int, default=10)
    ap.add_argument("--keys", default="")
    ap.add_argument("--out", default=None, help="directory for frame_NNN.ppm (omit: no files)")
    ap.add_argument("--devices", default="0")
    a = ap.parse_args()
    hs = host.HostScene.load(a.scene, base_dir=a.base_dir or os.path.dirname(os.path.abspath(a.scene)))
    boss = host.Boss(hs, a.width, a.height, devices=tuple(int(d) for d in a.devices.split(",")), max_depth=a.depth)
    if a.out:
        os.makedirs(a.out, exist_ok=True)
    n = [0]

    def present(bitmap, is_new):
        if is_new and a.out:
            host.write_ppm(os.path.join(a.out, f"frame_{n[0]:03d}.ppm"), bitmap)
        n[0] += int(is_new)

    v = host.Viewer(boss, present=present)
    for tick, batch in enumerate([None] + a.keys.split("|")):
        if batch:
            v.keys(batch)
        t0 = time.perf_counter()
        new = v.tick()
        ms = (time.perf_counter() - t0) * 1e3
        pos, ypr, _ = v.controls.camera()
        print(f"tick {tick}: keys={batch!r} new_frame={new} {ms:.2f} ms  pos={pos.tolist()} yaw/pitch/roll={ypr.tolist()}")
    boss.close()


if __name__ == "__main__":
    main()
