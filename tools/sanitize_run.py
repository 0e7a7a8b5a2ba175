#!/usr/bin/env python3
"""Tiny renders of every kernel instantiation (and a two-device handle where there are two GPUs), for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer
from _fixtures import earth_texels  # noqa: E402
earth = earth_texels()
for sid in (10, 0, 7, 8, 9):
    sc = BuiltinScene(sid, earth if sid in (2, 9) else None)
    cam = sc.camera(50, 29, 2, 50)  # not a multiple of the tile sizes
    for variant in (0, 1, 2, 3, 4):  # AUTO (= hit-queue), megakernel, wavefront, head/tail, hit-queue
        for flags in (0, 0x200, 0x100):  # scene in smem / in global memory / instrumented
            r = Renderer(sc.desc)
            r.render(cam, variant=variant, flags=flags)
            lin, s8, st = r.readback(linear=True, srgb8=True)
            r.close()
            print("scene", sid, "variant", variant, "flags", hex(flags), "rays", st.rays, "mean", float(lin.mean()), flush=True)
print("sanitize_run: done")
