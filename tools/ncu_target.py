#!/usr/bin/env python3
"""Short single-GPU target for ncu: two launches of the render kernel on the
Book 1 final scene (the second is the one to profile: `-s 1 -c 1`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, load_earth_fixture  # noqa: E402

sid = int(sys.argv[1]) if len(sys.argv) > 1 else 10
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
H = int(sys.argv[3]) if len(sys.argv) > 3 else 1080
spp = int(sys.argv[4]) if len(sys.argv) > 4 else 8
threads = int(sys.argv[5]) if len(sys.argv) > 5 else 0
bps = int(sys.argv[6]) if len(sys.argv) > 6 else 0
sc = BuiltinScene(sid, load_earth_fixture() if sid in (2, 9) else None)
cam = sc.camera(W, H, spp, 50)
r = Renderer(sc.desc)
for _ in range(2):
    r.render(cam, block_threads=threads, blocks_per_sm=bps)
    r.sync()
_, _, st = r.readback(linear=False)
print("rays", st.rays)
