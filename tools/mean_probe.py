#!/usr/bin/env python3
"""Development probe: image mean vs resolution (GPU) and vs the oracle at low resolution."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, _abi as A
from oracle import bindings as O
oracle = O.load_oracle()
sc = BuiltinScene(10)
spp = 8
r = Renderer(sc.desc)
imgs = {}
for W, H in [(480, 270), (960, 540), (1920, 1080), (3840, 2160)]:
    cam = sc.camera(W, H, spp, 50)
    r.render(cam, 0, spp)
    lin, _, st = r.readback()
    imgs[W] = lin
    k = W // 480
    small = lin.reshape(270, k, 480, k, 3).mean(axis=(1, 3))
    print(W, H, "mean", lin.mean(axis=(0, 1)), "rays/path", st.rays / (W * H * spp), flush=True)
    imgs[("s", W)] = small
cam = sc.camera(480, 270, spp, 50)
want = np.zeros((270, 480, 3)); ost = O.oracle_stats()
oracle.oracle_render(sc.desc, C.byref(cam), 0, spp, 1984, 1, 64, os.cpu_count(), want.ctypes.data, C.byref(ost))
want /= spp
print("oracle 480 mean", want.mean(axis=(0, 1)))
for W in (480, 960, 1920, 3840):
    d = imgs[("s", W)] - want
    print(W, "rmse vs oracle480", np.sqrt((d ** 2).mean()), "mean diff", d.mean(axis=(0, 1)))
    # where is the difference: 3x4 blocks
    blk = d.reshape(3, 90, 4, 120, 3).mean(axis=(1, 3))
    print(np.round(blk[..., 0], 3))
