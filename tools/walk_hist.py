#!/usr/bin/env python3
"""Histogram of walk lengths (box-pair steps per ray) of the hit-queue kernel, heads and tails apart (GPU box).

The instrumented kernel (RT_FLAG_STATS instantiation) counts, for ONE sample of the whole frame, how many child-pair
steps each lane's walk took (rt_debug_trace_path with pixel = -2 turns its per-path record buffer into 2 x 64 counters).
usage: walk_hist.py [--scene 10] [--width 1920] [--height 1080] [--sample 0]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, _abi as A  # noqa: E402
from _fixtures import earth_texels  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scene", type=int, default=10)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--sample", type=int, default=0)
a = ap.parse_args()
sc = BuiltinScene(a.scene, earth_texels() if a.scene in (2, 9) else None)
cam = sc.camera(a.width, a.height, 64, 50)
r = Renderer(sc.desc)
p = A.rt_render_params(sample_begin=0, sample_end=1, seed=1984, variant=A.RT_VARIANT_HITQUEUE, clear=1)
rec = np.zeros((256, 8), np.float32)
rc = r.lib.rt_debug_trace_path(r._h, C.byref(cam), C.byref(p), -2, a.sample, rec.ctypes.data, 256)
assert rc == 0, r.lib.rt_last_error()
h = rec.view(np.uint32).reshape(-1)[:128].astype(np.int64)
out = {}
for name, v in (("heads", h[:64]), ("tails", h[64:])):
    n = int(v.sum())
    steps = np.arange(64)
    mean = float((v * steps).sum() / max(n, 1))
    cdf = np.cumsum(v) / max(n, 1)
    out[name] = {"rays": n, "mean_steps": round(mean, 2), "share_0_1": round(float(v[:2].sum() / max(n, 1)), 3),
                 "share_2_3": round(float(v[2:4].sum() / max(n, 1)), 3), "median": int(np.searchsorted(cdf, 0.5)),
                 "p90": int(np.searchsorted(cdf, 0.9)), "p99": int(np.searchsorted(cdf, 0.99)),
                 "hist": [int(x) for x in v[:40]]}
    # expected lanes busy in a 32-lane round whose lanes draw their walk lengths independently from this histogram
    pmf = v / max(n, 1)
    cdfk = np.cumsum(pmf)
    emax = float(sum(1.0 - cdfk[k] ** 32 for k in range(63)))  # E[max] = sum_k P(max > k)
    out[name]["expected_round_length_32_lanes"] = round(emax, 2)
    out[name]["expected_lane_utilisation"] = round(mean / emax, 3) if emax > 0 else None
print(json.dumps({"scene": a.scene, "size": f"{a.width}x{a.height}", **out}))
