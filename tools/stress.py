#!/usr/bin/env python3
"""Randomised robustness run (GPU box): random scene / size / sample range / kernel variant / BVH mode /
block shape, each checked against the FP64 oracle.  Not part of the test-suite (takes minutes)."""
import ctypes as C, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, _abi as A
from _fixtures import earth_texels  # noqa: E402
from oracle import bindings as O
oracle = O.load_oracle()
earth = earth_texels()
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
t0 = time.time(); n = 0; worst = 1.0; nbad = 0
scenes = {sid: BuiltinScene(sid, earth if sid in (2, 9) else None) for sid in range(11)}
while time.time() - t0 < budget:
    sid = rnd.choice(list(range(11)))
    W, H = rnd.randint(1, 97), rnd.randint(1, 61)
    spp = rnd.randint(1, 6); s0 = rnd.randint(0, 3); depth = rnd.choice([1, 2, 5, 50])
    variant = rnd.choice([0, 1, 2, 3, 3]); bvh = rnd.choice([A.RT_BVH_SAH, A.RT_BVH_SAH, A.RT_BVH_REFERENCE, A.RT_BVH_NONE])
    threads = rnd.choice([0, 32, 64, 128, 256, 512, 640, 768]); flags = rnd.choice([0, 0, 0x200, 0x100])
    if sid == 9 and bvh == A.RT_BVH_NONE and W * H > 1500:
        bvh = A.RT_BVH_SAH  # the linear list over 3400 primitives is slow, not wrong
    sc = scenes[sid]
    cam = sc.camera(W, H, spp + s0, depth)
    want = np.zeros((H, W, 3)); ost = O.oracle_stats()
    oracle.oracle_render(sc.desc, C.byref(cam), s0, s0 + spp, 1984, 1, 64, os.cpu_count(), want.ctypes.data, C.byref(ost))
    r = Renderer(sc.desc, bvh=bvh)
    r.render(cam, s0, s0 + spp, variant=variant, block_threads=threads, flags=flags)
    got, _, st = r.readback()
    r.close()
    ref = want / (spp + s0)
    ok = (np.abs(got - ref) <= 1e-3 * np.abs(ref) + 1e-6).all(axis=2)
    frac = ok.mean(); bad = int((~ok).sum())
    n += 1; worst = min(worst, frac)
    tag = f"scene {sid} {W}x{H} spp [{s0},{s0 + spp}) depth {depth} variant {variant} bvh {bvh} threads {threads} flags {hex(flags)}"
    if bad > max(2, 0.004 * W * H) or abs(int(st.rays) - int(ost.rays)) > max(3, 0.004 * ost.rays):
        nbad += 1
        print("MISMATCH", tag, "bad", bad, "of", W * H, "rays", st.rays, ost.rays, flush=True)
    elif n % 20 == 0:
        print("ok", n, tag, f"match {frac:.4f}", flush=True)
print(f"stress: {n} cases in {time.time() - t0:.0f} s, {nbad} flagged, worst match fraction {worst:.4f}")
