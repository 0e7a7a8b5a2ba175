#!/usr/bin/env python3
"""Development probe: megakernel vs wavefront at the current code state (4K, Book 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer
sc = BuiltinScene(10)
cam = sc.camera(3840, 2160, 64, 50)
r = Renderer(sc.desc)
stream = torch.cuda.current_stream().cuda_stream
def run(**kw):
    r.render(cam, stream=stream, **kw); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.render(cam, stream=stream, **kw); r.render(cam, stream=stream, **kw); e1.record(); torch.cuda.synchronize()
    _, _, st = r.readback(linear=False)
    return st.rays / (e0.elapsed_time(e1) / 2) / 1e6
print("megakernel", f"{run(variant=1):.2f} Grays/s", flush=True)
for threads in (0, 768, 640, 512):
    try:
        print("head/tail threads", threads, f"{run(variant=3, block_threads=threads):.2f} Grays/s", flush=True)
    except Exception as e:
        print("head/tail threads", threads, "failed:", str(e)[:120])
if len(sys.argv) > 1 and sys.argv[1] == "ht":
    sys.exit(0)
for slots in (64, 96):
    for idle, leaf, refill in ((16, 16, 16), (16, 16, 8), (24, 16, 8), (12, 12, 6), (8, 16, 8)):
        fl = ((slots // 32) << 12) | (idle << 16) | (leaf << 21) | (refill << 26)
        print("wavefront", slots, idle, leaf, refill, f"{run(variant=2, flags=fl):.2f} Grays/s", flush=True)
