#!/usr/bin/env python3
"""Development probe: megakernel throughput vs block shape, for the library named by RT_B200_LIBRARY."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer
sc = BuiltinScene(10)
cam = sc.camera(3840, 2160, 16, 50)
r = Renderer(sc.desc)
stream = torch.cuda.current_stream().cuda_stream
for threads, bps in [(int(a), int(b)) for a, b in (x.split("x") for x in sys.argv[1:])]:
    try:
        r.render(cam, stream=stream, block_threads=threads, blocks_per_sm=bps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            r.render(cam, stream=stream, block_threads=threads, blocks_per_sm=bps)
        e1.record()
        torch.cuda.synchronize()
        _, _, st = r.readback(linear=False)
        ms = e0.elapsed_time(e1) / 2
        print(os.environ.get("RT_B200_LIBRARY", "default"), threads, bps, f"{ms:.2f} ms {st.rays / ms / 1e6:.2f} Grays/s", flush=True)
    except Exception as e:
        print("skip", threads, bps, str(e)[:100])
