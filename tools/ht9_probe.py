#!/usr/bin/env python3
"""Development probe: head/tail vs megakernel on scene 9 for several block sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, load_earth_fixture
sc = BuiltinScene(9, load_earth_fixture())
cam = sc.camera(1920, 1080, 32, 50)
r = Renderer(sc.desc)
stream = torch.cuda.current_stream().cuda_stream
for variant, threads in [(1, 512), (1, 384), (3, 640), (3, 512), (3, 384), (3, 256)]:
    r.render(cam, stream=stream, variant=variant, block_threads=threads); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.render(cam, stream=stream, variant=variant, block_threads=threads); e1.record(); torch.cuda.synchronize()
    _, _, st = r.readback(linear=False)
    print(f"scene 9 variant {variant} threads {threads}: {e0.elapsed_time(e1):.1f} ms {st.rays / e0.elapsed_time(e1) / 1e6:.2f} Grays/s", flush=True)
