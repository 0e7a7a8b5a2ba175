#!/usr/bin/env python3
"""Development probe: every kernel variant on the BASELINE scenes (CUDA-event timing)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, load_earth_fixture
earth = load_earth_fixture()
stream = torch.cuda.current_stream().cuda_stream
for sid, W, H, spp in [(10, 3840, 2160, 64), (0, 1920, 1080, 64), (7, 1024, 1024, 64), (8, 1024, 1024, 64), (9, 1920, 1080, 32)]:
    sc = BuiltinScene(sid, earth if sid in (2, 9) else None)
    cam = sc.camera(W, H, spp, 50)
    r = Renderer(sc.desc)
    for variant in (1, 2, 3):
        r.render(cam, stream=stream, variant=variant); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.render(cam, stream=stream, variant=variant); e1.record(); torch.cuda.synchronize()
        _, _, st = r.readback(linear=False)
        print(f"scene {sid} {W}x{H}x{spp} variant {variant}: {e0.elapsed_time(e1):.1f} ms {st.rays / e0.elapsed_time(e1) / 1e6:.2f} Grays/s smem={r.info().scene_in_smem}", flush=True)
    r.close()
