#!/usr/bin/env python3
"""Development probe: megakernel throughput for a list of rt_render_params.flags values (hex)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, load_earth_fixture
sid = int(os.environ.get("SCENE", "10"))
W, H, spp = (int(x) for x in os.environ.get("SIZE", "3840,2160,16").split(","))
sc = BuiltinScene(sid, load_earth_fixture() if sid in (2, 9) else None)
cam = sc.camera(W, H, spp, int(os.environ.get("DEPTH", "50")))
r = Renderer(sc.desc, max_leaf_prims=int(os.environ.get('LEAF', '0')))
stream = torch.cuda.current_stream().cuda_stream
for f in sys.argv[1:]:
    flags = int(f, 0)
    r.render(cam, stream=stream, flags=flags, variant=int(os.environ.get("VARIANT", "0")))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        r.render(cam, stream=stream, flags=flags, variant=int(os.environ.get("VARIANT", "0")))
    e1.record()
    torch.cuda.synchronize()
    _, _, st = r.readback(linear=False)
    ms = e0.elapsed_time(e1) / 3
    extra = f" node/ray {st.node_tests / st.rays:.2f} prim/ray {st.prim_tests / st.rays:.3f}" if flags & 0x100 else ""
    print(f"scene {sid} flags {f}: {ms:.2f} ms {st.rays / ms / 1e6:.2f} Grays/s" + extra, flush=True)
