#!/usr/bin/env python3
"""Development probe: where the end-to-end frame time goes (upload / render / readback / free)."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, _abi as A
W, H, spp = 3840, 2160, int(sys.argv[1]) if len(sys.argv) > 1 else 64
sc = BuiltinScene(10)
cam = sc.camera(W, H, spp, 50)
host = torch.empty(H * W * 3, dtype=torch.float32).pin_memory()
accum = torch.zeros(H * W * 3, dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
def t():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = t(); r = Renderer(sc.desc)
    t1 = t(); r.render(cam, stream=stream, accum_ptr=accum.data_ptr())
    t2 = t(); st = A.rt_stats(); rc = r.lib.rt_readback(r._h, C.c_void_p(accum.data_ptr()), C.c_void_p(host.data_ptr()), None, C.byref(st))
    t3 = t(); r.close()
    t4 = t()
    print(f"upload {1e3*(t1-t0):.1f} ms  render {1e3*(t2-t1):.1f} ms  readback {1e3*(t3-t2):.1f} ms  free {1e3*(t4-t3):.1f} ms", flush=True)
