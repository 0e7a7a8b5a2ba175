#!/usr/bin/env python3
"""A/B timing of kernel variants / upload options on the BASELINE scenes (GPU box, CUDA events, no profiler).

usage: ab_probe.py [--cases "10:3840x2160x64,0:1920x1080x64"] [--variants 3,4] [--upload-flags 0,1]
                   [--threads 0] [--flags 0] [--reps 3] [--leaf 0]
Prints one JSON line per (case, variant, upload flag, block size): best-of-reps ms, Grays/s, registers are in the
build log.  The library under test is RT_B200_LIBRARY or the in-tree one."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer, library_path  # noqa: E402
from _fixtures import earth_texels  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="10:3840x2160x64,0:1920x1080x64,8:1024x1024x64,9:1920x1080x32")
ap.add_argument("--variants", default="3,4")
ap.add_argument("--upload-flags", default="0")
ap.add_argument("--threads", default="0")
ap.add_argument("--flags", default="0")
ap.add_argument("--max-leaf", default="0")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tag", default="")
a = ap.parse_args()

earth = earth_texels()
stream = torch.cuda.current_stream().cuda_stream
for case in a.cases.split(","):
    sid, dims = case.split(":")
    sid = int(sid)
    W, H, spp = (int(x) for x in dims.split("x"))
    sc = BuiltinScene(sid, earth if sid in (2, 9) else None)
    cam = sc.camera(W, H, spp, 50)
    for uf in (int(x, 0) for x in a.upload_flags.split(",")):
        for ml in (int(x) for x in a.max_leaf.split(",")):
            r = Renderer(sc.desc, upload_flags=uf, max_leaf_prims=ml)
            for variant in (int(x) for x in a.variants.split(",")):
                for threads in (int(x) for x in a.threads.split(",")):
                    for flags in (int(x, 0) for x in a.flags.split(",")):
                        kw = dict(stream=stream, variant=variant, block_threads=threads, flags=flags)
                        try:
                            r.render(cam, **kw)
                            torch.cuda.synchronize()
                        except Exception as e:  # noqa: BLE001
                            print(json.dumps({"scene": sid, "variant": variant, "threads": threads, "error": str(e)[:200]}),
                                  flush=True)
                            continue
                        best = 1e30
                        for _ in range(a.reps):
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                            r.render(cam, **kw)
                            e1.record()
                            torch.cuda.synchronize()
                            best = min(best, e0.elapsed_time(e1))
                        _, _, st = r.readback(linear=False)
                        info = r.info()
                        print(json.dumps({"tag": a.tag, "lib": os.path.basename(library_path()), "scene": sid,
                                          "size": f"{W}x{H}x{spp}", "variant": variant, "upload_flags": uf,
                                          "max_leaf": ml, "threads": threads, "flags": flags, "ms": round(best, 2),
                                          "grays_s": round(st.rays / best / 1e6, 3), "rays": int(st.rays),
                                          "smem": info.scene_in_smem, "nodes": info.n_nodes}), flush=True)
            r.close()
