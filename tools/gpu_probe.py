#!/usr/bin/env python3
"""Development probe for the GPU box: parity fractions, kernel-variant timings,
reference-kernel timings.  Writes gpurun_out/probe_<tag>.json.  Not a test and
not the benchmark; numbers quoted in DESIGN.md come from bench.py / profiles/."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer  # noqa: E402
from _fixtures import earth_texels  # noqa: E402
from raytracinginoneweekendincuda_b200 import _abi as A  # noqa: E402
from oracle import bindings as O  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
res = {"gpu": torch.cuda.get_device_name(0), "cpus": os.cpu_count()}
earth = earth_texels()
oracle = O.load_oracle()


def oracle_render(sc, cam, s0, s1, seed=1984):
    out = np.zeros((cam.image_height, cam.image_width, 3))
    st = O.oracle_stats()
    oracle.oracle_render(sc.desc, C.byref(cam), s0, s1, seed, 1, 64, os.cpu_count(), out.ctypes.data, C.byref(st))
    return out, st


def timed(r, cam, reps=1, **kw):
    stream = torch.cuda.current_stream().cuda_stream
    r.render(cam, stream=stream, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r.render(cam, stream=stream, **kw)
    e1.record()
    torch.cuda.synchronize()
    _, _, st = r.readback(linear=False)
    return e0.elapsed_time(e1) / reps, st


what = set(sys.argv[1:]) or {"parity", "perf", "ref"}

if "parity" in what:
    rows = []
    for sid, W, H in [(10, 400, 225), (0, 400, 225), (7, 160, 160), (8, 160, 160), (9, 240, 135), (3, 160, 90), (5, 160, 90),
                      (2, 160, 90)]:
        sc = BuiltinScene(sid, earth if sid in (2, 9) else None)
        for spp in (1, 10):
            cam = sc.camera(W, H, spp, 50)
            want, ost = oracle_render(sc, cam, 0, spp)
            r = Renderer(sc.desc)
            r.render(cam)
            got, _, st = r.readback()
            r.close()
            ref = want / spp
            ok = (np.abs(got - ref) <= 1e-3 * np.abs(ref) + 1e-6).all(axis=2)
            rows.append({"scene": sid, "W": W, "H": H, "spp": spp, "match": float(ok.mean()), "bad_pixels": int((~ok).sum()),
                         "rays_gpu": int(st.rays), "rays_oracle": int(ost.rays),
                         "box_per_ray_ref": ost.box_tests / ost.rays, "sph_per_ray_ref": ost.sphere_tests / ost.rays,
                         "quad_per_ray_ref": ost.quad_tests / ost.rays})
            print(rows[-1], flush=True)
    res["parity"] = rows

if "perf" in what:
    rows = []
    sc = BuiltinScene(10)
    cam = sc.camera(3840, 2160, 16, 50)
    for leaf in (1, 2, 4):
        r = Renderer(sc.desc, max_leaf_prims=leaf)
        for threads, bps in [(256, 1), (512, 1), (768, 1), (256, 2), (256, 3), (384, 2), (128, 4), (128, 6)]:
            for flags in (0, 0x200):
                try:
                    ms, st = timed(r, cam, block_threads=threads, blocks_per_sm=bps, flags=flags)
                except Exception as e:  # noqa: BLE001
                    print("skip", threads, bps, flags, e)
                    continue
                rows.append({"scene": 10, "leaf": leaf, "threads": threads, "blocks_per_sm": bps, "smem": flags == 0,
                             "ms": ms, "rays": int(st.rays), "grays_s": st.rays / ms / 1e6})
                print(rows[-1], flush=True)
        ms, st = timed(r, cam, flags=0x100)
        rows.append({"scene": 10, "leaf": leaf, "stats": True, "ms": ms, "node_per_ray": st.node_tests / st.rays,
                     "prim_per_ray": st.prim_tests / st.rays, "rays_per_path": st.rays / max(1, st.paths)})
        print(rows[-1], flush=True)
        r.close()
    for sid, W, H, spp in [(0, 1920, 1080, 16), (8, 1024, 1024, 16), (7, 1024, 1024, 16), (9, 3840, 2160, 2)]:
        sc = BuiltinScene(sid, earth if sid in (2, 9) else None)
        cam = sc.camera(W, H, spp, 50)
        r = Renderer(sc.desc)
        ms, st = timed(r, cam)
        info = r.info()
        rows.append({"scene": sid, "W": W, "H": H, "spp": spp, "ms": ms, "rays": int(st.rays), "grays_s": st.rays / ms / 1e6,
                     "smem": info.scene_in_smem, "nodes": info.n_nodes})
        print(rows[-1], flush=True)
        r.close()
    res["perf"] = rows

if "ref" in what:
    rows = []
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu")
    for args in [(1200, 675, 10, 10), (3840, 2160, 10, 8), (1920, 1080, 0, 8), (1024, 1024, 8, 8), (1920, 1080, 9, 2)]:
        t0 = time.time()
        out = subprocess.run([exe] + [str(a) for a in args], cwd=os.path.dirname(exe), capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        row = json.loads(line[-1]) if line else {"error": out.stderr[-500:]}
        row["wall_s"] = time.time() - t0
        rows.append(row)
        print(row, flush=True)
    res["ref_gpu"] = rows

tag = "_".join(sorted(what))
with open(os.path.join(OUT, f"probe_{tag}.json"), "w") as f:
    json.dump(res, f, indent=1)
print("wrote probe", tag)

if "dump9" in what:
    # images for offline analysis of the scene-9 mismatches
    sc = BuiltinScene(9, earth)
    W, H, spp = 240, 135, 10
    cam = sc.camera(W, H, spp, 50)
    want, ost = oracle_render(sc, cam, 0, spp)
    r = Renderer(sc.desc)
    r.render(cam)
    got, _, st = r.readback()
    r.close()
    np.savez_compressed(os.path.join(OUT, "dump_scene9.npz"), gpu=got, oracle=want / spp)
    print("dumped scene 9")

if "sweep" in what:
    rows = []
    sc = BuiltinScene(10)
    cam = sc.camera(3840, 2160, 16, 50)
    r = Renderer(sc.desc)
    for threads, bps in [(512, 1), (256, 2), (128, 4), (128, 5), (128, 6), (256, 3), (64, 10), (384, 2), (1024, 1)]:
        for refill in (32, 24, 16, 12, 8, 4, 1):
            try:
                ms, st = timed(r, cam, reps=2, block_threads=threads, blocks_per_sm=bps, flags=refill << 16)
            except Exception as e:  # noqa: BLE001
                print("skip", threads, bps, refill, str(e)[:80])
                break
            rows.append({"threads": threads, "blocks_per_sm": bps, "refill": refill, "ms": ms, "grays_s": st.rays / ms / 1e6})
            print(rows[-1], flush=True)
    r.close()
    with open(os.path.join(OUT, "probe_sweep.json"), "w") as f:
        json.dump(rows, f, indent=1)

if "wave" in what:
    rows = []
    sc = BuiltinScene(10)
    cam = sc.camera(3840, 2160, 16, 50)
    r = Renderer(sc.desc)
    ms, st = timed(r, cam, reps=2)
    print({"variant": "mega", "ms": ms, "grays_s": st.rays / ms / 1e6}, flush=True)
    for threads, slots in ((512, 64), (512, 96), (448, 96), (384, 96), (384, 128), (320, 128)):
        for idle in (8, 16, 24):
            for leaf in (8, 16, 24):
                for refill in (1, 8, 16):
                    fl = ((slots // 32) << 12) | (idle << 16) | (leaf << 21) | (refill << 26)
                    try:
                        ms, st = timed(r, cam, reps=2, variant=2, block_threads=threads, flags=fl)
                    except Exception as e:  # noqa: BLE001
                        print("skip", threads, slots, idle, leaf, refill, str(e)[:80])
                        break
                    rows.append({"threads": threads, "slots": slots, "idle_exit": idle, "leaf_batch": leaf, "refill_min": refill,
                                 "ms": ms, "grays_s": st.rays / ms / 1e6})
                    print(rows[-1], flush=True)
    r.close()
    with open(os.path.join(OUT, "probe_wave.json"), "w") as f:
        json.dump(rows, f, indent=1)

if "diverge" in what:
    # per-sample comparison: find (pixel, sample) paths whose radiance differs, dump both paths
    out = {}
    for sid, W, H, nS in [(9, 240, 135, 10), (8, 160, 160, 6)]:
        sc = BuiltinScene(sid, earth if sid in (2, 9) else None)
        r = Renderer(sc.desc)
        rows = []
        nbad = 0
        for smp in range(nS):
            cam = sc.camera(W, H, 1, 50)
            want = np.zeros((H, W, 3))
            st = O.oracle_stats()
            oracle.oracle_render(sc.desc, C.byref(cam), smp, smp + 1, 1984, 1, 64, os.cpu_count(), want.ctypes.data, C.byref(st))
            r.render(cam, smp, smp + 1)
            got, _, gst = r.readback()
            ok = (np.abs(got - want) <= 1e-3 * np.abs(want) + 1e-6).all(axis=2)
            bad = np.argwhere(~ok)
            nbad += len(bad)
            for (j, i) in bad[:12]:
                orec = np.zeros((64, 8))
                n = oracle.oracle_trace_path(sc.desc, C.byref(cam), int(i), int(j), smp, 1984, orec.ctypes.data, 64)
                grec = np.zeros((64, 8), np.float32)
                p = A.rt_render_params(sample_begin=smp, sample_end=smp + 1, seed=1984, clear=1)
                rc = r.lib.rt_debug_trace_path(r._h, C.byref(cam), C.byref(p), int(j * W + i), smp, grec.ctypes.data, 64)
                assert rc == 0
                g = []
                for k in range(64):
                    if grec[k, 7] == 0:
                        break
                    hid = int(grec[k, 0:1].view(np.uint32)[0])
                    g.append({"hit_type": (hid >> 29) & 3, "hit_index": hid & 0x1fffffff, "t": float(grec[k, 1]),
                              "mat": int(grec[k, 2:3].view(np.int32)[0]), "front": float(grec[k, 3]),
                              "p": [float(x) for x in grec[k, 4:7]]})
                o = [{"hit": orec[k, 0], "t": orec[k, 1], "mat": int(orec[k, 2]), "front": orec[k, 3],
                      "p": list(orec[k, 4:7]), "mtype": int(orec[k, 7])} for k in range(n)]
                rows.append({"pixel": [int(i), int(j)], "sample": smp, "gpu_rgb": [float(x) for x in got[j, i]],
                             "oracle_rgb": [float(x) for x in want[j, i]], "gpu_path": g, "oracle_path": o})
        r.close()
        out[str(sid)] = {"W": W, "H": H, "samples": nS, "bad": nbad, "paths": rows}
        print("scene", sid, "bad (pixel,sample) pairs:", nbad, "of", W * H * nS, flush=True)
    with open(os.path.join(OUT, "probe_diverge.json"), "w") as f:
        json.dump(out, f)
