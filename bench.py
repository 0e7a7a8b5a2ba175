#!/usr/bin/env python3
"""Benchmark of the render hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own scene code on the host cores

One "step" = one frame of the workload: Book 1 final scene, 3840x2160, 1024 spp,
max_depth 50 (BASELINE.json configs[1]).  With N ranks (torchrun, one per GPU)
rank k renders global samples [k*spp/N, (k+1)*spp/N) of every pixel and the fp32
accumulators are summed with ONE NCCL reduce to rank 0 -- total work is fixed, so
"scaling" is "strong".

Printed JSON line (rank 0):
  value        Mrays/s incl. secondary rays, scene resident in HBM, render (+ reduce) only
  e2e          same metric through the public API with HOST buffers: rt_scene_upload from
               the host scene description, rt_render, reduce, rt_readback of the linear
               fp32 frame into pinned host memory -- every step
  roofline     FP32-issue roofline of the dominant kernel (RenderMega): the path is not
               HBM- or tensor-bound; an "hbm" sub-object is given beside it
  cpu_baseline the reference's own headers compiled for the host (oracle/_ref), all cores,
               on a bounded sample of the same frame
Only this file's cpu_baseline / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = {"scene": 10, "name": "book1_final", "width": 3840, "height": 2160, "spp": 1024, "max_depth": 50}
SEED = 1984
# algorithmic work per ray on the REFERENCE-topology BVH
FLOP_BOX, FLOP_SPHERE, FLOP_QUAD, FLOP_SHADE = 24.0, 30.0, 28.0, 120.0
BYTES_NODE = 32.0


def earth_texels():
    """Texels of the reference's earthmap.jpg as its RtwImage produces them: the committed fixture (the bench must
    not depend on /root/reference; the product decodes the JPEG itself when given a path)."""
    import numpy as np
    return np.ascontiguousarray(np.load(os.path.join(ROOT, "tests", "golden", "earthmap_rgb8.npz"))["rgb"])


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # development overrides; the driver never passes these (the JSON line names what ran)
    ap.add_argument("--width", type=int, default=WORKLOAD["width"])
    ap.add_argument("--height", type=int, default=WORKLOAD["height"])
    ap.add_argument("--spp", type=int, default=WORKLOAD["spp"])
    ap.add_argument("--scene", type=int, default=WORKLOAD["scene"])
    ap.add_argument("--block-threads", type=int, default=0)
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--flags", type=lambda x: int(x, 0), default=0)
    ap.add_argument("--upload-flags", type=lambda x: int(x, 0), default=0)
    ap.add_argument("--single-process", action="store_true",
                    help="with --gpus N > 1 and no torchrun: ONE process drives the N devices through the C ABI "
                         "(rt_upload_options.n_devices; reduce inside rt_readback)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------- CPU arms
def load_cpu_reference():
    """-> (kind, what, callable(W,H,s0,s1,threads) -> (rays, seconds)).  The CPU arm: the reference's own headers
    compiled for the host, -O3 -march=native (oracle/_ref/libref_stream_fast.so; the -O2 -ffp-contract=off build
    beside it is the bit-exact pin, not the baseline), else the FP64 oracle port."""
    from oracle import bindings as O
    from raytracinginoneweekendincuda_b200 import BuiltinScene
    import numpy as np
    sid = ARGS.scene
    earth = None
    if sid in (2, 9):
        earth = earth_texels()
    lib, flags = None, "g++ -O3 -march=native"
    if os.path.exists(O.ref_stream_path(fast=True)):
        # -march=native was resolved where the library was BUILT; this host may be a different CPU.  Try it in a child
        # process first (an illegal instruction kills the child, not the bench).
        probe = ("import ctypes as C, numpy as np, sys; sys.path.insert(0, %r); from oracle import bindings as O; "
                 "l = O.load_ref_stream(fast=True); o = np.zeros((8, 8, 3)); s = O.ref_stream_stats(); "
                 "sys.exit(l.ref_stream_render(10, 8, 8, 0, 1, 50, 1984, None, 0, 0, 1, o.ctypes.data, C.byref(s)))" % ROOT)
        try:
            if subprocess.run([sys.executable, "-c", probe], capture_output=True, timeout=120).returncode == 0:
                lib = O.load_ref_stream(fast=True)
        except Exception:  # noqa: BLE001
            lib = None
    if lib is None:
        lib, flags = O.load_ref_stream(fast=False), "g++ -O2 -ffp-contract=off (the pin build; no -O3 build present)"
    if lib is not None:
        def run_ref(W, H, s0, s1, threads):
            out = np.zeros((H, W, 3), np.float64)
            st = O.ref_stream_stats()
            ep, ew, eh = (earth.ctypes.data, earth.shape[1], earth.shape[0]) if earth is not None else (None, 0, 0)
            t0 = time.perf_counter()
            rc = lib.ref_stream_render(sid, W, H, s0, s1, WORKLOAD["max_depth"], SEED, ep, ew, eh, threads,
                                       out.ctypes.data, C.byref(st))
            dt = time.perf_counter() - t0
            assert rc == 0
            return int(st.rays), dt
        return "reference", f"reference headers (FP64) compiled for the host by {flags}, row-parallel std::thread", run_ref
    lib = O.load_oracle()
    sc = BuiltinScene(sid, earth)

    def run_port(W, H, s0, s1, threads):
        cam = sc.camera(W, H, WORKLOAD["spp"], WORKLOAD["max_depth"])
        out = np.zeros((H, W, 3), np.float64)
        st = O.oracle_stats()
        t0 = time.perf_counter()
        rc = lib.oracle_render(sc.desc, C.byref(cam), s0, s1, SEED, 1, 64, threads, out.ctypes.data, C.byref(st))
        dt = time.perf_counter() - t0
        assert rc == 0
        return int(st.rays), dt
    return "port", "oracle/rt_oracle.cpp (FP64 restatement, g++ -O2 -ffp-contract=off), row-parallel std::thread", run_port


def reference_topology_counts(scene_id=None):
    """n_box, n_sphere, n_quad per ray on the reference-topology BVH, measured by the
    FP64 oracle on a small sample of the workload (the algorithmic work definition
    of SURVEY.md 8d is implementation independent)."""
    from oracle import bindings as O
    from raytracinginoneweekendincuda_b200 import BuiltinScene
    import numpy as np
    lib = O.load_oracle()
    sid = ARGS.scene if scene_id is None else scene_id
    earth = None
    if sid in (2, 9):
        earth = earth_texels()
    sc = BuiltinScene(sid, earth)
    W, H = 480, 270
    cam = sc.camera(W, H, 1, WORKLOAD["max_depth"])
    out = np.zeros((H, W, 3), np.float64)
    st = O.oracle_stats()
    lib.oracle_render(sc.desc, C.byref(cam), 0, 1, SEED, 1, 64, os.cpu_count() or 1, out.ctypes.data, C.byref(st))
    r = max(1, st.rays)
    return {"n_box": st.box_tests / r, "n_sphere": st.sphere_tests / r, "n_quad": st.quad_tests / r,
            "rays_per_path": st.rays / max(1, st.paths)}


def cpu_baseline(budget_s=12.0):
    kind, what, run = load_cpu_reference()
    cores = os.cpu_count() or 1
    W, H = ARGS.width, ARGS.height
    rays1, t1 = run(W, H, 0, 1, cores)  # calibration sample
    n = int(max(1, min(16, round(budget_s / max(t1, 1e-3)))))
    rays, dt = run(W, H, 1, 1 + n, cores)
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": f"{W}x{H}, samples [1,{1 + n}) of {ARGS.spp}, {rays} rays in {dt:.2f} s", "what": what}


def gpu_reference(scene, W, H, spp):
    """The reference's own kernel.cu (FP64, cuRAND XORWOW, -arch=sm_100; oracle/_ref/ref_gpu, built by
    oracle/build_ref.py with argv/ray-counter/event-timing patches) on this GPU, on a bounded sample of the
    workload: the bar of BASELINE.json's ">= 10x the reference's CUDA kernel".  A reported baseline."""
    from oracle import bindings as O
    exe = O.ref_gpu_path()
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, str(W), str(H), str(scene), str(spp), str(SEED)],
                             cwd=os.path.dirname(exe), capture_output=True, text=True, timeout=600)
        row = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}
    return {"value": row["mrays_per_s"], "unit": "Mrays/s", "kind": "reference kernel.cu on this GPU (Render only, CUDA events)",
            "sample": f"{W}x{H}, {spp} spp, {row['rays']} rays in {row['render_ms']:.1f} ms",
            "render_init_ms": row["render_init_ms"]}


def run_reference_arm():
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, what, run = load_cpu_reference()
    cores = os.cpu_count() or 1
    W, H = ARGS.width, ARGS.height
    # one step = one sample per pixel of the frame (1/spp of the workload), all host threads
    for w in range(ARGS.warmup):
        run(W, H, w, w + 1, cores)
    total_rays, total_t = 0, 0.0
    for k in range(ARGS.steps):
        rays, dt = run(W, H, ARGS.warmup + k, ARGS.warmup + k + 1, cores)
        total_rays += rays
        total_t += dt
    v = total_rays / total_t / 1e6
    sample = f"{W}x{H}, 1 of {ARGS.spp} spp per step ({total_rays // max(1, ARGS.steps)} rays/step)"
    line = {
        "impl": "reference", "metric": "Mrays/s incl. secondary rays", "value": v, "unit": "Mrays/s",
        "n_gpus": ARGS.gpus, "steps": ARGS.steps, "warmup": ARGS.warmup, "ms_per_step": total_t / ARGS.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1),
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample, "what": what},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(n, single=False):
    if n > 1 and single:
        par = (f"samples split over {n} GPUs driven by ONE process through the C ABI (rt_upload_options.n_devices); the fp32 "
               "accumulators are summed on device 0 inside rt_readback")
    elif n > 1:
        par = f"samples split over {n} GPU(s), one process each, one NCCL reduce of the fp32 accumulator"
    else:
        par = "1 GPU"
    return {"workload": f"{WORKLOAD['name'] if ARGS.scene == 10 else 'scene%d' % ARGS.scene} {ARGS.width}x{ARGS.height} "
                        f"{ARGS.spp}spp max_depth {WORKLOAD['max_depth']}",
            "scene": ARGS.scene, "width": ARGS.width, "height": ARGS.height, "spp": ARGS.spp,
            "max_depth": WORKLOAD["max_depth"], "seed": SEED, "parallelism": par,
            "l2": "flushed between steps (256 MiB device memset inside the timed region); scene is shared-memory "
                  "resident by design, the 99.5 MB accumulator is written once per pixel per step"}


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self, gpus):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 8:
                continue
            try:
                if int(c[0]) >= gpus:
                    continue
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(names, c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------ GPU arm
OTHER_CONFIGS = [  # BASELINE.json configs[2..4] at their image sizes, spp bounded so the whole bench stays within minutes
    {"name": "configs[2] bouncing spheres (motion blur + checker)", "scene": 0, "width": 1920, "height": 1080, "spp": 256, "of": 512},
    {"name": "configs[3] Cornell smoke (quads, instances, media)", "scene": 8, "width": 1024, "height": 1024, "spp": 256, "of": 4096},
    {"name": "configs[4] Book 2 final", "scene": 9, "width": 3840, "height": 2160, "spp": 32, "of": 10000},
]


def flop_per_ray(topo):
    return FLOP_BOX * topo["n_box"] + FLOP_SPHERE * topo["n_sphere"] + FLOP_QUAD * topo["n_quad"] + FLOP_SHADE


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the shipping kernel, from the committed
    `ncu --set full` capture (profiles/r2_traffic.json, written by tools/ncu_summary.py): bytes or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            return json.load(f).get(kernel_key, {}).get("dram_bytes")
    except (OSError, ValueError):
        return None


def time_config(torch, Renderer, sc, cfg, local, peak_tflops, reps=3):
    """One of the other BASELINE configs on this GPU: CUDA-event time of a launch, roofline with that scene's own
    reference-topology counts, the reference's kernel.cu on the same frame beside it."""
    W, H, spp = cfg["width"], cfg["height"], cfg["spp"]
    cam = sc.camera(W, H, spp, WORKLOAD["max_depth"])
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
    r = Renderer(sc.desc, device=local)
    r.render(cam, stream=stream)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.render(cam, stream=stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    _, _, st = r.readback(linear=False)
    info = r.info()
    r.close()
    topo = reference_topology_counts(cfg["scene"])
    fpr = flop_per_ray(topo)
    achieved = int(st.rays) * fpr / (best * 1e-3) / 1e12
    out = {"config": cfg["name"], "scene": cfg["scene"], "size": f"{W}x{H}", "spp": spp, "spp_of_config": cfg["of"],
           "ms": best, "mrays_per_s": int(st.rays) / (best * 1e-3) / 1e6, "rays": int(st.rays),
           "kernel": {"name": "RenderHitQueue", "features": info.features, "scene_in_smem": info.scene_in_smem,
                      "block_threads": info.block_threads, "registers": info.registers, "nodes": info.n_nodes},
           "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                        "frac": achieved / peak_tflops if peak_tflops else None, "flop_per_ray": fpr,
                        "reference_topology": topo},
           "timing": "best of %d launches, CUDA events, 256 MiB L2 flush before each" % reps}
    ref = gpu_reference(cfg["scene"], W, H, 2 if cfg["scene"] == 9 else 4)
    if ref:
        out["gpu_reference"] = ref
    return out


def short_frame_e2e(torch, Renderer, local, steps):
    """BASELINE configs[0] (Book 1 final 1200x675, 10 spp: ~1 ms of kernel) end to end, where upload and readback are
    not hidden behind a second of rendering: host scene -> rt_scene_upload -> rt_render -> rt_readback of the
    quantised frame (what the reference writes to its PPM) into host memory."""
    from raytracinginoneweekendincuda_b200 import BuiltinScene
    sc = BuiltinScene(10)
    W, H, spp = 1200, 675, 10
    cam = sc.camera(W, H, spp, WORKLOAD["max_depth"])
    parts = {"upload_ms": 0.0, "render_ms": 0.0, "readback_ms": 0.0, "free_ms": 0.0}
    rays = 0
    total = 0.0
    for k in range(steps + 1):
        t = [time.perf_counter()]
        rr = Renderer(sc.desc, device=local)
        t.append(time.perf_counter())
        rr.render(cam)
        rr.sync()
        t.append(time.perf_counter())
        _, s8, st = rr.readback(linear=False, srgb8=True)
        t.append(time.perf_counter())
        rr.close()
        t.append(time.perf_counter())
        if k == 0:
            continue  # warm-up
        for name, (x, y) in zip(parts, zip(t, t[1:])):
            parts[name] += (y - x) * 1e3 / steps
        total += (t[-1] - t[0]) * 1e3 / steps
        rays = int(st.rays)
    return {"workload": "book1_final 1200x675 10spp (BASELINE configs[0])", "ms_per_frame": total, "rays": rays,
            "mrays_per_s": rays / (total * 1e-3) / 1e6, **parts, "h2d_bytes": sc.desc_bytes(), "d2h_bytes": W * H * 3 + 32,
            "what": "wall clock around upload + render + readback(srgb8) + free, host buffers"}


def host_build_times(reps=5):
    """what the host does before the path -- validation, instance baking, box recognition, hoisting, the
    binned-SAH build and packing (csrc/rt_pack.hpp, everything rt_scene_upload does short of the copy) -- timed through
    the host-only rt_scene_pack_info: best of `reps`, ms, for the headline scene and the largest one."""
    import ctypes as C
    from raytracinginoneweekendincuda_b200 import BuiltinScene, _abi as A
    out = {}
    for name, sid in (("book1_final (485 primitives)", 10), ("book2_final (3 409 primitives, 400 boxes)", 9)):
        sc = BuiltinScene(sid, earth_texels() if sid == 9 else None)
        info, opt = A.rt_pack_info(), A.rt_upload_options()
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            rc = sc.lib.rt_scene_pack_info(sc.desc, C.byref(opt), C.byref(info))
            best = min(best, (time.perf_counter() - t0) * 1e3)
        if rc == 0:
            out[name] = {"ms": round(best, 3), "nodes": info.n_nodes, "boxes": info.n_boxes, "hoisted": info.n_hoisted}
    return out


def run_b200_arm():
    import numpy as np
    import torch
    import torch.distributed as dist

    from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer
    from raytracinginoneweekendincuda_b200.multigpu import reduce_accumulators, sample_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the render path has no CPU fallback "
                         "(use --impl reference for the host baseline)")
    single = ARGS.single_process and world == 1 and ARGS.gpus > 1  # one process drives ARGS.gpus devices via the C ABI
    n_dev = ARGS.gpus if single else world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created: send stdout to
        # stderr while that happens, so that stdout carries the one JSON line and nothing else
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    W, H, spp = ARGS.width, ARGS.height, ARGS.spp
    earth = earth_texels() if ARGS.scene in (2, 9) else None
    sc = BuiltinScene(ARGS.scene, earth)  # host scene description (the caller's input)
    cam = sc.camera(W, H, spp, WORKLOAD["max_depth"])
    s0, s1 = sample_range(rank, world, spp)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    common = dict(seed=SEED, block_threads=ARGS.block_threads, blocks_per_sm=ARGS.blocks_per_sm, variant=ARGS.variant,
                  flags=ARGS.flags)
    if single:
        accum = None
        r = Renderer(sc.desc, devices=list(range(n_dev)), upload_flags=ARGS.upload_flags)
        kw = dict(common)
    else:
        accum = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)  # caller-owned accumulator (NCCL buffer)
        r = Renderer(sc.desc, device=local, upload_flags=ARGS.upload_flags)
        kw = dict(common, stream=stream, accum_ptr=accum.data_ptr())
    info_desc_bytes = sc.desc_bytes()

    def step_resident(time_kernel=None):
        flush.zero_()
        if time_kernel is not None:
            time_kernel[0].record()
        r.render(cam, s0, s1, clear=True, **kw)
        if single:
            r.readback(linear=False)  # rt_readback: waits for every device, reduces onto device 0
        if time_kernel is not None:
            time_kernel[1].record()
        if world > 1:
            reduce_accumulators(accum, dst=0)

    if single:  # (an NCCL communicator created inside rt_readback prints its banner on stdout as well)
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            step_resident()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    for _ in range(max(ARGS.warmup, 0)):
        step_resident()
    barrier()
    clocks = ClockSampler()
    if rank == 0:
        clocks.start()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ARGS.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(ARGS.steps):
        step_resident(kev[k])
    e1.record()
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / max(1, ARGS.steps)
    _, _, st = r.readback(linear=False)
    rays_rank = torch.tensor([int(st.rays)], dtype=torch.int64, device=dev)  # one step (clear=True resets counters)
    kms = torch.tensor([kernel_ms], dtype=torch.float64, device=dev)
    per_rank_ms = [kernel_ms]
    if single:
        tm = r.timing()
        per_rank_ms = [float(tm.render_ms[k]) for k in range(tm.n_devices)]
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(kms) for _ in range(world)]
        dist.all_gather(gathered, kms)
        per_rank_ms = [float(x.item()) for x in gathered]
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
        rays_all = rays_rank.clone()
        dist.all_reduce(rays_all, op=dist.ReduceOp.SUM)
    else:
        rays_all = rays_rank
    clk = clocks.stop(n_dev) if rank == 0 else None
    rays_step = int(rays_all.item())
    ms_step = float(ms_total.item()) / ARGS.steps
    value = rays_step / (ms_step * 1e-3) / 1e6
    info = r.info()

    # ---- N ranks: the reduced frame against a 1-GPU render of the same samples, outside the timing
    nrank_parity = None
    if (world > 1 or single) and not ARGS.no_parity:
        if world > 1:
            accum.zero_()
            r.render(cam, s0, s1, clear=True, **kw)
            reduce_accumulators(accum, dst=0)
            torch.cuda.synchronize()
            many = accum.clone() if rank == 0 else None
            dist.barrier(device_ids=[local])
        else:
            r.render(cam, s0, s1, clear=True, **kw)
            lin_many, _, _ = r.readback(linear=True)  # mean radiance, reduced on device 0
            many = torch.from_numpy(lin_many.reshape(-1)).to(dev) * float(spp)
        if rank == 0:
            # the same slices on ONE GPU, accumulated in rank order: what the N accumulators must sum to.  The peer-
            # memory reduction adds in device order too (bit-identical expected); NCCL may associate differently.
            one = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
            r1 = Renderer(sc.desc, device=local)
            for k in range(n_dev):
                b0, b1 = sample_range(k, n_dev, spp)
                r1.render(cam, b0, b1, clear=False, seed=SEED, stream=stream, accum_ptr=one.data_ptr(), variant=ARGS.variant)
            torch.cuda.synchronize()
            r1.close()
            diff = (many - one).abs()
            rel = diff / one.abs().clamp_min(1e-3)
            ok = bool((diff <= 1e-6 * one.abs() + 1e-7 * spp).all().item())
            nrank_parity = {"allclose_rtol_1e-6": ok, "max_rel_diff": float(rel.max().item()),
                            "bit_identical_fraction": float((many == one).float().mean().item()),
                            "pixels_compared": W * H,
                            "what": f"{n_dev}-GPU reduced sums vs the same {n_dev} sample slices rendered on 1 GPU and "
                                    "accumulated in rank order (all pixels of the frame)"}
            del one, many
    r.close()

    # ---- e2e: host description -> upload -> render -> reduce -> linear fp32 frame in pinned host memory
    e2e = None
    if not ARGS.no_e2e:
        host_frame = torch.empty(H * W * 3, dtype=torch.float32).pin_memory()
        A = sys.modules["raytracinginoneweekendincuda_b200._abi"]
        lib = r.lib
        trace = os.environ.get("RT_BENCH_TRACE")

        def step_e2e():
            t = [time.perf_counter()]
            flush.zero_()
            if single:
                rr = Renderer(sc.desc, devices=list(range(n_dev)), upload_flags=ARGS.upload_flags)
            else:
                rr = Renderer(sc.desc, device=local, upload_flags=ARGS.upload_flags)  # deep copy + bake + BVH + H2D
            t.append(time.perf_counter())
            rr.render(cam, s0, s1, clear=True, **kw)
            if trace:
                torch.cuda.synchronize()
            t.append(time.perf_counter())
            if world > 1:
                reduce_accumulators(accum, dst=0)
            if rank == 0:
                stt = A.rt_stats()
                rc = lib.rt_readback(rr._h, C.c_void_p(accum.data_ptr()) if accum is not None else None,
                                     C.c_void_p(host_frame.data_ptr()), None, C.byref(stt))
                assert rc == 0, lib.rt_last_error()
            else:
                rr.sync()
            t.append(time.perf_counter())
            rr.close()
            t.append(time.perf_counter())
            if trace:
                print("e2e step: upload %.1f render %.1f reduce+readback %.1f free %.1f ms" %
                      tuple(1e3 * (b - a) for a, b in zip(t, t[1:])), file=sys.stderr, flush=True)

        step_e2e()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(ARGS.steps):
            step_e2e()
        f1.record()
        barrier()
        ems = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        ems_step = float(ems.item()) / ARGS.steps
        e2e = {"value": rays_step / (ems_step * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": ems_step,
               "h2d_bytes_per_step": int(info_desc_bytes) * n_dev, "d2h_bytes_per_step": H * W * 3 * 4 + 32,
               "what": "rt_scene_upload(host scene) + rt_render + reduce + rt_readback(linear fp32 -> pinned host)"}
        assert float(host_frame[:3 * W].sum()) > 0.0 or rank != 0

    if rank != 0:
        if world > 1:
            dist.barrier(device_ids=[local])
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel, FP32-issue bound
    topo = reference_topology_counts()
    flop_ray = flop_per_ray(topo)
    bytes_ray = BYTES_NODE * (topo["n_box"] + topo["n_sphere"] + topo["n_quad"])
    peak = C.c_double()
    sms = C.c_double()
    assert r.lib.rt_measure_fp32_peak(local, C.byref(peak), C.byref(sms)) == 0
    kernel_ms_max = float(kms.item()) if not single else max(per_rank_ms)
    rays_kernel = int(rays_rank.item()) if not single else rays_step // n_dev
    achieved = rays_kernel * flop_ray / (kernel_ms_max * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_bytes = H * W * 3 * 4 * 2  # accumulator read-modify-write, once per pixel per launch
    kname = {1: "RenderMega", 2: "RenderWave", 3: "RenderHeadTail", 4: "RenderHitQueue"}.get(info.variant, "?")
    roofline = {
        "bound": "fp32", "kernel": kname,
        "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
        "frac": achieved / peak.value if peak.value else None,
        "traffic": ncu_traffic(f"{kname}<{info.features},{info.scene_in_smem},0>") if (W, H) == (3840, 2160) else None,
        "traffic_source": "profiles/r2_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch of this "
                          "kernel on this frame size (`ncu --set full`); the accumulator's read-modify-write, whatever the spp",
        "peak_source": "FFMA microbenchmark run live on this GPU (rt_measure_fp32_peak); MEASURED_PEAKS.json "
                       "holds only HBM and bf16-tensor peaks, neither of which bounds this path",
        "flop_per_ray": flop_ray, "l1_bytes_per_ray": bytes_ray, "reference_topology": topo,
        "kernel_ms": kernel_ms_max, "rays_per_launch": rays_kernel,
        "grays_per_s_kernel": rays_kernel / (kernel_ms_max * 1e-3) / 1e9,
        "registers": info.registers, "block_threads": info.block_threads,
        # secondary bound of SURVEY 8(d): node/primitive fetches (algorithmic bytes on the reference topology)
        # against the shared-memory/L1 bandwidth, nominal 128 B/clk/SM at the clock seen during the run
        "l1": {"achieved": rays_kernel * bytes_ray / (kernel_ms_max * 1e-3) / 1e9, "unit": "GB/s",
               "peak": sms.value * 128.0 * ((clk or {}).get("sm_mhz") or 1965.0) * 1e6 / 1e9,
               "peak_source": "nominal: SMs x 128 B/clk x sampled SM clock (not measured)"},
        "hbm": {"achieved": hbm_bytes / (kernel_ms_max * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_bytes / (kernel_ms_max * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                "algorithmic_bytes_per_launch": hbm_bytes},
    }
    roofline["l1"]["frac"] = roofline["l1"]["achieved"] / roofline["l1"]["peak"]
    launches_per_step = n_dev + (1 if single else 0)  # render kernels (+ the reduce/resolve kernel)
    line = {
        "metric": "Mrays/s incl. secondary rays", "value": value, "unit": "Mrays/s", "n_gpus": n_dev,
        "steps": ARGS.steps, "warmup": ARGS.warmup, "ms_per_step": ms_step, "ms_per_frame": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(n_dev, single), "rays_per_frame": rays_step,
        "clocks": clk, "e2e": e2e,
        "gpu_launches": ARGS.steps * launches_per_step + ((ARGS.steps + 1) * (launches_per_step + 1) if e2e else 0),
        "kernel": {"features": info.features, "scene_in_smem": info.scene_in_smem, "nodes": info.n_nodes,
                   "prims": info.n_prims_baked, "registers": info.registers, "block_threads": info.block_threads},
        "per_rank_kernel_ms": {"min": min(per_rank_ms), "max": max(per_rank_ms), "all": per_rank_ms},
        "roofline": roofline,
    }
    if nrank_parity is not None:
        line["nrank_parity"] = nrank_parity["allclose_rtol_1e-6"]
        line["nrank_parity_detail"] = nrank_parity
    if n_dev == 1 and not ARGS.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
        line["gpu_reference"] = gpu_reference(ARGS.scene, W, H, 8)
    if n_dev == 1 and not ARGS.no_configs:
        cfgs = []
        for cfg in OTHER_CONFIGS:
            scn = BuiltinScene(cfg["scene"], earth_texels() if cfg["scene"] in (2, 9) else None)
            cfgs.append(time_config(torch, Renderer, scn, cfg, local, peak.value))
        line["configs"] = cfgs
        line["e2e_short_frame"] = short_frame_e2e(torch, Renderer, local, max(3, ARGS.steps))
        line["host_build_ms"] = host_build_times()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


ARGS = parse_args()

if __name__ == "__main__":
    sys.exit(run_reference_arm() if ARGS.impl == "reference" else run_b200_arm())
