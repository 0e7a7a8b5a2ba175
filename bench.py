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
# dram__bytes_read.sum + dram__bytes_write.sum of one RenderMega launch on the 4K frame, from the
# `ncu --set full` capture summarised in profiles/r1_RenderMega_book1_raw_selected.txt (99.9 MB + 50.3 MB):
# the fp32 accumulator, read-modify-written once per pixel per launch, whatever the spp.
NCU_DRAM_BYTES_4K_LAUNCH = 99915520 + 50284800


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
    lib, flags = O.load_ref_stream(fast=True), "g++ -O3 -march=native"
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


def gpu_reference():
    """The reference's own kernel.cu (FP64, cuRAND XORWOW, -arch=sm_100; oracle/_ref/ref_gpu, built by
    oracle/build_ref.py with argv/ray-counter/event-timing patches) on this GPU, on a bounded sample of the
    workload: the bar of BASELINE.json's ">= 10x the reference's CUDA kernel".  A reported baseline."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gpu")
    if not os.path.exists(exe):
        return None
    spp = 8
    try:
        out = subprocess.run([exe, str(ARGS.width), str(ARGS.height), str(ARGS.scene), str(spp), str(SEED)],
                             cwd=os.path.dirname(exe), capture_output=True, text=True, timeout=600)
        row = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:200]}
    return {"value": row["mrays_per_s"], "unit": "Mrays/s", "kind": "reference kernel.cu on this GPU (Render only, CUDA events)",
            "sample": f"{ARGS.width}x{ARGS.height}, {spp} of {ARGS.spp} spp, {row['rays']} rays in {row['render_ms']:.1f} ms",
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


def workload_config(n):
    return {"workload": f"{WORKLOAD['name'] if ARGS.scene == 10 else 'scene%d' % ARGS.scene} {ARGS.width}x{ARGS.height} "
                        f"{ARGS.spp}spp max_depth {WORKLOAD['max_depth']}",
            "scene": ARGS.scene, "width": ARGS.width, "height": ARGS.height, "spp": ARGS.spp,
            "max_depth": WORKLOAD["max_depth"], "seed": SEED,
            "parallelism": f"samples split over {n} GPU(s), one NCCL reduce of the fp32 accumulator" if n > 1
            else "1 GPU",
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
    accum = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)  # caller-owned accumulator (NCCL buffer)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kw = dict(seed=SEED, stream=stream, accum_ptr=accum.data_ptr(), block_threads=ARGS.block_threads,
              blocks_per_sm=ARGS.blocks_per_sm, variant=ARGS.variant, flags=ARGS.flags)

    r = Renderer(sc.desc, device=local, upload_flags=ARGS.upload_flags)
    info_desc_bytes = sc.desc_bytes()

    def step_resident(time_kernel=None):
        flush.zero_()
        if time_kernel is not None:
            time_kernel[0].record()
        r.render(cam, s0, s1, clear=True, **kw)
        if time_kernel is not None:
            time_kernel[1].record()
        if world > 1:
            reduce_accumulators(accum, dst=0)

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
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
        rays_all = rays_rank.clone()
        dist.all_reduce(rays_all, op=dist.ReduceOp.SUM)
    else:
        rays_all = rays_rank
    clk = clocks.stop(world) if rank == 0 else None
    rays_step = int(rays_all.item())
    ms_step = float(ms_total.item()) / ARGS.steps
    value = rays_step / (ms_step * 1e-3) / 1e6
    info = r.info()
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
            rr = Renderer(sc.desc, device=local, upload_flags=ARGS.upload_flags)  # deep copy + bake + BVH + H2D of every table
            t.append(time.perf_counter())
            rr.render(cam, s0, s1, clear=True, **kw)
            if trace:
                torch.cuda.synchronize()
            t.append(time.perf_counter())
            if world > 1:
                reduce_accumulators(accum, dst=0)
            if rank == 0:
                stt = A.rt_stats()
                rc = lib.rt_readback(rr._h, C.c_void_p(accum.data_ptr()), C.c_void_p(host_frame.data_ptr()), None,
                                     C.byref(stt))
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
               "h2d_bytes_per_step": int(info_desc_bytes) * world, "d2h_bytes_per_step": H * W * 3 * 4 + 32,
               "what": "rt_scene_upload(host scene) + rt_render + reduce + rt_readback(linear fp32 -> pinned host)"}
        assert float(host_frame[:3 * W].sum()) > 0.0 or rank != 0

    if rank != 0:
        if world > 1:
            dist.barrier(device_ids=[local])
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (RenderMega), FP32-issue bound
    topo = reference_topology_counts()
    flop_ray = FLOP_BOX * topo["n_box"] + FLOP_SPHERE * topo["n_sphere"] + FLOP_QUAD * topo["n_quad"] + FLOP_SHADE
    bytes_ray = BYTES_NODE * (topo["n_box"] + topo["n_sphere"] + topo["n_quad"])
    peak = C.c_double()
    sms = C.c_double()
    assert r.lib.rt_measure_fp32_peak(local, C.byref(peak), C.byref(sms)) == 0
    kernel_ms_max = float(kms.item())
    rays_kernel = int(rays_rank.item())
    achieved = rays_kernel * flop_ray / (kernel_ms_max * 1e-3) / 1e12
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_bytes = H * W * 3 * 4 * 2  # accumulator read-modify-write, once per pixel per launch
    roofline = {
        "bound": "fp32", "kernel": {1: "RenderMega", 2: "RenderWave", 3: "RenderHeadTail", 4: "RenderHitQueue"}.get(info.variant, "?"),
        "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
        "frac": achieved / peak.value if peak.value else None,
        "traffic": NCU_DRAM_BYTES_4K_LAUNCH if (W, H) == (3840, 2160) else None,
        "peak_source": "FFMA microbenchmark run live on this GPU (rt_measure_fp32_peak); MEASURED_PEAKS.json "
                       "holds only HBM and bf16-tensor peaks, neither of which bounds this path",
        "flop_per_ray": flop_ray, "l1_bytes_per_ray": bytes_ray, "reference_topology": topo,
        "kernel_ms": kernel_ms_max, "rays_per_launch": rays_kernel,
        "grays_per_s_kernel": rays_kernel / (kernel_ms_max * 1e-3) / 1e9,
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
    line = {
        "metric": "Mrays/s incl. secondary rays", "value": value, "unit": "Mrays/s", "n_gpus": world,
        "steps": ARGS.steps, "warmup": ARGS.warmup, "ms_per_step": ms_step, "ms_per_frame": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(world), "rays_per_frame": rays_step,
        "clocks": clk, "e2e": e2e, "gpu_launches": ARGS.steps * world + (ARGS.steps * (world + 1) if e2e else 0),
        "kernel": {"features": info.features, "scene_in_smem": info.scene_in_smem, "nodes": info.n_nodes,
                   "prims": info.n_prims_baked},
        "roofline": roofline,
    }
    if world == 1 and not ARGS.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
        line["gpu_reference"] = gpu_reference()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(device_ids=[local])
        dist.destroy_process_group()
    return 0


ARGS = parse_args()

if __name__ == "__main__":
    sys.exit(run_reference_arm() if ARGS.impl == "reference" else run_b200_arm())
