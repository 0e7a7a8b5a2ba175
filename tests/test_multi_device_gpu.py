"""Multi-device and output-path entry points of the C ABI on real GPUs: one process driving N devices
(rt_upload_options.n_devices), the reduction on device 0 (peer-memory fused kernel and NCCL), progressive output,
per-device timing.  Tests that need two GPUs skip on a one-GPU box (`gpurun --gpus 2` runs them)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, A, oracle_render
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def render(sc, cam, **kw):
    r = Renderer(sc.desc, **kw)
    r.render(cam)
    t = r.timing()
    lin, s8, st = r.readback(linear=True, srgb8=True)
    info = r.info()
    r.close()
    return lin, s8, st, info, t


def test_a_device_list_of_one_is_the_single_device_path():
    sc = BuiltinScene(10)
    cam = sc.camera(96, 54, 4, 50)
    a, a8, sa, ia, ta = render(sc, cam)
    b, b8, sb, ib, tb = render(sc, cam, devices=[0])
    assert np.array_equal(a, b) and np.array_equal(a8, b8) and sa.rays == sb.rays
    assert ia.n_devices == ib.n_devices == 1 and ib.reduce_path == 0
    assert ta.n_devices == 1 and ta.render_ms[0] > 0.0
    assert ia.registers > 0 and ia.block_threads % 32 == 0


@pytest.mark.parametrize("flags,path", [(0, 1), (A.RT_UPLOAD_REDUCE_NCCL, 2)])
@pytest.mark.parametrize("sid,W,H,spp", [(10, 320, 180, 8), (9, 160, 90, 4)])
def test_two_devices_render_the_one_device_frame(earth, flags, path, sid, W, H, spp):
    """the N-device frame equals the 1-device frame up to fp32 summation order -- same sample set (the
    stream is keyed on the global sample index), ray for ray -- through both reductions: the peer-memory kernel that
    sums device 1's accumulator straight over NVLink, and ncclReduce."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    sc = BuiltinScene(sid, earth if sid == 9 else None)
    cam = sc.camera(W, H, spp, 50)
    one, one8, s1, _, _ = render(sc, cam)
    two, two8, s2, info, t = render(sc, cam, devices=[0, 1], upload_flags=flags)
    assert info.n_devices == 2 and info.reduce_path == path
    assert s1.rays == s2.rays
    assert np.allclose(one, two, rtol=3e-6, atol=1e-7)
    assert np.abs(one8.astype(int) - two8.astype(int)).max() <= 1
    assert t.n_devices == 2 and t.render_ms[0] > 0 and t.render_ms[1] > 0
    # device order does not matter, only the sample slices move
    rev, _, s3, _, _ = render(sc, cam, devices=[1, 0], upload_flags=flags)
    assert s3.rays == s1.rays and np.allclose(one, rev, rtol=3e-6, atol=1e-7)


def test_two_devices_importance_sampling_equals_one_device():
    """RT_FLAG_IMPORTANCE through a two-device handle: the light table travels with the arena, the draws are keyed on the
    global sample index, so the two-device frame is the one-device frame up to fp32 summation order."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    sc = BuiltinScene(7)
    cam = sc.camera(96, 96, 8, 50)
    out = []
    for devices in (None, [0, 1]):
        r = Renderer(sc.desc, devices=devices) if devices else Renderer(sc.desc)
        r.render(cam, flags=A.RT_FLAG_IMPORTANCE)
        lin, _, st = r.readback()
        r.close()
        out.append((lin, st.rays))
    assert out[0][1] == out[1][1]
    assert np.allclose(out[0][0], out[1][0], rtol=3e-6, atol=1e-7)


def test_two_devices_accumulate_across_calls():
    """Render, read back (reduce), render more samples, read back again: device 0 keeps the running total, the other
    accumulators restart from zero after each reduction."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    sc = BuiltinScene(10)
    cam = sc.camera(128, 72, 8, 50)
    full, _, sf, _, _ = render(sc, cam)
    r = Renderer(sc.desc, devices=[0, 1])
    r.render(cam, 0, 3, clear=True)
    r.readback(linear=False)
    r.render(cam, 3, 8, clear=False)
    got, _, st = r.readback()
    r.close()
    assert st.rays == sf.rays
    assert np.allclose(got, full, rtol=3e-6, atol=1e-7)


def test_the_callers_current_device_is_left_alone():
    import torch
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    torch.cuda.set_device(0)
    sc = BuiltinScene(10)
    cam = sc.camera(64, 36, 2, 50)
    r = Renderer(sc.desc, device=1)
    assert torch.cuda.current_device() == 0
    r.render(cam)
    lin, _, _ = r.readback()
    r.close()
    x = torch.zeros(4, device="cuda")
    assert x.device.index == 0 and torch.cuda.current_device() == 0 and lin.any()


def test_multi_device_argument_errors(lib):
    sc = BuiltinScene(10)
    h = C.c_void_p()
    ids = (C.c_int32 * 2)(0, 0)
    opt = A.rt_upload_options(n_devices=2, device_ids=ids)
    assert lib.rt_scene_upload(sc.desc, C.byref(opt), C.byref(h)) == A.RT_ERR_INVALID
    assert b"twice" in lib.rt_last_error()
    ids = (C.c_int32 * 2)(0, 99)
    opt = A.rt_upload_options(n_devices=2, device_ids=ids)
    assert lib.rt_scene_upload(sc.desc, C.byref(opt), C.byref(h)) == A.RT_ERR_INVALID
    cam = sc.camera(16, 8, 0, 50)  # spp 0: rt_readback would divide by it
    r = Renderer(sc.desc)
    p = A.rt_render_params(sample_begin=0, sample_end=1, seed=1)
    assert lib.rt_render(r._h, C.byref(cam), C.byref(p)) == A.RT_ERR_INVALID
    r.close()


@pytest.mark.parametrize("devices", [None, [0, 1]])
def test_progressive_output(devices):
    """rt_render_progressive: one callback per batch with the mean of the samples done so far; the last frame is
    the frame of a plain render; the callback of batch k runs while batch k+1 renders (not checked here)."""
    if devices and n_gpus() < len(devices):
        pytest.skip("needs 2 GPUs")
    sc = BuiltinScene(10)
    W, H, spp, batch = 160, 90, 10, 4
    cam = sc.camera(W, H, spp, 50)
    frames = []
    r = Renderer(sc.desc, devices=devices)
    r.render_progressive(cam, batch, lambda lin, s8, done, total: frames.append((lin.copy(), s8.copy(), done, total)),
                         linear=True, srgb8=True)
    last, last8, st = r.readback(linear=True, srgb8=True)  # the handle still holds the finished frame
    r.close()
    assert [f[2] for f in frames] == [4, 8, 10] and all(f[3] == spp for f in frames)
    plain, plain8, sp, _, _ = render(sc, cam)
    assert st.rays == sp.rays
    assert np.allclose(frames[-1][0], plain, rtol=3e-6, atol=1e-7)
    assert np.allclose(last, plain, rtol=3e-6, atol=1e-7)
    assert np.abs(frames[-1][1].astype(int) - plain8.astype(int)).max() <= 1
    # the first frame is the mean of samples [0, 4)
    r4 = Renderer(sc.desc)
    r4.render(sc.camera(W, H, 4, 50), 0, 4)
    first, _, _ = r4.readback()
    r4.close()
    assert np.allclose(frames[0][0], first, rtol=3e-6, atol=1e-7)


def test_resolve_handles_odd_sizes_and_unaligned_accumulators():
    """The vectorised reduce+resolve kernel on a frame whose float count is not a multiple of four and whose rows are
    not either, and on a caller-owned accumulator that is only 4-byte aligned (scalar instantiation)."""
    import torch
    sc = BuiltinScene(4)
    W, H, spp = 37, 21, 3
    cam = sc.camera(W, H, spp, 50)
    a, a8, _, _, _ = render(sc, cam)
    buf = torch.zeros(W * H * 3 + 1, dtype=torch.float32, device="cuda")
    r = Renderer(sc.desc)
    r.render(cam, accum_ptr=buf.data_ptr() + 4, stream=torch.cuda.current_stream().cuda_stream)
    b, b8, _ = r.readback(linear=True, srgb8=True, accum_ptr=buf.data_ptr() + 4)
    r.close()
    assert np.array_equal(a, b) and np.array_equal(a8, b8)
    g = np.sqrt(a[::-1])
    want = (np.float32(256.0) * np.clip(g, 0.0, np.float32(0.999))).astype(np.int32)
    assert np.abs(want - a8.astype(np.int32)).max() <= 1


def test_cli_one_and_two_gpus_write_the_same_picture(built, tmp_path):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cli = os.path.join(ROOT, "raytracinginoneweekendincuda_b200", "rt_cli")
    outs = []
    for g in (1, 2):
        out = tmp_path / f"g{g}.ppm"
        res = subprocess.run([cli, "--scene", "10", "--width", "200", "--height", "112", "--spp", "8", "--gpus", str(g),
                              "--p6", "--out", str(out)], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr[-2000:]
        outs.append(np.frombuffer(out.read_bytes()[len(b"P6\n200 112\n255\n"):], np.uint8).astype(int))
    d = np.abs(outs[0] - outs[1])
    assert d.max() <= 1 and (d != 0).mean() < 1e-3  # fp32 summation order may move a byte across a quantisation step


def test_cli_progressive_and_jpeg_texture(built, tmp_path):
    cli = os.path.join(ROOT, "raytracinginoneweekendincuda_b200", "rt_cli")
    jpg = os.path.join(ROOT, "oracle", "_ref", "earthmap.jpg")
    if not os.path.exists(jpg):
        pytest.skip("earthmap.jpg not available")
    out = tmp_path / "earth.ppm"
    res = subprocess.run([cli, "--scene", "2", "--width", "96", "--height", "54", "--spp", "6", "--earth", jpg,
                          "--progressive", "4", "--p6", "--out", str(out)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "4 / 6 samples" in res.stderr and "6 / 6 samples" in res.stderr
    img = np.frombuffer(out.read_bytes()[len(b"P6\n96 54\n255\n"):], np.uint8).reshape(54, 96, 3)
    centre = img[20:34, 40:56].reshape(-1, 3).astype(int)
    assert not (np.abs(centre - np.array([0, 255, 255])) < 8).all(axis=1).any(), "the globe rendered cyan: texture not loaded"
    # the same scene through the Python API with the texels decoded by the library
    from raytracinginoneweekendincuda_b200 import load_image
    sc = BuiltinScene(2, load_image(jpg))
    r = Renderer(sc.desc)
    r.render(sc.camera(96, 54, 6, 50))
    _, s8, _ = r.readback(linear=False, srgb8=True)
    r.close()
    assert np.abs(s8.astype(int) - img.astype(int)).max() <= 1
