"""Importance sampling of the lights (RT_FLAG_IMPORTANCE; SURVEY 8 row f4: phase 4 of the reference's roadmap,
/root/reference README.md:37-42 -- "importance sampling / PDFs / ONB" -- which the reference does not implement).

There is no reference render to compare with, so the anchors are:
  * the estimator is unbiased for the image the reference's own scattering produces: oracle with and without the
    mode converge to the same mean, the mode with less noise (CPU, this file);
  * the CUDA path reproduces the oracle's restatement (oracle/rt_oracle.cpp ScatterImportance: book 3's
    quad / sphere pdf_value + random, onb, mixture density) sample for sample (GPU, exact stream);
  * the CUDA path with the flag converges to the CUDA path without it (GPU).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import A, oracle_render
from raytracinginoneweekendincuda_b200 import BuiltinScene, Renderer

IMPORTANCE = 2  # oracle mode bit (oracle_render's `bvh` argument: bit 0 BVH, bit 1 importance sampling)


def scene_for(sid, earth):
    return BuiltinScene(sid, earth if sid in (2, 9) else None)


def test_light_tables(lib, earth):
    """The sampling targets: quads and spheres with a DiffuseLight material (kernel.cu:330-333 simple light: a sphere
    and a quad; :351,:370,:414,:457 the ceiling lights of the Cornell boxes and of the Book 2 final scene)."""
    for sid, n in ((10, 0), (0, 0), (5, 2), (6, 1), (7, 1), (8, 1), (9, 1)):
        sc = scene_for(sid, earth)
        i = A.rt_pack_info()
        o = A.rt_upload_options()
        assert sc.lib.rt_scene_pack_info(sc.desc, C.byref(o), C.byref(i)) == 0
        assert i.n_lights == n, (sid, i.n_lights)


@pytest.mark.parametrize("sid,W,H", [(7, 32, 32), (5, 48, 27)])
def test_oracle_importance_is_unbiased_and_less_noisy(oracle, sid, W, H):
    """Same mean image as the reference's scattering, lower variance: 4 batches of 128 spp each way."""
    sc = BuiltinScene(sid)
    batches, spp = 4, 128
    cam = sc.camera(W, H, batches * spp, 50)
    ref = np.array([oracle_render(oracle, sc, cam, k * spp, (k + 1) * spp, bvh=1)[0] / spp for k in range(batches)])
    imp = np.array([oracle_render(oracle, sc, cam, k * spp, (k + 1) * spp, bvh=1 | IMPORTANCE)[0] / spp
                    for k in range(batches)])
    noise_ref, noise_imp = ref.std(0).mean(), imp.std(0).mean()
    assert noise_imp < 0.6 * noise_ref, (noise_ref, noise_imp)
    # frame means agree within 4 standard errors of the noisier estimator
    se = ref.mean(axis=(1, 2, 3)).std() / np.sqrt(batches) + 1e-4
    assert abs(ref.mean() - imp.mean()) < 4 * se + 0.01 * ref.mean(), (ref.mean(), imp.mean(), se)


@pytest.mark.gpu
@pytest.mark.parametrize("sid,W,H,spp", [(5, 120, 68, 8), (6, 96, 96, 8), (7, 96, 96, 8), (8, 96, 96, 8), (9, 160, 90, 4)])
def test_gpu_importance_exact_stream_parity(oracle, earth, sid, W, H, spp):
    """Identical streams: >= 99.9 % of the pixels within 1e-3 relative of the FP64 restatement (the bar of the
    reference-scattering path, tests/test_parity_gpu.py)."""
    sc = scene_for(sid, earth)
    cam = sc.camera(W, H, spp, 50)
    want, ost = oracle_render(oracle, sc, cam, 0, spp, bvh=1 | IMPORTANCE)
    r = Renderer(sc.desc)
    r.render(cam, 0, spp, flags=A.RT_FLAG_IMPORTANCE)
    got, _, st = r.readback()
    info = r.info()
    r.close()
    assert info.variant == A.RT_VARIANT_HITQUEUE
    ref = want / spp
    ok = (np.abs(got.astype(np.float64) - ref) <= 1e-3 * np.abs(ref) + 1e-6).all(axis=2)
    assert ok.mean() >= 0.999, f"scene {sid}: {ok.mean() * 100:.3f}% of pixels within 1e-3"
    assert abs(int(st.rays) - int(ost.rays)) <= 2e-3 * ost.rays


@pytest.mark.gpu
@pytest.mark.parametrize("sid,noise_ratio", [(7, 0.6), (8, 0.8)])  # (the smoke scene's noise is mostly its media's)
def test_gpu_importance_converges_to_the_plain_render_with_less_noise(sid, noise_ratio):
    sc = BuiltinScene(sid)
    W = H = 64
    batches, spp = 8, 256
    cam = sc.camera(W, H, batches * spp, 50)

    def batch_means(flags):
        out = []
        for k in range(batches):
            r = Renderer(sc.desc)
            r.render(cam, k * spp, (k + 1) * spp, flags=flags)
            lin, _, _ = r.readback()
            r.close()
            out.append(lin.astype(np.float64) * batches)  # readback divides by samples_per_pixel = batches * spp
        return np.array(out)

    plain, imp = batch_means(0), batch_means(A.RT_FLAG_IMPORTANCE)
    assert imp.std(0).mean() < noise_ratio * plain.std(0).mean()
    se = plain.mean(axis=(1, 2, 3)).std() / np.sqrt(batches) + 1e-4
    assert abs(plain.mean() - imp.mean()) < 4 * se + 0.005 * plain.mean(), (plain.mean(), imp.mean(), se)


@pytest.mark.gpu
def test_gpu_importance_needs_the_hit_queue_kernel():
    sc = BuiltinScene(7)
    cam = sc.camera(32, 32, 2, 50)
    r = Renderer(sc.desc)
    with pytest.raises(Exception):
        r.render(cam, 0, 2, flags=A.RT_FLAG_IMPORTANCE, variant=A.RT_VARIANT_MEGAKERNEL)
    r.render(cam, 0, 2, flags=A.RT_FLAG_IMPORTANCE)  # AUTO picks the hit-queue kernel
    r.close()
