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


def test_light_densities_are_normalised_and_match_their_samplers(oracle):
    """Book 3's quad / sphere pdf_value and random (oracle LightPdf / LightDirection): (1) the density of a light,
    integrated over all directions from a point, is 1: uniform directions d, mean(pdf(d)) * 4 pi ~ 1; (2) every sampled
    direction hits the light it was drawn towards (pdf > 0); (3) importance weights of the sampler against its own
    density average to the light's solid angle measure: E[1 / pdf] over sampled directions = solid angle, and the two
    estimates of the solid angle agree.  Scene 5 has one sphere light and one quad light, scene 7 one quad."""
    rng = np.random.default_rng(7)
    n = 400000
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for sid, origin in ((7, (278.0, 100.0, 278.0)), (5, (4.0, 0.5, 3.0))):
        sc = BuiltinScene(sid)
        o = np.array(origin, np.float64)
        pdf = np.zeros(n)
        n_lights = oracle.oracle_light_pdf(sc.desc, o.ctypes.data, np.ascontiguousarray(v).ctypes.data, n, pdf.ctypes.data)
        assert n_lights == (1 if sid == 7 else 2)
        hit = pdf > 0
        se = (pdf * 4 * np.pi).std() / np.sqrt(n)
        assert abs(pdf.mean() * 4 * np.pi - 1.0) < 5 * se + 0.01, (sid, pdf.mean() * 4 * np.pi, se)
        omega_uniform = hit.mean() * 4 * np.pi  # solid angle of the union of the lights (they do not overlap here)
        omega_sampled = 0.0
        for light in range(n_lights):
            m = 50000
            r12 = rng.random((m, 2))
            dirs = np.zeros((m, 3))
            assert oracle.oracle_light_direction(sc.desc, light, o.ctypes.data, r12.ctypes.data, m, dirs.ctypes.data) == n_lights
            p = np.zeros(m)
            oracle.oracle_light_pdf(sc.desc, o.ctypes.data, dirs.ctypes.data, m, p.ctypes.data)
            assert (p > 0).mean() > 0.999, (sid, light)  # (a sample on the rim may miss by rounding)
            # p is the MEAN over the lights; the density of this light's sampler is n_lights * p where only it is hit
            omega_sampled += float(np.mean(1.0 / (n_lights * p[p > 0])))
        assert abs(omega_sampled - omega_uniform) < 0.05 * omega_uniform, (sid, omega_sampled, omega_uniform)


def test_the_reference_lambertian_lobe_is_two_cos_cubed_over_pi():
    """The density ScatterImportance weighs with: the direction of N + (uniform point in the unit ball)
    (Material.h:14-24,68-86) has density 2 cos^3(theta) / pi -- checked against a histogram of cos(theta)."""
    rng = np.random.default_rng(11)
    pts = rng.uniform(-1, 1, size=(600000, 3))
    pts = pts[(pts ** 2).sum(1) < 1.0]
    d = pts + np.array([0.0, 0.0, 1.0])
    c = d[:, 2] / np.linalg.norm(d, axis=1)
    hist, edges = np.histogram(c, bins=20, range=(0.0, 1.0), density=True)
    mid = 0.5 * (edges[1:] + edges[:-1])
    # density in cos(theta): p(omega) * 2 pi = 4 cos^3
    assert np.allclose(hist, 4.0 * mid ** 3, rtol=0.06, atol=0.02)


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
