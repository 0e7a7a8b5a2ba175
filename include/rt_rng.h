/*
 * rt_rng.h -- the render path's counter-based random stream.
 *
 * The reference keeps one cuRAND XORWOW state per pixel (48 B/pixel, 398 MB at
 * 4K) initialised by RenderInit (kernel.cu:110-119) and consumes it
 * sequentially (kernel.cu:140-143, Camera.h:15,80, Material.h:19-21,
 * Dielectric.h:41, ConstantMedium.h:79).  Here every uniform is a pure
 * function of
 *
 *     (seed, pixel, sample, slot, domain, dim)
 *
 * so there is no init kernel, no per-pixel state, and any GPU can render any
 * sample range and still reproduce the 1-GPU sample set.
 *
 *   pixel  = j*W + i, j = 0 is the bottom row (kernel.cu:131)
 *   sample = global sample index (kernel.cu:138)
 *   slot   = 0 for the camera draws of the sample (jitter u, jitter v, lens
 *            disk, shutter time -- kernel.cu:140-141, Camera.h:78-80),
 *            bounce+1 for everything drawn inside RayColor iteration `bounce`
 *   domain = 0 for the sequential draws of the slot (scatter), and
 *            1 + 2*medium_id + visit for the single draw a ConstantMedium makes
 *            inside BVH traversal (ConstantMedium.h:79) -- keyed so the image
 *            does not depend on traversal order (SURVEY.md trap T3)
 *   dim    = running index of the draw inside (slot, domain)
 *
 * Hash: PCG4D (Jarzynski & Olano, "Hash Functions for GPU Rendering", JCGT
 * 2020): one call turns a 4-word key into four 32-bit outputs, i.e. four
 * consecutive dims.  Bits -> float follows cuRAND's curand_uniform
 * (curand_uniform.h:69-72): x*2^-32 + 2^-33 in fp32, range (0,1] (trap T5).
 */
#ifndef RT_RNG_H
#define RT_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

struct rt_u4 {
    uint32_t x, y, z, w;
};

RT_HD rt_u4 rt_pcg4d(rt_u4 v)
{
    v.x = v.x * 1664525u + 1013904223u;
    v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u;
    v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w;
    v.y += v.z * v.x;
    v.z += v.x * v.y;
    v.w += v.y * v.z;
    v.x ^= v.x >> 16;
    v.y ^= v.y >> 16;
    v.z ^= v.z >> 16;
    v.w ^= v.w >> 16;
    v.x += v.y * v.w;
    v.y += v.z * v.x;
    v.z += v.x * v.y;
    v.w += v.y * v.z;
    return v;
}

/* Key word z: slot in bits 0-7, 4-dim block in bits 8-15, domain in bits 16-31. */
RT_HD uint32_t rt_rng_zword(uint32_t slot, uint32_t block, uint32_t domain)
{
    return (slot & 0xffu) | ((block & 0xffu) << 8) | (domain << 16);
}

RT_HD rt_u4 rt_rng_block(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot,
                         uint32_t domain, uint32_t block)
{
    rt_u4 k;
    k.x = pixel;
    k.y = sample;
    k.z = rt_rng_zword(slot, block, domain);
    k.w = seed;
    return rt_pcg4d(k);
}

/* 32 random bits -> float in (0,1], bit-for-bit cuRAND's _curand_uniform. */
RT_HD float rt_bits_to_u01(uint32_t bits)
{
    return (float)bits * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
}

RT_HD uint32_t rt_u4_lane(const rt_u4& v, uint32_t lane)
{
    return lane == 0 ? v.x : lane == 1 ? v.y : lane == 2 ? v.z : v.w;
}

/* Sequential view of one (pixel, sample, slot, domain) stream. */
struct rt_rng_stream {
    uint32_t seed, pixel, sample, slot, domain;
    uint32_t dim;
    rt_u4 cur;

    RT_HD void begin(uint32_t seed_, uint32_t pixel_, uint32_t sample_, uint32_t slot_,
                     uint32_t domain_)
    {
        seed = seed_;
        pixel = pixel_;
        sample = sample_;
        slot = slot_;
        domain = domain_;
        dim = 0;
    }
    RT_HD float next()
    {
        const uint32_t lane = dim & 3u;
        if (lane == 0) cur = rt_rng_block(seed, pixel, sample, slot, domain, dim >> 2);
        ++dim;
        return rt_bits_to_u01(rt_u4_lane(cur, lane));
    }
};

#endif /* RT_RNG_H */
