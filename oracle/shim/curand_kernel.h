// oracle/shim/curand_kernel.h -- the reference's `curandState*` plumbing mapped
// onto (a) a host XORWOW for the scene stream and (b) the render path's
// counter-based stream.  TEST INFRASTRUCTURE ONLY (oracle/build_ref.py).
//
// (a) curand_init(seed, 0, 0, &s) puts `s` in sequential XORWOW mode -- cuRAND's
//     generator restated from the CUDA 12.9 toolkit header (curand_kernel.h:
//     800-825 init, 863-874 step; curand_uniform.h:69-72 bits->float).  That is
//     the stream CreateWorld consumes (reference kernel.cu:105,184).
// (b) rtshim_key(&s, seed, pixel, sample) + rtshim_slot(&s, slot) put `s` in
//     keyed mode: curand_uniform(&s) returns dim 0,1,2,.. of the stream
//     (seed, pixel, sample, slot, domain 0) of include/rt_rng.h, restated here.
//     rtshim_medium_uniform(&s, id) returns the keyed draw of medium `id` for
//     its next visit in this slot (SURVEY.md traps T2/T3).
#pragma once
#include <cstdint>

struct curandState {
    int keyed = 0;
    // XORWOW
    uint32_t v[5] = {0, 0, 0, 0, 0}, d = 0;
    // keyed stream
    uint32_t seed = 0, pixel = 0, sample = 0, slot = 0, dim = 0;
    uint32_t out[4] = {0, 0, 0, 0};
    uint32_t visit[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long draws = 0;
};

inline void curand_init(unsigned long long seed, unsigned long long subsequence, unsigned long long offset,
                        curandState* s)
{
    (void)subsequence; // only subsequence 0 / offset 0 (no skip-ahead) is supported
    (void)offset;
    s->keyed = 0;
    const uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u;
    const uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    s->d = 6615241u + t1 + t0;
    s->v[0] = 123456789u + t0;
    s->v[1] = 362436069u ^ t0;
    s->v[2] = 521288629u + t1;
    s->v[3] = 88675123u ^ t1;
    s->v[4] = 5783321u + t0;
}

inline void rtshim_pcg4d(uint32_t v[4])
{
    for (int k = 0; k < 4; ++k) v[k] = v[k] * 1664525u + 1013904223u;
    v[0] += v[1] * v[3];
    v[1] += v[2] * v[0];
    v[2] += v[0] * v[1];
    v[3] += v[1] * v[2];
    for (int k = 0; k < 4; ++k) v[k] ^= v[k] >> 16;
    v[0] += v[1] * v[3];
    v[1] += v[2] * v[0];
    v[2] += v[0] * v[1];
    v[3] += v[1] * v[2];
}

inline float rtshim_to_uniform(uint32_t bits) { return (float)bits * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }

inline void rtshim_key(curandState* s, uint32_t seed, uint32_t pixel, uint32_t sample)
{
    s->keyed = 1;
    s->seed = seed;
    s->pixel = pixel;
    s->sample = sample;
}

inline void rtshim_slot(curandState* s, uint32_t slot)
{
    s->slot = slot;
    s->dim = 0;
    for (int k = 0; k < 8; ++k) s->visit[k] = 0;
}

inline float curand_uniform(curandState* s)
{
    ++s->draws;
    if (!s->keyed) {
        const uint32_t t = s->v[0] ^ (s->v[0] >> 2);
        s->v[0] = s->v[1];
        s->v[1] = s->v[2];
        s->v[2] = s->v[3];
        s->v[3] = s->v[4];
        s->v[4] = (s->v[4] ^ (s->v[4] << 4)) ^ (t ^ (t << 1));
        s->d += 362437u;
        return rtshim_to_uniform(s->v[4] + s->d);
    }
    const uint32_t lane = s->dim & 3u;
    if (lane == 0) {
        s->out[0] = s->pixel;
        s->out[1] = s->sample;
        s->out[2] = (s->slot & 0xffu) | (((s->dim >> 2) & 0xffu) << 8);
        s->out[3] = s->seed;
        rtshim_pcg4d(s->out);
    }
    ++s->dim;
    return rtshim_to_uniform(s->out[lane]);
}

inline float rtshim_medium_uniform(curandState* s, int mediumId)
{
    ++s->draws;
    const uint32_t visit = s->visit[mediumId & 7]++;
    uint32_t v[4] = {s->pixel, s->sample, (s->slot & 0xffu) | ((1u + 2u * (uint32_t)mediumId + visit) << 16), s->seed};
    rtshim_pcg4d(v);
    return rtshim_to_uniform(v[0]);
}

int rtshim_next_medium_id();
