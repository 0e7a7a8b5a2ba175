// rt_device.cu -- kernels and the device half of the C ABI (include/rt_abi.h).
//
// Replaces the reference's RenderInit + Render launches and the device-side
// object graph they walk (reference kernel.cu:110-154, launched :681-689).
//
// Two kernels render the same image: RenderMega (default) and RenderWave (an
// on-chip wavefront, kept for the comparison DESIGN.md 5.3 reports).
//
// Megakernel design (sm_100a, 148 SMs):
//   * persistent CTAs, one per SM, 768 threads; each warp pulls 8x4-pixel tiles
//     from a global atomic counter, so long tiles do not stall a whole block
//     the way the reference's 8x8 blocks do;
//   * inside a tile every lane owns one pixel and runs a flat state machine
//     "regenerate if dead -> trace one ray -> shade": a lane whose path ends
//     starts its next sample at once instead of idling until the slowest path
//     of the warp finishes (the reference nests spp loop > bounce loop > BVH
//     loop, kernel.cu:138-144 / :71-95 / BvhNode.h:113-155);
//   * when nodes + primitives + materials fit, they are staged once per CTA in
//     shared memory (Book 1: ~62 KB) and traversal runs on LDS.128; otherwise
//     LDG.E.128 through the read-only path with the set resident in L1/L2;
//   * the traversal stack lives in shared memory, [level][thread];
//   * no per-pixel RNG state: every uniform is a hash of
//     (seed, pixel, sample, slot, domain, dim) -- include/rt_rng.h;
//   * radiance is summed in fp32 registers per pixel in sample order and added
//     to the fp32 accumulator once per tile (deterministic, no atomics).
// No CPU fallback: without a CUDA device every entry point returns an error.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/rt_abi.h"
#include "rt_pack.hpp"
#include "rt_trace.cuh"

void rt_set_error(const char* fmt, ...); // rt_error.cpp

#define RT_CUDA(call)                                                                               \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            rt_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return RT_ERR_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

namespace {

using namespace rtdev;

constexpr int kMaxStackLevels = 32;
constexpr int kTileW = 8, kTileH = 4;

struct RenderArgs {
    float* accum;              // W*H*3 fp32 sums, row 0 = bottom
    unsigned long long* stats; // [0] rays [1] paths [2] node tests [3] prim tests
    unsigned int* tileCounter;
    int sampleBegin, sampleEnd;
    uint32_t seed;
    int tilesX, tilesY;
    int waveSlots, waveIdleExit, waveLeafBatch, waveRefillMin; // wavefront variant tuning
    int megaLeafMask; // megakernel: leaves are tested when (step & mask) == 0
    int stackLevels;  // traversal stack entries per thread (BVH depth + 3, at most 32)
    // test hook (STATS instantiations only): per-bounce records of one (pixel, sample) path
    int debugPixel, debugSample;
    float* debugOut; // [max_depth][8]: hit id, t, material, front, p.x, p.y, p.z, 1
    // byte sizes of the staged arrays (SMEM variant)
    uint32_t nodesBytes, spheresBytes, sphereMatBytes, movingBytes, quadsBytes, mediaBytes, materialsBytes;
};

__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t Stage(uint32_t& cursor, char* smem, const void* src, uint32_t bytes)
{
    const uint32_t at = cursor;
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(smem + at);
    for (uint32_t k = threadIdx.x; k < bytes / 16u; k += blockDim.x) d[k] = __ldg(&s[k]);
    cursor += (bytes + 15u) & ~15u;
    return at;
}

// Stages the scene arrays in shared memory (SMEM) or points at them in global memory.
template <bool SMEM>
__device__ __forceinline__ SceneView<SMEM> SetupScene(const DevScene& scene, const RenderArgs& args, char* smem, uint32_t smemBase,
                                                      uint32_t& cursor)
{
    SceneView<SMEM> sv;
    if constexpr (SMEM) {
        sv.nodes.a = smemBase + Stage(cursor, smem, scene.nodes, args.nodesBytes);
        sv.spheres.a = smemBase + Stage(cursor, smem, scene.spheres, args.spheresBytes);
        sv.sphere_material.a = smemBase + Stage(cursor, smem, scene.sphere_material, args.sphereMatBytes);
        sv.moving.a = smemBase + Stage(cursor, smem, scene.moving, args.movingBytes);
        sv.quads.a = smemBase + Stage(cursor, smem, scene.quads, args.quadsBytes);
        sv.media.a = smemBase + Stage(cursor, smem, scene.media, args.mediaBytes);
        sv.materials.a = smemBase + Stage(cursor, smem, scene.materials, args.materialsBytes);
        __syncthreads();
    } else {
        sv.nodes.a = reinterpret_cast<const char*>(scene.nodes);
        sv.spheres.a = reinterpret_cast<const char*>(scene.spheres);
        sv.sphere_material.a = reinterpret_cast<const char*>(scene.sphere_material);
        sv.moving.a = reinterpret_cast<const char*>(scene.moving);
        sv.quads.a = reinterpret_cast<const char*>(scene.quads);
        sv.media.a = reinterpret_cast<const char*>(scene.media);
        sv.materials.a = reinterpret_cast<const char*>(scene.materials);
    }
    sv.textures = scene.textures;
    sv.perlins = scene.perlins;
    sv.images = scene.images;
    sv.root_ref = scene.root_ref;

    return sv;
}

// 768 threads per SM (24 warps, 80 registers): measured 18 % faster than 512 x 87
// registers -- the kernel stalls on fixed-latency dependencies ("wait"), which more
// resident warps hide (profiles/README.md).
// The feature-complete instantiations need ~125 registers and stay at 512.
constexpr int MegaMaxThreads(int feat) { return feat == 0 ? 768 : 512; }
// The head/tail kernel holds no path state across rounds: with moving spheres and checker textures it
// still fits 80 registers (768 threads); the feature-complete instantiation runs at 640.
constexpr int HtMaxThreads(int feat) { return (feat & ~(RT_FEAT_MOVING | RT_FEAT_TEXTURE)) == 0 ? 768 : 640; }

template <int FEAT, bool SMEM, bool STATS>
__global__ void __launch_bounds__(MegaMaxThreads(FEAT), 1) RenderMega(const DevScene scene, const DevCamera cam, const RenderArgs args)
{
    extern __shared__ __align__(16) char smem[];
    const uint32_t smemBase = SmemAddr(smem);
    uint32_t cursor = blockDim.x * 4u * (uint32_t)args.stackLevels;

    const SceneView<SMEM> sv = SetupScene<SMEM>(scene, args, smem, smemBase, cursor);

    Stack stack;
    stack.base = smemBase + threadIdx.x * 4u;
    stack.stride = blockDim.x * 4u;

    const int lane = threadIdx.x & 31;
    const int nTiles = args.tilesX * args.tilesY;
    const f3 background = make_f3(cam.background[0], cam.background[1], cam.background[2]);
    const uint32_t leafMask = (uint32_t)args.megaLeafMask;
    unsigned long long nRays = 0, nPaths = 0, nNode = 0, nPrim = 0;

    while (true) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(args.tileCounter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= nTiles) break;
        const int tx = tile % args.tilesX, ty = tile / args.tilesX;
        const int i = tx * kTileW + (lane & (kTileW - 1));
        const int j = ty * kTileH + (lane / kTileW);
        const bool valid = i < cam.width && j < cam.height;
        const uint32_t pixel = (uint32_t)(j * cam.width + i);

        f3 sum = make_f3(0.0f, 0.0f, 0.0f);
        f3 throughput = make_f3(1.0f, 1.0f, 1.0f);
        Ray ray;
        ray.o.x = ray.o.y = ray.o.z = 0.0;
        ray.d.x = ray.d.y = ray.d.z = 1.0;
        ray.time = 0.0f;
        RaySlab slab = MakeSlab(ray);
        double a = 3.0;
        Trav tv;
        tv.Idle();
        tv.tMedium = 0.0;
        int sample = args.sampleBegin;
        int bounce = 0;
        bool alive = false;
        bool done = !valid;

        // Lanes run a flat state machine.  Phase A (below): every lane shades the hit
        // of its finished walk -- which either continues the path, or ends it and
        // starts the pixel's next sample -- and leaves with a fresh ray.  Phase B:
        // every lane walks the tree to completion.  (Leaving phase B early, once a
        // number of lanes wait, was measured and is slower: profiles/README.md.)
        while (true) {
            if (tv.ref == RT_TRAV_DONE && !done) {
                if (alive) {
                    ++nRays;
                    if (tv.hit == RT_HIT_NONE) {
                        sum = sum + throughput * background; // kernel.cu:74-79
                        alive = false;
                    } else {
                        Hit h;
                        FinalizeHit<FEAT, SMEM>(sv, ray, a, tv.hit, tv.t, tv.tMedium, h);
                        const uint32_t type = RT_HIT_TYPE(tv.hit);
                        if (STATS && args.debugOut && (int)pixel == args.debugPixel && sample == args.debugSample) {
                            float* o = args.debugOut + bounce * 8;
                            o[0] = __uint_as_float(tv.hit);
                            o[1] = tv.t;
                            o[2] = __int_as_float(h.material);
                            o[3] = h.front ? 1.0f : 0.0f;
                            o[4] = (float)h.p.x;
                            o[5] = (float)h.p.y;
                            o[6] = (float)h.p.z;
                            o[7] = 1.0f;
                        }
                        const bool sphereLike = type == RT_LEAF_SPHERE || type == RT_LEAF_MOVING;
                        const StreamKey rng = MakeKey(args.seed, pixel, (uint32_t)sample, (uint32_t)bounce + 1u);
                        f3 atten, emitted;
                        d3 dir;
                        const bool scattered = Scatter<FEAT, SMEM>(sv, h, ray.d, a, sphereLike, rng, atten, dir, emitted);
                        sum = sum + throughput * emitted; // kernel.cu:82-83
                        if (!scattered) {
                            alive = false;
                        } else {
                            throughput = throughput * atten; // kernel.cu:93-94
                            ray.o = h.p;
                            ray.d = dir;
                            if (++bounce >= cam.max_depth) alive = false; // kernel.cu:71,97
                        }
                    }
                    if (!alive) ++sample;
                }
                if (!alive) {
                    if (sample >= args.sampleEnd) {
                        done = true;
                    } else {
                        const StreamKey rng = MakeKey(args.seed, pixel, (uint32_t)sample, 0u);
                        ray = CameraRay(cam, i, j, rng);
                        throughput = make_f3(1.0f, 1.0f, 1.0f);
                        bounce = 0;
                        alive = true;
                        if (STATS) ++nPaths;
                    }
                }
                if (!done) {
                    slab = MakeSlab(ray);
                    a = fma(ray.d.x, ray.d.x, fma(ray.d.y, ray.d.y, ray.d.z * ray.d.z));
                    tv.Begin(sv.root_ref, stack);
                }
            }
            if (__all_sync(0xffffffffu, done)) break;
            // A lane that reaches a leaf waits for the warp's next leaf turn (every
            // leafPeriod-th step): the FP64 primitive tests then run for all the lanes that
            // piled up instead of for one or two lanes in nearly every step.
            uint32_t step = 0;
            while (tv.ref != RT_TRAV_DONE) {
                uint32_t nodeTests = 0, primTests = 0;
                ++step;
                if (!(tv.ref & RT_REF_LEAF))
                    TraceBox<SMEM>(sv, slab, 0.001f, stack, tv, nodeTests);
                else if ((step & leafMask) == 0u)
                    TraceLeaf<FEAT, SMEM>(sv, ray, a, slab.rcpA, 0.001f, stack, tv, args.seed, pixel, (uint32_t)sample,
                                          (uint32_t)bounce + 1u, primTests);
                if (STATS) {
                    nNode += nodeTests;
                    nPrim += primTests;
                }
            }
        }
        if (valid) {
            float* px = args.accum + (size_t)pixel * 3u;
            px[0] += sum.x;
            px[1] += sum.y;
            px[2] += sum.z;
        }
    }

    // one atomic per warp for the counters
    for (int off = 16; off > 0; off >>= 1) {
        nRays += __shfl_down_sync(0xffffffffu, nRays, off);
        if (STATS) {
            nPaths += __shfl_down_sync(0xffffffffu, nPaths, off);
            nNode += __shfl_down_sync(0xffffffffu, nNode, off);
            nPrim += __shfl_down_sync(0xffffffffu, nPrim, off);
        }
    }
    if (lane == 0) {
        atomicAdd(&args.stats[0], nRays);
        if (STATS) {
            atomicAdd(&args.stats[1], nPaths);
            atomicAdd(&args.stats[2], nNode);
            atomicAdd(&args.stats[3], nPrim);
        }
    }
}

// ---------------------------------------------------------------------------
// Wavefront variant, on chip.  The classic wavefront path tracer (ray-gen /
// extend / shade kernels exchanging rays through queues in device memory) would
// move ~200 B per ray through HBM -- 4 TB/s at 20 Grays/s -- for a scene that
// fits in shared memory.  Here the queues live in shared memory and belong to a
// warp: each warp owns a tile of 8 x S/8 pixels = S path slots (S = 64..128; SoA, 99 B each) and
// three compacted slot lists built with ballot + popc prefix sums:
//   ready  paths that have a ray to extend
//   shade  paths whose ray hit a surface (or a medium)
//   gen    paths that ended (miss / absorbed / light / depth) and need the
//          pixel's next camera sample
// and alternates three phases, each run by 32 lanes taking 32 list entries:
//   EXTEND  lanes pull slots off `ready` the moment they fall idle (persistent
//           threads over the warp's own queue), so box tests run on a full warp
//           while `ready` lasts; a lane that reaches a leaf waits until enough
//           lanes hold one, then the FP64 primitive tests run together;
//   SHADE   FinalizeHit + Scatter for 32 hits at a time;
//   GEN     background / next sample / camera ray for 32 ended paths at a time.
// A lane is no longer tied to a pixel, only a slot is: samples of a pixel are
// still taken in order and summed in the slot, so the image is bit-identical to
// the megakernel's.
#define RT_HIT_GEN_FIRST 0xfffffffeu /* slot has not started its first sample */
#define RT_HIT_GEN_ENDED 0xfffffffdu /* path ended on a surface                */

// One warp's pool: S path slots, structure of arrays (field stride = S), 88 B per
// slot, then the three slot lists.
struct Pool {
    double* O;     // [3][S] ray origin
    double* D;     // [3][S] ray direction
    double* TMED;  // [S]    medium scatter distance
    float* THR;    // [3][S] throughput
    float* SUM;    // [3][S] radiance sum of the pixel
    float* TIME;   // [S]
    float* T;      // [S]    hit distance (fp32)
    uint32_t* SB;  // [S]    sample << 8 | bounce
    uint32_t* HIT; // [S]    hit id, or RT_HIT_NONE / RT_HIT_GEN_*
    uint8_t* ready;
    uint8_t* shade;
    uint8_t* gen;
    int S;
    __device__ __forceinline__ Pool(char* p, int slots) : S(slots)
    {
        O = reinterpret_cast<double*>(p);
        D = O + 3 * S;
        TMED = D + 3 * S;
        THR = reinterpret_cast<float*>(TMED + S);
        SUM = THR + 3 * S;
        TIME = SUM + 3 * S;
        T = TIME + S;
        SB = reinterpret_cast<uint32_t*>(T + S);
        HIT = SB + S;
        ready = reinterpret_cast<uint8_t*>(HIT + S);
        shade = ready + S;
        gen = shade + S;
    }
    __device__ __forceinline__ d3 LoadO(int s) const { return make_d3(O[s], O[S + s], O[2 * S + s]); }
    __device__ __forceinline__ d3 LoadD(int s) const { return make_d3(D[s], D[S + s], D[2 * S + s]); }
    __device__ __forceinline__ void StoreO(int s, const d3& v) const
    {
        O[s] = v.x;
        O[S + s] = v.y;
        O[2 * S + s] = v.z;
    }
    __device__ __forceinline__ void StoreD(int s, const d3& v) const
    {
        D[s] = v.x;
        D[S + s] = v.y;
        D[2 * S + s] = v.z;
    }
};
__host__ __device__ constexpr int PoolBytes(int slots) { return slots * (7 * 8 + 10 * 4 + 3) + 16 - (slots * 3) % 16; }

template <int FEAT, bool SMEM, bool STATS>
__global__ void __launch_bounds__(512, 1) RenderWave(const DevScene scene, const DevCamera cam, const RenderArgs args)
{
    extern __shared__ __align__(16) char smem[];
    const uint32_t smemBase = SmemAddr(smem);
    uint32_t cursor = blockDim.x * 4u * (uint32_t)args.stackLevels;
    const SceneView<SMEM> sv = SetupScene<SMEM>(scene, args, smem, smemBase, cursor);

    Stack stack;
    stack.base = smemBase + threadIdx.x * 4u;
    stack.stride = blockDim.x * 4u;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu, ltMask = (1u << lane) - 1u;
    const int S = args.waveSlots, tileH = S / 8;
    const Pool pool(smem + ((cursor + 15u) & ~15u) + (uint32_t)warp * (uint32_t)PoolBytes(S), S);
    const int nTiles = args.tilesX * args.tilesY;
    const f3 background = make_f3(cam.background[0], cam.background[1], cam.background[2]);
    const int idleExit = args.waveIdleExit, leafBatch = args.waveLeafBatch, refillMin = args.waveRefillMin;
    unsigned long long nRays = 0, nPaths = 0, nNode = 0, nPrim = 0;

    while (true) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(args.tileCounter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= nTiles) break;
        const int tx = tile % args.tilesX, ty = tile / args.tilesX;
        const int px0 = tx * 8, py0 = ty * tileH;

        int nReady = 0, nShade = 0, nGen = 0;
        for (int s = lane; s < S; s += 32) {
            const bool valid = px0 + (s & 7) < cam.width && py0 + (s >> 3) < cam.height && args.sampleBegin < args.sampleEnd;
            pool.SUM[s] = pool.SUM[S + s] = pool.SUM[2 * S + s] = 0.0f;
            pool.SB[s] = (uint32_t)args.sampleBegin << 8;
            pool.HIT[s] = RT_HIT_GEN_FIRST;
            const unsigned m = __ballot_sync(FULL, valid);
            if (valid) pool.gen[nGen + __popc(m & ltMask)] = (uint8_t)s;
            nGen += __popc(m);
        }
        int nLive = nGen;

        // EXTEND state of this lane (kept in registers across the other phases).
        // An idle lane has tv.ref == RT_TRAV_DONE and mySlot < 0.
        int mySlot = -1;
        uint32_t myPixel = 0, mySB = 0;
        Ray ray;
        ray.o = make_d3(0.0, 0.0, 0.0);
        ray.d = make_d3(1.0, 1.0, 1.0);
        ray.time = 0.0f;
        RaySlab slab = MakeSlab(ray);
        double a = 3.0;
        Trav tv;
        tv.Idle();
        tv.tMedium = 0.0;
        unsigned idleMask = FULL;

        while (nLive > 0) {
            // ------------------------------------------------------------ EXTEND
            while (true) {
                const int nIdle = __popc(idleMask);
                if (nReady > 0 && (nIdle >= refillMin || nIdle == 32)) {
                    __syncwarp();
                    const int take = min(nIdle, nReady);
                    const int rank = __popc(idleMask & ltMask);
                    if (mySlot < 0 && rank < take) {
                        const int s = pool.ready[nReady - 1 - rank];
                        mySlot = s;
                        ray.o = pool.LoadO(s);
                        ray.d = pool.LoadD(s);
                        ray.time = pool.TIME[s];
                        mySB = pool.SB[s];
                        myPixel = (uint32_t)((py0 + (s >> 3)) * cam.width + px0 + (s & 7));
                        slab = MakeSlab(ray);
                        a = fma(ray.d.x, ray.d.x, fma(ray.d.y, ray.d.y, ray.d.z * ray.d.z));
                        tv.Begin(sv.root_ref, stack);
                    }
                    nReady -= take;
                    idleMask = __ballot_sync(FULL, mySlot < 0);
                }
                if (idleMask == FULL) break;
                if (nReady == 0 && nShade + nGen > 0 && __popc(idleMask) >= idleExit) break;

                // box steps, until `leafBatch` lanes wait at a leaf (or are done) or none is left
                const unsigned flying = ~idleMask;
                uint32_t nodeTests = 0, primTests = 0;
                while (true) {
                    const bool atBox = (tv.ref & RT_REF_LEAF) == 0u;
                    const unsigned boxMask = __ballot_sync(FULL, atBox);
                    if (boxMask == 0u || __popc(flying & ~boxMask) >= leafBatch) break;
                    if (atBox) TraceBox<SMEM>(sv, slab, 0.001f, stack, tv, nodeTests);
                }
                // the leaves that piled up, together
                if ((tv.ref & RT_REF_LEAF) != 0u && tv.ref != RT_TRAV_DONE)
                    TraceLeaf<FEAT, SMEM>(sv, ray, a, slab.rcpA, 0.001f, stack, tv, args.seed, myPixel, mySB >> 8, (mySB & 0xffu) + 1u,
                                          primTests);
                if (STATS) {
                    nNode += nodeTests;
                    nPrim += primTests;
                }
                // retire finished walks: misses to `gen`, hits to `shade`
                const bool fin = mySlot >= 0 && tv.ref == RT_TRAV_DONE;
                const unsigned finMask = __ballot_sync(FULL, fin);
                if (finMask != 0u) {
                    const bool miss = fin && tv.hit == RT_HIT_NONE;
                    const unsigned missMask = __ballot_sync(FULL, miss), hitMask = finMask & ~missMask;
                    if (fin) {
                        ++nRays;
                        pool.HIT[mySlot] = tv.hit;
                        pool.T[mySlot] = tv.t;
                        if (FEAT & RT_FEAT_MEDIUM) pool.TMED[mySlot] = tv.tMedium;
                        if (miss)
                            pool.gen[nGen + __popc(missMask & ltMask)] = (uint8_t)mySlot;
                        else
                            pool.shade[nShade + __popc(hitMask & ltMask)] = (uint8_t)mySlot;
                        mySlot = -1;
                    }
                    nGen += __popc(missMask);
                    nShade += __popc(hitMask);
                    idleMask |= finMask;
                }
            }

            // ------------------------------------------------------------- SHADE
            // full chunks of 32; a partial chunk only when nothing else can make progress
            const bool starving = nShade < 32 && nGen < 32;
            while (nShade >= 32 || (starving && nShade > 0)) {
                __syncwarp();
                const int n = min(nShade, 32);
                const bool on = lane < n;
                bool toReady = false, toGen = false;
                int s = 0;
                if (on) {
                    s = pool.shade[nShade - n + lane];
                    Ray r;
                    r.o = pool.LoadO(s);
                    r.d = pool.LoadD(s);
                    r.time = pool.TIME[s];
                    const uint32_t sb = pool.SB[s], hit = pool.HIT[s];
                    const uint32_t sample = sb >> 8, bounce = sb & 0xffu;
                    const uint32_t pixel = (uint32_t)((py0 + (s >> 3)) * cam.width + px0 + (s & 7));
                    const double ra = fma(r.d.x, r.d.x, fma(r.d.y, r.d.y, r.d.z * r.d.z));
                    Hit h;
                    FinalizeHit<FEAT, SMEM>(sv, r, ra, hit, pool.T[s], (FEAT & RT_FEAT_MEDIUM) ? pool.TMED[s] : 0.0, h);
                    if (STATS && args.debugOut && (int)pixel == args.debugPixel && (int)sample == args.debugSample) {
                        float* o = args.debugOut + bounce * 8;
                        o[0] = __uint_as_float(hit);
                        o[1] = pool.T[s];
                        o[2] = __int_as_float(h.material);
                        o[3] = h.front ? 1.0f : 0.0f;
                        o[4] = (float)h.p.x;
                        o[5] = (float)h.p.y;
                        o[6] = (float)h.p.z;
                        o[7] = 1.0f;
                    }
                    const uint32_t type = RT_HIT_TYPE(hit);
                    const bool sphereLike = type == RT_LEAF_SPHERE || type == RT_LEAF_MOVING;
                    const StreamKey rng = MakeKey(args.seed, pixel, sample, bounce + 1u);
                    f3 atten, emitted;
                    d3 dir;
                    const bool scattered = Scatter<FEAT, SMEM>(sv, h, r.d, ra, sphereLike, rng, atten, dir, emitted);
                    const f3 thr = make_f3(pool.THR[s], pool.THR[S + s], pool.THR[2 * S + s]);
                    if (!scattered) { // kernel.cu:82-83 (emission is black unless the path ends on a light)
                        pool.SUM[s] += thr.x * emitted.x;
                        pool.SUM[S + s] += thr.y * emitted.y;
                        pool.SUM[2 * S + s] += thr.z * emitted.z;
                    }
                    if (scattered && (int)bounce + 1 < cam.max_depth) { // kernel.cu:93-94, :71
                        pool.THR[s] = thr.x * atten.x;
                        pool.THR[S + s] = thr.y * atten.y;
                        pool.THR[2 * S + s] = thr.z * atten.z;
                        pool.StoreO(s, h.p);
                        pool.StoreD(s, dir);
                        pool.SB[s] = sb + 1u;
                        toReady = true;
                    } else {
                        pool.HIT[s] = RT_HIT_GEN_ENDED;
                        toGen = true;
                    }
                }
                nShade -= n;
                const unsigned rm = __ballot_sync(FULL, toReady), gm = __ballot_sync(FULL, toGen);
                if (toReady) pool.ready[nReady + __popc(rm & ltMask)] = (uint8_t)s;
                if (toGen) pool.gen[nGen + __popc(gm & ltMask)] = (uint8_t)s;
                nReady += __popc(rm);
                nGen += __popc(gm);
            }

            // --------------------------------------------------------------- GEN
            while (nGen >= 32 || (starving && nGen > 0)) {
                __syncwarp();
                const int n = min(nGen, 32);
                const bool on = lane < n;
                bool toReady = false, finished = false;
                int s = 0;
                if (on) {
                    s = pool.gen[nGen - n + lane];
                    const uint32_t hit = pool.HIT[s];
                    uint32_t sample = pool.SB[s] >> 8;
                    if (hit == RT_HIT_NONE) { // kernel.cu:74-79
                        pool.SUM[s] += pool.THR[s] * background.x;
                        pool.SUM[S + s] += pool.THR[S + s] * background.y;
                        pool.SUM[2 * S + s] += pool.THR[2 * S + s] * background.z;
                    }
                    if (hit != RT_HIT_GEN_FIRST) ++sample;
                    if ((int)sample >= args.sampleEnd) {
                        finished = true;
                    } else {
                        const int i = px0 + (s & 7), j = py0 + (s >> 3);
                        const StreamKey rng = MakeKey(args.seed, (uint32_t)(j * cam.width + i), sample, 0u);
                        const Ray r = CameraRay(cam, i, j, rng);
                        pool.StoreO(s, r.o);
                        pool.StoreD(s, r.d);
                        pool.TIME[s] = r.time;
                        pool.THR[s] = pool.THR[S + s] = pool.THR[2 * S + s] = 1.0f;
                        pool.SB[s] = sample << 8;
                        toReady = true;
                        if (STATS) ++nPaths;
                    }
                }
                nGen -= n;
                const unsigned rm = __ballot_sync(FULL, toReady), fm = __ballot_sync(FULL, finished);
                if (toReady) pool.ready[nReady + __popc(rm & ltMask)] = (uint8_t)s;
                nReady += __popc(rm);
                nLive -= __popc(fm);
            }
        }

        __syncwarp();
        for (int s = lane; s < S; s += 32) {
            const int i = px0 + (s & 7), j = py0 + (s >> 3);
            if (i < cam.width && j < cam.height) {
                float* px = args.accum + ((size_t)j * cam.width + i) * 3u;
                px[0] += pool.SUM[s];
                px[1] += pool.SUM[S + s];
                px[2] += pool.SUM[2 * S + s];
            }
        }
        __syncwarp();
    }

    for (int off = 16; off > 0; off >>= 1) {
        nRays += __shfl_down_sync(FULL, nRays, off);
        if (STATS) {
            nPaths += __shfl_down_sync(FULL, nPaths, off);
            nNode += __shfl_down_sync(FULL, nNode, off);
            nPrim += __shfl_down_sync(FULL, nPrim, off);
        }
    }
    if (lane == 0) {
        atomicAdd(&args.stats[0], nRays);
        if (STATS) {
            atomicAdd(&args.stats[1], nPaths);
            atomicAdd(&args.stats[2], nNode);
            atomicAdd(&args.stats[3], nPrim);
        }
    }
}

// ---------------------------------------------------------------------------
// Head/tail variant.  Measured on the megakernel (profiles/README.md): when every
// lane of a warp is at the same bounce -- max_depth 1 or 2 -- it runs 25 / 22
// Grays/s instead of 13.5, because camera rays are generated on full warps, the
// coherent primary rays of 32 neighbouring pixels finish their walks together,
// and every lane has something to shade.  This kernel keeps that synchrony
// without idling lanes whose path is over:
//   HEAD  all 32 lanes start the SAME sample of their pixels: camera ray, walk,
//         shade.  Paths that go on are not continued by their lane: their state
//         (ray, throughput, owner, sample, bounce: 64 B) is pushed on the warp's
//         queue in shared memory (ballot + popc compaction).
//   TAIL  whenever the queue holds 32 continuations, one round takes 32 of them:
//         one ray each, walk, shade, push back what survives.  Every round runs
//         on a full warp whatever the path lengths are.
// Radiance goes to per-pixel sums in shared memory (a tail adds to its owner's
// sum; two tails of one owner in a round are serialised in lane order, so the
// result is deterministic).  The order in which a pixel's paths are summed is no
// longer the sample order, so images equal the megakernel's up to fp32
// summation order, not bit for bit.
constexpr int kHtQueue = 64; // entries per warp
__host__ __device__ constexpr int HtWarpBytes(int feat)
{
    return kHtQueue * (6 * 8 + 3 * 4 + 4 + ((feat & RT_FEAT_MOVING) ? 4 : 0)) + 32 * 3 * 4;
}
#define RT_HT_MAX_SAMPLES (1 << 19) /* sample index relative to sample_begin is packed in 19 bits */

template <int FEAT, bool SMEM, bool STATS>
__global__ void __launch_bounds__(HtMaxThreads(FEAT), 1) RenderHeadTail(const DevScene scene, const DevCamera cam, const RenderArgs args)
{
    extern __shared__ __align__(16) char smem[];
    const uint32_t smemBase = SmemAddr(smem);
    uint32_t cursor = blockDim.x * 4u * (uint32_t)args.stackLevels;
    const SceneView<SMEM> sv = SetupScene<SMEM>(scene, args, smem, smemBase, cursor);

    Stack stack;
    stack.base = smemBase + threadIdx.x * 4u;
    stack.stride = blockDim.x * 4u;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu, ltMask = (1u << lane) - 1u;
    char* wbase = smem + ((cursor + 15u) & ~15u) + (uint32_t)warp * (uint32_t)HtWarpBytes(FEAT);
    double* QO = reinterpret_cast<double*>(wbase);            // [3][64]
    double* QD = QO + 3 * kHtQueue;                           // [3][64]
    float* QTHR = reinterpret_cast<float*>(QD + 3 * kHtQueue); // [3][64]
    uint32_t* QMETA = reinterpret_cast<uint32_t*>(QTHR + 3 * kHtQueue); // owner | bounce << 5 | (sample - begin) << 13
    float* QTIME = reinterpret_cast<float*>(QMETA + kHtQueue);          // [64], FEAT_MOVING only
    float* SUM = QTIME + ((FEAT & RT_FEAT_MOVING) ? kHtQueue : 0);      // [3][32]

    const int nTiles = args.tilesX * args.tilesY;
    const f3 background = make_f3(cam.background[0], cam.background[1], cam.background[2]);
    const uint32_t leafMask = (uint32_t)args.megaLeafMask;
    unsigned long long nRays = 0, nPaths = 0, nNode = 0, nPrim = 0;

    while (true) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(args.tileCounter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= nTiles) break;
        const int px0 = (tile % args.tilesX) * kTileW, py0 = (tile / args.tilesX) * kTileH;
        const bool valid = px0 + (lane & (kTileW - 1)) < cam.width && py0 + lane / kTileW < cam.height;
        SUM[lane] = SUM[32 + lane] = SUM[64 + lane] = 0.0f;
        __syncwarp();
        int nQ = 0;
        int headSample = args.sampleBegin;

        while (headSample < args.sampleEnd || nQ > 0) {
            // A round is a TAIL round when the queue could not take the survivors of another head
            // (or when there are no heads left); else a HEAD round.
            const bool tailRound = nQ > kHtQueue - 32 || headSample >= args.sampleEnd;
            Ray ray;
            f3 thr;
            uint32_t owner = (uint32_t)lane, sample = 0, bounce = 0;
            bool active;
            if (tailRound) {
                const int n = min(nQ, 32);
                active = lane < n;
                if (active) {
                    const int e = nQ - n + lane;
                    ray.o = make_d3(QO[e], QO[kHtQueue + e], QO[2 * kHtQueue + e]);
                    ray.d = make_d3(QD[e], QD[kHtQueue + e], QD[2 * kHtQueue + e]);
                    ray.time = (FEAT & RT_FEAT_MOVING) ? QTIME[e] : 0.0f;
                    thr = make_f3(QTHR[e], QTHR[kHtQueue + e], QTHR[2 * kHtQueue + e]);
                    const uint32_t meta = QMETA[e];
                    owner = meta & 31u;
                    bounce = (meta >> 5) & 0xffu;
                    sample = (uint32_t)args.sampleBegin + (meta >> 13);
                }
                nQ -= n;
                __syncwarp(); // all entries are read before any survivor is written back
            } else {
                active = valid;
                sample = (uint32_t)headSample;
                ++headSample;
            }
            const int oi = px0 + (int)(owner & (kTileW - 1)), oj = py0 + (int)(owner / kTileW);
            const uint32_t pixel = (uint32_t)(oj * cam.width + oi);
            if (!tailRound && active) {
                const StreamKey rng = MakeKey(args.seed, pixel, sample, 0u);
                ray = CameraRay(cam, oi, oj, rng);
                thr = make_f3(1.0f, 1.0f, 1.0f);
                if (STATS) ++nPaths;
            }

            // one ray per lane, walked to completion
            Trav tv;
            tv.Idle();
            tv.tMedium = 0.0;
            RaySlab slab;
            double a = 1.0;
            if (active) {
                slab = MakeSlab(ray);
                a = fma(ray.d.x, ray.d.x, fma(ray.d.y, ray.d.y, ray.d.z * ray.d.z));
                tv.Begin(sv.root_ref, stack);
                ++nRays;
            }
            uint32_t step = 0;
            while (tv.ref != RT_TRAV_DONE) {
                uint32_t nodeTests = 0, primTests = 0;
                ++step;
                if (!(tv.ref & RT_REF_LEAF))
                    TraceBox<SMEM>(sv, slab, 0.001f, stack, tv, nodeTests);
                else if ((step & leafMask) == 0u)
                    TraceLeaf<FEAT, SMEM>(sv, ray, a, slab.rcpA, 0.001f, stack, tv, args.seed, pixel, sample, bounce + 1u,
                                          primTests);
                if (STATS) {
                    nNode += nodeTests;
                    nPrim += primTests;
                }
            }

            // shade
            f3 add = make_f3(0.0f, 0.0f, 0.0f);
            bool hasAdd = false, survives = false;
            d3 newO = ray.o, newD = ray.d;
            if (active) {
                if (tv.hit == RT_HIT_NONE) {
                    add = thr * background; // kernel.cu:74-79
                    hasAdd = true;
                } else {
                    Hit h;
                    FinalizeHit<FEAT, SMEM>(sv, ray, a, tv.hit, tv.t, tv.tMedium, h);
                    const uint32_t type = RT_HIT_TYPE(tv.hit);
                    if (STATS && args.debugOut && (int)pixel == args.debugPixel && (int)sample == args.debugSample) {
                        float* o = args.debugOut + bounce * 8;
                        o[0] = __uint_as_float(tv.hit);
                        o[1] = tv.t;
                        o[2] = __int_as_float(h.material);
                        o[3] = h.front ? 1.0f : 0.0f;
                        o[4] = (float)h.p.x;
                        o[5] = (float)h.p.y;
                        o[6] = (float)h.p.z;
                        o[7] = 1.0f;
                    }
                    const bool sphereLike = type == RT_LEAF_SPHERE || type == RT_LEAF_MOVING;
                    const StreamKey rng = MakeKey(args.seed, pixel, sample, bounce + 1u);
                    f3 atten, emitted;
                    d3 dir;
                    const bool scattered = Scatter<FEAT, SMEM>(sv, h, ray.d, a, sphereLike, rng, atten, dir, emitted);
                    if (!scattered) { // kernel.cu:82-83: emission is black unless the path ends on a light
                        add = thr * emitted;
                        hasAdd = emitted.x != 0.0f || emitted.y != 0.0f || emitted.z != 0.0f;
                    } else if ((int)bounce + 1 < cam.max_depth) { // kernel.cu:93-94, :71
                        thr = thr * atten;
                        newO = h.p;
                        newD = dir;
                        survives = true;
                    }
                }
            }

            // radiance to the owner's sum; two contributions to one owner are applied in lane order
            {
                const unsigned am = __ballot_sync(FULL, hasAdd);
                int rank = 0;
                if (tailRound && hasAdd) rank = __popc(__match_any_sync(am, owner) & ltMask); // heads: owner == lane
                for (int r = 0;; ++r) {
                    if (hasAdd && rank == r) {
                        SUM[owner] += add.x;
                        SUM[32 + owner] += add.y;
                        SUM[64 + owner] += add.z;
                    }
                    __syncwarp();
                    if (__ballot_sync(FULL, hasAdd && rank > r) == 0u) break;
                }
            }
            // survivors back on the queue
            {
                const unsigned sm_ = __ballot_sync(FULL, survives);
                if (survives) {
                    const int e = nQ + __popc(sm_ & ltMask);
                    QO[e] = newO.x;
                    QO[kHtQueue + e] = newO.y;
                    QO[2 * kHtQueue + e] = newO.z;
                    QD[e] = newD.x;
                    QD[kHtQueue + e] = newD.y;
                    QD[2 * kHtQueue + e] = newD.z;
                    QTHR[e] = thr.x;
                    QTHR[kHtQueue + e] = thr.y;
                    QTHR[2 * kHtQueue + e] = thr.z;
                    if (FEAT & RT_FEAT_MOVING) QTIME[e] = ray.time;
                    QMETA[e] = owner | ((bounce + 1u) << 5) | ((sample - (uint32_t)args.sampleBegin) << 13);
                }
                nQ += __popc(sm_);
                __syncwarp();
            }
        }

        if (valid) {
            float* px = args.accum + ((size_t)(py0 + lane / kTileW) * cam.width + px0 + (lane & (kTileW - 1))) * 3u;
            px[0] += SUM[lane];
            px[1] += SUM[32 + lane];
            px[2] += SUM[64 + lane];
        }
        __syncwarp();
    }

    for (int off = 16; off > 0; off >>= 1) {
        nRays += __shfl_down_sync(FULL, nRays, off);
        if (STATS) {
            nPaths += __shfl_down_sync(FULL, nPaths, off);
            nNode += __shfl_down_sync(FULL, nNode, off);
            nPrim += __shfl_down_sync(FULL, nPrim, off);
        }
    }
    if (lane == 0) {
        atomicAdd(&args.stats[0], nRays);
        if (STATS) {
            atomicAdd(&args.stats[1], nPaths);
            atomicAdd(&args.stats[2], nNode);
            atomicAdd(&args.stats[3], nPrim);
        }
    }
}

// kernel.cu:147-153 (mean, sqrt gamma) + :712-718 (clamp to [0,0.999], *256),
// fused with the row flip to PPM order (kernel.cu:699: top row first).
__global__ void ResolveKernel(const float* __restrict__ accum, float* __restrict__ linearOut, uint8_t* __restrict__ srgbOut,
                              int width, int height, float invSpp)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = width * height;
    if (idx >= n) return;
    const float r = accum[idx * 3 + 0] * invSpp, g = accum[idx * 3 + 1] * invSpp, b = accum[idx * 3 + 2] * invSpp;
    if (linearOut) {
        linearOut[idx * 3 + 0] = r;
        linearOut[idx * 3 + 1] = g;
        linearOut[idx * 3 + 2] = b;
    }
    if (srgbOut) {
        const int i = idx % width, j = idx / width;
        const int o = ((height - 1 - j) * width + i) * 3;
        const float c[3] = {sqrtf(r), sqrtf(g), sqrtf(b)};
        for (int k = 0; k < 3; ++k) {
            const float v = c[k] < 0.0f ? 0.0f : (c[k] > 0.999f ? 0.999f : c[k]);
            srgbOut[o + k] = (uint8_t)(int)(256.0f * v);
        }
    }
}

// Roofline denominator: dependent-free FFMA chains, 2 flop per FFMA per lane.
__global__ void FmaPeakKernel(float* out, int iters)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f,
          a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a0 = fmaf(a0, m, c);
            a1 = fmaf(a1, m, c);
            a2 = fmaf(a2, m, c);
            a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c);
            a5 = fmaf(a5, m, c);
            a6 = fmaf(a6, m, c);
            a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

using KernelFn = void (*)(const DevScene, const DevCamera, const RenderArgs);

template <int FEAT> KernelFn PickKernel(int variant, bool smem, bool stats)
{
    const bool wave = variant == RT_VARIANT_WAVEFRONT;
    if (variant == RT_VARIANT_HEADTAIL) {
        if (smem) return stats ? RenderHeadTail<FEAT, true, true> : RenderHeadTail<FEAT, true, false>;
        return stats ? RenderHeadTail<FEAT, false, true> : RenderHeadTail<FEAT, false, false>;
    }
    if (wave) {
        if (smem) return stats ? RenderWave<FEAT, true, true> : RenderWave<FEAT, true, false>;
        return stats ? RenderWave<FEAT, false, true> : RenderWave<FEAT, false, false>;
    }
    if (smem) return stats ? RenderMega<FEAT, true, true> : RenderMega<FEAT, true, false>;
    return stats ? RenderMega<FEAT, false, true> : RenderMega<FEAT, false, false>;
}

// Instantiations: spheres only / + moving spheres and textures / everything.
constexpr int kFeatSpheres = 0;
constexpr int kFeatMotion = RT_FEAT_MOVING | RT_FEAT_TEXTURE;
constexpr int kFeatAll = RT_FEAT_MOVING | RT_FEAT_QUAD | RT_FEAT_MEDIUM | RT_FEAT_TEXTURE | RT_FEAT_TEXTURE_HEAVY;

KernelFn PickKernelForFeatures(int features, int variant, bool smem, bool stats, int* picked)
{
    if (features == 0) {
        *picked = kFeatSpheres;
        return PickKernel<kFeatSpheres>(variant, smem, stats);
    }
    if ((features & ~kFeatMotion) == 0) {
        *picked = kFeatMotion;
        return PickKernel<kFeatMotion>(variant, smem, stats);
    }
    *picked = kFeatAll;
    return PickKernel<kFeatAll>(variant, smem, stats);
}

// Frame-sized device buffers (accumulator, readback staging) are recycled across
// handles: cudaMalloc/cudaFree of ~100 MB blocks was measured at up to 0.5 s per
// upload/free cycle, which is what an application rendering frame after frame does.
struct BigBlock {
    int device;
    size_t bytes;
    void* ptr;
};
std::mutex gBigMutex;
std::vector<BigBlock> gBigCache;
constexpr size_t kBigCacheEntries = 8;

cudaError_t BigMalloc(int device, void** out, size_t bytes)
{
    {
        std::lock_guard<std::mutex> lock(gBigMutex);
        for (size_t k = 0; k < gBigCache.size(); ++k)
            if (gBigCache[k].device == device && gBigCache[k].bytes == bytes) {
                *out = gBigCache[k].ptr;
                gBigCache.erase(gBigCache.begin() + (long)k);
                return cudaSuccess;
            }
    }
    return cudaMalloc(out, bytes);
}

void BigFree(int device, void* ptr, size_t bytes)
{
    if (!ptr) return;
    {
        std::lock_guard<std::mutex> lock(gBigMutex);
        if (gBigCache.size() < kBigCacheEntries) {
            gBigCache.push_back(BigBlock{device, bytes, ptr});
            return;
        }
    }
    cudaFree(ptr);
}

// Host image of the device arena: every table of the scene back to back, 256-byte
// aligned, so that one allocation and ONE host-to-device copy upload the scene
// (and one ncclBroadcast would replicate it).
struct ArenaBuilder {
    std::vector<char> bytes;
    template <class T> size_t Add(const std::vector<T>& v)
    {
        // at least one zeroed record so that staging code can always read 16 bytes
        const size_t n = v.empty() ? 1 : v.size();
        const size_t at = (bytes.size() + 255) / 256 * 256;
        bytes.resize(at + (n * sizeof(T) + 15) / 16 * 16, 0);
        if (!v.empty()) std::memcpy(bytes.data() + at, v.data(), v.size() * sizeof(T));
        return at;
    }
    size_t Reserve(size_t n)
    {
        const size_t at = (bytes.size() + 255) / 256 * 256;
        bytes.resize(at + (n + 15) / 16 * 16, 0);
        return at;
    }
};

} // namespace

struct rt_scene_s {
    int device = 0;
    DevScene dev{};
    rtpack::Packed* host = nullptr; // kept for sizes / info
    char* arena = nullptr; // one device block: scene tables, images, counters, debug records
    size_t arenaBytes = 0;
    uint64_t deviceBytes = 0;
    bool fitsSmem = false;
    uint32_t stagedBytes = 0;
    int smCount = 0;
    int maxSmemOptin = 0;
    // accumulator + counters
    float* accum = nullptr;
    size_t accumFloats = 0;
    float* linearStage = nullptr;
    uint8_t* srgbStage = nullptr;
    size_t linearStageFloats = 0, srgbStageBytes = 0;
    unsigned long long* stats = nullptr;
    unsigned int* tileCounter = nullptr;
    float* debugOut = nullptr; // test hook, 256*8 floats
    int debugPixel = -1, debugSample = -1;
    cudaStream_t lastStream = nullptr;
    rt_camera lastCam{};
    bool rendered = false;
    int pickedFeatures = 0;
    int pickedVariant = 0;
};

extern "C" {

int rt_scene_upload(const rt_scene_desc* scene, const rt_upload_options* opt, rt_scene_handle* out)
{
    if (!scene || !out) {
        rt_set_error("rt_scene_upload: NULL argument");
        return RT_ERR_INVALID;
    }
    *out = nullptr;
    rt_upload_options o{};
    if (opt) o = *opt;
    // host work first (validation, baking, BVH): a malformed scene is reported
    // as such even on a machine without a GPU
    rtpack::Packed* packed = nullptr;
    try {
        rtpack::Packer pk(*scene, o);
        pk.Run();
        packed = new rtpack::Packed(std::move(pk.out));
    } catch (const std::exception& e) {
        rt_set_error("rt_scene_upload: %s", e.what());
        return RT_ERR_INVALID;
    }
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev <= 0) {
        delete packed;
        rt_set_error("rt_scene_upload: no CUDA device available (this library has no CPU path)");
        return RT_ERR_NO_DEVICE;
    }
    if (o.device < 0 || o.device >= nDev) {
        delete packed;
        rt_set_error("rt_scene_upload: device %d out of range (have %d)", o.device, nDev);
        return RT_ERR_INVALID;
    }
    RT_CUDA(cudaSetDevice(o.device));
    rt_scene_s* h = new rt_scene_s();
    h->device = o.device;
    h->host = packed;
    // two attributes, not cudaGetDeviceProperties: the full query costs tens of ms per call
    RT_CUDA(cudaDeviceGetAttribute(&h->smCount, cudaDevAttrMultiProcessorCount, o.device));
    RT_CUDA(cudaDeviceGetAttribute(&h->maxSmemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, o.device));

    ArenaBuilder ab;
    const size_t oNodes = ab.Add(packed->nodes), oSpheres = ab.Add(packed->spheres);
    const size_t oSphereMat = ab.Add(packed->sphere_material), oMoving = ab.Add(packed->moving);
    const size_t oQuads = ab.Add(packed->quads), oMedia = ab.Add(packed->media), oMaterials = ab.Add(packed->materials);
    const size_t oTextures = ab.Add(packed->textures), oPerlins = ab.Add(packed->perlins);
    std::vector<size_t> oImage(packed->image_bytes.size(), 0);
    for (size_t k = 0; k < packed->image_bytes.size(); ++k)
        if (!packed->image_bytes[k].empty()) oImage[k] = ab.Add(packed->image_bytes[k]);
    const size_t oImages = ab.Reserve(std::max<size_t>(1, packed->image_bytes.size()) * sizeof(DevImage));
    const size_t oStats = ab.Reserve(4 * sizeof(unsigned long long)), oTile = ab.Reserve(sizeof(unsigned int));
    const size_t oDebug = ab.Reserve(256 * 8 * sizeof(float));
    h->arenaBytes = (ab.bytes.size() + 255) / 256 * 256;
    ab.bytes.resize(h->arenaBytes, 0);
    {
        void* p = nullptr;
        const cudaError_t e = BigMalloc(h->device, &p, h->arenaBytes);
        if (e != cudaSuccess) {
            rt_set_error("rt_scene_upload: cudaMalloc of %zu bytes failed: %s", h->arenaBytes, cudaGetErrorString(e));
            rt_scene_free(h);
            return RT_ERR_CUDA;
        }
        h->arena = static_cast<char*>(p);
    }
    for (size_t k = 0; k < packed->image_bytes.size(); ++k) { // image table: device addresses inside the arena
        DevImage im{};
        im.width = packed->image_w[k];
        im.height = packed->image_h[k];
        im.rgb = packed->image_bytes[k].empty() ? nullptr : reinterpret_cast<const uint8_t*>(h->arena + oImage[k]);
        std::memcpy(ab.bytes.data() + oImages + k * sizeof(DevImage), &im, sizeof im);
    }
    {
        const cudaError_t e = cudaMemcpy(h->arena, ab.bytes.data(), h->arenaBytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            rt_set_error("rt_scene_upload: host-to-device copy failed: %s", cudaGetErrorString(e));
            rt_scene_free(h);
            return RT_ERR_CUDA;
        }
    }
    h->deviceBytes = h->arenaBytes;
    h->dev.nodes = reinterpret_cast<const DevNode*>(h->arena + oNodes);
    h->dev.spheres = reinterpret_cast<const DevSphere*>(h->arena + oSpheres);
    h->dev.sphere_material = reinterpret_cast<const int32_t*>(h->arena + oSphereMat);
    h->dev.moving = reinterpret_cast<const DevMovingSphere*>(h->arena + oMoving);
    h->dev.quads = reinterpret_cast<const DevQuad*>(h->arena + oQuads);
    h->dev.media = reinterpret_cast<const DevMedium*>(h->arena + oMedia);
    h->dev.materials = reinterpret_cast<const DevMaterial*>(h->arena + oMaterials);
    h->dev.textures = reinterpret_cast<const DevTexture*>(h->arena + oTextures);
    h->dev.perlins = reinterpret_cast<const DevPerlin*>(h->arena + oPerlins);
    h->dev.images = reinterpret_cast<const DevImage*>(h->arena + oImages);
    h->stats = reinterpret_cast<unsigned long long*>(h->arena + oStats);
    h->tileCounter = reinterpret_cast<unsigned int*>(h->arena + oTile);
    h->debugOut = reinterpret_cast<float*>(h->arena + oDebug);
    h->dev.root_ref = packed->root_ref;
    h->dev.n_nodes = (int)packed->nodes.size();
    h->dev.n_spheres = (int)packed->spheres.size();
    h->dev.n_moving = (int)packed->moving.size();
    h->dev.n_quads = (int)packed->quads.size();
    h->dev.n_media = (int)packed->media.size();
    h->dev.n_materials = (int)packed->materials.size();
    h->dev.n_textures = (int)packed->textures.size();
    h->dev.features = packed->features;

    auto pad16 = [](size_t b) { return (uint32_t)((b + 15) / 16 * 16); };
    h->stagedBytes = pad16(std::max<size_t>(1, packed->nodes.size()) * sizeof(DevNode)) +
                     pad16(std::max<size_t>(1, packed->spheres.size()) * sizeof(DevSphere)) +
                     pad16(std::max<size_t>(1, packed->sphere_material.size()) * sizeof(int32_t)) +
                     pad16(std::max<size_t>(1, packed->moving.size()) * sizeof(DevMovingSphere)) +
                     pad16(std::max<size_t>(1, packed->quads.size()) * sizeof(DevQuad)) +
                     pad16(std::max<size_t>(1, packed->media.size()) * sizeof(DevMedium)) +
                     pad16(std::max<size_t>(1, packed->materials.size()) * sizeof(DevMaterial));

    *out = h;
    return RT_OK;
}

int rt_render(rt_scene_handle h, const rt_camera* cam, const rt_render_params* p)
{
    if (!h || !cam || !p) {
        rt_set_error("rt_render: NULL argument");
        return RT_ERR_INVALID;
    }
    if (cam->image_width <= 0 || cam->image_height <= 0 || cam->max_depth <= 0 || cam->max_depth > 254 ||
        p->sample_end < p->sample_begin || p->sample_begin < 0) {
        rt_set_error("rt_render: bad camera or sample range (max_depth must be 1..254)");
        return RT_ERR_INVALID;
    }
    if ((long long)cam->image_width * cam->image_height > 0x7fffffffLL / 3) {
        rt_set_error("rt_render: image too large");
        return RT_ERR_INVALID;
    }
    if (p->variant < RT_VARIANT_AUTO || p->variant > RT_VARIANT_HEADTAIL) {
        rt_set_error("rt_render: unknown variant %d", p->variant);
        return RT_ERR_INVALID;
    }
    // AUTO: the head/tail kernel (measured +17..39 % over the megakernel on scenes 0, 7, 8, 9, 10) unless the
    // sample range exceeds its packed sample index.
    int variant = p->variant;
    if (variant == RT_VARIANT_AUTO)
        variant = (long long)p->sample_end - p->sample_begin <= RT_HT_MAX_SAMPLES ? RT_VARIANT_HEADTAIL : RT_VARIANT_MEGAKERNEL;
    const bool wave = variant == RT_VARIANT_WAVEFRONT;
    const bool headTail = variant == RT_VARIANT_HEADTAIL;
    if (headTail && (long long)p->sample_end - p->sample_begin > RT_HT_MAX_SAMPLES) {
        rt_set_error("rt_render: the head/tail variant takes at most %d samples per call", RT_HT_MAX_SAMPLES);
        return RT_ERR_INVALID;
    }
    if (wave && p->sample_end >= (1 << 24)) { // its slots pack sample << 8 | bounce
        rt_set_error("rt_render: the wavefront variant takes sample indices below 2^24");
        return RT_ERR_INVALID;
    }
    RT_CUDA(cudaSetDevice(h->device));
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(p->stream);
    const size_t nFloats = (size_t)cam->image_width * cam->image_height * 3;
    float* accum = p->accum;
    if (!accum) {
        if (h->accumFloats != nFloats) {
            BigFree(h->device, h->accum, h->accumFloats * sizeof(float));
            h->accum = nullptr;
            h->accumFloats = 0;
            RT_CUDA(BigMalloc(h->device, reinterpret_cast<void**>(&h->accum), nFloats * sizeof(float)));
            h->accumFloats = nFloats;
            RT_CUDA(cudaMemsetAsync(h->accum, 0, nFloats * sizeof(float), stream));
        }
        accum = h->accum;
    }
    if (p->clear) {
        RT_CUDA(cudaMemsetAsync(accum, 0, nFloats * sizeof(float), stream));
        RT_CUDA(cudaMemsetAsync(h->stats, 0, 4 * sizeof(unsigned long long), stream));
    }
    RT_CUDA(cudaMemsetAsync(h->tileCounter, 0, sizeof(unsigned int), stream));

    const DevCamera dc = rtpack::MakeCamera(*cam);
    RenderArgs a{};
    a.accum = accum;
    a.stats = h->stats;
    a.tileCounter = h->tileCounter;
    a.sampleBegin = p->sample_begin;
    a.sampleEnd = p->sample_end;
    a.seed = p->seed;
    a.debugPixel = h->debugPixel;
    a.debugSample = h->debugSample;
    a.debugOut = h->debugPixel >= 0 ? h->debugOut : nullptr;
    // tiles: 8x4 pixels per warp (megakernel: one pixel per lane) or 8x8 (wavefront: 64 path slots)
    // leaf turn every 2nd step (measured best: 1 -> 11.3, 2 -> 11.5, 4 -> 11.2 Grays/s); development knob in
    // flags bits 4-5: 1 = every step, 2 = every 4th, 3 = every 8th
    static const int kLeafMasks[4] = {1, 0, 3, 7};
    a.megaLeafMask = kLeafMasks[(p->flags >> 4) & 3];
    // tuning knobs of the wavefront variant (development): flags bits 12-15 slots/32,
    // 16-20 idle-exit, 21-25 leaf batch, 26-30 refill minimum
    int waveSlots = ((p->flags >> 12) & 0xf) * 32;
    a.waveIdleExit = (p->flags >> 16) & 0x1f;
    a.waveLeafBatch = (p->flags >> 21) & 0x1f;
    a.waveRefillMin = (p->flags >> 26) & 0x1f;
    if (a.waveIdleExit <= 0) a.waveIdleExit = 16;
    if (a.waveLeafBatch <= 0) a.waveLeafBatch = 16;
    if (a.waveRefillMin <= 0) a.waveRefillMin = 8;
    auto pad16 = [](size_t b) { return (uint32_t)((b + 15) / 16 * 16); };
    const rtpack::Packed& pk = *h->host;
    a.nodesBytes = pad16(std::max<size_t>(1, pk.nodes.size()) * sizeof(DevNode));
    a.spheresBytes = pad16(std::max<size_t>(1, pk.spheres.size()) * sizeof(DevSphere));
    a.sphereMatBytes = pad16(std::max<size_t>(1, pk.sphere_material.size()) * sizeof(int32_t));
    a.movingBytes = pad16(std::max<size_t>(1, pk.moving.size()) * sizeof(DevMovingSphere));
    a.quadsBytes = pad16(std::max<size_t>(1, pk.quads.size()) * sizeof(DevQuad));
    a.mediaBytes = pad16(std::max<size_t>(1, pk.media.size()) * sizeof(DevMedium));
    a.materialsBytes = pad16(std::max<size_t>(1, pk.materials.size()) * sizeof(DevMaterial));

    const int featClassEarly = h->dev.features == 0 ? 0 : ((h->dev.features & ~(RT_FEAT_MOVING | RT_FEAT_TEXTURE)) == 0 ? (RT_FEAT_MOVING | RT_FEAT_TEXTURE) : 31);
    const int maxThreads = wave ? 512 : (headTail ? HtMaxThreads(featClassEarly) : MegaMaxThreads(h->dev.features == 0 ? 0 : 1));
    int threads = p->block_threads > 0 ? p->block_threads : maxThreads;
    threads = std::max(32, std::min(maxThreads, (threads / 32) * 32));
    int blocksPerSm = p->blocks_per_sm > 0 ? p->blocks_per_sm : 1;
    const int stackLevels = std::max(3, std::min(kMaxStackLevels, h->host->max_depth + 3)); // + sentinel slot
    a.stackLevels = stackLevels;
    if (wave) {
        threads = std::min(threads, 512); // __launch_bounds__(512, 1)
        blocksPerSm = 1;
    }
    const int featClass = h->dev.features == 0 ? kFeatSpheres : ((h->dev.features & ~kFeatMotion) == 0 ? kFeatMotion : kFeatAll);
    if (headTail && p->block_threads <= 0 && !(p->flags & 0x200)) {
        // the largest block whose stacks + queues still leave room for the scene in shared memory
        // (if none does, the scene stays in global memory and the block is as large as registers allow)
        threads = maxThreads;
        for (int cand = maxThreads; cand >= 384; cand -= 128)
            if ((size_t)cand * 4 * stackLevels + (size_t)(cand / 32) * HtWarpBytes(featClass) + 16 + h->stagedBytes <=
                (size_t)h->maxSmemOptin) {
                threads = cand;
                break;
            }
    }
    const size_t stackBytes = (size_t)threads * 4 * stackLevels;
    size_t poolBytes = headTail ? (size_t)(threads / 32) * HtWarpBytes(featClass) + 16 : 0;
    if (wave) {
        // 64 slots (an 8x8 tile) measured best: 96 or 128 fill SHADE/GEN chunks better but cost more shared
        // memory traffic and longer tile tails (profiles/r1_wavefront_parameter_sweep.json)
        auto bytesFor = [&](int slots) { return (size_t)(threads / 32) * PoolBytes(slots) + 16; };
        if (waveSlots < 64 || waveSlots > 128) waveSlots = 64;
        poolBytes = bytesFor(waveSlots);
    }
    a.waveSlots = waveSlots;
    const int tileH = wave ? waveSlots / 8 : kTileH;
    a.tilesX = (cam->image_width + kTileW - 1) / kTileW;
    a.tilesY = (cam->image_height + tileH - 1) / tileH;
    const bool wantStats = (p->flags & 0x100) != 0 || a.debugOut != nullptr;
    const bool smem = !(p->flags & 0x200) &&
                      stackBytes + poolBytes + h->stagedBytes <= (size_t)h->maxSmemOptin / (size_t)blocksPerSm;
    const size_t smemBytes = stackBytes + poolBytes + (smem ? h->stagedBytes : 0);
    if (smemBytes > (size_t)h->maxSmemOptin) {
        rt_set_error("rt_render: block of %d threads needs %zu B of shared memory (max %d)", threads, smemBytes,
                     h->maxSmemOptin);
        return RT_ERR_INVALID;
    }
    KernelFn fn = PickKernelForFeatures(h->dev.features, variant, smem, wantStats, &h->pickedFeatures);
    h->pickedVariant = variant;
    RT_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    const int nTiles = a.tilesX * a.tilesY;
    const int warpsPerBlock = threads / 32;
    int blocks = h->smCount * blocksPerSm;
    blocks = std::max(1, std::min(blocks, (nTiles + warpsPerBlock - 1) / warpsPerBlock));
    h->fitsSmem = smem;
    fn<<<blocks, threads, smemBytes, stream>>>(h->dev, dc, a);
    RT_CUDA(cudaGetLastError());
    h->lastStream = stream;
    h->lastCam = *cam;
    h->rendered = true;
    return RT_OK;
}

int rt_accum_ptr(rt_scene_handle h, float** dev_ptr, uint64_t* n_floats)
{
    if (!h || !dev_ptr) {
        rt_set_error("rt_accum_ptr: NULL argument");
        return RT_ERR_INVALID;
    }
    *dev_ptr = h->accum;
    if (n_floats) *n_floats = h->accumFloats;
    return h->accum ? RT_OK : RT_ERR_STATE;
}

int rt_sync(rt_scene_handle h)
{
    if (!h) {
        rt_set_error("rt_sync: NULL handle");
        return RT_ERR_INVALID;
    }
    RT_CUDA(cudaSetDevice(h->device));
    RT_CUDA(cudaStreamSynchronize(h->lastStream));
    return RT_OK;
}

int rt_readback(rt_scene_handle h, const float* accum, float* linear_rgb, uint8_t* srgb8, rt_stats* stats)
{
    if (!h) {
        rt_set_error("rt_readback: NULL handle");
        return RT_ERR_INVALID;
    }
    if (!h->rendered) {
        rt_set_error("rt_readback: nothing rendered yet");
        return RT_ERR_STATE;
    }
    RT_CUDA(cudaSetDevice(h->device));
    const int W = h->lastCam.image_width, H = h->lastCam.image_height;
    const size_t nFloats = (size_t)W * H * 3;
    const float* src = accum ? accum : h->accum;
    if (!src && (linear_rgb || srgb8)) {
        rt_set_error("rt_readback: no accumulator (the render used a caller-owned one: pass it as `accum`)");
        return RT_ERR_STATE;
    }
    cudaStream_t stream = h->lastStream;
    if (linear_rgb || srgb8) {
        if (linear_rgb && h->linearStageFloats < nFloats) {
            BigFree(h->device, h->linearStage, h->linearStageFloats * sizeof(float));
            h->linearStage = nullptr;
            h->linearStageFloats = 0;
            RT_CUDA(BigMalloc(h->device, reinterpret_cast<void**>(&h->linearStage), nFloats * sizeof(float)));
            h->linearStageFloats = nFloats;
        }
        if (srgb8 && h->srgbStageBytes < nFloats) {
            BigFree(h->device, h->srgbStage, h->srgbStageBytes);
            h->srgbStage = nullptr;
            h->srgbStageBytes = 0;
            RT_CUDA(BigMalloc(h->device, reinterpret_cast<void**>(&h->srgbStage), nFloats));
            h->srgbStageBytes = nFloats;
        }
        const int n = W * H;
        const float invSpp = 1.0f / (float)h->lastCam.samples_per_pixel;
        ResolveKernel<<<(n + 255) / 256, 256, 0, stream>>>(src, linear_rgb ? h->linearStage : nullptr,
                                                           srgb8 ? h->srgbStage : nullptr, W, H, invSpp);
        RT_CUDA(cudaGetLastError());
        if (linear_rgb)
            RT_CUDA(cudaMemcpyAsync(linear_rgb, h->linearStage, nFloats * sizeof(float), cudaMemcpyDeviceToHost, stream));
        if (srgb8) RT_CUDA(cudaMemcpyAsync(srgb8, h->srgbStage, nFloats, cudaMemcpyDeviceToHost, stream));
    }
    if (stats) {
        unsigned long long s[4];
        RT_CUDA(cudaMemcpyAsync(s, h->stats, sizeof s, cudaMemcpyDeviceToHost, stream));
        RT_CUDA(cudaStreamSynchronize(stream));
        stats->rays = s[0];
        stats->paths = s[1];
        stats->node_tests = s[2];
        stats->prim_tests = s[3];
    }
    RT_CUDA(cudaStreamSynchronize(stream));
    return RT_OK;
}

int rt_scene_free(rt_scene_handle h)
{
    if (!h) return RT_OK;
    cudaSetDevice(h->device);
    if (h->rendered) cudaStreamSynchronize(h->lastStream); // recycled buffers must be idle
    BigFree(h->device, h->arena, h->arenaBytes);
    BigFree(h->device, h->accum, h->accumFloats * sizeof(float));
    BigFree(h->device, h->linearStage, h->linearStageFloats * sizeof(float));
    BigFree(h->device, h->srgbStage, h->srgbStageBytes);
    delete h->host;
    delete h;
    return RT_OK;
}

int rt_release_cached_memory(void)
{
    std::vector<BigBlock> blocks;
    {
        std::lock_guard<std::mutex> lock(gBigMutex);
        blocks.swap(gBigCache);
    }
    for (const BigBlock& b : blocks) {
        if (cudaSetDevice(b.device) == cudaSuccess) cudaFree(b.ptr);
    }
    return RT_OK;
}

int rt_scene_get_info(rt_scene_handle h, rt_scene_info* info)
{
    if (!h || !info) {
        rt_set_error("rt_scene_get_info: NULL argument");
        return RT_ERR_INVALID;
    }
    std::memset(info, 0, sizeof *info);
    info->n_prims_baked = h->dev.n_spheres + h->dev.n_moving + h->dev.n_quads;
    info->n_nodes = h->dev.n_nodes;
    info->n_media = h->dev.n_media;
    info->max_depth_bvh = h->host->max_depth;
    info->features = h->dev.features;
    info->scene_in_smem = h->fitsSmem ? 1 : 0;
    info->variant = h->pickedVariant;
    info->device_bytes = h->deviceBytes;
    for (int k = 0; k < 8; ++k) info->medium_visits[k] = h->host->medium_visits[k];
    return RT_OK;
}

int rt_debug_trace_path(rt_scene_handle h, const rt_camera* cam, const rt_render_params* p, int32_t pixel,
                        int32_t sample, float* records, int32_t max_records)
{
    if (!h || !cam || !p || !records || max_records <= 0 || max_records > 256) {
        rt_set_error("rt_debug_trace_path: bad argument");
        return RT_ERR_INVALID;
    }
    RT_CUDA(cudaSetDevice(h->device));
    RT_CUDA(cudaMemset(h->debugOut, 0, 256 * 8 * sizeof(float)));
    h->debugPixel = pixel;
    h->debugSample = sample;
    rt_render_params q = *p;
    q.sample_begin = sample;
    q.sample_end = sample + 1;
    const int rc = rt_render(h, cam, &q);
    h->debugPixel = h->debugSample = -1;
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaStreamSynchronize(h->lastStream));
    RT_CUDA(cudaMemcpy(records, h->debugOut, (size_t)max_records * 8 * sizeof(float), cudaMemcpyDeviceToHost));
    return RT_OK;
}

float rt_rng_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t domain, uint32_t dim)
{
    const rt_u4 b = rt_rng_block(seed, pixel, sample, slot, domain, dim >> 2);
    return rt_bits_to_u01(rt_u4_lane(b, dim & 3u));
}

// P3 text: the reference's format (kernel.cu:696-723), byte for byte; digits come from a
// 256-entry table instead of a printf per pixel (4K: 95 MB of text in ~0.1 s instead of ~2 s).
static int WritePpm(const char* path, const uint8_t* srgb8, int32_t width, int32_t height, bool binary, const char* who)
{
    if (!path || !srgb8 || width <= 0 || height <= 0) {
        rt_set_error("%s: bad argument", who);
        return RT_ERR_INVALID;
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) {
        rt_set_error("%s: cannot open %s", who, path);
        return RT_ERR_INVALID;
    }
    const size_t n = (size_t)width * height;
    bool ok = std::fprintf(f, "%s\n%d %d\n255\n", binary ? "P6" : "P3", width, height) > 0;
    if (binary) {
        ok = ok && std::fwrite(srgb8, 1, n * 3, f) == n * 3;
    } else {
        char digits[256][4];
        uint8_t len[256];
        for (int v = 0; v < 256; ++v) len[v] = (uint8_t)std::snprintf(digits[v], sizeof digits[v], "%d", v);
        std::vector<char> buf;
        buf.reserve((1u << 20) + 16);
        for (size_t k = 0; k < n && ok; ++k) {
            for (int c = 0; c < 3; ++c) {
                const uint8_t v = srgb8[3 * k + c];
                buf.insert(buf.end(), digits[v], digits[v] + len[v]);
                buf.push_back(c == 2 ? '\n' : ' ');
            }
            if (buf.size() >= (1u << 20)) {
                ok = std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
                buf.clear();
            }
        }
        ok = ok && std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) {
        rt_set_error("%s: write to %s failed", who, path);
        return RT_ERR_INVALID;
    }
    return RT_OK;
}

int rt_write_ppm(const char* path, const uint8_t* srgb8, int32_t width, int32_t height)
{
    return WritePpm(path, srgb8, width, height, false, "rt_write_ppm");
}

int rt_write_ppm_binary(const char* path, const uint8_t* srgb8, int32_t width, int32_t height)
{
    return WritePpm(path, srgb8, width, height, true, "rt_write_ppm_binary");
}

int rt_measure_fp32_peak(int32_t device, double* tflops, double* sm_count)
{
    if (!tflops) {
        rt_set_error("rt_measure_fp32_peak: NULL argument");
        return RT_ERR_INVALID;
    }
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || device < 0 || device >= nDev) {
        rt_set_error("rt_measure_fp32_peak: no such CUDA device");
        return RT_ERR_NO_DEVICE;
    }
    RT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    RT_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    float* out = nullptr;
    RT_CUDA(cudaMalloc(reinterpret_cast<void**>(&out), (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    RT_CUDA(cudaEventCreate(&e0));
    RT_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        RT_CUDA(cudaEventRecord(e0));
        FmaPeakKernel<<<blocks, threads>>>(out, iters);
        RT_CUDA(cudaEventRecord(e1));
        RT_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        RT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return RT_OK;
}

int rt_abi_sizeof(const char* name)
{
    const std::string n = name ? name : "";
    if (n == "rt_prim") return (int)sizeof(rt_prim);
    if (n == "rt_xform") return (int)sizeof(rt_xform);
    if (n == "rt_object") return (int)sizeof(rt_object);
    if (n == "rt_material") return (int)sizeof(rt_material);
    if (n == "rt_texture") return (int)sizeof(rt_texture);
    if (n == "rt_perlin") return (int)sizeof(rt_perlin);
    if (n == "rt_image") return (int)sizeof(rt_image);
    if (n == "rt_scene_desc") return (int)sizeof(rt_scene_desc);
    if (n == "rt_camera") return (int)sizeof(rt_camera);
    if (n == "rt_upload_options") return (int)sizeof(rt_upload_options);
    if (n == "rt_render_params") return (int)sizeof(rt_render_params);
    if (n == "rt_stats") return (int)sizeof(rt_stats);
    if (n == "rt_scene_info") return (int)sizeof(rt_scene_info);
    return -1;
}

} // extern "C"
