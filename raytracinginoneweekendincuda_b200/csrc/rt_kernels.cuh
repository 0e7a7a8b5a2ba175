// rt_kernels.cuh -- the render kernels (included by rt_device.cu only).
//
// Four kernels render the same image from the same device functions (rt_trace.cuh):
//   RenderHitQueue  what RT_VARIANT_AUTO runs (rt_kernel_hq.cuh): heads + queued HITS, shading on full warps
//   RenderHeadTail  round-1 shipping kernel: heads + queued continuation RAYS
//   RenderMega      persistent-thread megakernel
//   RenderWave      on-chip wavefront (kept for the comparison DESIGN.md reports)
// All replace the reference's RenderInit + Render launches (reference kernel.cu:110-154, launched :681-689).
#pragma once

#include "rt_trace.cuh"

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
    uint32_t nodesBytes, spheresBytes, sphereMatBytes, movingBytes, quadsBytes, boxesBytes, mediaBytes, materialsBytes, matParamsBytes;
    uint32_t perlinBytes; // sizeof(DevPerlin) when Perlin table 0 is staged in shared memory, else 0
    int stageNodesOnly; // scene in global memory, node table staged (SceneView::nodes_shared)
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

// The node table on its way into shared memory: the ref of an internal node (index of its child pair) becomes the
// pair's shared ADDRESS, so a traversal step loads from [ref + imm] without a base register or a shift.  Leaf refs
// (bit 31 set) stay as they are; shared addresses never reach bit 31.
__device__ __forceinline__ uint32_t StageNodes(uint32_t& cursor, char* smem, uint32_t smemBase, const void* src, uint32_t bytes)
{
    const uint32_t at = cursor;
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(smem + at);
    for (uint32_t k = threadIdx.x; k < bytes / 16u; k += blockDim.x) {
        uint4 v = __ldg(&s[k]);
        if ((k & 1u) == 0u && !(v.w & RT_REF_LEAF)) v.w = smemBase + at + v.w * 32u; // DevNode: {c.xyz, ref}{e.xyz, aux}
        d[k] = v;
    }
    cursor += (bytes + 15u) & ~15u;
    return at;
}

// Stages the scene arrays in shared memory (SMEM) or points at them in global memory.
template <bool SMEM>
__device__ __forceinline__ SceneView<SMEM> SetupScene(const DevScene& scene, const RenderArgs& args, char* smem, uint32_t smemBase,
                                                      uint32_t& cursor)
{
    SceneView<SMEM> sv;
    const uint32_t nodesAt = cursor;
    (void)nodesAt;
    if constexpr (SMEM) {
        sv.nodes.a = smemBase + StageNodes(cursor, smem, smemBase, scene.nodes, args.nodesBytes);
        sv.spheres.a = smemBase + Stage(cursor, smem, scene.spheres, args.spheresBytes);
        sv.sphere_material.a = smemBase + Stage(cursor, smem, scene.sphere_material, args.sphereMatBytes);
        sv.moving.a = smemBase + Stage(cursor, smem, scene.moving, args.movingBytes);
        sv.quads.a = smemBase + Stage(cursor, smem, scene.quads, args.quadsBytes);
        sv.boxes.a = smemBase + Stage(cursor, smem, scene.boxes, args.boxesBytes);
        sv.media.a = smemBase + Stage(cursor, smem, scene.media, args.mediaBytes);
        sv.materials.a = smemBase + Stage(cursor, smem, scene.materials, args.materialsBytes);
        sv.mat_params.a = smemBase + Stage(cursor, smem, scene.mat_params, args.matParamsBytes);
        __syncthreads();
    } else {
        if (args.stageNodesOnly) {
            StageNodes(cursor, smem, smemBase, scene.nodes, args.nodesBytes);
            __syncthreads();
        }
        sv.nodes.a = reinterpret_cast<const char*>(scene.nodes);
        sv.spheres.a = reinterpret_cast<const char*>(scene.spheres);
        sv.sphere_material.a = reinterpret_cast<const char*>(scene.sphere_material);
        sv.moving.a = reinterpret_cast<const char*>(scene.moving);
        sv.quads.a = reinterpret_cast<const char*>(scene.quads);
        sv.boxes.a = reinterpret_cast<const char*>(scene.boxes);
        sv.media.a = reinterpret_cast<const char*>(scene.media);
        sv.materials.a = reinterpret_cast<const char*>(scene.materials);
        sv.mat_params.a = reinterpret_cast<const char*>(scene.mat_params);
    }
    sv.textures = scene.textures;
    sv.perlins = scene.perlins;
    sv.perlin0 = scene.perlins;
    if (args.perlinBytes) { // the marble texture's lattice tables (Perlin.h:22-34): 5 KB, 56 x 4 loads per shaded hit
        const uint32_t at = Stage(cursor, smem, scene.perlins, args.perlinBytes);
        sv.perlin0 = reinterpret_cast<const DevPerlin*>(smem + at);
        __syncthreads();
    }
    sv.images = scene.images;
    sv.uv_frames = scene.uv_frames;
    sv.lights = scene.lights;
    sv.n_lights = scene.n_lights;
    sv.arena = scene.arena;
    sv.root_ref = scene.root_ref;
    sv.nodes_shared = false;
    if constexpr (SMEM) {
        if (!(sv.root_ref & RT_REF_LEAF)) sv.root_ref = sv.nodes.a + sv.root_ref * 32u;
    } else if (args.stageNodesOnly) {
        sv.nodes_shared = true; // (the node table was staged first: it starts where the cursor stood on entry)
        if (!(sv.root_ref & RT_REF_LEAF)) sv.root_ref = smemBase + nodesAt + sv.root_ref * 32u;
    }
    sv.n_hoisted = scene.n_hoisted;
    sv.hoisted = scene.hoisted;
    return sv;
}

// 768 threads per SM (24 warps, 80 registers): measured 18 % faster than 512 x 87
// registers -- the kernel stalls on fixed-latency dependencies ("wait"), which more
// resident warps hide (profiles/README.md).
// The feature-complete instantiations need ~125 registers and stay at 512.
constexpr int MegaMaxThreads(int feat) { return feat == 0 ? 768 : 512; }
// The head/tail kernel holds no path state across rounds: with moving spheres and checker textures it
// still fits 80 registers (768 threads); the feature-complete instantiation runs at 640.
// Build options for the A/B.  The register file is per SM sub-partition (16 K registers each), so what counts is the
// fullest sub-partition: 768 threads = 6 warps each -> 80 registers; 800 or 832 put 7 warps on one -> 72 registers and
// spills: measured -6.6 % / -8 % on Book 1 (profiles/r2_ab_za.jsonl).  640 = 5 warps each -> 96 registers; 672 / 704 ->
// 80: -11 % / -14 % on the Book 2 final scene.
#ifndef RT_HT_THREADS_SMALL
#define RT_HT_THREADS_SMALL 768
#endif
#ifndef RT_HT_THREADS_LARGE
#define RT_HT_THREADS_LARGE 640
#endif
constexpr int HtMaxThreads(int feat) { return (feat & ~(RT_FEAT_MOVING | RT_FEAT_TEXTURE)) == 0 ? RT_HT_THREADS_SMALL : RT_HT_THREADS_LARGE; }

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
                    uint32_t hoistTests = 0;
                    BeginWalk<FEAT, SMEM>(sv, ray, a, slab.rcpA, 0.001f, stack, tv, args.seed, pixel, (uint32_t)sample,
                                          (uint32_t)bounce + 1u, hoistTests);
                    if (STATS) nPrim += hoistTests;
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
                        uint32_t hoistTests = 0;
                        BeginWalk<FEAT, SMEM>(sv, ray, a, slab.rcpA, 0.001f, stack, tv, args.seed, myPixel, mySB >> 8,
                                              (mySB & 0xffu) + 1u, hoistTests);
                        if (STATS) nPrim += hoistTests;
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
                uint32_t hoistTests = 0;
                BeginWalk<FEAT, SMEM>(sv, ray, a, slab.rcpA, 0.001f, stack, tv, args.seed, pixel, sample, bounce + 1u, hoistTests);
                if (STATS) nPrim += hoistTests;
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

} // namespace
