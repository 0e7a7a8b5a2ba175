// rt_kernel_hq.cuh -- RenderHitQueue: synchronous heads + queued HITS (included by rt_device.cu only).
//
// Round 1's RenderHeadTail queued continuation RAYS: a round walked 32 rays and then shaded the ~56 % of them
// that had hit something, with the three materials' rejection loops run one after the other -- ncu put Scatter on
// 7.6 of 32 lanes (profiles/r1_RenderHeadTail_book1_source_lines.txt).  The RayColor loop (reference
// kernel.cu:71-95: Hit -> Emitted/Scatter -> next ray) is a cycle, and where the cycle is cut to park a path on
// the queue decides which half runs on a full warp.  This kernel cuts it AFTER the walk:
//
//   TAIL round  pop 32 hit records -> FinalizeHit + Scatter on 32 lanes -> walk the scattered rays (all but the
//               absorbed / light-terminated / depth-capped ones) -> the ones that hit again are pushed, the ones
//               that escape add throughput x background at once;
//   HEAD round  all 32 lanes start the SAME sample of their pixels: camera ray, walk, push the hits.
//
// A record is {hit point estimate p0 = o + t d (FP64), direction d (FP64), throughput, owner lane | bounce |
// sample, hit id}: 68 bytes.  FinalizeHit refines the hit distance with an FP64 Newton step anyway, so it starts
// from p0 with t = 0 instead of from the ray origin with the fp32 t -- one word less per record than (o, t), and for
// a medium p0 IS the scatter point.  Draw keys are (pixel, sample, bounce), so the sample set and every path are
// the ones the other kernels trace; radiance goes to per-pixel sums in shared memory in a deterministic order.
#pragma once

#include "rt_kernels.cuh"

#ifndef RT_HQ_WALK_UNROLL
#define RT_HQ_WALK_UNROLL 1 /* build option (A/B): RT_HQ_BOX_STEPS box steps + one leaf step per loop turn: 0 never, 1 small kernels, 2 all */
#endif

namespace {

template <int FEAT> constexpr bool kHqWalkUnroll = RT_HQ_WALK_UNROLL == 2 || (RT_HQ_WALK_UNROLL == 1 && !(FEAT & RT_FEAT_TEXTURE_HEAVY));

#ifndef RT_HQ_BOX_STEPS
#define RT_HQ_BOX_STEPS 3 /* box steps per leaf step of the unrolled walk loop: 2, 3 or 4.  Measured (64 spp, Grays/s; Book 1 4K /
                              scene 0 / scene 7): 2: 21.51 / 14.38 / 19.29, 3: 21.78 / 14.75 / 19.30, 4: 21.56 / 14.62 / 19.32 */
#endif

#ifndef RT_HQ_STACKED_HOIST
#define RT_HQ_STACKED_HOIST 1 /* build option (A/B): hoisted items through the walk loop's leaf step (BeginWalkStacked):
                                 0 never, 1 the feature-complete kernel, 2 every kernel */
#endif
// Measured (64 spp; Book 1 4K / scene 0 / 7 / 8 / 9, Grays/s): never 21.24 / 14.33 / 17.59 / 12.97 / 4.59; every kernel
// 20.84 / 14.33 / 17.38 / 12.87 / 4.97 (profiles/r2_ab_p.jsonl).  The feature-complete kernel shrinks from 5 344 to
// 4 048 instructions and gains 8 %; the small kernels lose 1-2 % to the extra loop turns.
// RT_HQ_BINS: the 64 record slots of a warp as TWO stacks, one from each end -- hits on spheres from the bottom, hits on
// quads / box faces / media from the top -- and a tail round pops from the fuller one first (the other tops the round
// up when it holds fewer than 32).  The hit type is in the hit id, so binning costs no load; what it buys is rounds
// whose lanes shade the same primitive type and start walks of similar length (a ray leaving the ground boxes of the
// Book 2 final scene escapes in a few steps, one inside the 1000-sphere cluster takes dozens).  Scenes with quads only.
// Build option, OFF: measured on the B200 with parity green (profiles/r2_ab_ze.jsonl, Grays/s bins vs one stack): simple
// light 10.89 vs 11.03, Cornell boxes 18.95 vs 19.40, Cornell smoke 14.35 vs 14.64, Book 2 final 5.22 vs 5.24 (1080p) and
// 6.17 vs 6.21 (4K) -- with 32 to 63 records per warp neither stack is often full enough to fill a round on its own, and
// the mixed rounds that remain pay for the second ballot.  north_star's "per-material shade queues", tried and not taken.
#ifndef RT_HQ_BINS
#define RT_HQ_BINS 0
#endif
template <int FEAT> constexpr bool kHqBins = RT_HQ_BINS != 0 && (FEAT & RT_FEAT_QUAD) != 0;

template <int FEAT>
constexpr bool kHqStackedHoist = RT_HQ_STACKED_HOIST == 2 || (RT_HQ_STACKED_HOIST == 1 && (FEAT & RT_FEAT_TEXTURE_HEAVY) != 0);

constexpr int kHqQueue = 64; // records per warp: a tail pops 32 before it pushes at most 32; a head needs 32 free
__host__ __device__ constexpr int HqWarpBytes(int feat)
{
    return kHqQueue * (6 * 8 + 3 * 4 + 4 + 4 + ((feat & RT_FEAT_MOVING) ? 4 : 0)) + 32 * 3 * 4;
}

template <int FEAT, bool SMEM, bool STATS>
__global__ void __launch_bounds__(HtMaxThreads(FEAT), 1) RenderHitQueue(const DevScene scene, const DevCamera cam, const RenderArgs args)
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
    char* wbase = smem + ((cursor + 15u) & ~15u) + (uint32_t)warp * (uint32_t)HqWarpBytes(FEAT);
    double* QP = reinterpret_cast<double*>(wbase);                       // [3][64] hit point estimate
    double* QD = QP + 3 * kHqQueue;                                      // [3][64] direction of the ray that hit
    float* QTHR = reinterpret_cast<float*>(QD + 3 * kHqQueue);           // [3][64] throughput up to the hit
    uint32_t* QMETA = reinterpret_cast<uint32_t*>(QTHR + 3 * kHqQueue);  // owner | bounce << 5 | (sample - begin) << 13
    uint32_t* QHIT = QMETA + kHqQueue;                                   // RT_HIT_* id
    float* QTIME = reinterpret_cast<float*>(QHIT + kHqQueue);            // [64], FEAT_MOVING only
    float* SUM = QTIME + ((FEAT & RT_FEAT_MOVING) ? kHqQueue : 0);       // [3][32]

    const int nTiles = args.tilesX * args.tilesY;
    const f3 background = make_f3(cam.background[0], cam.background[1], cam.background[2]);
    const uint32_t leafMask = (uint32_t)args.megaLeafMask;
    (void)leafMask;
    unsigned long long nPaths = 0, nNode = 0, nPrim = 0;

    while (true) {
        uint32_t nRays = 0; // per tile: a 64-bit running count would sit in two registers (or a spill slot) all the way
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(args.tileCounter, 1u);
        tile = __shfl_sync(FULL, tile, 0);
        if (tile >= nTiles) break;
        const int px0 = (tile % args.tilesX) * kTileW, py0 = (tile / args.tilesX) * kTileH;
        const bool valid = px0 + (lane & (kTileW - 1)) < cam.width && py0 + lane / kTileW < cam.height;
        SUM[lane] = SUM[32 + lane] = SUM[64 + lane] = 0.0f;
        __syncwarp();
        int nQ = 0;
        int nTop = 0; // kHqBins: records in the stack that grows down from slot 63 (the other one holds nQ - nTop)
        (void)nTop;
        int headSample = args.sampleBegin;

        while (headSample < args.sampleEnd || nQ > 0) {
            // TAIL when a full warp of hits waits (or no head is left); else HEAD (the queue then has room for 32)
            const bool tailRound = nQ >= 32 || headSample >= args.sampleEnd;
            Ray ray;
            f3 thr;
            uint32_t owner = (uint32_t)lane, sample = 0, bounce = 0;
            bool walk = false;
            f3 add = make_f3(0.0f, 0.0f, 0.0f);
            bool hasAdd = false;
            uint32_t pixel;
            if (tailRound) {
                const int n = min(nQ, 32);
                const bool on = lane < n;
                uint32_t hit = 0;
                int e = nQ - n + lane;
                if constexpr (kHqBins<FEAT>) {
                    // the fuller stack first; lanes beyond it take the newest records of the other one
                    const int nBot = nQ - nTop;
                    const bool topFirst = nTop > nBot;
                    const int takeTop = topFirst ? min(n, nTop) : n - min(n, nBot);
                    const int takeBot = n - takeTop;
                    const int k = topFirst ? lane : lane - takeBot; // index among the lanes that read the top stack
                    const bool fromTop = topFirst ? lane < takeTop : lane >= takeBot;
                    e = fromTop ? kHqQueue - nTop + k : nBot - takeBot + (topFirst ? lane - takeTop : lane);
                    nTop -= takeTop;
                }
                if (on) {
                    ray.o = make_d3(QP[e], QP[kHqQueue + e], QP[2 * kHqQueue + e]);
                    ray.d = make_d3(QD[e], QD[kHqQueue + e], QD[2 * kHqQueue + e]);
                    ray.time = (FEAT & RT_FEAT_MOVING) ? QTIME[e] : 0.0f;
                    thr = make_f3(QTHR[e], QTHR[kHqQueue + e], QTHR[2 * kHqQueue + e]);
                    const uint32_t meta = QMETA[e];
                    hit = QHIT[e];
                    owner = meta & 31u;
                    bounce = (meta >> 5) & 0xffu;
                    sample = (uint32_t)args.sampleBegin + (meta >> 13);
                }
                nQ -= n;
                __syncwarp(); // every record is read before any new one is written over it
                pixel = (uint32_t)((py0 + (int)(owner / kTileW)) * cam.width + px0 + (int)(owner & (kTileW - 1)));
                if (on) {
                    // shade: 32 hits, one per lane (kernel.cu:81-94)
                    const double aIn = fma(ray.d.x, ray.d.x, fma(ray.d.y, ray.d.y, ray.d.z * ray.d.z));
                    Hit h;
                    FinalizeHit<FEAT, SMEM>(sv, ray, aIn, hit, 0.0f, 0.0, h); // from p0: t = 0
                    const uint32_t type = RT_HIT_TYPE(hit);
                    if (STATS && args.debugOut && (int)pixel == args.debugPixel && (int)sample == args.debugSample) {
                        float* o = args.debugOut + bounce * 8; // o[0], o[1] (hit id, t) were written when the hit was queued
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
                    const bool scattered = Scatter<FEAT, SMEM>(sv, h, ray.d, aIn, sphereLike, rng, atten, dir, emitted);
                    if (!scattered) { // kernel.cu:82-83: emission is black unless the path ends on a light
                        add = thr * emitted;
                        hasAdd = emitted.x != 0.0f || emitted.y != 0.0f || emitted.z != 0.0f;
                    } else if ((int)bounce + 1 < cam.max_depth) { // kernel.cu:93-94, :71
                        thr = thr * atten;
                        ray.o = h.p;
                        ray.d = dir;
                        ++bounce;
                        walk = true;
                    }
                }
            } else {
                sample = (uint32_t)headSample;
                ++headSample;
                const int oi = px0 + (lane & (kTileW - 1)), oj = py0 + lane / kTileW;
                pixel = (uint32_t)(oj * cam.width + oi);
                if (valid) {
                    const StreamKey rng = MakeKey(args.seed, pixel, sample, 0u);
                    ray = CameraRay(cam, oi, oj, rng);
                    thr = make_f3(1.0f, 1.0f, 1.0f);
                    walk = true;
                    if (STATS) ++nPaths;
                }
            }

            // one ray per lane, walked to completion
            Trav tv;
            tv.Idle();
            tv.tMedium = 0.0;
            RaySlab slab;
            double a = 1.0;
            if (walk) {
                slab = MakeSlab(ray);
                a = fma(ray.d.x, ray.d.x, fma(ray.d.y, ray.d.y, ray.d.z * ray.d.z));
                if constexpr (kHqStackedHoist<FEAT>) {
                    BeginWalkStacked<SMEM>(sv, stack, tv);
                } else {
                    uint32_t hoistTests = 0;
                    BeginWalk<FEAT, SMEM>(sv, ray, a, slab.rcpA, 0.001f, stack, tv, args.seed, pixel, sample, bounce + 1u,
                                          hoistTests);
                    if (STATS) nPrim += hoistTests;
                }
                ++nRays;
            }
            // One turn of the loop = RT_HQ_BOX_STEPS (three) box steps, then one leaf step.  A lane that reaches a leaf waits
            // for the leaf step (the FP64 primitive tests then run for every lane that piled up in the box steps before); a
            // lane that is done (RT_TRAV_DONE has the leaf bit and is excluded) idles until the slowest walk ends.  Same
            // schedule as the `step & mask` leaf turn of the other kernels, without the step counter and with the loop
            // test, the reconvergence point and the branch paid once per three box steps instead of once per step.
            // Measured, two box steps against the plain loop (4K Book 1 / scene 0 / 7 / 8 / 9, 64 spp): +0.2 / +1.8 / +8.3 /
            // +0.5 / -6.9 % -- a gain wherever the kernel is small, a loss for the feature-complete instantiation (its
            // code does not fit the instruction cache as it is; measured again after it shrank: -5 %), so that one keeps
            // the plain loop.  Three steps against two: +1.3 % Book 1, +2.6 % scene 0.
            uint32_t walkTests = 0; // STATS: box tests of this lane's walk (histogram hook, rt_debug_trace_path pixel = -2)
            (void)walkTests;
            if constexpr (kHqWalkUnroll<FEAT>) {
            while (tv.ref != RT_TRAV_DONE) {
                uint32_t nodeTests = 0, primTests = 0;
                if (!(tv.ref & RT_REF_LEAF)) TraceBox<SMEM>(sv, slab, 0.001f, stack, tv, nodeTests);
                if (!(tv.ref & RT_REF_LEAF)) TraceBox<SMEM>(sv, slab, 0.001f, stack, tv, nodeTests);
#if RT_HQ_BOX_STEPS >= 3 /* (profiles/r2_ab_y.jsonl) */
                if (!(tv.ref & RT_REF_LEAF)) TraceBox<SMEM>(sv, slab, 0.001f, stack, tv, nodeTests);
#endif
#if RT_HQ_BOX_STEPS >= 4
                if (!(tv.ref & RT_REF_LEAF)) TraceBox<SMEM>(sv, slab, 0.001f, stack, tv, nodeTests);
#endif
                if ((tv.ref & RT_REF_LEAF) && tv.ref != RT_TRAV_DONE)
                    TraceLeaf<FEAT, SMEM>(sv, ray, a, slab.rcpA, 0.001f, stack, tv, args.seed, pixel, sample, bounce + 1u,
                                          primTests);
                if (STATS) {
                    nNode += nodeTests;
                    nPrim += primTests;
                    walkTests += nodeTests;
                }
            }
            } else {
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
                    walkTests += nodeTests;
                }
            }
            }
            if (STATS && args.debugOut && args.debugPixel == -2 && walk) // histogram of box-pair steps per walk: heads | tails
                atomicAdd(reinterpret_cast<unsigned int*>(args.debugOut) + (tailRound ? 64u : 0u) + min(63u, walkTests / 2u), 1u);
            if (walk && tv.hit == RT_HIT_NONE) { // kernel.cu:74-79
                add = thr * background;
                hasAdd = true;
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
            // hits on the queue
            {
                const bool push = walk && tv.hit != RT_HIT_NONE;
                const unsigned pm = __ballot_sync(FULL, push);
                int e = nQ + __popc(pm & ltMask);
                int pushedTop = 0;
                if constexpr (kHqBins<FEAT>) {
                    const uint32_t ht = RT_HIT_TYPE(tv.hit);
                    const bool toTop = push && ht != RT_LEAF_SPHERE && ht != RT_LEAF_MOVING;
                    const unsigned tm = __ballot_sync(FULL, toTop);
                    pushedTop = __popc(tm);
                    e = toTop ? kHqQueue - 1 - nTop - __popc(tm & ltMask) : (nQ - nTop) + __popc((pm & ~tm) & ltMask);
                }
                if (push) {
                    double tHit = (double)tv.t;
                    if ((FEAT & RT_FEAT_MEDIUM) && RT_HIT_TYPE(tv.hit) == RT_LEAF_MEDIUM) tHit = tv.tMedium;
                    QP[e] = fma(tHit, ray.d.x, ray.o.x);
                    QP[kHqQueue + e] = fma(tHit, ray.d.y, ray.o.y);
                    QP[2 * kHqQueue + e] = fma(tHit, ray.d.z, ray.o.z);
                    QD[e] = ray.d.x;
                    QD[kHqQueue + e] = ray.d.y;
                    QD[2 * kHqQueue + e] = ray.d.z;
                    QTHR[e] = thr.x;
                    QTHR[kHqQueue + e] = thr.y;
                    QTHR[2 * kHqQueue + e] = thr.z;
                    if (FEAT & RT_FEAT_MOVING) QTIME[e] = ray.time;
                    QMETA[e] = owner | (bounce << 5) | ((sample - (uint32_t)args.sampleBegin) << 13);
                    QHIT[e] = tv.hit;
                    if (STATS && args.debugOut && (int)pixel == args.debugPixel && (int)sample == args.debugSample) {
                        args.debugOut[bounce * 8] = __uint_as_float(tv.hit);
                        args.debugOut[bounce * 8 + 1] = tv.t;
                    }
                }
                nQ += __popc(pm);
                if constexpr (kHqBins<FEAT>) nTop += pushedTop;
                __syncwarp();
            }
        }

        if (valid) {
            float* px = args.accum + ((size_t)(py0 + lane / kTileW) * cam.width + px0 + (lane & (kTileW - 1))) * 3u;
            px[0] += SUM[lane];
            px[1] += SUM[32 + lane];
            px[2] += SUM[64 + lane];
        }
        nRays = __reduce_add_sync(FULL, nRays);
        if (lane == 0) atomicAdd(&args.stats[0], (unsigned long long)nRays);
        __syncwarp();
    }

    for (int off = 16; off > 0; off >>= 1) {
        if (STATS) {
            nPaths += __shfl_down_sync(FULL, nPaths, off);
            nNode += __shfl_down_sync(FULL, nNode, off);
            nPrim += __shfl_down_sync(FULL, nPrim, off);
        }
    }
    if (lane == 0) {
        if (STATS) {
            atomicAdd(&args.stats[1], nPaths);
            atomicAdd(&args.stats[2], nNode);
            atomicAdd(&args.stats[3], nPrim);
        }
    }
}

} // namespace
