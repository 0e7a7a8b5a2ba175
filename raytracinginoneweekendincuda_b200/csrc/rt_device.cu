// rt_device.cu -- kernels and the device half of the C ABI (include/rt_abi.h).
//
// Replaces the reference's RenderInit + Render launches and the device-side
// object graph they walk (reference kernel.cu:110-154, launched :681-689).
//
// The kernels live in rt_kernels.cuh / rt_kernel_hq.cuh; this file is the host side: arena upload, launch
// configuration, resolve + readback, multi-device fan-out.
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

#include "rt_kernels.cuh"
#include "rt_kernel_hq.cuh"

namespace {

using namespace rtdev;

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
    if (variant == RT_VARIANT_HITQUEUE) {
        if (smem) return stats ? RenderHitQueue<FEAT, true, true> : RenderHitQueue<FEAT, true, false>;
        return stats ? RenderHitQueue<FEAT, false, true> : RenderHitQueue<FEAT, false, false>;
    }
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

static uint32_t Pad16(size_t b) { return (uint32_t)((b + 15) / 16 * 16); }

// Bytes a CTA stages in shared memory: every table SetupScene<true> copies, each padded to 16 (an empty table
// still holds one zeroed record).
static uint32_t StagedBytes(const rtpack::Packed& pk)
{
    return Pad16(std::max<size_t>(1, pk.nodes.size()) * sizeof(DevNode)) +
           Pad16(std::max<size_t>(1, pk.spheres.size()) * sizeof(DevSphere)) +
           Pad16(std::max<size_t>(1, pk.sphere_material.size()) * sizeof(int32_t)) +
           Pad16(std::max<size_t>(1, pk.moving.size()) * sizeof(DevMovingSphere)) +
           Pad16(std::max<size_t>(1, pk.quads.size()) * sizeof(DevQuad)) +
           Pad16(std::max<size_t>(1, pk.media.size()) * sizeof(DevMedium)) +
           Pad16(std::max<size_t>(1, pk.materials.size()) * sizeof(DevMaterial)) +
           Pad16(std::max<size_t>(1, pk.mat_params.size()) * sizeof(double));
}

int rt_scene_pack_info(const rt_scene_desc* scene, const rt_upload_options* opt, rt_pack_info* out)
{
    if (!scene || !out) {
        rt_set_error("rt_scene_pack_info: NULL argument");
        return RT_ERR_INVALID;
    }
    rt_upload_options o{};
    if (opt) o = *opt;
    try {
        rtpack::Packer pk(*scene, o);
        pk.Run();
        const rtpack::Packed& p = pk.out;
        std::memset(out, 0, sizeof *out);
        out->n_nodes = (int32_t)p.nodes.size();
        out->n_spheres = (int32_t)p.spheres.size();
        out->n_moving = (int32_t)p.moving.size();
        out->n_quads = (int32_t)p.quads.size();
        out->n_media = (int32_t)p.media.size();
        out->n_materials = (int32_t)p.materials.size();
        out->n_mat_params = (int32_t)p.mat_params.size();
        out->max_depth_bvh = p.max_depth;
        out->features = p.features;
        out->n_hoisted = p.n_hoisted;
        for (int k = 0; k < RT_MAX_HOISTED; ++k) out->hoisted[k] = p.hoisted[k];
        out->staged_bytes = StagedBytes(p);
    } catch (const std::exception& e) {
        rt_set_error("rt_scene_pack_info: %s", e.what());
        return RT_ERR_INVALID;
    }
    return RT_OK;
}

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
    const size_t oMatParams = ab.Add(packed->mat_params);
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
    h->dev.mat_params = reinterpret_cast<const double*>(h->arena + oMatParams);
    h->dev.textures = reinterpret_cast<const DevTexture*>(h->arena + oTextures);
    h->dev.perlins = reinterpret_cast<const DevPerlin*>(h->arena + oPerlins);
    h->dev.images = reinterpret_cast<const DevImage*>(h->arena + oImages);
    h->stats = reinterpret_cast<unsigned long long*>(h->arena + oStats);
    h->tileCounter = reinterpret_cast<unsigned int*>(h->arena + oTile);
    h->debugOut = reinterpret_cast<float*>(h->arena + oDebug);
    h->dev.root_ref = packed->root_ref;
    h->dev.n_hoisted = packed->n_hoisted;
    for (int k = 0; k < RT_MAX_HOISTED; ++k) h->dev.hoisted[k] = packed->hoisted[k];
    h->dev.n_nodes = (int)packed->nodes.size();
    h->dev.n_spheres = (int)packed->spheres.size();
    h->dev.n_moving = (int)packed->moving.size();
    h->dev.n_quads = (int)packed->quads.size();
    h->dev.n_media = (int)packed->media.size();
    h->dev.n_materials = (int)packed->materials.size();
    h->dev.n_textures = (int)packed->textures.size();
    h->dev.features = packed->features;

    h->stagedBytes = StagedBytes(*packed);

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
    if (p->variant < RT_VARIANT_AUTO || p->variant > RT_VARIANT_HITQUEUE) {
        rt_set_error("rt_render: unknown variant %d", p->variant);
        return RT_ERR_INVALID;
    }
    // AUTO: the hit-queue kernel unless the sample range exceeds its packed sample index.
    int variant = p->variant;
    if (variant == RT_VARIANT_AUTO)
        variant = (long long)p->sample_end - p->sample_begin <= RT_HT_MAX_SAMPLES ? RT_VARIANT_HITQUEUE : RT_VARIANT_MEGAKERNEL;
    const bool wave = variant == RT_VARIANT_WAVEFRONT;
    const bool hitQueue = variant == RT_VARIANT_HITQUEUE;
    const bool headTail = variant == RT_VARIANT_HEADTAIL || hitQueue; // the two queue kernels share their launch shape
    if (headTail && (long long)p->sample_end - p->sample_begin > RT_HT_MAX_SAMPLES) {
        rt_set_error("rt_render: the head/tail and hit-queue variants take at most %d samples per call", RT_HT_MAX_SAMPLES);
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
    a.matParamsBytes = pad16(std::max<size_t>(1, pk.mat_params.size()) * sizeof(double));

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
    const size_t warpBytes = hitQueue ? HqWarpBytes(featClass) : HtWarpBytes(featClass);
    if (headTail && p->block_threads <= 0 && !(p->flags & 0x200)) {
        // the largest block whose stacks + queues still leave room for the scene in shared memory
        // (if none does, the scene stays in global memory and the block is as large as registers allow)
        threads = maxThreads;
        for (int cand = maxThreads; cand >= 384; cand -= 128)
            if ((size_t)cand * 4 * stackLevels + (size_t)(cand / 32) * warpBytes + 16 + h->stagedBytes <=
                (size_t)h->maxSmemOptin) {
                threads = cand;
                break;
            }
    }
    const size_t stackBytes = (size_t)threads * 4 * stackLevels;
    size_t poolBytes = headTail ? (size_t)(threads / 32) * warpBytes + 16 : 0;
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
    if (n == "rt_pack_info") return (int)sizeof(rt_pack_info);
    return -1;
}

} // extern "C"
