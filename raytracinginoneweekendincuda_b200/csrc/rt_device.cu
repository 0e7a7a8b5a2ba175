// rt_device.cu -- the device half of the C ABI (include/rt_abi.h): scene arena upload, launch configuration,
// multi-device fan-out, accumulator reduction, resolve + readback, progressive output.
//
// Replaces the reference's RenderInit + Render launches and the host reads of its managed framebuffer
// (reference kernel.cu:110-154, launched :676-691, read :696-723).  The render kernels live in rt_kernels.cuh /
// rt_kernel_hq.cuh.  No CPU fallback: without a CUDA device every entry point returns an error.
//
// Multi-device: ONE host process drives N devices.  The scene is packed once and its arena -- one
// position-independent block -- is copied to every device; rt_render gives device k the k-th slice of the sample
// range on that device's own stream (the random stream is keyed on the global sample index, so the union is the
// 1-device sample set); rt_readback sums the fp32 accumulators on device 0 and resolves the frame there:
//   * peer path (default where every device can map device 0's peers -- NVLink / NVSwitch): ReduceResolveKernel
//     reads the N accumulators straight from peer memory, sums them in device order (deterministic), and applies
//     mean / gamma / quantise / row flip in the same pass -- collective and epilogue are one kernel, no staging copy;
//   * NCCL path (RT_UPLOAD_REDUCE_NCCL, or no peer access): one ncclReduce to device 0, then the same kernel.
//     libnccl.so.2 is dlopen'ed on first use, so single-device hosts carry no NCCL dependency.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
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

constexpr int kMaxDevices = 16;

// Every entry point selects the device it works on and puts the caller's current device back on the way out: a
// host that drives several GPUs itself (torch, a multi-device C++ application) must not find it changed.
struct DeviceGuard {
    int saved = -1;
    DeviceGuard() { cudaGetDevice(&saved); }
    ~DeviceGuard()
    {
        if (saved >= 0) cudaSetDevice(saved);
    }
};

// ------------------------------------------------------------------ output
// Reduce + resolve in one pass (reference kernel.cu:147-153: mean, sqrt gamma; :712-718: clamp to [0,0.999], *256;
// :699: top row first).  `set` holds the fp32 accumulators of the participating devices -- device 0's own and, on
// the peer path, the others' mapped over NVLink: they are summed in device order, so the N-device frame differs
// from the 1-device frame by fp32 summation order only, and is the same from run to run.  The sum is written back
// to device 0's accumulator when several were read (what a reduce leaves there).  Each thread moves VEC
// consecutive floats: 128-bit loads and stores on the accumulators and the linear output, a 32-bit store of four
// quantised bytes when a row is a multiple of four floats.
struct AccumSet {
    const float* p[kMaxDevices];
    int n;
};

template <int VEC>
__global__ void ReduceResolveKernel(const AccumSet set, float* __restrict__ sumOut, float* __restrict__ linearOut,
                                    uint8_t* __restrict__ srgbOut, int width, int height, float invSpp)
{
    const long long nFloats = (long long)width * height * 3;
    const long long first = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (first >= nFloats) return;
    float v[VEC];
    if (VEC == 4 && first + 4 <= nFloats) {
        float4 a = *reinterpret_cast<const float4*>(set.p[0] + first);
        for (int k = 1; k < set.n; ++k) {
            const float4 b = *reinterpret_cast<const float4*>(set.p[k] + first);
            a.x += b.x;
            a.y += b.y;
            a.z += b.z;
            a.w += b.w;
        }
        if (sumOut) *reinterpret_cast<float4*>(sumOut + first) = a;
        v[0] = a.x * invSpp;
        v[1 % VEC] = a.y * invSpp;
        v[2 % VEC] = a.z * invSpp;
        v[3 % VEC] = a.w * invSpp;
        if (linearOut) *reinterpret_cast<float4*>(linearOut + first) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
    } else {
        for (int e = 0; e < VEC; ++e) {
            if (first + e >= nFloats) break;
            float a = set.p[0][first + e];
            for (int k = 1; k < set.n; ++k) a += set.p[k][first + e];
            if (sumOut) sumOut[first + e] = a;
            v[e] = a * invSpp;
            if (linearOut) linearOut[first + e] = v[e];
        }
    }
    if (!srgbOut) return;
    const long long rowLen = (long long)width * 3;
    uint8_t q[VEC];
    for (int e = 0; e < VEC; ++e) {
        float g = sqrtf(v[e]);
        g = g < 0.0f ? 0.0f : (g > 0.999f ? 0.999f : g);
        q[e] = (uint8_t)(int)(256.0f * g);
    }
    const long long row = first / rowLen;
    if (VEC == 4 && (rowLen & 3) == 0 && first + 4 <= nFloats) {
        const long long o = (height - 1 - row) * rowLen + (first - row * rowLen);
        *reinterpret_cast<uchar4*>(srgbOut + o) = make_uchar4(q[0], q[1 % VEC], q[2 % VEC], q[3 % VEC]);
    } else {
        for (int e = 0; e < VEC; ++e) {
            const long long f = first + e;
            if (f >= nFloats) break;
            const long long r = f / rowLen;
            srgbOut[(height - 1 - r) * rowLen + (f - r * rowLen)] = q[e];
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
// Cornell-box class (quads + media, solid colours): instantiated for the shipping kernel only -- without the texture
// and moving-sphere code it is half the size of the feature-complete one.
constexpr int kFeatBox = RT_FEAT_QUAD | RT_FEAT_MEDIUM;

template <int FEAT> KernelFn PickHitQueue(bool smem, bool stats)
{
    if (smem) return stats ? RenderHitQueue<FEAT, true, true> : RenderHitQueue<FEAT, true, false>;
    return stats ? RenderHitQueue<FEAT, false, true> : RenderHitQueue<FEAT, false, false>;
}
KernelFn PickHitQueueBox(bool smem, bool stats) { return PickHitQueue<kFeatBox>(smem, stats); }

// RT_FLAG_IMPORTANCE: two more classes of the hit-queue kernel, with the light-sampling code in their Scatter.
constexpr int kFeatBoxImportance = kFeatBox | RT_FEAT_IMPORTANCE, kFeatAllImportance = kFeatAll | RT_FEAT_IMPORTANCE;

KernelFn PickKernelForFeatures(int features, int variant, bool smem, bool stats, int* picked, bool importance = false)
{
    if (importance) {
        const bool boxClass = (features & ~kFeatBox) == 0;
        *picked = boxClass ? kFeatBoxImportance : kFeatAllImportance;
        return boxClass ? PickHitQueue<kFeatBoxImportance>(smem, stats) : PickHitQueue<kFeatAllImportance>(smem, stats);
    }
    if (features == 0) {
        *picked = kFeatSpheres;
        return PickKernel<kFeatSpheres>(variant, smem, stats);
    }
    if ((features & ~kFeatMotion) == 0) {
        *picked = kFeatMotion;
        return PickKernel<kFeatMotion>(variant, smem, stats);
    }
    if ((features & ~kFeatBox) == 0 && variant == RT_VARIANT_HITQUEUE) {
        *picked = kFeatBox;
        return PickHitQueueBox(smem, stats);
    }
    *picked = kFeatAll;
    return PickKernel<kFeatAll>(variant, smem, stats);
}

// Frame-sized device buffers (accumulator, readback staging) and scene arenas are recycled across handles:
// cudaMalloc/cudaFree of ~100 MB blocks was measured at up to 0.5 s per upload/free cycle, which is what an
// application rendering frame after frame does.  The cache is stream-ordered: a block is parked together with an
// event recorded on the stream that last used it, and whoever takes it makes its own stream wait for that event,
// so a block is never handed to new work while old work on another stream still touches it.  When the cache is
// full the oldest block is returned to the driver.
struct BigBlock {
    int device;
    size_t bytes;
    void* ptr;
    cudaEvent_t idle; // may be nullptr (block known to be idle)
};
std::mutex gBigMutex;
std::vector<BigBlock> gBigCache;
constexpr size_t kBigCacheEntries = 8;

// The caller has made `device` current.
cudaError_t BigMalloc(int device, void** out, size_t bytes, cudaStream_t stream)
{
    BigBlock got{};
    bool found = false;
    {
        std::lock_guard<std::mutex> lock(gBigMutex);
        for (size_t k = 0; k < gBigCache.size(); ++k)
            if (gBigCache[k].device == device && gBigCache[k].bytes == bytes) {
                got = gBigCache[k];
                gBigCache.erase(gBigCache.begin() + (long)k);
                found = true;
                break;
            }
    }
    if (found) {
        if (got.idle) {
            const cudaError_t e = cudaStreamWaitEvent(stream, got.idle, 0);
            cudaEventDestroy(got.idle);
            if (e != cudaSuccess) {
                cudaFree(got.ptr);
                return e;
            }
        }
        *out = got.ptr;
        return cudaSuccess;
    }
    return cudaMalloc(out, bytes);
}

// The caller has made `device` current; `stream` is the stream that last used the block.
void BigFree(int device, void* ptr, size_t bytes, cudaStream_t stream)
{
    if (!ptr) return;
    BigBlock blk{device, bytes, ptr, nullptr};
    if (cudaEventCreateWithFlags(&blk.idle, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(blk.idle, stream) != cudaSuccess) {
        if (blk.idle) cudaEventDestroy(blk.idle);
        cudaStreamSynchronize(stream);
        cudaFree(ptr);
        return;
    }
    BigBlock evicted{};
    bool evict = false;
    {
        std::lock_guard<std::mutex> lock(gBigMutex);
        if (gBigCache.size() >= kBigCacheEntries) {
            evicted = gBigCache.front();
            gBigCache.erase(gBigCache.begin());
            evict = true;
        }
        gBigCache.push_back(blk);
    }
    if (evict) {
        int cur = -1;
        cudaGetDevice(&cur);
        if (cudaSetDevice(evicted.device) == cudaSuccess) {
            if (evicted.idle) {
                cudaEventSynchronize(evicted.idle);
                cudaEventDestroy(evicted.idle);
            }
            cudaFree(evicted.ptr);
        }
        if (cur >= 0) cudaSetDevice(cur);
    }
}

// Host image of the device arena: every table of the scene back to back, 256-byte aligned and free of absolute
// addresses (image texels are referenced by offset), so that one allocation and ONE host-to-device copy upload the
// scene, and the same bytes serve every device.
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

// ------------------------------------------------------------------- NCCL
// The five NCCL entry points the reduce path uses, resolved from libnccl.so.2 at first use (if the host process has
// already loaded an NCCL -- torch does -- dlopen hands back that one).  Types are spelled out here so that neither the
// build nor single-device hosts need NCCL headers or libraries.
typedef struct ncclComm* NcclComm;
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
constexpr int kNcclFloat32 = 7, kNcclSum = 0; // ncclFloat32, ncclSum (nccl.h: ncclDataType_t, ncclRedOp_t)
std::mutex gNcclMutex;
NcclApi gNccl;

bool LoadNccl()
{
    std::lock_guard<std::mutex> lock(gNcclMutex);
    if (gNccl.ok) return true;
    if (!gNccl.lib) gNccl.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!gNccl.lib) gNccl.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!gNccl.lib) {
        rt_set_error("multi-device reduce: cannot load libnccl.so.2 (%s)", dlerror());
        return false;
    }
    auto sym = [&](const char* name) { return dlsym(gNccl.lib, name); };
    gNccl.CommInitAll = reinterpret_cast<int (*)(NcclComm*, int, const int*)>(sym("ncclCommInitAll"));
    gNccl.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(sym("ncclCommDestroy"));
    gNccl.Reduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t)>(sym("ncclReduce"));
    gNccl.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
    gNccl.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
    gNccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
    gNccl.ok = gNccl.CommInitAll && gNccl.CommDestroy && gNccl.Reduce && gNccl.GroupStart && gNccl.GroupEnd && gNccl.GetErrorString;
    if (!gNccl.ok) rt_set_error("multi-device reduce: libnccl.so.2 lacks an expected entry point");
    return gNccl.ok;
}

#define RT_NCCL(call)                                                                                   \
    do {                                                                                                \
        const int r_ = (call);                                                                          \
        if (r_ != 0) {                                                                                  \
            rt_set_error("%s failed: %s (%s:%d)", #call, gNccl.GetErrorString(r_), __FILE__, __LINE__); \
            return RT_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)

uint32_t Pad16(size_t b) { return (uint32_t)((b + 15) / 16 * 16); }

// Bytes a CTA stages in shared memory: every table SetupScene<true> copies, each padded to 16 (an empty table
// still holds one zeroed record).
uint32_t StagedBytes(const rtpack::Packed& pk)
{
    return Pad16(std::max<size_t>(1, pk.nodes.size()) * sizeof(DevNode)) +
           Pad16(std::max<size_t>(1, pk.spheres.size()) * sizeof(DevSphere)) +
           Pad16(std::max<size_t>(1, pk.sphere_material.size()) * sizeof(int32_t)) +
           Pad16(std::max<size_t>(1, pk.moving.size()) * sizeof(DevMovingSphere)) +
           Pad16(std::max<size_t>(1, pk.quads.size()) * sizeof(DevQuad)) +
           Pad16(std::max<size_t>(1, pk.boxes.size()) * sizeof(DevBox)) +
           Pad16(std::max<size_t>(1, pk.media.size()) * sizeof(DevMedium)) +
           Pad16(std::max<size_t>(1, pk.materials.size()) * sizeof(DevMaterial)) +
           Pad16(std::max<size_t>(1, pk.mat_params.size()) * sizeof(double));
}


// One device's share of a handle.
struct DeviceCtx {
    int device = 0;
    DevScene dev{};
    char* arena = nullptr; // one device block: scene tables, images, counters, debug records
    float* accum = nullptr;
    size_t accumFloats = 0;
    unsigned long long* stats = nullptr; // [0] rays [1] paths [2] node tests [3] prim tests
    unsigned int* tileCounter = nullptr;
    float* debugOut = nullptr; // test hook, 256*8 floats
    cudaStream_t ownStream = nullptr;  // multi-device handles: this device's stream
    cudaStream_t lastStream = nullptr; // the stream the last render ran on
    cudaEvent_t evStart = nullptr, evStop = nullptr; // around the render launch(es)
    int smCount = 0, maxSmemOptin = 0;
    int sampleBegin = 0, sampleEnd = 0; // slice of the last render
};

} // namespace

struct rt_scene_s {
    std::vector<DeviceCtx> devs;
    rtpack::Packed* host = nullptr; // kept for sizes / info
    size_t arenaBytes = 0;
    uint32_t stagedBytes = 0;
    int uploadFlags = 0;
    // multi-device
    int reducePath = 0;      // 0 none, 1 peer-memory fused kernel, 2 NCCL
    bool peerCapable = false;
    std::vector<NcclComm> comms;
    // output staging on device 0: two sets (progressive output double-buffers them)
    float* linearStage[2] = {nullptr, nullptr};
    uint8_t* srgbStage[2] = {nullptr, nullptr};
    size_t linearStageFloats[2] = {0, 0}, srgbStageBytes[2] = {0, 0};
    float* hostLinear[2] = {nullptr, nullptr}; // pinned, progressive output
    uint8_t* hostSrgb[2] = {nullptr, nullptr};
    size_t hostLinearFloats = 0, hostSrgbBytes = 0;
    cudaStream_t copyStream = nullptr; // device 0: progressive device-to-host copies
    cudaEvent_t evReduce0 = nullptr, evReduce1 = nullptr, evResolve1 = nullptr, evReduced = nullptr;
    bool timedReadback = false;
    int debugPixel = -1, debugSample = -1;
    rt_camera lastCam{};
    bool rendered = false;
    bool fitsSmem = false, nodesInSmem = false;
    int pickedFeatures = 0, pickedVariant = 0, pickedThreads = 0, pickedRegisters = 0;
};

namespace {

void FreeDevice(rt_scene_s* h, DeviceCtx& d)
{
    if (cudaSetDevice(d.device) != cudaSuccess) return;
    cudaStream_t s = d.lastStream ? d.lastStream : d.ownStream;
    if (h->rendered) cudaStreamSynchronize(s);
    BigFree(d.device, d.arena, h->arenaBytes, s);
    BigFree(d.device, d.accum, d.accumFloats * sizeof(float), s);
    if (d.evStart) cudaEventDestroy(d.evStart);
    if (d.evStop) cudaEventDestroy(d.evStop);
    if (d.ownStream) cudaStreamDestroy(d.ownStream);
    d = DeviceCtx{};
}

// Output stage buffers on device 0 (the caller has made it current).
int EnsureStage(rt_scene_s* h, int set, size_t nFloats, bool linear, bool srgb, cudaStream_t stream)
{
    DeviceCtx& d0 = h->devs[0];
    if (linear && h->linearStageFloats[set] < nFloats) {
        BigFree(d0.device, h->linearStage[set], h->linearStageFloats[set] * sizeof(float), stream);
        h->linearStage[set] = nullptr;
        h->linearStageFloats[set] = 0;
        RT_CUDA(BigMalloc(d0.device, reinterpret_cast<void**>(&h->linearStage[set]), nFloats * sizeof(float), stream));
        h->linearStageFloats[set] = nFloats;
    }
    if (srgb && h->srgbStageBytes[set] < nFloats) {
        BigFree(d0.device, h->srgbStage[set], h->srgbStageBytes[set], stream);
        h->srgbStage[set] = nullptr;
        h->srgbStageBytes[set] = 0;
        RT_CUDA(BigMalloc(d0.device, reinterpret_cast<void**>(&h->srgbStage[set]), nFloats, stream));
        h->srgbStageBytes[set] = nFloats;
    }
    return RT_OK;
}

// Queues, on device 0's stream: [multi-device: the reduction of all accumulators onto device 0] + the resolve of the
// frame into stage set `set`.  `src` = a caller-owned accumulator (single device) or nullptr.  Leaves device 0 current.
int QueueReduceResolve(rt_scene_s* h, const float* src, int set, bool linear, bool srgb, float invSpp)
{
    const int nDev = (int)h->devs.size();
    DeviceCtx& d0 = h->devs[0];
    const int W = h->lastCam.image_width, H = h->lastCam.image_height;
    const size_t nFloats = (size_t)W * H * 3;
    RT_CUDA(cudaSetDevice(d0.device));
    cudaStream_t s0 = d0.lastStream;
    AccumSet acc{};
    acc.n = 1;
    acc.p[0] = src ? src : d0.accum;
    float* sumOut = nullptr;
    const bool multi = nDev > 1 && !src;
    if (h->timedReadback) RT_CUDA(cudaEventRecord(h->evReduce0, s0));
    if (multi) {
        // device 0's stream waits for every device's render
        for (int k = 1; k < nDev; ++k) RT_CUDA(cudaStreamWaitEvent(s0, h->devs[k].evStop, 0));
        if (h->reducePath == 0) h->reducePath = (h->peerCapable && !(h->uploadFlags & RT_UPLOAD_REDUCE_NCCL)) ? 1 : 2;
        if (h->reducePath == 1) {
            acc.n = nDev;
            for (int k = 1; k < nDev; ++k) acc.p[k] = h->devs[k].accum; // peer-mapped: read over NVLink by the kernel
            sumOut = d0.accum;
        } else {
            if (!LoadNccl()) return RT_ERR_UNSUPPORTED;
            if (h->comms.empty()) {
                std::vector<int> ids;
                for (const DeviceCtx& d : h->devs) ids.push_back(d.device);
                h->comms.assign((size_t)nDev, nullptr);
                RT_NCCL(gNccl.CommInitAll(h->comms.data(), nDev, ids.data()));
            }
            // every rank's collective runs on that device's render stream, after its kernel
            RT_NCCL(gNccl.GroupStart());
            for (int k = 0; k < nDev; ++k) {
                DeviceCtx& d = h->devs[k];
                RT_CUDA(cudaSetDevice(d.device));
                RT_NCCL(gNccl.Reduce(d.accum, k == 0 ? d.accum : nullptr, nFloats, kNcclFloat32, kNcclSum, 0, h->comms[(size_t)k],
                                     d.lastStream));
            }
            RT_NCCL(gNccl.GroupEnd());
            RT_CUDA(cudaSetDevice(d0.device));
        }
    }
    if (h->timedReadback && h->reducePath == 2) RT_CUDA(cudaEventRecord(h->evReduce1, s0));
    if (linear || srgb || sumOut) {
        const int rc = EnsureStage(h, set, nFloats, linear, srgb, s0);
        if (rc != RT_OK) return rc;
        float* lin = linear ? h->linearStage[set] : nullptr;
        uint8_t* s8 = srgb ? h->srgbStage[set] : nullptr;
        bool aligned = (reinterpret_cast<uintptr_t>(lin) & 15u) == 0;
        for (int k = 0; k < acc.n; ++k) aligned = aligned && (reinterpret_cast<uintptr_t>(acc.p[k]) & 15u) == 0;
        if (aligned) {
            const long long threads = ((long long)nFloats + 3) / 4;
            ReduceResolveKernel<4><<<(unsigned)((threads + 255) / 256), 256, 0, s0>>>(acc, sumOut, lin, s8, W, H, invSpp);
        } else {
            ReduceResolveKernel<1><<<(unsigned)((nFloats + 255) / 256), 256, 0, s0>>>(acc, sumOut, lin, s8, W, H, invSpp);
        }
        RT_CUDA(cudaGetLastError());
    }
    if (h->timedReadback && h->reducePath != 2) RT_CUDA(cudaEventRecord(h->evReduce1, s0));
    if (multi) {
        // device 0 now holds the total: the partial sums of the others are spent.  Clearing them (after the reduce has
        // read them) keeps "render more samples, reduce again" correct.
        cudaEvent_t reduced = h->evReduced;
        RT_CUDA(cudaEventRecord(reduced, s0));
        for (int k = 1; k < nDev; ++k) {
            DeviceCtx& d = h->devs[k];
            RT_CUDA(cudaSetDevice(d.device));
            RT_CUDA(cudaStreamWaitEvent(d.lastStream, reduced, 0));
            RT_CUDA(cudaMemsetAsync(d.accum, 0, nFloats * sizeof(float), d.lastStream));
        }
        RT_CUDA(cudaSetDevice(d0.device));
    }
    return RT_OK;
}

// Launch configuration + launch of samples [begin, end) on one device (made current by the caller).
int LaunchOn(rt_scene_s* h, DeviceCtx& d, const rt_camera* cam, const rt_render_params* p, int begin, int end, float* accum,
             cudaStream_t stream, bool clear)
{
    const size_t nFloats = (size_t)cam->image_width * cam->image_height * 3;
    if (clear) {
        RT_CUDA(cudaMemsetAsync(accum, 0, nFloats * sizeof(float), stream));
        RT_CUDA(cudaMemsetAsync(d.stats, 0, 4 * sizeof(unsigned long long), stream));
    }
    RT_CUDA(cudaMemsetAsync(d.tileCounter, 0, sizeof(unsigned int), stream));

    int variant = p->variant;
    const bool importance = (p->flags & RT_FLAG_IMPORTANCE) != 0;
    if (variant == RT_VARIANT_AUTO)
        variant = ((long long)end - begin <= RT_HT_MAX_SAMPLES || importance) ? RT_VARIANT_HITQUEUE : RT_VARIANT_MEGAKERNEL;
    if (importance && variant != RT_VARIANT_HITQUEUE) {
        rt_set_error("rt_render: RT_FLAG_IMPORTANCE is implemented by the hit-queue kernel (RT_VARIANT_AUTO / _HITQUEUE)");
        return RT_ERR_UNSUPPORTED;
    }
    if (importance && (long long)end - begin > RT_HT_MAX_SAMPLES) {
        rt_set_error("rt_render: RT_FLAG_IMPORTANCE takes at most %d samples per device and call", RT_HT_MAX_SAMPLES);
        return RT_ERR_INVALID;
    }
    const bool wave = variant == RT_VARIANT_WAVEFRONT;
    const bool hitQueue = variant == RT_VARIANT_HITQUEUE;
    const bool queued = variant == RT_VARIANT_HEADTAIL || hitQueue; // the two queue kernels share their launch shape

    const DevCamera dc = rtpack::MakeCamera(*cam);
    RenderArgs a{};
    a.accum = accum;
    a.stats = d.stats;
    a.tileCounter = d.tileCounter;
    a.sampleBegin = begin;
    a.sampleEnd = end;
    a.seed = p->seed;
    a.debugPixel = h->debugPixel;
    a.debugSample = h->debugSample;
    a.debugOut = (h->debugPixel >= 0 || h->debugPixel == -2) ? d.debugOut : nullptr; // -2: walk-length histogram
    // leaf turn every 2nd step (measured best: 1 -> 11.3, 2 -> 11.5, 4 -> 11.2 Grays/s in round 1; again in round 2);
    // development knob in flags bits 4-5: 1 = every step, 2 = every 4th, 3 = every 8th
    static const int kLeafMasks[4] = {1, 0, 3, 7};
    a.megaLeafMask = kLeafMasks[(p->flags >> 4) & 3];
    // tuning knobs of the wavefront variant (development): flags bits 12-15 slots/32, 16-20 idle-exit, 21-25 leaf
    // batch, 26-30 refill minimum
    int waveSlots = ((p->flags >> 12) & 0xf) * 32;
    a.waveIdleExit = (p->flags >> 16) & 0x1f;
    a.waveLeafBatch = (p->flags >> 21) & 0x1f;
    a.waveRefillMin = (p->flags >> 26) & 0x1f;
    if (a.waveIdleExit <= 0) a.waveIdleExit = 16;
    if (a.waveLeafBatch <= 0) a.waveLeafBatch = 16;
    if (a.waveRefillMin <= 0) a.waveRefillMin = 8;
    const rtpack::Packed& pk = *h->host;
    a.nodesBytes = Pad16(std::max<size_t>(1, pk.nodes.size()) * sizeof(DevNode));
    a.spheresBytes = Pad16(std::max<size_t>(1, pk.spheres.size()) * sizeof(DevSphere));
    a.sphereMatBytes = Pad16(std::max<size_t>(1, pk.sphere_material.size()) * sizeof(int32_t));
    a.movingBytes = Pad16(std::max<size_t>(1, pk.moving.size()) * sizeof(DevMovingSphere));
    a.quadsBytes = Pad16(std::max<size_t>(1, pk.quads.size()) * sizeof(DevQuad));
    a.boxesBytes = Pad16(std::max<size_t>(1, pk.boxes.size()) * sizeof(DevBox));
    a.mediaBytes = Pad16(std::max<size_t>(1, pk.media.size()) * sizeof(DevMedium));
    a.materialsBytes = Pad16(std::max<size_t>(1, pk.materials.size()) * sizeof(DevMaterial));
    a.matParamsBytes = Pad16(std::max<size_t>(1, pk.mat_params.size()) * sizeof(double));

    int featClass = d.dev.features == 0 ? kFeatSpheres : ((d.dev.features & ~kFeatMotion) == 0 ? kFeatMotion : kFeatAll);
    if (hitQueue && featClass == kFeatAll && (d.dev.features & ~kFeatBox) == 0) featClass = kFeatBox;
    if (importance) featClass = (d.dev.features & ~kFeatBox) == 0 ? kFeatBoxImportance : kFeatAllImportance;
    const int maxThreads = wave ? 512 : (queued ? HtMaxThreads(featClass) : MegaMaxThreads(featClass));
    int threads = p->block_threads > 0 ? p->block_threads : maxThreads;
    threads = std::max(32, std::min(maxThreads, (threads / 32) * 32));
    int blocksPerSm = p->blocks_per_sm > 0 ? p->blocks_per_sm : 1;
    // + sentinel slot; at least the hoisted refs + root + sentinel (BeginWalkStacked parks them there before the tree)
    const int stackLevels = std::max(RT_MAX_HOISTED + 2, std::min(kMaxStackLevels, h->host->max_depth + 3));
    a.stackLevels = stackLevels;
    if (wave) blocksPerSm = 1;
    const size_t warpBytes = hitQueue ? HqWarpBytes(featClass) : HtWarpBytes(featClass);
    if (queued && p->block_threads <= 0 && !(p->flags & RT_FLAG_SCENE_IN_GLOBAL)) {
        // the largest block whose stacks + queues still leave room for the scene in shared memory
        // (if none does, the scene stays in global memory and the block is as large as registers allow)
        threads = maxThreads;
        for (int cand = maxThreads; cand >= 384; cand -= 32)
            if ((size_t)cand * 4 * stackLevels + (size_t)(cand / 32) * warpBytes + 16 + h->stagedBytes <= (size_t)d.maxSmemOptin) {
                threads = cand;
                break;
            }
    }
    // a scene that does not fit as a whole: the node table alone, if that fits beside a full-size block
    bool nodesOnly = false;
    if (hitQueue && !(p->flags & (RT_FLAG_SCENE_IN_GLOBAL | RT_FLAG_NODES_IN_GLOBAL)) &&
        (size_t)threads * 4 * stackLevels + (size_t)(threads / 32) * warpBytes + 16 + h->stagedBytes > (size_t)d.maxSmemOptin) {
        for (int cand = p->block_threads > 0 ? threads : maxThreads; cand >= std::max(512, maxThreads - 128); cand -= 32) {
            if ((size_t)cand * 4 * stackLevels + (size_t)(cand / 32) * warpBytes + 32 + a.nodesBytes <= (size_t)d.maxSmemOptin) {
                threads = cand;
                nodesOnly = true;
                break;
            }
            if (p->block_threads > 0) break;
        }
    }
    a.stageNodesOnly = nodesOnly ? 1 : 0;
    const size_t stackBytes = (size_t)threads * 4 * stackLevels;
    size_t poolBytes = queued ? (size_t)(threads / 32) * warpBytes + 16 : 0;
    if (wave) {
        // 64 slots (an 8x8 tile) measured best: 96 or 128 fill SHADE/GEN chunks better but cost more shared
        // memory traffic and longer tile tails (profiles/r1_wavefront_parameter_sweep.json)
        if (waveSlots < 64 || waveSlots > 128) waveSlots = 64;
        poolBytes = (size_t)(threads / 32) * PoolBytes(waveSlots) + 16;
    }
    a.waveSlots = waveSlots;
    const int tileH = wave ? waveSlots / 8 : kTileH;
    a.tilesX = (cam->image_width + kTileW - 1) / kTileW;
    a.tilesY = (cam->image_height + tileH - 1) / tileH;
    const bool wantStats = (p->flags & RT_FLAG_STATS) != 0 || a.debugOut != nullptr;
    const bool smem = !(p->flags & RT_FLAG_SCENE_IN_GLOBAL) &&
                      stackBytes + poolBytes + h->stagedBytes <= (size_t)d.maxSmemOptin / (size_t)blocksPerSm;
    size_t smemBytes = stackBytes + poolBytes + (smem ? h->stagedBytes : (nodesOnly ? a.nodesBytes + 16 : 0));
    // Perlin table 0 beside it when the kernel has noise textures and there is room.  Where the
    // primitives stay in global memory, stacks, queues and the node table leave L1 ~28 KB, which the 5 KB table (224 loads
    // per marble hit) shared with every primitive, material and texel fetch: measured on the Book 2 final scene 5.23 ->
    // 5.57 Grays/s at 1920x1080, 6.19 -> 6.78 at 4K.  Small scenes: 13.3 vs 13.0 (scene 3), 11.1 vs 10.8 (scene 5) against
    // the same generic loads from global memory (profiles/r2_ab_zf.jsonl).
    if (hitQueue && (d.dev.features & RT_FEAT_TEXTURE_HEAVY) && !pk.perlins.empty() && !(p->flags & RT_FLAG_SCENE_IN_GLOBAL) &&
        smemBytes + sizeof(DevPerlin) + 16 <= (size_t)d.maxSmemOptin / (size_t)blocksPerSm) {
        a.perlinBytes = (uint32_t)sizeof(DevPerlin);
        smemBytes += sizeof(DevPerlin) + 16;
    }
    if (smemBytes > (size_t)d.maxSmemOptin) {
        rt_set_error("rt_render: block of %d threads needs %zu B of shared memory (max %d)", threads, smemBytes, d.maxSmemOptin);
        return RT_ERR_INVALID;
    }
    KernelFn fn = PickKernelForFeatures(d.dev.features, variant, smem, wantStats, &h->pickedFeatures, importance);
    h->pickedVariant = variant;
    h->pickedThreads = threads;
    RT_CUDA(cudaFuncSetAttribute(reinterpret_cast<const void*>(fn), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    if (&d == &h->devs[0]) {
        cudaFuncAttributes fa{};
        if (cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(fn)) == cudaSuccess) h->pickedRegisters = fa.numRegs;
    }
    const int nTiles = a.tilesX * a.tilesY;
    const int warpsPerBlock = threads / 32;
    int blocks = d.smCount * blocksPerSm;
    blocks = std::max(1, std::min(blocks, (nTiles + warpsPerBlock - 1) / warpsPerBlock));
    h->fitsSmem = smem;
    h->nodesInSmem = smem || nodesOnly;
    if (end > begin) {
        fn<<<blocks, threads, smemBytes, stream>>>(d.dev, dc, a);
        RT_CUDA(cudaGetLastError());
    }
    return RT_OK;
}

int CheckRenderArgs(rt_scene_handle h, const rt_camera* cam, const rt_render_params* p, const char* who)
{
    if (!h || !cam || !p) {
        rt_set_error("%s: NULL argument", who);
        return RT_ERR_INVALID;
    }
    if (cam->image_width <= 0 || cam->image_height <= 0 || cam->max_depth <= 0 || cam->max_depth > 254 ||
        p->sample_end < p->sample_begin || p->sample_begin < 0) {
        rt_set_error("%s: bad camera or sample range (max_depth must be 1..254)", who);
        return RT_ERR_INVALID;
    }
    if (cam->samples_per_pixel <= 0) {
        rt_set_error("%s: samples_per_pixel must be positive (rt_readback divides by it)", who);
        return RT_ERR_INVALID;
    }
    if ((long long)cam->image_width * cam->image_height > 0x7fffffffLL / 3) {
        rt_set_error("%s: image too large", who);
        return RT_ERR_INVALID;
    }
    if (p->variant < RT_VARIANT_AUTO || p->variant > RT_VARIANT_HITQUEUE) {
        rt_set_error("%s: unknown variant %d", who, p->variant);
        return RT_ERR_INVALID;
    }
    const long long n = (long long)p->sample_end - p->sample_begin;
    const long long perDevice = (n + (long long)h->devs.size() - 1) / (long long)h->devs.size();
    if ((p->variant == RT_VARIANT_HEADTAIL || p->variant == RT_VARIANT_HITQUEUE) && perDevice > RT_HT_MAX_SAMPLES) {
        rt_set_error("%s: the head/tail and hit-queue variants take at most %d samples per device and call", who, RT_HT_MAX_SAMPLES);
        return RT_ERR_INVALID;
    }
    if (p->variant == RT_VARIANT_WAVEFRONT && p->sample_end >= (1 << 24)) { // its slots pack sample << 8 | bounce
        rt_set_error("%s: the wavefront variant takes sample indices below 2^24", who);
        return RT_ERR_INVALID;
    }
    if (h->devs.size() > 1 && (p->stream || p->accum)) {
        rt_set_error("%s: a multi-device handle renders on its own streams into its own accumulators "
                     "(rt_render_params.stream / .accum must be NULL)", who);
        return RT_ERR_INVALID;
    }
    return RT_OK;
}

} // namespace

extern "C" {

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
        out->n_boxes = (int32_t)p.boxes.size();
        out->n_lights = (int32_t)p.lights.size();
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
    rt_scene_s* h = new rt_scene_s();
    h->host = packed; // from here on rt_scene_free(h) releases everything, on every path out
    h->uploadFlags = o.flags;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev <= 0) {
        rt_scene_free(h);
        rt_set_error("rt_scene_upload: no CUDA device available (this library has no CPU path)");
        return RT_ERR_NO_DEVICE;
    }
    std::vector<int> ids;
    if (o.n_devices > 1) {
        if (o.n_devices > kMaxDevices) {
            rt_scene_free(h);
            rt_set_error("rt_scene_upload: at most %d devices", kMaxDevices);
            return RT_ERR_INVALID;
        }
        for (int k = 0; k < o.n_devices; ++k) ids.push_back(o.device_ids ? o.device_ids[k] : k);
    } else {
        ids.push_back(o.n_devices == 1 && o.device_ids ? o.device_ids[0] : o.device);
    }
    for (size_t k = 0; k < ids.size(); ++k) {
        bool dup = false;
        for (size_t j = 0; j < k; ++j) dup = dup || ids[j] == ids[k];
        if (ids[k] < 0 || ids[k] >= nDev || dup) {
            rt_scene_free(h);
            rt_set_error("rt_scene_upload: device %d out of range or listed twice (have %d)", ids[k], nDev);
            return RT_ERR_INVALID;
        }
    }
    DeviceGuard guard;

    // the arena: packed ONCE, the same bytes for every device
    ArenaBuilder ab;
    const size_t oNodes = ab.Add(packed->nodes), oSpheres = ab.Add(packed->spheres);
    const size_t oSphereMat = ab.Add(packed->sphere_material), oMoving = ab.Add(packed->moving);
    const size_t oQuads = ab.Add(packed->quads), oBoxes = ab.Add(packed->boxes);
    const size_t oMedia = ab.Add(packed->media), oMaterials = ab.Add(packed->materials);
    const size_t oMatParams = ab.Add(packed->mat_params);
    const size_t oTextures = ab.Add(packed->textures), oPerlins = ab.Add(packed->perlins);
    const size_t oUvFrames = ab.Add(packed->uv_frames);
    const size_t oLights = ab.Add(packed->lights);
    const size_t oHoisted = ab.Add(std::vector<uint32_t>(packed->hoisted, packed->hoisted + RT_MAX_HOISTED));
    std::vector<DevImage> images(std::max<size_t>(1, packed->image_bytes.size()));
    std::memset(images.data(), 0, images.size() * sizeof(DevImage));
    for (size_t k = 0; k < packed->image_bytes.size(); ++k)
        if (!packed->image_bytes[k].empty()) {
            images[k].offset = (uint32_t)ab.Add(packed->image_bytes[k]);
            images[k].width = packed->image_w[k];
            images[k].height = packed->image_h[k];
        }
    const size_t oImages = ab.Add(images);
    const size_t oStats = ab.Reserve(4 * sizeof(unsigned long long)), oTile = ab.Reserve(sizeof(unsigned int));
    const size_t oDebug = ab.Reserve(256 * 8 * sizeof(float));
    h->arenaBytes = (ab.bytes.size() + 255) / 256 * 256;
    ab.bytes.resize(h->arenaBytes, 0);
    if (h->arenaBytes >= (1ull << 32)) {
        rt_scene_free(h);
        rt_set_error("rt_scene_upload: scene arena of %zu bytes exceeds the 4 GiB the texel offsets address", h->arenaBytes);
        return RT_ERR_UNSUPPORTED;
    }
    h->stagedBytes = StagedBytes(*packed);

    h->devs.resize(ids.size());
    for (size_t k = 0; k < ids.size(); ++k) {
        DeviceCtx& d = h->devs[k];
        d.device = ids[k];
        cudaError_t e = cudaSetDevice(d.device);
        // two attributes, not cudaGetDeviceProperties: the full query costs tens of ms per call
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.smCount, cudaDevAttrMultiProcessorCount, d.device);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d.maxSmemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.device);
        if (e == cudaSuccess && ids.size() > 1) e = cudaStreamCreateWithFlags(&d.ownStream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&d.evStart);
        if (e == cudaSuccess) e = cudaEventCreate(&d.evStop);
        void* p = nullptr;
        if (e == cudaSuccess) e = BigMalloc(d.device, &p, h->arenaBytes, d.ownStream);
        d.arena = static_cast<char*>(p);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d.arena, ab.bytes.data(), h->arenaBytes, cudaMemcpyHostToDevice, d.ownStream);
        if (e != cudaSuccess) {
            rt_set_error("rt_scene_upload: device %d: %s", d.device, cudaGetErrorString(e));
            rt_scene_free(h);
            return RT_ERR_CUDA;
        }
        d.lastStream = d.ownStream;
        d.dev.nodes = reinterpret_cast<const DevNode*>(d.arena + oNodes);
        d.dev.spheres = reinterpret_cast<const DevSphere*>(d.arena + oSpheres);
        d.dev.sphere_material = reinterpret_cast<const int32_t*>(d.arena + oSphereMat);
        d.dev.moving = reinterpret_cast<const DevMovingSphere*>(d.arena + oMoving);
        d.dev.quads = reinterpret_cast<const DevQuad*>(d.arena + oQuads);
        d.dev.boxes = reinterpret_cast<const DevBox*>(d.arena + oBoxes);
        d.dev.media = reinterpret_cast<const DevMedium*>(d.arena + oMedia);
        d.dev.materials = reinterpret_cast<const DevMaterial*>(d.arena + oMaterials);
        d.dev.mat_params = reinterpret_cast<const double*>(d.arena + oMatParams);
        d.dev.textures = reinterpret_cast<const DevTexture*>(d.arena + oTextures);
        d.dev.perlins = reinterpret_cast<const DevPerlin*>(d.arena + oPerlins);
        d.dev.images = reinterpret_cast<const DevImage*>(d.arena + oImages);
        d.dev.uv_frames = reinterpret_cast<const DevUvFrame*>(d.arena + oUvFrames);
        d.dev.lights = reinterpret_cast<const DevLight*>(d.arena + oLights);
        d.dev.n_lights = (int)packed->lights.size();
        d.dev.arena = reinterpret_cast<const uint8_t*>(d.arena);
        d.stats = reinterpret_cast<unsigned long long*>(d.arena + oStats);
        d.tileCounter = reinterpret_cast<unsigned int*>(d.arena + oTile);
        d.debugOut = reinterpret_cast<float*>(d.arena + oDebug);
        d.dev.root_ref = packed->root_ref;
        d.dev.n_hoisted = packed->n_hoisted;
        d.dev.hoisted = reinterpret_cast<const uint32_t*>(d.arena + oHoisted);
        d.dev.n_nodes = (int)packed->nodes.size();
        d.dev.n_spheres = (int)packed->spheres.size();
        d.dev.n_moving = (int)packed->moving.size();
        d.dev.n_quads = (int)packed->quads.size();
        d.dev.n_boxes = (int)packed->boxes.size();
        d.dev.n_media = (int)packed->media.size();
        d.dev.n_materials = (int)packed->materials.size();
        d.dev.n_textures = (int)packed->textures.size();
        d.dev.features = packed->features;
    }
    // the pageable host arena dies with this call: the copies must have left it
    for (DeviceCtx& d : h->devs) {
        cudaSetDevice(d.device);
        const cudaError_t e = cudaStreamSynchronize(d.ownStream);
        if (e != cudaSuccess) {
            rt_set_error("rt_scene_upload: device %d: %s", d.device, cudaGetErrorString(e));
            rt_scene_free(h);
            return RT_ERR_CUDA;
        }
    }
    if (h->devs.size() > 1) {
        // peer path: device 0 maps the others' memory (NVLink / NVSwitch on a B200 box)
        h->peerCapable = true;
        cudaSetDevice(h->devs[0].device);
        for (size_t k = 1; k < h->devs.size(); ++k) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, h->devs[0].device, h->devs[k].device) != cudaSuccess || !can) {
                h->peerCapable = false;
                break;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(h->devs[k].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) h->peerCapable = false;
            cudaGetLastError(); // (clears "already enabled")
        }
    }
    {
        cudaSetDevice(h->devs[0].device);
        cudaError_t e = cudaEventCreate(&h->evReduce0);
        if (e == cudaSuccess) e = cudaEventCreate(&h->evReduce1);
        if (e == cudaSuccess) e = cudaEventCreate(&h->evResolve1);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->evReduced, cudaEventDisableTiming);
        if (e != cudaSuccess) {
            rt_set_error("rt_scene_upload: %s", cudaGetErrorString(e));
            rt_scene_free(h);
            return RT_ERR_CUDA;
        }
    }
    *out = h;
    return RT_OK;
}

int rt_render(rt_scene_handle h, const rt_camera* cam, const rt_render_params* p)
{
    const int rc0 = CheckRenderArgs(h, cam, p, "rt_render");
    if (rc0 != RT_OK) return rc0;
    DeviceGuard guard;
    const int nDev = (int)h->devs.size();
    const size_t nFloats = (size_t)cam->image_width * cam->image_height * 3;
    const long long n = (long long)p->sample_end - p->sample_begin;
    for (int k = 0; k < nDev; ++k) {
        DeviceCtx& d = h->devs[k];
        RT_CUDA(cudaSetDevice(d.device));
        cudaStream_t stream = nDev > 1 ? d.ownStream : reinterpret_cast<cudaStream_t>(p->stream);
        float* accum = p->accum;
        bool fresh = false;
        if (!accum) {
            if (d.accumFloats != nFloats) {
                BigFree(d.device, d.accum, d.accumFloats * sizeof(float), d.lastStream);
                d.accum = nullptr;
                d.accumFloats = 0;
                RT_CUDA(BigMalloc(d.device, reinterpret_cast<void**>(&d.accum), nFloats * sizeof(float), stream));
                d.accumFloats = nFloats;
                fresh = true;
            }
            accum = d.accum;
        }
        // device k of G renders the k-th slice of the sample range (bench.py / multigpu.py use the same split)
        const int begin = p->sample_begin + (int)((n * k) / nDev), end = p->sample_begin + (int)((n * (k + 1)) / nDev);
        d.sampleBegin = begin;
        d.sampleEnd = end;
        RT_CUDA(cudaEventRecord(d.evStart, stream));
        const int rc = LaunchOn(h, d, cam, p, begin, end, accum, stream, p->clear != 0 || fresh);
        if (rc != RT_OK) return rc;
        RT_CUDA(cudaEventRecord(d.evStop, stream));
        d.lastStream = stream;
    }
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
    *dev_ptr = h->devs[0].accum;
    if (n_floats) *n_floats = h->devs[0].accumFloats;
    return h->devs[0].accum ? RT_OK : RT_ERR_STATE;
}

int rt_sync(rt_scene_handle h)
{
    if (!h) {
        rt_set_error("rt_sync: NULL handle");
        return RT_ERR_INVALID;
    }
    DeviceGuard guard;
    for (DeviceCtx& d : h->devs) {
        RT_CUDA(cudaSetDevice(d.device));
        RT_CUDA(cudaStreamSynchronize(d.lastStream));
    }
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
    DeviceGuard guard;
    DeviceCtx& d0 = h->devs[0];
    const int nDev = (int)h->devs.size();
    const int W = h->lastCam.image_width, H = h->lastCam.image_height;
    const size_t nFloats = (size_t)W * H * 3;
    if (accum && nDev > 1) {
        rt_set_error("rt_readback: a multi-device handle reduces its own accumulators (`accum` must be NULL)");
        return RT_ERR_INVALID;
    }
    if (!accum && !d0.accum && (linear_rgb || srgb8)) {
        rt_set_error("rt_readback: no accumulator (the render used a caller-owned one: pass it as `accum`)");
        return RT_ERR_STATE;
    }
    RT_CUDA(cudaSetDevice(d0.device));
    cudaStream_t s0 = d0.lastStream;
    h->timedReadback = true;
    const bool needFrame = linear_rgb || srgb8;
    if (needFrame || (nDev > 1 && d0.accum)) {
        const float invSpp = 1.0f / (float)h->lastCam.samples_per_pixel;
        const int rc = QueueReduceResolve(h, accum, 0, linear_rgb != nullptr, srgb8 != nullptr, invSpp);
        if (rc != RT_OK) return rc;
        if (linear_rgb)
            RT_CUDA(cudaMemcpyAsync(linear_rgb, h->linearStage[0], nFloats * sizeof(float), cudaMemcpyDeviceToHost, s0));
        if (srgb8) RT_CUDA(cudaMemcpyAsync(srgb8, h->srgbStage[0], nFloats, cudaMemcpyDeviceToHost, s0));
        RT_CUDA(cudaEventRecord(h->evResolve1, s0));
    } else {
        h->timedReadback = false;
    }
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (DeviceCtx& d : h->devs) {
            unsigned long long s[4];
            RT_CUDA(cudaSetDevice(d.device));
            RT_CUDA(cudaMemcpyAsync(s, d.stats, sizeof s, cudaMemcpyDeviceToHost, d.lastStream));
            RT_CUDA(cudaStreamSynchronize(d.lastStream));
            stats->rays += s[0];
            stats->paths += s[1];
            stats->node_tests += s[2];
            stats->prim_tests += s[3];
        }
    }
    for (DeviceCtx& d : h->devs) {
        RT_CUDA(cudaSetDevice(d.device));
        RT_CUDA(cudaStreamSynchronize(d.lastStream));
    }
    return RT_OK;
}

int rt_render_progressive(rt_scene_handle h, const rt_camera* cam, const rt_render_params* p, int32_t batch,
                          int32_t want_linear, int32_t want_srgb8, rt_progress_fn fn, void* user)
{
    const int rc0 = CheckRenderArgs(h, cam, p, "rt_render_progressive");
    if (rc0 != RT_OK) return rc0;
    if (batch <= 0 || !fn || (!want_linear && !want_srgb8) || p->accum) {
        rt_set_error("rt_render_progressive: needs batch > 0, a callback, an output format, and the handle's own accumulator");
        return RT_ERR_INVALID;
    }
    if (!p->clear) {
        rt_set_error("rt_render_progressive: the mean is taken over the samples of this call: rt_render_params.clear must be 1");
        return RT_ERR_INVALID;
    }
    DeviceGuard guard;
    DeviceCtx& d0 = h->devs[0];
    const size_t nFloats = (size_t)cam->image_width * cam->image_height * 3;
    RT_CUDA(cudaSetDevice(d0.device));
    if (!h->copyStream) RT_CUDA(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
    if (want_linear && h->hostLinearFloats < nFloats) {
        for (int k = 0; k < 2; ++k) {
            if (h->hostLinear[k]) cudaFreeHost(h->hostLinear[k]);
            h->hostLinear[k] = nullptr;
        }
        h->hostLinearFloats = 0;
        for (int k = 0; k < 2; ++k) RT_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->hostLinear[k]), nFloats * sizeof(float), cudaHostAllocDefault));
        h->hostLinearFloats = nFloats;
    }
    if (want_srgb8 && h->hostSrgbBytes < nFloats) {
        for (int k = 0; k < 2; ++k) {
            if (h->hostSrgb[k]) cudaFreeHost(h->hostSrgb[k]);
            h->hostSrgb[k] = nullptr;
        }
        h->hostSrgbBytes = 0;
        for (int k = 0; k < 2; ++k) RT_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h->hostSrgb[k]), nFloats, cudaHostAllocDefault));
        h->hostSrgbBytes = nFloats;
    }
    cudaEvent_t resolved[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    for (int k = 0; k < 2; ++k) {
        RT_CUDA(cudaEventCreateWithFlags(&resolved[k], cudaEventDisableTiming));
        RT_CUDA(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming));
    }
    const int total = p->sample_end - p->sample_begin;
    int rc = RT_OK, done = 0, pendingDone = 0, b = 0;
    bool pending = false;
    auto deliver = [&](int set, int samples) -> int {
        RT_CUDA(cudaEventSynchronize(copied[set]));
        fn(user, want_linear ? h->hostLinear[set] : nullptr, want_srgb8 ? h->hostSrgb[set] : nullptr, samples, total);
        return RT_OK;
    };
    h->timedReadback = false;
    while (done < total && rc == RT_OK) {
        const int set = b & 1;
        rt_render_params q = *p;
        q.sample_begin = p->sample_begin + done;
        q.sample_end = std::min(p->sample_end, q.sample_begin + batch);
        q.clear = b == 0 ? 1 : 0;
        rc = rt_render(h, cam, &q);
        if (rc != RT_OK) break;
        done = q.sample_end - p->sample_begin;
        RT_CUDA(cudaSetDevice(d0.device));
        cudaStream_t s0 = d0.lastStream;
        // the resolve runs on the render stream, between two batches (it reads the accumulator the next batch
        // adds to); only the device-to-host copy overlaps the next batch.  Stage set `set` was last read by the copy
        // of batch b-2.
        if (b >= 2) RT_CUDA(cudaStreamWaitEvent(s0, copied[set], 0));
        rc = QueueReduceResolve(h, nullptr, set, want_linear != 0, want_srgb8 != 0, 1.0f / (float)done);
        if (rc != RT_OK) break;
        RT_CUDA(cudaEventRecord(resolved[set], s0));
        RT_CUDA(cudaStreamWaitEvent(h->copyStream, resolved[set], 0));
        if (want_linear)
            RT_CUDA(cudaMemcpyAsync(h->hostLinear[set], h->linearStage[set], nFloats * sizeof(float), cudaMemcpyDeviceToHost, h->copyStream));
        if (want_srgb8) RT_CUDA(cudaMemcpyAsync(h->hostSrgb[set], h->srgbStage[set], nFloats, cudaMemcpyDeviceToHost, h->copyStream));
        RT_CUDA(cudaEventRecord(copied[set], h->copyStream));
        // hand over the PREVIOUS batch while this one renders
        if (pending) rc = deliver(set ^ 1, pendingDone);
        pending = true;
        pendingDone = done;
        ++b;
    }
    if (rc == RT_OK && pending) rc = deliver((b - 1) & 1, pendingDone);
    cudaSetDevice(d0.device);
    cudaStreamSynchronize(h->copyStream);
    for (int k = 0; k < 2; ++k) {
        if (resolved[k]) cudaEventDestroy(resolved[k]);
        if (copied[k]) cudaEventDestroy(copied[k]);
    }
    if (rc == RT_OK) rc = rt_sync(h);
    rt_camera c = *cam;
    c.samples_per_pixel = std::max(1, total); // a later rt_readback of this handle reports the same mean
    h->lastCam = c;
    return rc;
}

int rt_get_timing(rt_scene_handle h, rt_timing* out)
{
    if (!h || !out) {
        rt_set_error("rt_get_timing: NULL argument");
        return RT_ERR_INVALID;
    }
    if (!h->rendered) {
        rt_set_error("rt_get_timing: nothing rendered yet");
        return RT_ERR_STATE;
    }
    DeviceGuard guard;
    std::memset(out, 0, sizeof *out);
    out->n_devices = (int32_t)h->devs.size();
    for (size_t k = 0; k < h->devs.size() && k < 16; ++k) {
        DeviceCtx& d = h->devs[k];
        RT_CUDA(cudaSetDevice(d.device));
        RT_CUDA(cudaStreamSynchronize(d.lastStream));
    }
    RT_CUDA(cudaSetDevice(h->devs[0].device));
    if (h->timedReadback) {
        RT_CUDA(cudaEventElapsedTime(&out->reduce_ms, h->evReduce0, h->evReduce1));
        RT_CUDA(cudaEventElapsedTime(&out->resolve_ms, h->evReduce1, h->evResolve1));
    }
    for (size_t k = 0; k < h->devs.size() && k < 16; ++k) {
        DeviceCtx& d = h->devs[k];
        RT_CUDA(cudaSetDevice(d.device));
        if (cudaEventElapsedTime(&out->render_ms[k], d.evStart, d.evStop) != cudaSuccess) {
            cudaGetLastError();
            out->render_ms[k] = 0.0f;
        }
    }
    return RT_OK;
}

int rt_scene_free(rt_scene_handle h)
{
    if (!h) return RT_OK;
    DeviceGuard guard;
    if (!h->devs.empty()) {
        DeviceCtx& d0 = h->devs[0];
        if (cudaSetDevice(d0.device) == cudaSuccess) {
            if (h->copyStream) {
                cudaStreamSynchronize(h->copyStream);
                cudaStreamDestroy(h->copyStream);
            }
            cudaStream_t s0 = d0.lastStream ? d0.lastStream : d0.ownStream;
            for (int k = 0; k < 2; ++k) {
                BigFree(d0.device, h->linearStage[k], h->linearStageFloats[k] * sizeof(float), s0);
                BigFree(d0.device, h->srgbStage[k], h->srgbStageBytes[k], s0);
                if (h->hostLinear[k]) cudaFreeHost(h->hostLinear[k]);
                if (h->hostSrgb[k]) cudaFreeHost(h->hostSrgb[k]);
            }
            if (h->evReduce0) cudaEventDestroy(h->evReduce0);
            if (h->evReduce1) cudaEventDestroy(h->evReduce1);
            if (h->evResolve1) cudaEventDestroy(h->evResolve1);
            if (h->evReduced) cudaEventDestroy(h->evReduced);
        }
    }
    for (NcclComm c : h->comms)
        if (c && gNccl.ok) gNccl.CommDestroy(c);
    for (DeviceCtx& d : h->devs) FreeDevice(h, d);
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
    if (blocks.empty()) return RT_OK;
    DeviceGuard guard;
    for (const BigBlock& b : blocks) {
        if (cudaSetDevice(b.device) == cudaSuccess) {
            if (b.idle) {
                cudaEventSynchronize(b.idle);
                cudaEventDestroy(b.idle);
            }
            cudaFree(b.ptr);
        }
    }
    return RT_OK;
}

int rt_scene_get_info(rt_scene_handle h, rt_scene_info* info)
{
    if (!h || !info) {
        rt_set_error("rt_scene_get_info: NULL argument");
        return RT_ERR_INVALID;
    }
    const DevScene& dev = h->devs[0].dev;
    std::memset(info, 0, sizeof *info);
    info->n_prims_baked = dev.n_spheres + dev.n_moving + dev.n_quads;
    info->n_nodes = dev.n_nodes;
    info->n_media = dev.n_media;
    info->max_depth_bvh = h->host->max_depth;
    info->features = dev.features;
    info->scene_in_smem = h->fitsSmem ? 1 : 0;
    info->nodes_in_smem = h->nodesInSmem ? 1 : 0;
    info->variant = h->pickedVariant;
    info->n_devices = (int32_t)h->devs.size();
    info->reduce_path = h->reducePath;
    info->block_threads = h->pickedThreads;
    info->registers = h->pickedRegisters;
    info->device_bytes = h->arenaBytes;
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
    if (h->devs.size() != 1) {
        rt_set_error("rt_debug_trace_path: single-device handles only");
        return RT_ERR_INVALID;
    }
    DeviceGuard guard;
    DeviceCtx& d = h->devs[0];
    RT_CUDA(cudaSetDevice(d.device));
    RT_CUDA(cudaMemset(d.debugOut, 0, 256 * 8 * sizeof(float)));
    h->debugPixel = pixel;
    h->debugSample = sample;
    rt_render_params q = *p;
    q.sample_begin = sample;
    q.sample_end = sample + 1;
    const int rc = rt_render(h, cam, &q);
    h->debugPixel = h->debugSample = -1;
    if (rc != RT_OK) return rc;
    RT_CUDA(cudaSetDevice(d.device));
    RT_CUDA(cudaStreamSynchronize(d.lastStream));
    RT_CUDA(cudaMemcpy(records, d.debugOut, (size_t)max_records * 8 * sizeof(float), cudaMemcpyDeviceToHost));
    return RT_OK;
}

float rt_rng_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t domain, uint32_t dim)
{
    const rt_u4 b = rt_rng_block(seed, pixel, sample, slot, domain, dim >> 2);
    return rt_bits_to_u01(rt_u4_lane(b, dim & 3u));
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
    if (n == "rt_timing") return (int)sizeof(rt_timing);
    return -1;
}

} // extern "C"
