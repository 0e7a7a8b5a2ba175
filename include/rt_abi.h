/*
 * rt_abi.h -- C ABI of the B200 path-tracing render path.
 *
 * This is the drop-in boundary for the one hot path of
 * eazuooz/RayTracinginOneWeekendinCUDA:
 *
 *     Render -> RayColor -> BvhNode::Hit / Material::Scatter
 *     (reference RayTracinginOneWeekend/kernel.cu:65-154)
 *
 * The reference has no FFI; the path sits behind (i) the scene-type
 * constructors used at kernel.cu:203-508, (ii) the Camera constructor
 * (Camera.h:36-46) and (iii) the launch
 *     Render<<<grid,8x8>>>(fb, W, H, spp, camera, world, randState)
 * (kernel.cu:122-124,685-687) followed by host reads of the framebuffer.
 * (i)+(ii) are kept as host C++ classes (include/rt/scene.hpp); (iii) is
 * replaced by rt_scene_upload / rt_render / rt_readback below.
 *
 * Everything crossing this boundary is plain C: pointers, sizes, PODs.
 * The scene description is the reference's object graph written out flat
 * and in FP64 (exactly the numbers the reference's constructors hold), so
 * that the FP64 oracle and the GPU path consume the same bytes.
 *
 * Error convention: every entry point returns RT_OK (0) or a negative
 * rt_status; nothing calls exit() (the reference does: kernel.cu:31-40).
 * rt_last_error() returns a thread-local message for the last failure.
 */
#ifndef RT_ABI_H
#define RT_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 2

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID = -1,     /* bad argument / malformed scene          */
    RT_ERR_CUDA = -2,        /* a CUDA runtime call failed               */
    RT_ERR_NO_DEVICE = -3,   /* no usable CUDA device (no CPU fallback)  */
    RT_ERR_UNSUPPORTED = -4, /* scene uses something the path lacks      */
    RT_ERR_STATE = -5        /* call order (e.g. readback before render) */
} rt_status;

/* ---- primitives (Sphere.h:8-21, MovingSphere.h:20-36, Quad.h:25-37) ---- */
enum { RT_PRIM_SPHERE = 0, RT_PRIM_MOVING_SPHERE = 1, RT_PRIM_QUAD = 2 };

typedef struct rt_prim {
    int32_t type;        /* RT_PRIM_*                                          */
    int32_t material;    /* index into rt_scene_desc.materials                 */
    int32_t first_xform; /* instance chain, outermost first (Instance.h:28,71) */
    int32_t xform_count; /* 0 = not instanced                                  */
    double a[3];         /* sphere: centre | moving: centre0 | quad: Q         */
    double b[3];         /* moving: centre1                  | quad: u         */
    double c[3];         /*                                  | quad: v         */
    double radius;       /* spheres                                            */
    double time0, time1; /* moving sphere shutter interval                     */
} rt_prim;

/* ---- instance wrappers (Instance.h:28-64 Translate, :71-159 RotateY) ---- */
enum { RT_XFORM_TRANSLATE = 0, RT_XFORM_ROTATE_Y = 1 };

typedef struct rt_xform {
    int32_t type;    /* RT_XFORM_*                                       */
    int32_t _pad;
    double v[3];     /* translate: offset | rotate_y: {sin, cos, degrees} */
} rt_xform;

/* ---- top-level list[] entries: what the reference hands to BvhNode ---- */
enum {
    RT_OBJ_PRIM = 0,  /* one primitive (possibly instanced)                    */
    RT_OBJ_LIST = 1,  /* owning HittableList tested linearly (HittableList.h)  */
    RT_OBJ_MEDIUM = 2 /* ConstantMedium over a boundary (ConstantMedium.h)     */
};

typedef struct rt_object {
    int32_t kind;           /* RT_OBJ_*                                        */
    int32_t first_prim;     /* prims of the object / of the medium's boundary  */
    int32_t prim_count;
    int32_t phase_material; /* medium: its Isotropic material                  */
    int32_t medium_id;      /* medium: construction order, keys its RNG draws  */
    int32_t _pad;
    double density;         /* medium: rho (ConstantMedium.h:22)               */
    double bbox[6];         /* xmin,xmax,ymin,ymax,zmin,zmax as the reference's
                               constructor chain computes it (AABB.h:26-48)    */
} rt_object;

/* ---- materials (Material.h:45-167, Metal.h:9-35, Dielectric.h:10-69) ---- */
enum {
    RT_MAT_LAMBERTIAN = 0,
    RT_MAT_METAL = 1,
    RT_MAT_DIELECTRIC = 2,
    RT_MAT_DIFFUSE_LIGHT = 3,
    RT_MAT_ISOTROPIC = 4
};

typedef struct rt_material {
    int32_t type;     /* RT_MAT_*                                     */
    int32_t texture;  /* lambertian / light / isotropic: texture index */
    double albedo[3]; /* metal                                         */
    double fuzz;      /* metal, already min(fuzz,1) (Metal.h:13)       */
    double ior;       /* dielectric                                    */
} rt_material;

/* ---- textures (Texture.h:35-176) ---- */
enum { RT_TEX_SOLID = 0, RT_TEX_CHECKER = 1, RT_TEX_IMAGE = 2, RT_TEX_NOISE = 3 };

typedef struct rt_texture {
    int32_t type;    /* RT_TEX_*                                           */
    int32_t even;    /* checker: texture index                             */
    int32_t odd;     /* checker: texture index                             */
    int32_t image;   /* image: index into images, -1 = none (-> cyan)      */
    int32_t perlin;  /* noise: index into perlins                          */
    int32_t _pad;
    double color[3]; /* solid                                              */
    double scale;    /* checker: 1/scale is applied (Texture.h:65) | noise */
} rt_texture;

/* Perlin tables (Perlin.h:22-34, 84-117): 256 unit vectors + 3 perms. */
typedef struct rt_perlin {
    double ranvec[256][3];
    int32_t perm_x[256];
    int32_t perm_y[256];
    int32_t perm_z[256];
} rt_perlin;

/* RGB8 texels as RtwImage hands them to ImageTexture (RtwImage.h:51-105):
 * already linearised and re-quantised; row 0 = top of the image. */
typedef struct rt_image {
    int32_t width, height;
    const uint8_t* rgb; /* width*height*3 bytes, may be NULL (-> cyan) */
} rt_image;

typedef struct rt_scene_desc {
    int32_t abi_version; /* RT_ABI_VERSION */
    int32_t n_objects, n_prims, n_xforms, n_materials, n_textures, n_perlins, n_images;
    const rt_object* objects; /* in list[] construction order (kernel.cu:207-508) */
    const rt_prim* prims;
    const rt_xform* xforms;
    const rt_material* materials;
    const rt_texture* textures;
    const rt_perlin* perlins;
    const rt_image* images;
} rt_scene_desc;

/* Camera surface (Camera.h:36-46 + kernel.cu:189-197), in the current book's
 * parameterisation; the reference's `aperture` is
 *     aperture = 2 * focus_dist * tan(defocus_angle/2)      (degrees in)
 * so aperture 0.1 at focus_dist 10 is defocus_angle = 2*atan(0.005) rad.
 * If aperture >= 0 it overrides defocus_angle (exact reference mapping). */
typedef struct rt_camera {
    int32_t image_width, image_height; /* aspect = width/height (kernel.cu:536) */
    int32_t samples_per_pixel;
    int32_t max_depth;                 /* reference hard-codes 50 (kernel.cu:71) */
    double vfov;                       /* degrees */
    double lookfrom[3], lookat[3], vup[3];
    double defocus_angle;              /* degrees */
    double focus_dist;
    double aperture;                   /* < 0: derive from defocus_angle */
    double time0, time1;               /* shutter */
    double background[3];              /* constant colour (kernel.cu:77,197) */
} rt_camera;

/* Kernel variants (rt_render_params.variant).  Megakernel and wavefront render the same image bit for
 * bit; head/tail and hit-queue the same up to the fp32 order in which a pixel's paths are summed (hit-queue
 * also refines hit points from a different starting value: equal to ~1e-14 relative). */
enum {
    RT_VARIANT_AUTO = 0,       /* hit-queue (megakernel beyond 2^19 samples per call)              */
    RT_VARIANT_MEGAKERNEL = 1, /* persistent-thread megakernel                                    */
    RT_VARIANT_WAVEFRONT = 2,  /* on-chip wavefront: extend / shade / gen over warp-local queues  */
    RT_VARIANT_HEADTAIL = 3,   /* synchronous heads (new samples) + queued continuation rays       */
    RT_VARIANT_HITQUEUE = 4    /* synchronous heads + queued hits: shading runs on full warps      */
};

/* BVH the device traverses (rt_render_params.bvh / rt_upload_options.bvh). */
enum {
    RT_BVH_SAH = 0,       /* binned-SAH tree over baked world-space primitives  */
    RT_BVH_REFERENCE = 1, /* the reference's median-split topology (BvhNode.h)  */
    RT_BVH_NONE = 2       /* linear list: the reference's own BVH==list check   */
};

typedef struct rt_upload_options {
    int32_t device; /* CUDA device ordinal (used when n_devices <= 1) */
    int32_t bvh;    /* RT_BVH_* */
    int32_t max_leaf_prims; /* 0 = default */
    int32_t flags;  /* RT_UPLOAD_* */
    /* Multi-device, single process (the reference's one launch site, kernel.cu:678-689, fanned out):
     * n_devices > 1 replicates the packed scene on device_ids[0..n_devices) -- packed ONCE on the host, one copy per
     * device.  rt_render then gives device k the k-th slice of the sample range on its own stream, and rt_readback
     * sums the fp32 accumulators on device_ids[0] (one reduction over NVLink) before it resolves the frame. */
    int32_t n_devices;         /* 0 or 1: single device */
    int32_t _pad;
    const int32_t* device_ids; /* NULL: devices 0 .. n_devices-1 */
} rt_upload_options;

enum {
    RT_UPLOAD_NO_HOIST = 1, /* keep scene-sized primitives / media inside the BVH (A/B test of the hoisting) */
    RT_UPLOAD_REDUCE_NCCL = 2, /* multi-device: sum the accumulators with ncclReduce (libnccl.so.2 is dlopen'ed on
                                  first use) instead of the fused peer-memory reduce + resolve kernel             */
    RT_UPLOAD_WHOLE_LISTS = 4, /* keep a small owning list (MakeBox: six quads) as ONE BVH item instead of one item
                                  per primitive (A/B test: fewer nodes, longer leaves; measured slower)            */
    RT_UPLOAD_NO_BOXES = 8     /* do not recognise closed six-quad boxes (MakeBox, Instance.h:166-184): test their
                                  quads one by one as the reference does (A/B test of the slab-tested box)          */
};

typedef struct rt_render_params {
    int32_t sample_begin, sample_end; /* global sample indices rendered by this call:
                                         GPU k of G renders [k*spp/G,(k+1)*spp/G)   */
    uint32_t seed;                    /* reference uses 1984 (kernel.cu:118)         */
    int32_t variant;                  /* RT_VARIANT_*                                */
    int32_t clear;                    /* 1: zero the accumulator + ray counter first */
    int32_t block_threads;            /* 0 = default                                 */
    int32_t blocks_per_sm;            /* 0 = default                                 */
    int32_t flags;                    /* RT_FLAG_*                                   */
    void* stream;                     /* cudaStream_t to launch on, NULL = default   */
    float* accum;                     /* optional caller-owned device buffer W*H*3
                                         fp32 (e.g. a torch tensor for NCCL); NULL =
                                         the handle's own accumulator                */
} rt_render_params;

enum {
    RT_FLAG_STATS = 0x100,          /* instrumented kernel: fills rt_stats.paths/node_tests/prim_tests */
    RT_FLAG_SCENE_IN_GLOBAL = 0x200, /* do not stage the scene in shared memory (A/B test)              */
    RT_FLAG_NODES_IN_GLOBAL = 0x400, /* a scene too large to stage: do not stage its node table either  */
    RT_FLAG_IMPORTANCE = 0x800       /* importance sampling of the lights (phase 4 of the reference's roadmap, README.md:37-42,
                                        not implemented there): every Lambertian / Isotropic bounce draws its direction
                                        from 1/2 (the reference's own scattering density) + 1/2 (density of the directions
                                        towards the quads and spheres that carry a DiffuseLight material) and is weighed
                                        with scattering density / mixture density -- the same image in expectation as
                                        without the flag, with less noise where lights are small.  Hit-queue kernel only. */
    /* bits 4-5 and 12-30 are development tuning knobs of the kernels (csrc/rt_device.cu)             */
};

typedef struct rt_scene_s* rt_scene_handle;

/* Per-ray traversal statistics gathered by an instrumented render (debug). */
typedef struct rt_stats {
    uint64_t rays;       /* world.Hit queries = RayColor loop iterations (kernel.cu:71-74) */
    uint64_t paths;      /* camera samples started                                         */
    uint64_t node_tests; /* box tests executed on device                                   */
    uint64_t prim_tests; /* primitive tests executed on device                             */
} rt_stats;

/* Deep-copies the host scene, bakes instances, builds the device BVH and
 * uploads everything to `opt->device`. Replaces CreateWorld<<<1,1>>>
 * (kernel.cu:176-543, launched :667) + the cudaMallocs at :628-649. */
int rt_scene_upload(const rt_scene_desc* scene, const rt_upload_options* opt, rt_scene_handle* out);

/* Replaces RenderInit + Render launches (kernel.cu:681-689). Asynchronous on
 * p->stream; adds fp32 linear radiance SUMS for samples [begin,end) into the
 * accumulator (row 0 = bottom row, like the reference framebuffer). */
int rt_render(rt_scene_handle scene, const rt_camera* cam, const rt_render_params* p);

/* Progressive output (replaces the reference's render-then-write-PPM sequence, kernel.cu:681-723, for
 * long renders): the sample range of `p` is rendered in batches of `batch` samples; after every batch the frame so far
 * (multi-device: reduced on device 0) is resolved -- mean over the samples done, and gamma + quantise when srgb8 is
 * wanted -- into one of two pinned host buffers, and `fn` is called with it while the NEXT batch is already
 * rendering.  The buffers belong to the handle and are valid until the following callback returns. */
typedef void (*rt_progress_fn)(void* user, const float* linear_rgb, const uint8_t* srgb8, int32_t samples_done,
                               int32_t samples_total);
int rt_render_progressive(rt_scene_handle scene, const rt_camera* cam, const rt_render_params* p, int32_t batch,
                          int32_t want_linear, int32_t want_srgb8, rt_progress_fn fn, void* user);

/* Device-side times of the last rt_render / rt_readback of this handle (CUDA events on each device's stream). */
typedef struct rt_timing {
    int32_t n_devices;
    int32_t _pad;
    float render_ms[16]; /* per device: its kernel launch(es) of the last rt_render */
    float reduce_ms;     /* multi-device: the accumulator reduction of the last rt_readback (0 if single) */
    float resolve_ms;    /* resolve kernel + device-to-host copies of the last rt_readback */
} rt_timing;
int rt_get_timing(rt_scene_handle scene, rt_timing* out);

/* Device pointer + float count of the handle's accumulator (for collectives). */
int rt_accum_ptr(rt_scene_handle scene, float** dev_ptr, uint64_t* n_floats);

/* Blocks until the handle's queued work is done. */
int rt_sync(rt_scene_handle scene);

/* Replaces the host reads of the managed framebuffer (kernel.cu:707).
 * linear_rgb (W*H*3 floats, may be NULL): mean radiance = sum / cam.samples_per_pixel.
 * srgb8 (W*H*3 bytes, may be NULL): sqrt-gamma (kernel.cu:150-152) then
 *   clamp[0,0.999]*256 quantise (kernel.cu:712-718); row 0 = TOP row (PPM order).
 * `accum`: device buffer to read instead of the handle's (may be NULL). */
int rt_readback(rt_scene_handle scene, const float* accum, float* linear_rgb, uint8_t* srgb8,
                rt_stats* stats);

int rt_scene_free(rt_scene_handle scene);

/* Scene arenas and frame-sized buffers of freed handles are kept (up to 8 blocks per process) and reused by
 * the next upload / render of the same size; this returns them to the driver. */
int rt_release_cached_memory(void);

/* Scene summary after upload (prim/node counts, bytes, kernel class). */
typedef struct rt_scene_info {
    int32_t n_prims_baked, n_nodes, n_media, max_depth_bvh;
    int32_t features;      /* RT_FEAT_* bitmask that selected the kernel instantiation */
    int32_t scene_in_smem; /* 1 when nodes+prims+materials are staged in shared memory */
    int32_t variant;       /* RT_VARIANT_* the last rt_render ran (what AUTO resolved to)    */
    int32_t n_devices;     /* devices the scene is resident on                              */
    int32_t reduce_path;   /* multi-device: 0 = none yet, 1 = peer-memory fused kernel, 2 = NCCL */
    int32_t block_threads, registers; /* launch shape and registers/thread of the last rt_render's kernel */
    int32_t nodes_in_smem; /* 1 when at least the BVH node table is staged in shared memory           */
    int32_t _pad;
    uint64_t device_bytes;
    int32_t medium_visits[8]; /* T2: reference-topology visit multiplicity per medium_id */
} rt_scene_info;
int rt_scene_get_info(rt_scene_handle scene, rt_scene_info* info);

/* Host only (no CUDA call): what rt_scene_upload would build for this scene -- baking, BVH, hoisting, packing --
 * as counts.  Lets hosts and CPU-only tests inspect the packer. */
typedef struct rt_pack_info {
    int32_t n_nodes, n_spheres, n_moving, n_quads, n_media, n_materials, n_mat_params;
    int32_t max_depth_bvh; /* levels the traversal stack must hold (joins of mixed leaves included) */
    int32_t features;      /* RT_FEAT_* */
    int32_t n_hoisted;     /* scene-sized items tested before the tree (see RT_UPLOAD_NO_HOIST) */
    uint32_t hoisted[4];   /* their leaf refs: type in bits 30..28 (0 sphere, 1 moving, 2 quad, 3 medium, 4 box) */
    uint64_t staged_bytes; /* nodes + primitives + materials: what a CTA stages in shared memory when it fits */
    int32_t n_boxes;       /* closed six-quad boxes tested by one slab test (their quads are counted in n_quads) */
    int32_t n_lights;      /* quads / spheres with a DiffuseLight material: the sampling targets of RT_FLAG_IMPORTANCE */
} rt_pack_info;
int rt_scene_pack_info(const rt_scene_desc* scene, const rt_upload_options* opt, rt_pack_info* out);

const char* rt_last_error(void);

/* Test hook: one uniform of the render stream, exactly as the kernel draws it:
 * (0,1], keyed on (seed, pixel, sample, slot, domain, dim). */
float rt_rng_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t domain,
                     uint32_t dim);

/* Test hook: renders `sample` of the whole frame with the instrumented kernel and
 * returns, for the path of (pixel, sample), one 8-float record per bounce:
 * {hit id bits, t, material bits, front face, p.x, p.y, p.z, 1}; unused records are 0.
 * pixel = -2 (hit-queue kernel): instead, the first 128 words of `records` are uint32 counters -- a histogram of the
 * child-pair steps per walk of that sample, camera rays in [0, 64), scattered rays in [64, 128) (tools/walk_hist.py). */
int rt_debug_trace_path(rt_scene_handle scene, const rt_camera* cam, const rt_render_params* p, int32_t pixel,
                        int32_t sample, float* records, int32_t max_records);

/* Measures the device's fp32 FMA throughput with dependent-free FFMA chains
 * (the roofline denominator of this path: it is FP32-issue bound, not HBM). */
int rt_measure_fp32_peak(int32_t device, double* tflops, double* sm_count);

/* sizeof() of an ABI struct by name, for binding self-checks; -1 if unknown. */
int rt_abi_sizeof(const char* name);

/* Writes the reference's P3 text PPM (kernel.cu:696-723) from srgb8 (top row first). */
int rt_write_ppm(const char* path, const uint8_t* srgb8, int32_t width, int32_t height);
/* Same pixels as binary P6 (3 bytes per pixel instead of ~11 of text; not in the reference). */
int rt_write_ppm_binary(const char* path, const uint8_t* srgb8, int32_t width, int32_t height);

#ifdef __cplusplus
}
#endif
#endif /* RT_ABI_H */
