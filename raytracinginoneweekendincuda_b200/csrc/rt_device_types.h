// rt_device_types.h -- layout of the scene in HBM / shared memory.
//
// Everything the kernel reads is a flat array of 16-byte-aligned records read
// with 128-bit loads.  World-space only: the reference's instance wrappers
// (Translate / RotateY, reference Instance.h:28-159) are baked into the
// primitives at upload, so the device never transforms a ray.
//
// Precision policy (DESIGN.md "numerics"): positions -- ray origins, sphere
// centres, quad corners, plane offsets -- are FP64; directions, normals,
// colours, BVH boxes and all shading are fp32.  B200 issues FP64 at half the
// fp32 rate, which makes this affordable; it is what keeps discrete decisions
// (hit/miss at tMin, face tests) equal to the FP64 reference.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define RT_ALIGN(n) __align__(n)
#else
#define RT_ALIGN(n) alignas(n)
#endif

// RT_BVH4 (build option, A/B only -- VERDICT r1 #3): the binary SAH tree collapsed into nodes of up to four children
// (four adjacent DevNode records = 128 bytes per visit; an unused slot has a negative half-extent and never hits).
// Measured against the binary tree and not shipped: DESIGN.md 5.4, profiles/r2_ab_w.jsonl.
#ifndef RT_BVH4
#define RT_BVH4 0
#endif

// ---- BVH node: 32 bytes = two 128-bit loads.  Children of an internal node
// are adjacent (index c and c+1), so one step fetches 64 contiguous bytes.
// The box is stored as centre and half-extent: the slab test then needs no
// per-axis min/max (entry/exit = t_centre -+ |e/d|), which moves 12 of its 20
// min/max instructions per child pair from the half-rate ALU pipe -- the busiest
// pipe of the kernel (ncu: 70 %) -- to the FMA pipe (20 %).
struct RT_ALIGN(16) DevNode {
    float c[3];
    uint32_t ref; // what lies below this box: see RT_REF_*
    float e[3];
    uint32_t aux; // unused (0); keeps the record at 32 bytes
};

// ref encoding.  Internal: index of the first of its two children.  Leaf:
//   bit 31      = 1
//   bits 30..28 = leaf type (RT_LEAF_*)
//   bits 27..18 = count-1 (1..1024 records)
//   bits 17..0  = first index in that type's array
#define RT_REF_LEAF 0x80000000u
#define RT_LEAF_SPHERE 0u
#define RT_LEAF_MOVING 1u
#define RT_LEAF_QUAD 2u
#define RT_LEAF_MEDIUM 3u
#define RT_LEAF_BOX 4u /* a closed six-quad box (DevBox); its hits are reported as hits of its quads */
#define RT_REF_TYPE(r) (((r) >> 28) & 7u)
#define RT_REF_COUNT(r) ((((r) >> 18) & 1023u) + 1u)
#define RT_REF_FIRST(r) ((r)&0x3ffffu)
#define RT_REF_MAKE_LEAF(type, first, count) \
    (RT_REF_LEAF | ((uint32_t)(type) << 28) | (((uint32_t)(count)-1u) << 18) | (uint32_t)(first))
#define RT_MAX_LEAF_PRIMS 1024
#define RT_MAX_HOISTED 4 /* leaf refs tested before the tree is entered (rt_pack.hpp: hoisting) */
#define RT_MAX_PRIMS_PER_TYPE (1 << 18)

// Hit identifier carried out of traversal: type in bits 30..29, index below.
#define RT_HIT_NONE 0xffffffffu
#define RT_HIT_MAKE(type, index) (((uint32_t)(type) << 29) | (uint32_t)(index))
#define RT_HIT_TYPE(h) (((h) >> 29) & 3u)
#define RT_HIT_INDEX(h) ((h)&0x1fffffffu)

// ---- Sphere (reference Sphere.h:8-21): 32 bytes, all FP64 (the quadratic's
// b and c are formed in FP64).  The material index lives in a parallel int
// array that is read once per ray, after traversal.
struct RT_ALIGN(16) DevSphere {
    double cx, cy, cz;
    double radius;
};

// ---- MovingSphere (reference MovingSphere.h:20-36): 64 bytes.
// centre(time) = c0 + ((time - time0) * inv_dt) * dc, dc = c1 - c0.
struct RT_ALIGN(16) DevMovingSphere {
    double c0x, c0y, c0z;
    float radius;
    int32_t material;
    double dcx, dcy, dcz;
    float time0;
    float inv_dt; // 1/(time1-time0)
};

// ---- Quad (reference Quad.h:25-37): 96 bytes.  Plane (n, D) in FP64 so that
// t = (D - n.O)/(n.d) does not lose the origin's position on the plane.  The interior coordinates of Quad.h:72-73,
// alpha = w.(p x v) and beta = w.(u x p) with p the hit point relative to Q, are triple products: alpha = p.(v x w),
// beta = p.(w x u).  The two constant vectors are formed once on the host (FP64, rounded to fp32), so a test costs two
// dot products instead of two cross products and two dot products.
struct RT_ALIGN(16) DevQuad {
    double qx, qy, qz;
    double D;
    double nx, ny, nz;
    float ax, ay, az; // v x w
    float bx, by, bz; // w x u
    int32_t material;
    int32_t _p[3];
};
static_assert(sizeof(DevQuad) == 96, "DevQuad layout");

// ---- Box: the six quads MakeBox builds (reference Instance.h:166-184), possibly under Translate / RotateY -- any
// closed parallelepiped of six quads.  128 bytes.  The reference tests the six quads one after the other
// (HittableList.h:39-57); a line meets a convex box in at most two of them, the entry and the exit of the three
// slabs, so ONE slab test over the three pairs of parallel faces finds the face that the six tests would find:
// pair p has the unit normal n[p] and its two faces lie at n[p].x = lo[p] < hi[p].  A hit is reported as a hit
// of the face's own DevQuad (first_quad + face number), which is what FinalizeHit refines and shades.
struct RT_ALIGN(16) DevBox {
    double n[3][3];
    double lo[3], hi[3];
    uint32_t first_quad; // the six quads, contiguous in DevScene.quads
    uint32_t faces;      // 3 bits per slab face: quad number (0..5) of pair p's lo face at bits 6p, hi face at 6p+3
};

// ---- ConstantMedium (reference ConstantMedium.h:18-50): 32 bytes.
struct RT_ALIGN(16) DevMedium {
    uint32_t boundary_ref;  // leaf-style ref to the boundary primitives
    int32_t phase_material; // its Isotropic
    int32_t medium_id;      // keys the RNG domain
    int32_t visits;         // reference-topology visit multiplicity (SURVEY trap T2)
    double neg_inv_density; // -1/rho
    double _p;
};

// ---- Light: a sampling target of the importance-sampling mode (RT_FLAG_IMPORTANCE): a quad or a sphere
// whose material is a DiffuseLight.  96 bytes, FP64, read once per diffuse bounce.
struct RT_ALIGN(16) DevLight {
    double a[3]; // quad: Q | sphere: centre
    double b[3]; // quad: u | sphere: radius, 0, 0
    double c[3]; // quad: v
    uint32_t hit; // RT_HIT_MAKE(type, index) of the primitive in its device table
    uint32_t _p;
    double area;  // quad: |u x v|
};

// ---- Material: 16 bytes = one 128-bit load.  Book 1 has one material per sphere (486 of them), and the table is
// staged in shared memory next to the BVH: at 32 bytes it was a quarter of the staged scene.  The FP64 parameter of
// the two materials that have one (metal fuzz, dielectric index: they enter the scattered direction, which is FP64
// end to end) lives in a side table indexed by `index`.
//   tt = type (RT_MAT_*, bits 0-2) | index << 3
//   index: lambertian / light / isotropic -> texture + 1 (0 = the solid colour in r,g,b)
//          metal / dielectric            -> slot in DevScene.mat_params
struct RT_ALIGN(16) DevMaterial {
    float r, g, b; // metal albedo, or the colour when the texture is solid (folded in)
    uint32_t tt;
};
#define RT_MAT_TT(type, index) ((uint32_t)(type) | ((uint32_t)(index) << 3))
#define RT_MAT_TT_TYPE(tt) ((int)((tt)&7u))
#define RT_MAT_TT_INDEX(tt) ((tt) >> 3)

// ---- Texture: 48 bytes.
struct RT_ALIGN(16) DevTexture {
    int32_t type; // RT_TEX_*
    int32_t even, odd;
    int32_t index; // image or perlin index
    float r, g, b;
    float scale;      // noise scale
    double inv_scale; // checker 1/scale (Texture.h:65)
    double _p;
};

// ---- Perlin tables (reference Perlin.h:22-34): 256 float4 vectors + 3x256 bytes.
struct RT_ALIGN(16) DevPerlin {
    float ranvec[256][4];
    uint8_t perm_x[256];
    uint8_t perm_y[256];
    uint8_t perm_z[256];
    uint8_t _p[256];
};

// ---- UV frame of a sphere baked out of a RotateY chain: the reference takes (u,v) from the OBJECT-space normal
// (Sphere.h:42-44 inside the instance, Instance.h:136-150 rotates only P and Normal back), so an image-textured sphere
// under RotateY keeps its texture turned with it.  The accumulated yaw takes the world normal back:
// x' = c x - s z, z' = s x + c z.  Spheres that need one carry its index + 1 in bits 20..30 of their material word.
struct DevUvFrame {
    float s, c;
};
#define RT_MATERIAL_INDEX_BITS 20
#define RT_MATERIAL_INDEX_MASK 0xfffff

// Texels live in the scene arena; the record holds their OFFSET from the arena base, so that the arena is one
// position-independent block that any device can receive as it is.
struct RT_ALIGN(16) DevImage {
    uint32_t offset; // bytes from DevScene.arena to the first texel (RGB8, row 0 = top)
    int32_t width, height; // 0 x 0: no image (-> cyan, Texture.h:113-114)
    uint32_t _p;
};

// Scene feature bits: select the kernel instantiation.
#define RT_FEAT_MOVING 1
#define RT_FEAT_QUAD 2
#define RT_FEAT_MEDIUM 4
#define RT_FEAT_TEXTURE 8        /* any non-solid texture */
#define RT_FEAT_TEXTURE_HEAVY 16 /* image or Perlin-noise textures (their code costs ~25 registers) */
#define RT_FEAT_IMPORTANCE 32    /* not a scene property: the kernel class RT_FLAG_IMPORTANCE renders with */

struct DevScene {
    const DevNode* nodes;
    const DevSphere* spheres;
    const int32_t* sphere_material;
    const DevMovingSphere* moving;
    const DevQuad* quads;
    const DevBox* boxes;
    const DevMedium* media;
    const DevMaterial* materials;
    const double* mat_params; // metal fuzz / dielectric index of refraction, FP64
    const DevTexture* textures;
    const DevPerlin* perlins;
    const DevImage* images;
    const DevUvFrame* uv_frames;
    const DevLight* lights; // sampling targets of RT_FLAG_IMPORTANCE
    const uint8_t* arena; // base of the scene arena (image texel offsets are relative to it)
    uint32_t root_ref;
    const uint32_t* hoisted; // RT_MAX_HOISTED leaf refs in the arena, tested before the tree is entered
    int32_t n_hoisted;
    int32_t n_nodes, n_spheres, n_moving, n_quads, n_boxes, n_media, n_materials, n_textures, n_lights;
    int32_t features;
};

// Camera frame (reference Camera.h:36-71), FP64 as computed on the host.
struct DevCamera {
    double origin[3], llc[3], horiz[3], vert[3];
    double u[3], v[3];
    double lens_radius;
    double inv_width, inv_height;
    float time0, time1;
    float background[3];
    int32_t width, height, max_depth;
};
