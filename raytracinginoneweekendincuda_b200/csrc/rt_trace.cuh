// rt_trace.cuh -- device functions of the render path: traversal,
// intersection, hit finalisation, textures, scattering.
//
// Follows the algorithm of the reference's Render -> RayColor -> BvhNode::Hit /
// Material::Scatter (reference kernel.cu:65-154 and the class headers cited at
// each function) over the flat device layout of rt_device_types.h.  Written
// for sm_100a: 128-bit loads (LDS.128 when the scene is staged in shared
// memory, LDG.E.128 through the read-only path otherwise), a per-thread
// traversal stack in shared memory laid out [level][thread] so that lanes
// never conflict on a bank, and the mixed precision described in DESIGN.md:
// FP64 for positions and for the cancellation-prone terms of the sphere and
// plane equations, fp32 for everything else.
#pragma once

#include <stdint.h>

#include "../../include/rt_abi.h"
#include "../../include/rt_rng.h"
#include "rt_device_types.h"

#define RT_DEV __device__ __forceinline__

namespace rtdev {

struct f3 {
    float x, y, z;
};
struct d3 {
    double x, y, z;
};

RT_DEV f3 make_f3(float x, float y, float z)
{
    f3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
RT_DEV f3 operator+(f3 a, f3 b) { return make_f3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV f3 operator-(f3 a, f3 b) { return make_f3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV f3 operator*(float s, f3 a) { return make_f3(s * a.x, s * a.y, s * a.z); }
RT_DEV f3 operator*(f3 a, f3 b) { return make_f3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_DEV f3 operator-(f3 a) { return make_f3(-a.x, -a.y, -a.z); }
RT_DEV float dot(f3 a, f3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
RT_DEV f3 cross(f3 a, f3 b) { return make_f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// Vec3.h:117-120 with :96-99: v * (1/len)
RT_DEV f3 unit(f3 a) { return (1.0f / sqrtf(dot(a, a))) * a; }

// SFU approximations (1-2 ulp), one instruction each instead of the ~9 of an IEEE
// division or root.  Used only where an FP64 refinement follows (roots of the
// winning primitive) or where a conservative margin absorbs it (slab test).
RT_DEV float RcpApprox(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
RT_DEV float SqrtApprox(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// FP64 vectors for the geometric chain hit point -> normal -> scattered
// direction -> next hit point.  An fp32 link anywhere in that chain is a 6e-8
// perturbation that diffuse inter-reflection between small spheres multiplies
// by ~d/r per bounce and the marble texture by ~1e2 per unit of position
// (measured: reference scene 9 loses parity after ~9 bounces inside its sphere
// cluster), so the chain is FP64 end to end; B200 issues FP64 at half rate.
RT_DEV d3 make_d3(double x, double y, double z)
{
    d3 r;
    r.x = x;
    r.y = y;
    r.z = z;
    return r;
}
RT_DEV d3 operator+(d3 a, d3 b) { return make_d3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV d3 operator-(d3 a, d3 b) { return make_d3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV d3 operator*(double s, d3 a) { return make_d3(s * a.x, s * a.y, s * a.z); }
RT_DEV d3 operator-(d3 a) { return make_d3(-a.x, -a.y, -a.z); }
RT_DEV double dot(d3 a, d3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
RT_DEV f3 to_f3(d3 a) { return make_f3((float)a.x, (float)a.y, (float)a.z); }
// 1/x, 1/sqrt(x), sqrt(x) to ~1e-14 relative from the fp32 SFU seed and one
// (two for the reciprocal root) Newton steps: no FP64 division or root sequence.
RT_DEV double RcpD(double x)
{
    const double r = (double)RcpApprox((float)x);
    const double e = fma(-x, r, 1.0);
    return fma(fma(e, e, e), r, r);
}
RT_DEV double RsqrtD(double x)
{
    double y = (double)rsqrtf((float)x);
    const double e = fma(-x * y, y, 1.0); // 1 - x y^2
    y = fma(y * e, fma(0.375, e, 0.5), y); // y (1 + e/2 + 3 e^2/8)
    return y;
}
RT_DEV double SqrtD(double x) { return x > 1e-290 ? x * RsqrtD(x) : 0.0; }

// ------------------------------------------------------------------ memory
// Scene arrays are addressed either as 32-bit shared-memory addresses (SMEM)
// or as generic pointers to global memory.
template <bool SMEM> struct Base;
template <> struct Base<true> {
    uint32_t a;
};
template <> struct Base<false> {
    const char* a;
};

template <bool SMEM> RT_DEV float4 Ld4(Base<SMEM> b, uint32_t off);
template <> RT_DEV float4 Ld4<true>(Base<true> b, uint32_t off)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(b.a + off));
    return v;
}
template <> RT_DEV float4 Ld4<false>(Base<false> b, uint32_t off) { return __ldg(reinterpret_cast<const float4*>(b.a + off)); }

template <bool SMEM> RT_DEV double2 LdD2(Base<SMEM> b, uint32_t off);
template <> RT_DEV double2 LdD2<true>(Base<true> b, uint32_t off)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(b.a + off));
    return v;
}
template <> RT_DEV double2 LdD2<false>(Base<false> b, uint32_t off) { return __ldg(reinterpret_cast<const double2*>(b.a + off)); }

template <bool SMEM> RT_DEV double LdD(Base<SMEM> b, uint32_t off);
template <> RT_DEV double LdD<true>(Base<true> b, uint32_t off)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(b.a + off));
    return v;
}
template <> RT_DEV double LdD<false>(Base<false> b, uint32_t off) { return __ldg(reinterpret_cast<const double*>(b.a + off)); }

template <bool SMEM> RT_DEV int32_t LdI(Base<SMEM> b, uint32_t off);
template <> RT_DEV int32_t LdI<true>(Base<true> b, uint32_t off)
{
    int32_t v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(b.a + off));
    return v;
}
template <> RT_DEV int32_t LdI<false>(Base<false> b, uint32_t off) { return __ldg(reinterpret_cast<const int32_t*>(b.a + off)); }

template <bool SMEM> struct SceneView {
    Base<SMEM> nodes, spheres, sphere_material, moving, quads, boxes, media, materials, mat_params;
    // never staged: textures and their tables
    const DevTexture* textures;
    const DevPerlin* perlins;
    const DevPerlin* perlin0; // table 0 staged in shared memory (generic pointer), or perlins: SetupScene
    const DevImage* images;
    const DevUvFrame* uv_frames;
    const DevLight* lights; // RT_FLAG_IMPORTANCE: sampling targets
    int n_lights;
    const uint8_t* arena;
    uint32_t root_ref;
    const uint32_t* hoisted; // leaf refs every ray tests before it enters the tree (global memory: a uniform load)
    int n_hoisted;
    // Scene too large for shared memory as a whole, but its NODE table fits beside the stacks and queues: the nodes
    // alone are staged -- traversal steps are then LDS hits while the primitives still come through L1 -- and refs of
    // internal nodes are shared addresses as in the fully staged case.  A warp-uniform flag, only looked at by the
    // global-memory instantiations.  (Measured on scene 9 built with RT_UPLOAD_WHOLE_LISTS, 78 KB of nodes: +1.8 %.)
    bool nodes_shared;
};

// Per-thread traversal stack in shared memory, [level][thread]: entry(level) = base + level*stride, so the lanes of
// a warp never share a bank.  The walk keeps the ADDRESS of its next free entry (Trav::sp), not a level index: a
// push is one store and one add, a pop one add and one load (the index form cost an IMAD and a re-read of the thread
// id per access -- the compiler rematerialised the base inside the loop rather than spend a register on it).
struct Stack {
    uint32_t base;   // shared address of this thread's level-0 slot
    uint32_t stride; // bytes between levels = 4*blockDim.x
};
RT_DEV void StackStore(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v)); }
RT_DEV uint32_t StackLoad(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ---------------------------------------------------------------------- ray
struct Ray {
    d3 o;       // FP64 origin
    d3 d;       // direction (primary: FP64; scattered: the fp32 value widened)
    float time;
};

// What traversal needs besides the ray: fp32 copies for the slab test.
// RT_SLAB_FFMA (build option; default on): also keep |1/d|, so that entry/exit = fma(-+e, |1/d|, t_centre) -- 9 FFMA
// per box instead of 3 FFMA + 3 FMUL + 6 FADD -- at the price of three more live registers in the traversal loop.
// Measured on the hit-queue kernel (4K Book 1, 64 spp): 19.49 -> 20.30 Grays/s (profiles/README.md, round 2).
#ifndef RT_SLAB_FFMA
#define RT_SLAB_FFMA 1
#endif
struct RaySlab {
    f3 inv;     // 1/d
    f3 ood;     // o/d
#if RT_SLAB_FFMA
    f3 ainv;    // |1/d|
#endif
    float rcpA; // 1/|d|^2, for the sphere roots
};

RT_DEV RaySlab MakeSlab(const Ray& r)
{
    RaySlab s;
    const f3 df = make_f3((float)r.d.x, (float)r.d.y, (float)r.d.z);
    // AABB.h:77-93 divides by d; d == 0 gives +-inf and the NaNs of 0*inf are
    // dropped by fminf/fmaxf exactly as fmin/fmax do in the reference.
    s.inv = make_f3(RcpApprox(df.x), RcpApprox(df.y), RcpApprox(df.z));
    s.rcpA = RcpApprox(fmaf(df.x, df.x, fmaf(df.y, df.y, df.z * df.z)));
    s.ood = make_f3((float)r.o.x * s.inv.x, (float)r.o.y * s.inv.y, (float)r.o.z * s.inv.z);
#if RT_SLAB_FFMA
    s.ainv = make_f3(fabsf(s.inv.x), fabsf(s.inv.y), fabsf(s.inv.z));
#endif
    return s;
}

// AABB.h:68-98 on a centre/half-extent box: per axis t_c = c*inv - o*inv and
// p = e*inv, entry = t_c - |p|, exit = t_c + |p| (the |.| is a free source modifier
// of FADD), so no per-axis min/max.  Returns whether the box is hit within
// [tmin, tmax]; `tn` is the entry distance.
RT_DEV bool SlabEntry(const float4 c, const float4 e, const RaySlab& s, float tmin, float tmax, float& tn)
{
    const float cx = fmaf(c.x, s.inv.x, -s.ood.x);
    const float cy = fmaf(c.y, s.inv.y, -s.ood.y);
    const float cz = fmaf(c.z, s.inv.z, -s.ood.z);
#if RT_SLAB_FFMA
    // e * |1/d| = NaN only for 0 * inf (a flat box seen edge-on): fminf/fmaxf drop it, like the product form below
    tn = fmaxf(fmaxf(fmaf(-e.x, s.ainv.x, cx), fmaf(-e.y, s.ainv.y, cy)), fmaxf(fmaf(-e.z, s.ainv.z, cz), tmin));
    const float tf = fminf(fminf(fmaf(e.x, s.ainv.x, cx), fmaf(e.y, s.ainv.y, cy)), fminf(fmaf(e.z, s.ainv.z, cz), tmax));
#else
    const float px = fabsf(e.x * s.inv.x), py = fabsf(e.y * s.inv.y), pz = fabsf(e.z * s.inv.z);
    tn = fmaxf(fmaxf(cx - px, cy - py), fmaxf(cz - pz, tmin));
    const float tf = fminf(fminf(cx + px, cy + py), fminf(cz + pz, tmax));
#endif
    return tf >= tn;
}

// ------------------------------------------------------------- primitives
// Primitive tests return the hit distance, or RT_MISS.  Distances may be
// negative (a medium's boundary is queried over (-inf, +inf),
// ConstantMedium.h:60), so the sentinel is a value no test can produce.
#define RT_MISS (-3.0e38f)
// Sphere.h:22-70.  Roots of a t^2 + 2 b t + c with b, c and the discriminant
// in FP64 (the centre may be 1000 units away from a hit point whose position
// matters to 1e-3); the square root and the roots themselves are fp32 -- the
// winner is refined by FinalizeSphere.  Root order and the open interval
// (tmin, tmax) follow the reference.  Returns t or RT_MISS.
// TM = float for surface queries (tmin = 0.001), double for a medium's boundary
// queries: its second one starts at t1 + 1e-4 (ConstantMedium.h:63), which fp32
// cannot hold beyond t ~ 1000.  rcpA ~ 1/a.
template <class TM>
RT_DEV float SphereRoots(double ocx, double ocy, double ocz, double radius, const Ray& r, double a, float rcpA, TM tmin, float tmax)
{
    const double b = fma(ocx, r.d.x, fma(ocy, r.d.y, ocz * r.d.z));
    const double c = fma(ocx, ocx, fma(ocy, ocy, fma(ocz, ocz, -radius * radius)));
    const double disc = fma(b, b, -a * c);
    if (!(disc > 0.0)) return RT_MISS;
    const float s = SqrtApprox((float)disc);
    const float bf = (float)b, cf = (float)c;
    // cancellation-free pair: q has the larger magnitude; roots q/a and c/q
    const bool pos = bf > 0.0f;
    const float q = pos ? -(bf + s) : (s - bf);
    const float ta = q * rcpA, tb = cf * RcpApprox(q);
    const float t0 = pos ? ta : tb, t1 = pos ? tb : ta; // t0 <= t1
    if (t0 < tmax && (TM)t0 > tmin) return t0;
    if (t1 < tmax && (TM)t1 > tmin) return t1;
    return RT_MISS;
}

template <bool SMEM, class TM>
RT_DEV float HitSphere(const SceneView<SMEM>& sv, uint32_t index, const Ray& r, double a, float rcpA, TM tmin, float tmax)
{
    const double2 s0 = LdD2<SMEM>(sv.spheres, index * 32u);
    const double2 s1 = LdD2<SMEM>(sv.spheres, index * 32u + 16u);
    return SphereRoots<TM>(r.o.x - s0.x, r.o.y - s0.y, r.o.z - s1.x, s1.y, r, a, rcpA, tmin, tmax);
}

// MovingSphere.h:44-102: centre lerped by the ray's time, then Sphere.
template <bool SMEM> RT_DEV d3 MovingCentre(const SceneView<SMEM>& sv, uint32_t index, float time, double& radius)
{
    const double2 m0 = LdD2<SMEM>(sv.moving, index * 64u);
    const double2 m1 = LdD2<SMEM>(sv.moving, index * 64u + 16u);
    const double2 m2 = LdD2<SMEM>(sv.moving, index * 64u + 32u);
    const double2 m3 = LdD2<SMEM>(sv.moving, index * 64u + 48u);
    const float rad = __int_as_float(__double2loint(m1.y));
    const float time0 = __int_as_float(__double2loint(m3.y));
    const float invDt = __int_as_float(__double2hiint(m3.y));
    const double frac = (double)((time - time0) * invDt);
    radius = (double)rad;
    d3 c;
    c.x = fma(frac, m2.x, m0.x);
    c.y = fma(frac, m2.y, m0.y);
    c.z = fma(frac, m3.x, m1.x);
    return c;
}

template <bool SMEM, class TM>
RT_DEV float HitMoving(const SceneView<SMEM>& sv, uint32_t index, const Ray& r, double a, float rcpA, TM tmin, float tmax)
{
    double radius;
    const d3 c = MovingCentre<SMEM>(sv, index, r.time, radius);
    return SphereRoots<TM>(r.o.x - c.x, r.o.y - c.y, r.o.z - c.z, radius, r, a, rcpA, tmin, tmax);
}

// Quad.h:54-99.  Plane terms in FP64, interior test in fp32.  Closed interval
// [tmin, tmax] and closed [0,1] for alpha/beta as in the reference.  Returns t
// or RT_MISS; alpha/beta are written on a hit.
template <bool SMEM, class TM>
RT_DEV float HitQuad(const SceneView<SMEM>& sv, uint32_t index, const Ray& r, TM tmin, float tmax, float& alpha, float& beta)
{
    const uint32_t off = index * 96u;
    const double2 q0 = LdD2<SMEM>(sv.quads, off);        // qx qy
    const double2 q1 = LdD2<SMEM>(sv.quads, off + 16u);  // qz D
    const double2 q2 = LdD2<SMEM>(sv.quads, off + 32u);  // nx ny
    const double2 q3 = LdD2<SMEM>(sv.quads, off + 48u);  // nz | ax ay
    const double nz = q3.x;
    const double denom = fma(q2.x, r.d.x, fma(q2.y, r.d.y, nz * r.d.z));
    if (fabs(denom) < 1e-8) return RT_MISS;
    const double num = q1.y - fma(q2.x, r.o.x, fma(q2.y, r.o.y, nz * r.o.z));
    const float t = (float)num * RcpApprox((float)denom);
    if ((TM)t < tmin || t > tmax) return RT_MISS;
    const float4 q4 = Ld4<SMEM>(sv.quads, off + 64u); // az bx by bz   (after ax ay in q3.y)
    const float ax = __int_as_float(__double2loint(q3.y)), ay = __int_as_float(__double2hiint(q3.y));
    const double td = (double)t;
    const f3 planar = make_f3((float)(fma(td, r.d.x, r.o.x) - q0.x), (float)(fma(td, r.d.y, r.o.y) - q0.y),
                              (float)(fma(td, r.d.z, r.o.z) - q1.x));
    // Quad.h:72-73 as triple products: alpha = p.(v x w), beta = p.(w x u)
    const float al = dot(planar, make_f3(ax, ay, q4.x));
    const float be = dot(planar, make_f3(q4.y, q4.z, q4.w));
    if (!(0.0f <= al && al <= 1.0f) || !(0.0f <= be && be <= 1.0f)) return RT_MISS;
    alpha = al;
    beta = be;
    return t;
}

// One closed box (DevBox, rt_device_types.h): MakeBox's six quads (Instance.h:166-184) as they lie in the world.  The
// reference runs Quad::Hit (Quad.h:54-99) on each of the six and keeps the closest (HittableList.h:39-57); a line
// meets a convex box in the entry and the exit of its three slabs and nowhere else, so the closest face in [tmin, tmax]
// is the entry face if the entry lies in the interval, else the exit face if that does.  Per pair of faces
// t = (D - n.O)/(n.d) as in Quad.h:62, plane terms in FP64; a pair the ray runs parallel to (|n.d| < 1e-8, Quad.h:59)
// is hit nowhere and only decides whether the line lies between its two planes.  Returns t (fp32; FinalizeHit refines
// it against the face's own plane) or RT_MISS; `quad` = index of the face's DevQuad.
template <bool SMEM, class TM>
RT_DEV float HitBox(const SceneView<SMEM>& sv, uint32_t index, const Ray& r, TM tmin, float tmax, uint32_t& quad)
{
    const uint32_t off = index * 128u;
    const double2 b0 = LdD2<SMEM>(sv.boxes, off);        // n0x n0y
    const double2 b1 = LdD2<SMEM>(sv.boxes, off + 16u);  // n0z n1x
    const double2 b2 = LdD2<SMEM>(sv.boxes, off + 32u);  // n1y n1z
    const double2 b3 = LdD2<SMEM>(sv.boxes, off + 48u);  // n2x n2y
    const double2 b4 = LdD2<SMEM>(sv.boxes, off + 64u);  // n2z lo0
    const double2 b5 = LdD2<SMEM>(sv.boxes, off + 80u);  // lo1 lo2
    const double2 b6 = LdD2<SMEM>(sv.boxes, off + 96u);  // hi0 hi1
    const double2 b7 = LdD2<SMEM>(sv.boxes, off + 112u); // hi2 | first_quad, faces
    const uint32_t firstQuad = (uint32_t)__double2loint(b7.y), faces = (uint32_t)__double2hiint(b7.y);
    const double nx[3] = {b0.x, b1.y, b3.x}, ny[3] = {b0.y, b2.x, b3.y}, nz[3] = {b1.x, b2.y, b4.x};
    const double lo[3] = {b4.y, b5.x, b5.y}, hi[3] = {b6.x, b6.y, b7.x};
    float tEnter = -3.402823466e+38f, tExit = 3.402823466e+38f;
    uint32_t fEnter = 0, fExit = 0;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const double den = fma(nx[p], r.d.x, fma(ny[p], r.d.y, nz[p] * r.d.z));
        const double s = fma(nx[p], r.o.x, fma(ny[p], r.o.y, nz[p] * r.o.z));
        if (fabs(den) < 1e-8) {
            if (s < lo[p] || s > hi[p]) return RT_MISS;
            continue;
        }
        const float inv = RcpApprox((float)den);
        const float ta = (float)(lo[p] - s) * inv, tb = (float)(hi[p] - s) * inv;
        const bool loFirst = ta < tb;
        const float tn = loFirst ? ta : tb, tf = loFirst ? tb : ta;
        const uint32_t pair = faces >> (6 * p);
        const uint32_t fn = (loFirst ? pair : pair >> 3) & 7u, ff = (loFirst ? pair >> 3 : pair) & 7u;
        if (tn > tEnter) {
            tEnter = tn;
            fEnter = fn;
        }
        if (tf < tExit) {
            tExit = tf;
            fExit = ff;
        }
    }
    if (tEnter > tExit) return RT_MISS;
    if (!((TM)tEnter < tmin) && !(tEnter > tmax)) { // closed interval, Quad.h:64
        quad = firstQuad + fEnter;
        return tEnter;
    }
    if (!((TM)tExit < tmin) && !(tExit > tmax)) {
        quad = firstQuad + fExit;
        return tExit;
    }
    return RT_MISS;
}

// The same slab test for a ConstantMedium whose boundary is a box (ConstantMedium.h:60-63): the first boundary query,
// over (-inf, inf), returns the entry of the line into the box, the second, from t1 + 1e-4, its exit.  FP64 throughout
// (the scatter point is the origin of the rest of the path).  False when the line misses the box.
template <bool SMEM>
RT_DEV bool BoxSpan(const SceneView<SMEM>& sv, uint32_t index, const Ray& r, double& t1, double& t2)
{
    const uint32_t off = index * 128u;
    const double2 b0 = LdD2<SMEM>(sv.boxes, off);
    const double2 b1 = LdD2<SMEM>(sv.boxes, off + 16u);
    const double2 b2 = LdD2<SMEM>(sv.boxes, off + 32u);
    const double2 b3 = LdD2<SMEM>(sv.boxes, off + 48u);
    const double2 b4 = LdD2<SMEM>(sv.boxes, off + 64u);
    const double2 b5 = LdD2<SMEM>(sv.boxes, off + 80u);
    const double2 b6 = LdD2<SMEM>(sv.boxes, off + 96u);
    const double hi2 = LdD<SMEM>(sv.boxes, off + 112u);
    const double nx[3] = {b0.x, b1.y, b3.x}, ny[3] = {b0.y, b2.x, b3.y}, nz[3] = {b1.x, b2.y, b4.x};
    const double lo[3] = {b4.y, b5.x, b5.y}, hi[3] = {b6.x, b6.y, hi2};
    t1 = -1.0e300;
    t2 = 1.0e300;
#pragma unroll
    for (int p = 0; p < 3; ++p) {
        const double den = fma(nx[p], r.d.x, fma(ny[p], r.d.y, nz[p] * r.d.z));
        const double s = fma(nx[p], r.o.x, fma(ny[p], r.o.y, nz[p] * r.o.z));
        if (fabs(den) < 1e-8) {
            if (s < lo[p] || s > hi[p]) return false;
            continue;
        }
        const double inv = RcpD(den);
        const double ta = (lo[p] - s) * inv, tb = (hi[p] - s) * inv;
        // (explicit compare-and-select: fmin / fmax of doubles carry NaN handling, ~8 instructions each -- ncu put these
        // two lines at 11 % of the Cornell-smoke kernel's instructions)
        const bool asc = ta < tb;
        const double tn = asc ? ta : tb, tf = asc ? tb : ta;
        t1 = tn > t1 ? tn : t1;
        t2 = tf < t2 ? tf : t2;
    }
    return t1 <= t2;
}

// One leaf-style run of primitives, closest hit with a shrinking tmax
// (HittableList.h:39-57).  Returns the hit id or RT_HIT_NONE; t in tmax.
template <int FEAT, bool SMEM, class TM>
RT_DEV uint32_t HitRun(const SceneView<SMEM>& sv, uint32_t ref, const Ray& r, double a, float rcpA, TM tmin, float& tmax,
                       uint32_t& primTests)
{
    const uint32_t type = RT_REF_TYPE(ref), first = RT_REF_FIRST(ref), count = RT_REF_COUNT(ref);
    uint32_t hit = RT_HIT_NONE;
#pragma unroll 1 // leaves hold 1-2 primitives: an unrolled body only bloats the divergent leaf path
    for (uint32_t i = 0; i < count; ++i) {
        float t;
        ++primTests;
        // (surface queries only: a medium whose boundary is a box takes BoxSpan, never this loop -- and the
        // feature-complete kernel is short of instruction cache, not of branches)
        if (sizeof(TM) == sizeof(float) && (FEAT & RT_FEAT_QUAD) && type == RT_LEAF_BOX) {
            uint32_t quad = 0;
            t = HitBox<SMEM, TM>(sv, first + i, r, tmin, tmax, quad);
            if (t != RT_MISS) {
                tmax = t;
                hit = RT_HIT_MAKE(RT_LEAF_QUAD, quad);
            }
            continue;
        }
        if ((FEAT & RT_FEAT_QUAD) && type == RT_LEAF_QUAD) {
            float al, be;
            t = HitQuad<SMEM, TM>(sv, first + i, r, tmin, tmax, al, be);
        } else if ((FEAT & RT_FEAT_MOVING) && type == RT_LEAF_MOVING) {
            t = HitMoving<SMEM, TM>(sv, first + i, r, a, rcpA, tmin, tmax);
        } else {
            t = HitSphere<SMEM, TM>(sv, first + i, r, a, rcpA, tmin, tmax);
        }
        if (t != RT_MISS) {
            tmax = t;
            hit = RT_HIT_MAKE(type, first + i);
        }
    }
    return hit;
}

// ------------------------------------------------------------ hit records
// Hittable.h:11-31, filled for the winning primitive only.
struct Hit {
    d3 p;          // FP64 hit point
    d3 n;          // FP64 shading normal (faces the ray)
    f3 outward;    // geometric outward normal, fp32 copy (sphere UV)
    float u, v;    // quad: alpha, beta; sphere: filled on demand
    bool front;
    int32_t material;
    int32_t uvFrame; // sphere under a RotateY chain with an image texture: index + 1 into uv_frames, else 0
};

RT_DEV void SetFaceNormal(Hit& h, const d3& dir, const d3& outward) // Hittable.h:26-30
{
    h.front = dot(dir, outward) < 0.0;
    h.n = h.front ? outward : -outward;
    h.outward = to_f3(outward);
}

// The fp32 root of a sphere re-solved in FP64: one Newton step on
// f(t) = a t^2 + 2 b t + c puts the hit point on the sphere to ~1e-14 relative.
RT_DEV double RefineSphereT(d3 c, double radius, const Ray& r, double a, float t)
{
    const double ocx = r.o.x - c.x, ocy = r.o.y - c.y, ocz = r.o.z - c.z;
    const double b = fma(ocx, r.d.x, fma(ocy, r.d.y, ocz * r.d.z));
    const double cc = fma(ocx, ocx, fma(ocy, ocy, fma(ocz, ocz, -radius * radius)));
    const double td = (double)t;
    const double f = fma(fma(a, td, 2.0 * b), td, cc);
    const double fp = 2.0 * fma(a, td, b);
    return td - (double)((float)f * RcpApprox((float)fp));
}

// t = (D - n.O)/(n.d) of a quad refined once in FP64 (Quad.h:62).
template <bool SMEM> RT_DEV double RefineQuadT(const SceneView<SMEM>& sv, uint32_t index, const Ray& r)
{
    const uint32_t off = index * 96u;
    const double2 q1 = LdD2<SMEM>(sv.quads, off + 16u);
    const double2 q2 = LdD2<SMEM>(sv.quads, off + 32u);
    const double nz = LdD2<SMEM>(sv.quads, off + 48u).x;
    const double denom = fma(q2.x, r.d.x, fma(q2.y, r.d.y, nz * r.d.z));
    const double num = q1.y - fma(q2.x, r.o.x, fma(q2.y, r.o.y, nz * r.o.z));
    const float invDen = RcpApprox((float)denom);
    double td = (double)((float)num * invDen);
    td += (double)((float)fma(-td, denom, num) * invDen);
    td += (double)((float)fma(-td, denom, num) * invDen);
    return td;
}

template <bool SMEM> RT_DEV d3 SphereCentre(const SceneView<SMEM>& sv, uint32_t index, double& radius)
{
    const double2 s0 = LdD2<SMEM>(sv.spheres, index * 32u);
    const double2 s1 = LdD2<SMEM>(sv.spheres, index * 32u + 16u);
    radius = s1.y;
    return make_d3(s0.x, s0.y, s1.x);
}

// FP64 distance of a surface hit found by HitRun (hit id + fp32 t).
template <int FEAT, bool SMEM> RT_DEV double RefineT(const SceneView<SMEM>& sv, uint32_t hit, const Ray& r, double a, float t)
{
    const uint32_t type = RT_HIT_TYPE(hit), index = RT_HIT_INDEX(hit);
    if ((FEAT & RT_FEAT_QUAD) && type == RT_LEAF_QUAD) return RefineQuadT<SMEM>(sv, index, r);
    double radius;
    d3 c;
    if ((FEAT & RT_FEAT_MOVING) && type == RT_LEAF_MOVING)
        c = MovingCentre<SMEM>(sv, index, r.time, radius);
    else
        c = SphereCentre<SMEM>(sv, index, radius);
    return RefineSphereT(c, radius, r, a, t);
}

RT_DEV void FinalizeSphereAt(d3 c, double radius, const Ray& r, double a, float t, Hit& h)
{
    const double td = RefineSphereT(c, radius, r, a, t);
    h.p.x = fma(td, r.d.x, r.o.x);
    h.p.y = fma(td, r.d.y, r.o.y);
    h.p.z = fma(td, r.d.z, r.o.z);
    const double invR = RcpD(radius); // Sphere.h:54 with Vec3.h:96-99: (P - C) * (1/r)
    SetFaceNormal(h, r.d, invR * (h.p - c));
}

// Material word of a sphere: material index, and (scenes with image textures only) its UV frame above it.
template <int FEAT> RT_DEV void SetSphereMaterial(Hit& h, int32_t word)
{
    if (FEAT & RT_FEAT_TEXTURE_HEAVY) {
        h.material = word & RT_MATERIAL_INDEX_MASK;
        h.uvFrame = word >> RT_MATERIAL_INDEX_BITS;
    } else {
        h.material = word;
    }
}

template <int FEAT, bool SMEM>
RT_DEV void FinalizeHit(const SceneView<SMEM>& sv, const Ray& r, double a, uint32_t hit, float t, double tMedium, Hit& h)
{
    const uint32_t type = RT_HIT_TYPE(hit), index = RT_HIT_INDEX(hit);
    h.u = 0.0f;
    h.v = 0.0f;
    h.uvFrame = 0;
    if ((FEAT & RT_FEAT_MEDIUM) && type == RT_LEAF_MEDIUM) {
        // ConstantMedium.h:86-91: arbitrary normal, front face, phase material
        const float4 m0 = Ld4<SMEM>(sv.media, index * 32u);
        h.p.x = fma(tMedium, r.d.x, r.o.x);
        h.p.y = fma(tMedium, r.d.y, r.o.y);
        h.p.z = fma(tMedium, r.d.z, r.o.z);
        h.n = make_d3(1.0, 0.0, 0.0);
        h.outward = make_f3(1.0f, 0.0f, 0.0f);
        h.front = true;
        h.material = __float_as_int(m0.y);
    } else if ((FEAT & RT_FEAT_QUAD) && type == RT_LEAF_QUAD) {
        const uint32_t off = index * 96u;
        const double2 q0 = LdD2<SMEM>(sv.quads, off);
        const double2 q1 = LdD2<SMEM>(sv.quads, off + 16u);
        const double2 q2 = LdD2<SMEM>(sv.quads, off + 32u);
        const double2 q3 = LdD2<SMEM>(sv.quads, off + 48u);
        const float4 q4 = Ld4<SMEM>(sv.quads, off + 64u);
        const double td = RefineQuadT<SMEM>(sv, index, r);
        h.p.x = fma(td, r.d.x, r.o.x);
        h.p.y = fma(td, r.d.y, r.o.y);
        h.p.z = fma(td, r.d.z, r.o.z);
        const f3 planar = make_f3((float)(h.p.x - q0.x), (float)(h.p.y - q0.y), (float)(h.p.z - q1.x));
        h.u = dot(planar, make_f3(__int_as_float(__double2loint(q3.y)), __int_as_float(__double2hiint(q3.y)), q4.x));
        h.v = dot(planar, make_f3(q4.y, q4.z, q4.w));
        h.material = LdI<SMEM>(sv.quads, off + 80u);
        SetFaceNormal(h, r.d, make_d3(q2.x, q2.y, q3.x));
    } else if ((FEAT & RT_FEAT_MOVING) && type == RT_LEAF_MOVING) {
        double radius;
        const d3 c = MovingCentre<SMEM>(sv, index, r.time, radius);
        FinalizeSphereAt(c, radius, r, a, t, h);
        SetSphereMaterial<FEAT>(h, __double2hiint(LdD2<SMEM>(sv.moving, index * 64u + 16u).y));
    } else {
        double radius;
        const d3 c = SphereCentre<SMEM>(sv, index, radius);
        FinalizeSphereAt(c, radius, r, a, t, h);
        SetSphereMaterial<FEAT>(h, LdI<SMEM>(sv.sphere_material, index * 4u));
    }
}

// ---------------------------------------------------------------- media
// ConstantMedium.h:52-94.  Two boundary queries, clip to [tmin,tmax], one
// keyed uniform per visit.  `visits` > 1 reproduces the reference testing a
// span-1 BVH leaf twice (SURVEY.md trap T2).  Distances are FP64: the scatter
// point is the origin of the rest of the path.  Returns true with the scatter
// distance in tOut (always >= tmin > 0).
// The body holds two boundary queries (each over every primitive type) and is reached from both leaf sites: ~40 % of
// the feature-complete kernel's code, and ncu showed that kernel waiting for instruction fetches
// (stall_no_instruction 5.2 cycles per issue on scene 9).
// (Measured, round 2: keeping this body out of line -- one copy instead of two -- shrank the feature-complete kernel
// from 7 000 to 4 100 instructions but forced the scene view into local memory for the call and lost 20-28 % on
// scenes 7, 8, 9; the code-size problem is handled by the runtime loop over the hoisted items instead.)
#ifndef RT_MEDIUM_INLINE
#define RT_MEDIUM_INLINE 1 /* build option for the A/B */
#endif
#if RT_MEDIUM_INLINE
#define RT_MEDIUM_FN __device__ __forceinline__
#else
#define RT_MEDIUM_FN __device__ __noinline__
#endif
template <int FEAT, bool SMEM>
RT_MEDIUM_FN bool HitMedium(const SceneView<SMEM>& sv, uint32_t index, const Ray& r, double a, float rcpA, float tminF, float tmaxF,
                      uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t& primTests, double& tOut)
{
    const float4 m0 = Ld4<SMEM>(sv.media, index * 32u);
    const uint32_t bref = (uint32_t)__float_as_int(m0.x);
    const uint32_t mediumId = (uint32_t)__float_as_int(m0.z);
    const int visits = __float_as_int(m0.w);
    double t1, t2;
    if (RT_REF_TYPE(bref) == RT_LEAF_SPHERE && RT_REF_COUNT(bref) == 1u) {
        // A single sphere as the boundary (both media of scene 9, kernel.cu:476-482): the two queries of
        // ConstantMedium.h:60-63 are the two roots of ONE quadratic -- the first over (-inf, inf) returns the smaller
        // root (Sphere.h:31-38), the second, from t1 + 1e-4, the larger one if it lies beyond.  Solved once, in FP64.
        double radius;
        const d3 c = SphereCentre<SMEM>(sv, RT_REF_FIRST(bref), radius);
        const double ocx = r.o.x - c.x, ocy = r.o.y - c.y, ocz = r.o.z - c.z;
        const double b = fma(ocx, r.d.x, fma(ocy, r.d.y, ocz * r.d.z));
        const double cc = fma(ocx, ocx, fma(ocy, ocy, fma(ocz, ocz, -radius * radius)));
        const double disc = fma(b, b, -a * cc);
        primTests += 2;
        if (!(disc > 0.0)) return false;
        const double sq = SqrtD(disc);
        const double q = b > 0.0 ? -(b + sq) : (sq - b); // the root pair without cancellation: q/a and cc/q
        const double ta = q * RcpD(a), tb = cc * RcpD(q);
        t1 = ta < tb ? ta : tb;
        t2 = ta < tb ? tb : ta;
        if (!(t2 > t1 + 0.0001)) return false;
    } else if ((FEAT & RT_FEAT_QUAD) && RT_REF_TYPE(bref) == RT_LEAF_BOX && RT_REF_COUNT(bref) == 1u) {
        // A box as the boundary (the smoke boxes of the Cornell scene, kernel.cu:424-431): entry and exit of one slab test.
        primTests += 2;
        if (!BoxSpan<SMEM>(sv, RT_REF_FIRST(bref), r, t1, t2)) return false;
        if (!(t2 >= t1 + 0.0001)) return false; // second query: closed at its tMin (Quad.h:64)
    } else {
        const float big = 3.402823466e+38f;
        float t1f = big;
        const uint32_t h1 = HitRun<FEAT, SMEM, double>(sv, bref, r, a, rcpA, -1.0e300, t1f, primTests);
        if (h1 == RT_HIT_NONE) return false;
        t1 = RefineT<FEAT, SMEM>(sv, h1, r, a, t1f);
        float t2f = big;
        // the candidates of the second query are fp32 roots: the entry root must not pass
        // as "beyond t1 + 1e-4" because its fp32 value lies above the refined t1
        const uint32_t h2 = HitRun<FEAT, SMEM, double>(sv, bref, r, a, rcpA, fmax(t1, (double)t1f) + 0.0001, t2f, primTests);
        if (h2 == RT_HIT_NONE) return false;
        t2 = RefineT<FEAT, SMEM>(sv, h2, r, a, t2f);
    }
    const double negInvDensity = LdD2<SMEM>(sv.media, index * 32u + 16u).x;
    const double invLength = RsqrtD(a), rayLength = a * invLength;
    double tmax = (double)tmaxF;
    bool any = false;
    for (int v = 0; v < visits; ++v) {
        double e1 = t1, e2 = t2;
        if (e1 < (double)tminF) e1 = (double)tminF;
        if (e2 > tmax) e2 = tmax;
        if (e1 >= e2) break;
        if (e1 < 0.0) e1 = 0.0;
        const double inside = (e2 - e1) * rayLength;
        const rt_u4 k = rt_rng_block(seed, pixel, sample, slot, 1u + 2u * mediumId + (uint32_t)v, 0);
        // ConstantMedium.h:79: log() of a float -- in the reference's device code that is CUDA's logf (<= 1 ulp), and
        // so it is here.  (Round 1 took the correctly rounded value through the FP64 log to agree with the host
        // oracle bit for bit: ~120 FP64 instructions, twice per ray for scene 9's mist.  A last-bit difference in a
        // scatter distance moves the scatter point by <= 6e-8 relative in ~15 % of the medium hits; the full-size
        // exact-stream tiles of scenes 8 and 9 hold the 99.9 % bar with it.)
        const double hitDistance = negInvDensity * (double)logf(rt_bits_to_u01(k.x));
        if (hitDistance > inside) continue;
        tmax = e1 + hitDistance * invLength;
        any = true;
    }
    tOut = tmax;
    return any;
}

// ---------------------------------------------------------------- traversal
// BvhNode.h:101-158 re-designed: children boxes are fetched as one 64-byte
// pair and tested together, the nearer child is entered first and the other
// pushed, leaves are contiguous typed runs.  The closest hit is the same
// minimum over all primitives the reference computes (its order of visiting
// them does not matter: media draws are keyed, not sequential).
//
// Traversal is a resumable state machine -- one node (or one leaf) per Step --
// so that the kernel can interleave lanes that are still walking the tree with
// lanes that are shading or starting a new path.
#define RT_TRAV_DONE 0xffffffffu
struct Trav {
    uint32_t ref; // node / leaf to visit next, RT_TRAV_DONE when the walk is over.  Internal node: index of its child
                  // pair (scene in global memory) or the pair's shared-memory ADDRESS (scene staged: SetupScene
                  // rewrites the refs while it copies the nodes, so a visit needs no base and no shift)
    uint32_t sp;  // shared address of the next free stack entry
    float t;      // closest hit so far (fp32; refined by FinalizeHit)
    uint32_t hit; // RT_HIT_* id or RT_HIT_NONE
    double tMedium; // FP64 scatter distance when `hit` is a medium
    // The bottom stack slot holds RT_TRAV_DONE, so popping never tests for an empty stack.
    RT_DEV void Begin(uint32_t root, const Stack& stack)
    {
        ref = root;
        StackStore(stack.base, RT_TRAV_DONE);
        sp = stack.base + stack.stride;
        t = 3.402823466e+38f;
        hit = RT_HIT_NONE;
    }
    RT_DEV void Idle()
    {
        ref = RT_TRAV_DONE;
        sp = 0;
        t = 3.402823466e+38f;
        hit = RT_HIT_NONE;
    }
};

RT_DEV void TravPop(const Stack& stack, Trav& tv)
{
    tv.sp -= stack.stride;
    tv.ref = StackLoad(tv.sp);
}
RT_DEV void TravPush(const Stack& stack, Trav& tv, uint32_t v)
{
    StackStore(tv.sp, v);
    tv.sp += stack.stride;
}

// The child pair of an internal node: four 128-bit loads.
template <bool SMEM> RT_DEV void LoadPair(const SceneView<SMEM>& sv, uint32_t ref, float4& lo0, float4& hi0, float4& lo1, float4& hi1);
template <> RT_DEV void LoadPair<true>(const SceneView<true>&, uint32_t ref, float4& lo0, float4& hi0, float4& lo1, float4& hi1)
{
    Base<true> b;
    b.a = ref; // already an address
    lo0 = Ld4<true>(b, 0u);
    hi0 = Ld4<true>(b, 16u);
    lo1 = Ld4<true>(b, 32u);
    hi1 = Ld4<true>(b, 48u);
}
template <> RT_DEV void LoadPair<false>(const SceneView<false>& sv, uint32_t ref, float4& lo0, float4& hi0, float4& lo1, float4& hi1)
{
    if (sv.nodes_shared) {
        Base<true> b;
        b.a = ref; // an address: see SceneView::nodes_shared
        lo0 = Ld4<true>(b, 0u);
        hi0 = Ld4<true>(b, 16u);
        lo1 = Ld4<true>(b, 32u);
        hi1 = Ld4<true>(b, 48u);
        return;
    }
    const uint32_t off = ref * 32u;
    lo0 = Ld4<false>(sv.nodes, off);
    hi0 = Ld4<false>(sv.nodes, off + 16u);
    lo1 = Ld4<false>(sv.nodes, off + 32u);
    hi1 = Ld4<false>(sv.nodes, off + 48u);
}

// (Measured and dropped: a `prefetch.global.L1` of the leaf's first record, issued by the box step that selects the leaf
// as the next visit, for scenes whose primitives stay in global memory -- Book 2 final 5.22 -> 4.72 Grays/s: ten more
// instructions per box step cost more than the hidden latency returns.  profiles/r2_ab_zd.jsonl.)
#if RT_BVH4
// A/B build (rt_device_types.h RT_BVH4): one collapsed node = four children, 8 x 128-bit loads, four slab tests, the hits
// ordered by entry distance with a five-exchange network, the nearest entered and the others parked far-to-near.
template <bool SMEM>
RT_DEV void TraceBox(const SceneView<SMEM>& sv, const RaySlab& slab, float tmin, const Stack& stack, Trav& tv, uint32_t& nodeTests)
{
    float4 lo0, hi0, lo1, hi1, lo2, hi2, lo3, hi3;
    LoadPair<SMEM>(sv, tv.ref, lo0, hi0, lo1, hi1);
    LoadPair<SMEM>(sv, tv.ref + (SMEM || sv.nodes_shared ? 64u : 2u), lo2, hi2, lo3, hi3); // shared address | node index
    nodeTests += 4;
    const float inf = 3.402823466e+38f;
    float e0, e1, e2, e3;
    if (!SlabEntry(lo0, hi0, slab, tmin, tv.t, e0)) e0 = inf;
    if (!SlabEntry(lo1, hi1, slab, tmin, tv.t, e1)) e1 = inf;
    if (!SlabEntry(lo2, hi2, slab, tmin, tv.t, e2)) e2 = inf;
    if (!SlabEntry(lo3, hi3, slab, tmin, tv.t, e3)) e3 = inf;
    uint32_t r0 = (uint32_t)__float_as_int(lo0.w), r1 = (uint32_t)__float_as_int(lo1.w);
    uint32_t r2 = (uint32_t)__float_as_int(lo2.w), r3 = (uint32_t)__float_as_int(lo3.w);
#define RT_CSWAP(ea, eb, ra, rb)                   \
    {                                              \
        const bool s_ = eb < ea;                   \
        const float lo_ = fminf(ea, eb), hi_ = fmaxf(ea, eb); \
        const uint32_t a_ = s_ ? rb : ra, b_ = s_ ? ra : rb;  \
        ea = lo_, eb = hi_, ra = a_, rb = b_;      \
    }
    RT_CSWAP(e0, e1, r0, r1)
    RT_CSWAP(e2, e3, r2, r3)
    RT_CSWAP(e0, e2, r0, r2)
    RT_CSWAP(e1, e3, r1, r3)
    RT_CSWAP(e1, e2, r1, r2)
#undef RT_CSWAP
    if (e3 < inf) TravPush(stack, tv, r3);
    if (e2 < inf) TravPush(stack, tv, r2);
    if (e1 < inf) TravPush(stack, tv, r1);
    if (e0 < inf)
        tv.ref = r0;
    else
        TravPop(stack, tv);
}
#else
// One internal node: both children boxes tested, nearer one entered first.
template <bool SMEM>
RT_DEV void TraceBox(const SceneView<SMEM>& sv, const RaySlab& slab, float tmin, const Stack& stack, Trav& tv, uint32_t& nodeTests)
{
    float4 lo0, hi0, lo1, hi1;
    LoadPair<SMEM>(sv, tv.ref, lo0, hi0, lo1, hi1);
    nodeTests += 2;
    float e0, e1;
    const bool h0 = SlabEntry(lo0, hi0, slab, tmin, tv.t, e0);
    const bool h1 = SlabEntry(lo1, hi1, slab, tmin, tv.t, e1);
    const uint32_t r0 = (uint32_t)__float_as_int(lo0.w), r1 = (uint32_t)__float_as_int(lo1.w);
    if (h0 && h1) {
        const bool swap = e1 < e0;
        TravPush(stack, tv, swap ? r0 : r1);
        tv.ref = swap ? r1 : r0;
    } else if (h0 || h1) {
        tv.ref = h0 ? r0 : r1;
    } else {
        TravPop(stack, tv);
    }
}
#endif

// One leaf: a typed run of primitives, or a medium.
template <int FEAT, bool SMEM>
RT_DEV void TestLeaf(const SceneView<SMEM>& sv, uint32_t ref, const Ray& r, double a, float rcpA, float tmin, Trav& tv,
                     uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t& primTests)
{
    if ((FEAT & RT_FEAT_MEDIUM) && RT_REF_TYPE(ref) == RT_LEAF_MEDIUM) {
        const uint32_t m = RT_REF_FIRST(ref);
        double tm;
        if (HitMedium<FEAT, SMEM>(sv, m, r, a, rcpA, tmin, tv.t, seed, pixel, sample, slot, primTests, tm)) {
            tv.t = (float)tm;
            tv.tMedium = tm;
            tv.hit = RT_HIT_MAKE(RT_LEAF_MEDIUM, m);
        }
    } else {
        const uint32_t h = HitRun<FEAT, SMEM, float>(sv, ref, r, a, rcpA, tmin, tv.t, primTests);
        if (h != RT_HIT_NONE) tv.hit = h;
    }
}

template <int FEAT, bool SMEM>
RT_DEV void TraceLeaf(const SceneView<SMEM>& sv, const Ray& r, double a, float rcpA, float tmin, const Stack& stack, Trav& tv,
                      uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t& primTests)
{
    TestLeaf<FEAT, SMEM>(sv, tv.ref, r, a, rcpA, tmin, tv, seed, pixel, sample, slot, primTests);
    TravPop(stack, tv);
}

// Starts a walk: the hoisted items (scene-sized primitives / media kept out of the tree, rt_pack.hpp) are tested
// first -- in a kernel whose lanes start their walks together this is straight-line code on a full warp -- and the
// distance they return bounds the walk through the tree.
template <int FEAT, bool SMEM>
RT_DEV void BeginWalk(const SceneView<SMEM>& sv, const Ray& r, double a, float rcpA, float tmin, const Stack& stack, Trav& tv,
                      uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t& primTests)
{
    tv.Begin(sv.root_ref, stack);
#pragma unroll 1 // ONE copy of the leaf test here (it is the largest piece of code of the feature-complete kernel)
    for (int k = 0; k < sv.n_hoisted; ++k)
        TestLeaf<FEAT, SMEM>(sv, __ldg(&sv.hoisted[k]), r, a, rcpA, tmin, tv, seed, pixel, sample, slot, primTests);
}

// The same start with ONE copy of the leaf test in the kernel: the hoisted refs go on the traversal stack above the
// root, the first of them into tv.ref, and the walk loop's own leaf step tests them -- still on a full warp, because
// every lane of a round starts its walk with the same refs in the same order.  They occupy the stack only until the
// root is popped, so the stack needs max(tree depth, n_hoisted) levels, not the sum.  For the instantiations whose
// leaf test is large (media: two boundary queries inlined) the second copy in BeginWalk cost more in instruction
// fetch stalls than the loop overhead it saved.
template <bool SMEM> RT_DEV void BeginWalkStacked(const SceneView<SMEM>& sv, const Stack& stack, Trav& tv)
{
    tv.Begin(sv.root_ref, stack);
    if (sv.n_hoisted > 0) {
        TravPush(stack, tv, sv.root_ref);
        for (int k = sv.n_hoisted - 1; k >= 1; --k) TravPush(stack, tv, __ldg(&sv.hoisted[k]));
        tv.ref = __ldg(&sv.hoisted[0]);
    }
}

// ----------------------------------------------------------------- textures
// Perlin.h:38-139.  Lattice cell and fractions from the FP64 point (octave 6
// scales it by 64, where fp32 would have lost the fraction); the rest fp32.
// (Plain loads, not __ldg: `t` may point at the copy of table 0 that SetupScene stages in shared memory.)
RT_DEV float PerlinNoise(const DevPerlin* t, double px, double py, double pz)
{
    const double fx = floor(px), fy = floor(py), fz = floor(pz);
    const float u = (float)(px - fx), v = (float)(py - fy), w = (float)(pz - fz);
    const int i = (int)fx, j = (int)fy, k = (int)fz;
    const float uu = u * u * (3.0f - 2.0f * u);
    const float vv = v * v * (3.0f - 2.0f * v);
    const float ww = w * w * (3.0f - 2.0f * w);
    float accum = 0.0f;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                const int hsh = t->perm_x[(i + di) & 255] ^ t->perm_y[(j + dj) & 255] ^ t->perm_z[(k + dk) & 255];
                const float4 c = *reinterpret_cast<const float4*>(&t->ranvec[hsh][0]);
                const float wu = di ? uu : 1.0f - uu, wv = dj ? vv : 1.0f - vv, wwt = dk ? ww : 1.0f - ww;
                accum += wu * wv * wwt * (c.x * (u - di) + c.y * (v - dj) + c.z * (w - dk));
            }
    return accum;
}

RT_DEV float PerlinTurb(const DevPerlin* t, d3 p, int depth)
{
    float accum = 0.0f, weight = 1.0f;
    for (int i = 0; i < depth; ++i) {
        accum += weight * PerlinNoise(t, p.x, p.y, p.z);
        weight *= 0.5f;
        p.x *= 2.0;
        p.y *= 2.0;
        p.z *= 2.0;
    }
    return fabsf(accum);
}

// Sphere.h:73-81
RT_DEV void SphereUV(f3 p, float& u, float& v)
{
    const float pi = 3.14159265358979323846f;
    const float theta = acosf(-p.y);
    const float phi = atan2f(-p.z, p.x) + pi;
    u = phi / (2.0f * pi);
    v = theta / pi;
}

// Texture.h:29-176
template <int FEAT, bool SMEM> RT_DEV f3 TextureValue(const SceneView<SMEM>& sv, int tex, const Hit& h, bool sphereLike)
{
    for (int guard = 0; guard < 16; ++guard) {
        const DevTexture* t = &sv.textures[tex];
        const int type = __ldg(&t->type);
        if (type == RT_TEX_SOLID) return make_f3(__ldg(&t->r), __ldg(&t->g), __ldg(&t->b));
        if (type == RT_TEX_CHECKER) {
            // Texture.h:70-81: the parity decision is taken on the FP64 point
            const double inv = __ldg(&t->inv_scale);
            const int xi = (int)floor(inv * h.p.x), yi = (int)floor(inv * h.p.y), zi = (int)floor(inv * h.p.z);
            tex = ((xi + yi + zi) % 2 == 0) ? __ldg(&t->even) : __ldg(&t->odd);
            continue;
        }
        if (!(FEAT & RT_FEAT_TEXTURE_HEAVY)) return make_f3(0.0f, 0.0f, 0.0f); // no image / noise texture in this scene
        if (type == RT_TEX_IMAGE) {
            // Texture.h:110-133: nearest texel, v flipped, cyan when there is no image
            const int idx = __ldg(&t->index);
            if (idx < 0) return make_f3(0.0f, 1.0f, 1.0f);
            const DevImage im = sv.images[idx];
            if (im.height <= 0 || im.width <= 0) return make_f3(0.0f, 1.0f, 1.0f);
            float u = h.u, v = h.v;
            if (sphereLike) {
                f3 n = h.outward;
                if (h.uvFrame) { // back to the object space the reference computes (u,v) in
                    const float fs = __ldg(&sv.uv_frames[h.uvFrame - 1].s), fc = __ldg(&sv.uv_frames[h.uvFrame - 1].c);
                    n = make_f3(fc * n.x - fs * n.z, n.y, fs * n.x + fc * n.z);
                }
                SphereUV(n, u, v);
            }
            u = fminf(fmaxf(u, 0.0f), 1.0f);
            v = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
            int i = (int)(u * (float)im.width), j = (int)(v * (float)im.height);
            if (i >= im.width) i = im.width - 1;
            if (j >= im.height) j = im.height - 1;
            const uint8_t* px = sv.arena + im.offset + ((size_t)j * im.width + i) * 3;
            const float cs = 1.0f / 255.0f;
            return make_f3(cs * __ldg(px), cs * __ldg(px + 1), cs * __ldg(px + 2));
        }
        // Texture.h:159-165: marble
        const int pidx = __ldg(&t->index);
        const DevPerlin* pt = pidx == 0 ? sv.perlin0 : &sv.perlins[pidx];
        // the phase is ~scale*z (tens of radians): reduce it mod 2 pi in FP64, then sinf
        const double phase = fma((double)__ldg(&t->scale), h.p.z, 10.0 * (double)PerlinTurb(pt, h.p, 7));
        const double k = rint(phase * 0.15915494309189535);
        const float red = (float)fma(-k, 1.2246467991473532e-16 * 2.0, fma(-k, 6.283185307179586, phase));
        const float g = 0.5f * (1.0f + __sinf(red)); // |red| <= pi after the FP64 reduction: MUFU.SIN is good to ~4e-7 there
        return make_f3(g, g, g);
    }
    return make_f3(0.0f, 0.0f, 0.0f);
}

// ---------------------------------------------------------------- scatter
// The sequential draws of one (pixel, sample, slot) stream, domain 0 (rt_rng.h).
// Draw number `dim` is lane dim&3 of block dim>>2; the two users below walk the
// blocks explicitly, so no lane selects or per-draw bookkeeping are executed.
struct StreamKey {
    uint32_t seed, pixel, sample, slot;
    RT_DEV rt_u4 Block(uint32_t block) const { return rt_rng_block(seed, pixel, sample, slot, 0, block); }
};
RT_DEV StreamKey MakeKey(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot)
{
    StreamKey k;
    k.seed = seed;
    k.pixel = pixel;
    k.sample = sample;
    k.slot = slot;
    return k;
}

// One candidate of Material.h:14-24 from three 32-bit draws, in two steps.  BallPreTest decides in fp32 on 2u-1 formed
// in one FMA from the bits (within 1e-7 of the exact value; the undecided band is 1e-5 wide): 0 = outside, 1 = inside,
// 2 = too close to call.  BallPoint builds the candidate exactly in FP64 from the fp32 uniform (2x-1 of a 24-bit x) and
// settles the undecided case on the exact length.
RT_DEV int BallPreTest(uint32_t bx, uint32_t by, uint32_t bz)
{
    const float k1 = 4.6566128730773926e-10f, k0 = 2.3283064365386963e-10f - 1.0f; // 2^-31, 2^-32 - 1
    const float fx = fmaf((float)bx, k1, k0), fy = fmaf((float)by, k1, k0), fz = fmaf((float)bz, k1, k0);
    const float l2 = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
    return l2 >= 1.00001f ? 0 : (l2 > 0.99999f ? 2 : 1);
}
RT_DEV bool BallPoint(uint32_t bx, uint32_t by, uint32_t bz, int pre, d3& p)
{
    const float x = rt_bits_to_u01(bx), y = rt_bits_to_u01(by), z = rt_bits_to_u01(bz);
    p = make_d3(fma(2.0, (double)x, -1.0), fma(2.0, (double)y, -1.0), fma(2.0, (double)z, -1.0));
    return pre == 1 || dot(p, p) < 1.0;
}
RT_DEV bool BallCandidate(uint32_t bx, uint32_t by, uint32_t bz, d3& p)
{
    const int pre = BallPreTest(bx, by, bz);
    return pre != 0 && BallPoint(bx, by, bz, pre, p);
}

// Material.h:14-24, drawing from the start of the slot's stream: candidate k
// uses draws 3k..3k+2, i.e. four candidates per three blocks.
//
// A lane returns the FIRST candidate of its sequence that lies in the ball (52 % each), and a warp waits for its
// unluckiest lane: ~5.8 candidate evaluations for 1.9 needed on average.  RT_BALL_EAGER evaluates the first two
// candidates' draws and fp32 pre-tests unconditionally -- the second costs no issue slot whenever any lane of the
// warp needs it, which is nearly always -- so that only the 23 % of the lanes that fail both enter the loop:
// 2 + E[max over ~7 lanes] = 4.9 evaluations per call instead of 5.8.  Same draws, same result.
#ifndef RT_BALL_EAGER
#define RT_BALL_EAGER 2 /* candidates pre-tested on every lane before the sequential loop: 0, 2 or 4 */
#endif
// RT_BALL_STRUCTURED (default): the same search written with ONE exit.  In the form above every candidate is a
// `return` with its own copy of BallPoint behind it: the lanes of a warp leave through eight different exits and the
// FP64 construction of the point runs once per exit that was taken, each time on the few lanes that took it (ncu, Cornell
// smoke scene: BallPoint on 3.9 lanes, the whole function 21.6 % of the kernel's instructions on 6.5 lanes).  Here the
// search only selects the accepted candidate's three 32-bit draws -- two candidates on every lane, then nested ifs
// that close before the loop turns -- and the point is built once, after the warp has reconverged.  Same draws, same
// candidate, same point.
#ifndef RT_BALL_STRUCTURED
#define RT_BALL_STRUCTURED 1
#endif
RT_DEV bool BallAccept(uint32_t bx, uint32_t by, uint32_t bz)
{
    const int pre = BallPreTest(bx, by, bz);
    if (pre == 2) { // within 1e-5 of the surface (one candidate in ~1e5): decided on the exact FP64 length
        d3 p;
        return BallPoint(bx, by, bz, 2, p);
    }
    return pre == 1;
}
#if RT_BALL_STRUCTURED == 4
// A/B: FOUR candidates (three blocks) on every lane before the loop; 5 % of the lanes go on.  Measured slower than two
// in this form as well: Book 1 -4.8 %, scene 0 -6.8 %, scene 9 -6.0 % (profiles/r2_ab_zc.jsonl).
RT_DEV d3 RandomInUnitSphere(const StreamKey& key)
{
    uint32_t bx, by, bz;
    bool found;
    {
        const rt_u4 b0 = key.Block(0u), b1 = key.Block(1u), b2 = key.Block(2u);
        const bool ok0 = BallAccept(b0.x, b0.y, b0.z), ok1 = BallAccept(b0.w, b1.x, b1.y);
        const bool ok2 = BallAccept(b1.z, b1.w, b2.x), ok3 = BallAccept(b2.y, b2.z, b2.w);
        bx = ok0 ? b0.x : (ok1 ? b0.w : (ok2 ? b1.z : b2.y));
        by = ok0 ? b0.y : (ok1 ? b1.x : (ok2 ? b1.w : b2.z));
        bz = ok0 ? b0.z : (ok1 ? b1.y : (ok2 ? b2.x : b2.w));
        found = ok0 || ok1 || ok2 || ok3;
    }
    uint32_t blk = 3u;
    while (!found) {
        const rt_u4 c0 = key.Block(blk);
        bx = c0.x, by = c0.y, bz = c0.z;
        found = BallAccept(bx, by, bz);
        if (!found) {
            const rt_u4 c1 = key.Block(blk + 1u);
            bx = c0.w, by = c1.x, bz = c1.y;
            found = BallAccept(bx, by, bz);
            if (!found) {
                const rt_u4 c2 = key.Block(blk + 2u);
                bx = c1.z, by = c1.w, bz = c2.x;
                found = BallAccept(bx, by, bz);
                if (!found) {
                    bx = c2.y, by = c2.z, bz = c2.w;
                    found = BallAccept(bx, by, bz);
                }
            }
        }
        blk += 3u;
    }
    d3 p;
    BallPoint(bx, by, bz, 1, p);
    return p;
}
#elif RT_BALL_STRUCTURED
RT_DEV d3 RandomInUnitSphere(const StreamKey& key)
{
    const rt_u4 b0 = key.Block(0u);
    rt_u4 bp = key.Block(1u); // the second block of the current group of three (four candidates)
    const bool ok0 = BallAccept(b0.x, b0.y, b0.z), ok1 = BallAccept(b0.w, bp.x, bp.y); // both, on every lane
    uint32_t bx = ok0 ? b0.x : b0.w, by = ok0 ? b0.y : bp.x, bz = ok0 ? b0.z : bp.y;
    bool found = ok0 || ok1;
    uint32_t blk = 2u;
    while (!found) { // 23 % of the lanes; each turn: candidates 2, 3 of this group, then 0, 1 of the next
        const rt_u4 b2 = key.Block(blk);
        bx = bp.z, by = bp.w, bz = b2.x;
        found = BallAccept(bx, by, bz);
        if (!found) {
            bx = b2.y, by = b2.z, bz = b2.w;
            found = BallAccept(bx, by, bz);
            if (!found) {
                const rt_u4 c0 = key.Block(blk + 1u);
                bx = c0.x, by = c0.y, bz = c0.z;
                found = BallAccept(bx, by, bz);
                if (!found) {
                    bp = key.Block(blk + 2u);
                    bx = c0.w, by = bp.x, bz = bp.y;
                    found = BallAccept(bx, by, bz);
                }
            }
        }
        blk += 3u;
    }
    d3 p;
    BallPoint(bx, by, bz, 1, p);
    return p;
}
#else
RT_DEV d3 RandomInUnitSphere(const StreamKey& key)
{
    d3 p;
#if RT_BALL_EAGER
    {
        const rt_u4 b0 = key.Block(0u), b1 = key.Block(1u);
        const int pre0 = BallPreTest(b0.x, b0.y, b0.z), pre1 = BallPreTest(b0.w, b1.x, b1.y); // both, on every lane
#if RT_BALL_EAGER >= 4
        const rt_u4 b2 = key.Block(2u);
        const int pre2 = BallPreTest(b1.z, b1.w, b2.x), pre3 = BallPreTest(b2.y, b2.z, b2.w);
#endif
        if (pre0 != 0 && BallPoint(b0.x, b0.y, b0.z, pre0, p)) return p;
        if (pre1 != 0 && BallPoint(b0.w, b1.x, b1.y, pre1, p)) return p;
#if RT_BALL_EAGER >= 4
        if (pre2 != 0 && BallPoint(b1.z, b1.w, b2.x, pre2, p)) return p;
        if (pre3 != 0 && BallPoint(b2.y, b2.z, b2.w, pre3, p)) return p;
#else
        const rt_u4 b2 = key.Block(2u);
        if (BallCandidate(b1.z, b1.w, b2.x, p)) return p;
        if (BallCandidate(b2.y, b2.z, b2.w, p)) return p;
#endif
    }
    for (uint32_t blk = 3;; blk += 3) {
#else
    for (uint32_t blk = 0;; blk += 3) {
#endif
        const rt_u4 b0 = key.Block(blk);
        if (BallCandidate(b0.x, b0.y, b0.z, p)) return p;
        const rt_u4 b1 = key.Block(blk + 1u);
        if (BallCandidate(b0.w, b1.x, b1.y, p)) return p;
        const rt_u4 b2 = key.Block(blk + 2u);
        if (BallCandidate(b1.z, b1.w, b2.x, p)) return p;
        if (BallCandidate(b2.y, b2.z, b2.w, p)) return p;
    }
}
#endif

RT_DEV d3 Reflect(d3 v, d3 n) { return v - (2.0 * dot(v, n)) * n; } // Vec3.h:122-125

// ------------------------------------------------------ importance sampling
// RT_FLAG_IMPORTANCE (phase 4 of the reference's roadmap, README.md:37-42 -- PDFs, mixture density,
// sampling of lights, orthonormal basis -- which the reference does not implement).  The machinery is the one of
// "Ray Tracing: The Rest of Your Life" (quad::pdf_value / random, sphere::pdf_value / random, onb, mixture_pdf),
// applied to the scattering the reference has: its Lambertian sends the ray to N + (point in the unit ball)
// (Material.h:68-86), a direction density of 2 cos^3(theta) / pi, weighed with the albedo alone; Isotropic
// (Material.h:151-162) is uniform, 1 / 4 pi.  A diffuse bounce draws from
//     1/2 (that density) + 1/2 (density of the directions towards the lights)
// and carries (scattering density) / (mixture density): the same image in expectation, less noise wherever the
// lights are small.  Restated in FP64 by oracle/rt_oracle.cpp ScatterImportance (the parity target).
//
// Density with which LightDirection produces `dir` from `o` (book 3: pdf_value); 0 when the ray misses the light.
template <int FEAT, bool SMEM>
RT_DEV double LightPdf(const SceneView<SMEM>& sv, const DevLight* l, const d3& o, const d3& dir, double len2)
{
    const uint32_t hit = __ldg(&l->hit), index = RT_HIT_INDEX(hit);
    Ray r;
    r.o = o;
    r.d = dir;
    r.time = 0.0f;
    if ((FEAT & RT_FEAT_QUAD) && RT_HIT_TYPE(hit) == RT_LEAF_QUAD) {
        float al, be;
        if (HitQuad<SMEM, double>(sv, index, r, 0.001, 3.402823466e+38f, al, be) == RT_MISS) return 0.0;
        const double t = RefineQuadT<SMEM>(sv, index, r);
        const double2 q2 = LdD2<SMEM>(sv.quads, index * 96u + 32u);
        const double nz = LdD2<SMEM>(sv.quads, index * 96u + 48u).x;
        const double cosine = fabs(fma(q2.x, dir.x, fma(q2.y, dir.y, nz * dir.z))) * RsqrtD(len2);
        return t * t * len2 / (cosine * __ldg(&l->area));
    }
    if (RT_HIT_TYPE(hit) != RT_LEAF_SPHERE) return 0.0;
    if (HitSphere<SMEM, double>(sv, index, r, len2, RcpApprox((float)len2), 0.001, 3.402823466e+38f) == RT_MISS) return 0.0;
    const double cx = __ldg(&l->a[0]) - o.x, cy = __ldg(&l->a[1]) - o.y, cz = __ldg(&l->a[2]) - o.z;
    const double radius = __ldg(&l->b[0]);
    const double cosMax = SqrtD(1.0 - radius * radius * RcpD(fma(cx, cx, fma(cy, cy, cz * cz))));
    return RcpD(6.283185307179586 * (1.0 - cosMax));
}

// A direction from `o` towards light `l` (book 3: quad::random; sphere::random = onb + random_to_sphere).
RT_DEV d3 LightDirection(const DevLight* l, const d3& o, double r1, double r2)
{
    const d3 a = make_d3(__ldg(&l->a[0]), __ldg(&l->a[1]), __ldg(&l->a[2]));
    const d3 b = make_d3(__ldg(&l->b[0]), __ldg(&l->b[1]), __ldg(&l->b[2]));
    if (RT_HIT_TYPE(__ldg(&l->hit)) == RT_LEAF_QUAD) {
        const d3 c = make_d3(__ldg(&l->c[0]), __ldg(&l->c[1]), __ldg(&l->c[2]));
        return make_d3(fma(r2, c.x, fma(r1, b.x, a.x)) - o.x, fma(r2, c.y, fma(r1, b.y, a.y)) - o.y,
                       fma(r2, c.z, fma(r1, b.z, a.z)) - o.z);
    }
    const d3 dir = a - o;
    const double dist2 = dot(dir, dir);
    const d3 w = RsqrtD(dist2) * dir;
    // onb: a = |w.x| > 0.9 ? y : x; v = unit(w x a); u = w x v
    d3 v = fabs(w.x) > 0.9 ? make_d3(-w.z, 0.0, w.x) : make_d3(0.0, w.z, -w.y);
    v = RsqrtD(dot(v, v)) * v;
    const d3 u = make_d3(w.y * v.z - w.z * v.y, w.z * v.x - w.x * v.z, w.x * v.y - w.y * v.x);
    const double z = 1.0 + r2 * (SqrtD(1.0 - b.x * b.x * RcpD(dist2)) - 1.0);
    double sn, cs;
    sincospi(2.0 * r1, &sn, &cs);
    const double rad = SqrtD(1.0 - z * z);
    return (cs * rad) * u + ((sn * rad) * v + z * w);
}

// Lambertian / Isotropic under RT_FLAG_IMPORTANCE.  The four draws (strategy, light, two for the point on it) are
// block 0 of the bounce's own keyed domain (32), so the sequential ball draws of domain 0 keep their positions.
// Returns false when the direction carries nothing.
template <int FEAT, bool SMEM>
RT_DEV bool ScatterImportance(const SceneView<SMEM>& sv, const Hit& h, bool lambertian, const StreamKey& rng, d3& dir, float& weight)
{
    const rt_u4 k = rt_rng_block(rng.seed, rng.pixel, rng.sample, rng.slot, 32u, 0u);
    const int nLights = sv.n_lights;
    if (nLights > 0 && rt_bits_to_u01(k.x) < 0.5f) {
        int li = (int)((double)rt_bits_to_u01(k.y) * (double)nLights);
        if (li > nLights - 1) li = nLights - 1;
        dir = LightDirection(&sv.lights[li], h.p, (double)rt_bits_to_u01(k.z), (double)rt_bits_to_u01(k.w));
    } else {
        const d3 ball = RandomInUnitSphere(rng);
        if (lambertian) {
            dir = h.n + ball;
            if (fabs(dir.x) < 1e-8 && fabs(dir.y) < 1e-8 && fabs(dir.z) < 1e-8) dir = h.n;
        } else {
            dir = RsqrtD(dot(ball, ball)) * ball;
        }
    }
    const double len2 = dot(dir, dir);
    if (!(len2 > 0.0)) return false;
    double pMat = 0.07957747154594767; // 1 / 4 pi
    if (lambertian) {
        const double c = dot(dir, h.n) * RsqrtD(len2);
        pMat = c > 0.0 ? 0.6366197723675814 * c * c * c : 0.0; // 2 cos^3 / pi
    }
    double pdf = pMat;
    if (nLights > 0) {
        double pLight = 0.0;
        for (int i = 0; i < nLights; ++i) pLight += LightPdf<FEAT, SMEM>(sv, &sv.lights[i], h.p, dir, len2);
        pdf = 0.5 * (pLight / (double)nLights) + 0.5 * pMat;
    }
    if (!(pMat > 0.0) || !(pdf > 0.0)) return false;
    weight = (float)(pMat / pdf);
    return true;
}

// Returns false when the path ends (light, or absorbed by metal).  On true,
// `dir` is the scattered direction (not normalised, like the reference) and
// `atten` the attenuation.  `emitted` is Material::Emitted (black unless light).
// `a` = |dirIn|^2.
template <int FEAT, bool SMEM>
RT_DEV bool Scatter(const SceneView<SMEM>& sv, const Hit& h, const d3& dirIn, double a, bool sphereLike, const StreamKey& rng,
                    f3& atten, d3& dir, f3& emitted)
{
    const float4 m0 = Ld4<SMEM>(sv.materials, (uint32_t)h.material * 16u);
    const uint32_t tt = (uint32_t)__float_as_int(m0.w);
    const int type = RT_MAT_TT_TYPE(tt);
    const bool hasParam = type == RT_MAT_METAL || type == RT_MAT_DIELECTRIC;
    const int tex = hasParam ? -1 : (int)RT_MAT_TT_INDEX(tt) - 1;
    double param = 0.0; // fuzz | ior, FP64
    if (hasParam) param = LdD<SMEM>(sv.mat_params, RT_MAT_TT_INDEX(tt) * 8u);
    emitted = make_f3(0.0f, 0.0f, 0.0f);
    f3 colour = make_f3(m0.x, m0.y, m0.z);
    if ((FEAT & RT_FEAT_TEXTURE) && tex >= 0 && type != RT_MAT_METAL && type != RT_MAT_DIELECTRIC)
        colour = TextureValue<FEAT, SMEM>(sv, tex, h, sphereLike);
    if constexpr ((FEAT & RT_FEAT_IMPORTANCE) != 0) {
        if (type == RT_MAT_LAMBERTIAN || type == RT_MAT_ISOTROPIC) {
            float weight = 0.0f;
            if (!ScatterImportance<FEAT, SMEM>(sv, h, type == RT_MAT_LAMBERTIAN, rng, dir, weight)) return false;
            atten = weight * colour;
            return true;
        }
    }
    // Lambertian, Metal and Isotropic all start with RandomInUnitSphere from the top of the slot's stream
    // (Material.h:75,157, Metal.h:27): ONE rejection loop serves the three, instead of one copy per material that
    // the lanes of a warp would run one after the other.
    d3 ball = make_d3(0.0, 0.0, 0.0);
    if (type == RT_MAT_LAMBERTIAN || type == RT_MAT_METAL || type == RT_MAT_ISOTROPIC) ball = RandomInUnitSphere(rng);
    if (type == RT_MAT_LAMBERTIAN) { // Material.h:68-86
        dir = h.n + ball;
        if (fabs(dir.x) < 1e-8 && fabs(dir.y) < 1e-8 && fabs(dir.z) < 1e-8) dir = h.n;
        atten = colour;
        return true;
    }
    if (type == RT_MAT_METAL) { // Metal.h:18-30 (draws even when fuzz == 0)
        const d3 reflected = Reflect(RsqrtD(a) * dirIn, h.n);
        dir = reflected + param * ball;
        atten = colour;
        return dot(dir, h.n) > 0.0;
    }
    if (type == RT_MAT_DIELECTRIC) { // Dielectric.h:18-68, Vec3.h:127-141
        atten = make_f3(1.0f, 1.0f, 1.0f);
        const double ratio = h.front ? RcpD(param) : param;
        const d3 ud = RsqrtD(a) * dirIn;
        const double cosTheta = fmin(-dot(ud, h.n), 1.0);
        const float cosF = (float)cosTheta, ratioF = (float)ratio;
        const float sinTheta = sqrtf(1.0f - cosF * cosF);
        bool reflect = ratioF * sinTheta > 1.0f;
        if (!reflect) { // no draw on total internal reflection
            float r0 = (1.0f - ratioF) / (1.0f + ratioF);
            r0 = r0 * r0;
            const float k = 1.0f - cosF;
            const float k2 = k * k;
            reflect = r0 + (1.0f - r0) * (k2 * k2 * k) > rt_bits_to_u01(rng.Block(0).x);
        }
        if (reflect) {
            dir = Reflect(ud, h.n);
        } else {
            const d3 perp = ratio * (ud + cosTheta * h.n);
            const d3 para = -SqrtD(fabs(1.0 - dot(perp, perp))) * h.n;
            dir = perp + para;
        }
        return true;
    }
    if (type == RT_MAT_ISOTROPIC) { // Material.h:151-162
        dir = RsqrtD(dot(ball, ball)) * ball;
        atten = colour;
        return true;
    }
    emitted = colour; // DiffuseLight, Material.h:116-128
    return false;
}

// ------------------------------------------------------------------- camera
// kernel.cu:140-142 + Camera.h:10-19,76-85: jitter (int + float sum in fp32),
// lens disk by rejection (always drawn), shutter time (always drawn).
RT_DEV Ray CameraRay(const DevCamera& cam, int i, int j, const StreamKey& rng)
{
    // draws 0,1: jitter; then pairs (2,3),(4,5),.. until one lies in the disk; then the time
    rt_u4 b = rng.Block(0);
    const float fu = (float)i + rt_bits_to_u01(b.x);
    const float fv = (float)j + rt_bits_to_u01(b.y);
    const double s = (double)fu * cam.inv_width; // kernel.cu:140-141 divides; the reciprocal is exact to 1 ulp of FP64
    const double t = (double)fv * cam.inv_height;
    // The first three disk candidates -- draws (2,3), (4,5), (6,7), i.e. the rest of block 0 and all of block 1 -- are
    // formed on every lane at once (a few fp32 operations each) and the first one inside the disk is selected; the
    // time is the draw that follows it.  78.5 % of the candidates are accepted, so 1 % of the lanes go on to the
    // sequential loop; the round-1 form entered it for 21 %, and a warp waited for its unluckiest lane.
    const rt_u4 b1 = rng.Block(1);
    const float x0 = 2.0f * rt_bits_to_u01(b.z) - 1.0f, y0 = 2.0f * rt_bits_to_u01(b.w) - 1.0f;
    const float x1 = 2.0f * rt_bits_to_u01(b1.x) - 1.0f, y1 = 2.0f * rt_bits_to_u01(b1.y) - 1.0f;
    const float x2 = 2.0f * rt_bits_to_u01(b1.z) - 1.0f, y2 = 2.0f * rt_bits_to_u01(b1.w) - 1.0f;
    const bool in0 = x0 * x0 + y0 * y0 < 1.0f, in1 = x1 * x1 + y1 * y1 < 1.0f, in2 = x2 * x2 + y2 * y2 < 1.0f;
    float px = in0 ? x0 : (in1 ? x1 : x2), py = in0 ? y0 : (in1 ? y1 : y2);
    uint32_t timeBits = in0 ? b1.x : b1.z; // (in2: the first draw of block 2, fetched below)
    if (!in0 && !in1) {
        if (in2) {
            timeBits = rng.Block(2).x;
        } else {
            for (uint32_t blk = 2;; ++blk) {
                b = rng.Block(blk);
                px = 2.0f * rt_bits_to_u01(b.x) - 1.0f;
                py = 2.0f * rt_bits_to_u01(b.y) - 1.0f;
                if (px * px + py * py < 1.0f) {
                    timeBits = b.z;
                    break;
                }
                px = 2.0f * rt_bits_to_u01(b.z) - 1.0f;
                py = 2.0f * rt_bits_to_u01(b.w) - 1.0f;
                if (px * px + py * py < 1.0f) {
                    timeBits = rng.Block(blk + 1u).x;
                    break;
                }
            }
        }
    }
    const double rdx = cam.lens_radius * (double)px, rdy = cam.lens_radius * (double)py;
    const double offx = cam.u[0] * rdx + cam.v[0] * rdy;
    const double offy = cam.u[1] * rdx + cam.v[1] * rdy;
    const double offz = cam.u[2] * rdx + cam.v[2] * rdy;
    Ray r;
    r.time = cam.time0 + rt_bits_to_u01(timeBits) * (cam.time1 - cam.time0);
    r.o.x = cam.origin[0] + offx;
    r.o.y = cam.origin[1] + offy;
    r.o.z = cam.origin[2] + offz;
    // primary directions stay FP64 (scattered ones are fp32 values widened)
    r.d.x = cam.llc[0] + s * cam.horiz[0] + t * cam.vert[0] - cam.origin[0] - offx;
    r.d.y = cam.llc[1] + s * cam.horiz[1] + t * cam.vert[1] - cam.origin[1] - offy;
    r.d.z = cam.llc[2] + s * cam.horiz[2] + t * cam.vert[2] - cam.origin[2] - offz;
    return r;
}

} // namespace rtdev
