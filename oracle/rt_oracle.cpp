// rt_oracle.cpp -- CPU restatement of the reference's render path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product (the package, the C-ABI
// library, the CLI) may include, link or call this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// do, and only as the checker.
//
// What it is: the algorithm of
//     Render -> RayColor -> BvhNode::Hit / Material::Scatter
// (reference RayTracinginOneWeekend/kernel.cu:65-154 and the headers it
// includes) written again as plain scalar C++ over the flat scene description
// of include/rt_abi.h.  Arithmetic is FP64 with the reference's expression
// order (built with -ffp-contract=off), traversal uses the reference's BVH
// topology and visit order, and every function cites the reference lines it
// follows.  The one deliberate difference is the random stream: the reference
// draws from a per-pixel cuRAND XORWOW state; this file draws from the
// counter-based stream specified in include/rt_rng.h (restated independently
// below), which is what "identical RNG streams" means for the parity tests.
//
// Pinning: oracle/build_ref.py compiles the reference's own headers
// (unmodified except for the sequencing patches listed there) against the same
// stream into oracle/_ref/ref_stream; tests/test_oracle_pin.py checks this
// file against it bit-for-bit, and against the golden renders committed under
// tests/golden/ (made by ref_stream) where /root/reference is absent.
//
// The float instantiation (precision=32) is a numerics study tool: it answers
// "how often does fp32 flip a discrete decision" without a GPU.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/rt_abi.h"

namespace {

const double kPiValue = 3.1415926535897932385;

// ----------------------------------------------------------------- counters
struct Stats {
    uint64_t rays = 0, paths = 0, box_tests = 0, sphere_tests = 0, quad_tests = 0, medium_tests = 0, draws = 0;
    void operator+=(const Stats& o)
    {
        rays += o.rays;
        paths += o.paths;
        box_tests += o.box_tests;
        sphere_tests += o.sphere_tests;
        quad_tests += o.quad_tests;
        medium_tests += o.medium_tests;
        draws += o.draws;
    }
};

// --------------------------------------------------- random stream (rt_rng.h)
// Independent restatement of the stream spec: PCG4D over
// (pixel, sample, slot | block<<8 | domain<<16, seed), four dims per block,
// bits -> (0,1] float like curand_uniform (curand_uniform.h:69-72).
struct Stream {
    uint32_t seed = 0, pixel = 0, sample = 0, slot = 0, domain = 0, dim = 0;
    uint32_t out[4] = {0, 0, 0, 0};
    Stats* stats = nullptr;

    static void Pcg4d(uint32_t v[4])
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
    static float ToUniform(uint32_t bits) { return (float)bits * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }
    static float At(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t domain, uint32_t dim)
    {
        uint32_t v[4] = {pixel, sample, (slot & 0xffu) | (((dim >> 2) & 0xffu) << 8) | (domain << 16), seed};
        Pcg4d(v);
        return ToUniform(v[dim & 3u]);
    }
    void Begin(uint32_t slot_)
    {
        slot = slot_;
        domain = 0;
        dim = 0;
    }
    float Next()
    {
        const uint32_t lane = dim & 3u;
        if (lane == 0) {
            out[0] = pixel;
            out[1] = sample;
            out[2] = (slot & 0xffu) | (((dim >> 2) & 0xffu) << 8) | (domain << 16);
            out[3] = seed;
            Pcg4d(out);
        }
        ++dim;
        if (stats) ++stats->draws;
        return ToUniform(out[lane]);
    }
    // The single draw of medium `id` on its `visit`-th test in this bounce.
    float Medium(uint32_t id, uint32_t visit) const
    {
        if (stats) ++stats->draws;
        return At(seed, pixel, sample, slot, 1u + 2u * id + visit, 0);
    }
    // Importance sampling: draw `dim` (0 strategy, 1 light index, 2, 3 the point on the light) of the bounce's own
    // keyed domain, so that the sequential scatter draws of domain 0 keep their positions.
    float Importance(uint32_t d) const
    {
        if (stats) ++stats->draws;
        return At(seed, pixel, sample, slot, 32u, d);
    }
};

// --------------------------------------------------------------------- math
template <class R> struct V3 {
    R e[3];
    V3() : e{0, 0, 0} {}
    V3(R x, R y, R z) : e{x, y, z} {}
    R operator[](int i) const { return e[i]; }
    R& operator[](int i) { return e[i]; }
    V3 operator-() const { return V3(-e[0], -e[1], -e[2]); }
    R LengthSquared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    R Length() const { return std::sqrt(LengthSquared()); }
};
template <class R> V3<R> operator+(const V3<R>& a, const V3<R>& b) { return V3<R>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template <class R> V3<R> operator-(const V3<R>& a, const V3<R>& b) { return V3<R>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template <class R> V3<R> operator*(const V3<R>& a, const V3<R>& b) { return V3<R>(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
template <class R> V3<R> operator*(R s, const V3<R>& v) { return V3<R>(s * v[0], s * v[1], s * v[2]); }
template <class R> V3<R> operator*(const V3<R>& v, R s) { return s * v; }
// Vec3.h:96-99: division multiplies by the reciprocal.
template <class R> V3<R> operator/(const V3<R>& v, R s) { return (R(1) / s) * v; }
template <class R> R Dot(const V3<R>& a, const V3<R>& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class R> V3<R> Cross(const V3<R>& a, const V3<R>& b)
{
    return V3<R>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
template <class R> V3<R> Unit(const V3<R>& v) { return v / v.Length(); }
template <class R> V3<R> FromD(const double* p) { return V3<R>((R)p[0], (R)p[1], (R)p[2]); }

template <class R> struct Limits;
template <> struct Limits<double> {
    static double Max() { return DBL_MAX; }
};
template <> struct Limits<float> {
    static float Max() { return FLT_MAX; }
};

template <class R> struct Ray {
    V3<R> o, d;
    R time = 0;
    V3<R> At(R t) const { return o + t * d; }
};

template <class R> struct Box {
    R lo[3], hi[3];
    // AABB.h:68-98: three divides, branch-free fmin/fmax, strict tMax > tMin.
    bool Hit(const Ray<R>& r, R tMin, R tMax) const
    {
        for (int a = 0; a < 3; ++a) {
            const R inv = R(1) / r.d[a];
            const R t0 = (lo[a] - r.o[a]) * inv;
            const R t1 = (hi[a] - r.o[a]) * inv;
            tMin = std::fmax(tMin, std::fmin(t0, t1));
            tMax = std::fmin(tMax, std::fmax(t0, t1));
        }
        return tMax > tMin;
    }
};

// Hittable.h:11-31
template <class R> struct HitRec {
    V3<R> p, n;
    R t = 0, u = 0, v = 0;
    bool front = false;
    int material = -1;
    // Hittable.h:26-30
    void SetFaceNormal(const Ray<R>& r, const V3<R>& outward)
    {
        front = Dot(r.d, outward) < R(0);
        n = front ? outward : -outward;
    }
};

// ------------------------------------------------------------------ the scene
template <class R> struct Prim {
    int type, material, first_xform, xform_count;
    V3<R> a, b, c;
    R radius, time0, time1;
    // Quad.h:33-36 cached constants
    V3<R> normal, w;
    R D;
};

template <class R> struct Xform {
    int type;
    V3<R> offset;
    R sin_t, cos_t;
};

struct BvhNodeRef {
    // child >= 0: index of an internal node; child < 0: leaf object ~child
    int left, right;
    int box; // index into node boxes
};

template <class R> struct Scene {
    std::vector<Prim<R>> prims;
    std::vector<Xform<R>> xforms;
    std::vector<rt_object> objects;
    std::vector<Box<R>> object_box;
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    const rt_perlin* perlins = nullptr;
    const rt_image* images = nullptr;
    // reference-topology BVH
    std::vector<BvhNodeRef> nodes;
    std::vector<Box<R>> node_box;
    int root = -1;
    std::vector<int> medium_visits;
    // Importance sampling (the reference's roadmap, README.md:37-42, unimplemented there): the sampling
    // targets are the quads and spheres with a DiffuseLight material, in primitive order (media boundaries excluded).
    bool importance = false;
    std::vector<int> lights;
};

// BvhNode.h:50-90 + DeviceSort/BoxCompare :170-193, on object indices.
template <class R>
int BuildReferenceBvh(Scene<R>& s, std::vector<int>& order, int start, int end, const std::vector<double>& bbox)
{
    // node box: union starting from the empty interval (Interval.h:13-17)
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = start; i < end; ++i) {
        const double* b = &bbox[(size_t)order[i] * 6];
        for (int a = 0; a < 3; ++a) {
            lo[a] = lo[a] <= b[2 * a] ? lo[a] : b[2 * a];
            hi[a] = hi[a] >= b[2 * a + 1] ? hi[a] : b[2 * a + 1];
        }
    }
    // AABB.h:101-107
    const double sx = hi[0] - lo[0], sy = hi[1] - lo[1], sz = hi[2] - lo[2];
    const int axis = (sx > sy) ? (sx > sz ? 0 : 2) : (sy > sz ? 1 : 2);

    const int me = (int)s.nodes.size();
    s.nodes.push_back(BvhNodeRef{0, 0, me});
    Box<R> nb;
    for (int a = 0; a < 3; ++a) {
        nb.lo[a] = (R)lo[a];
        nb.hi[a] = (R)hi[a];
    }
    s.node_box.push_back(nb);

    const int span = end - start;
    int left, right;
    if (span == 1) {
        left = right = ~order[start];
    } else if (span == 2) {
        left = ~order[start];
        right = ~order[start + 1];
    } else {
        for (int i = start + 1; i < end; ++i) {
            const int key = order[i];
            const double keyMin = bbox[(size_t)key * 6 + 2 * axis];
            int j = i - 1;
            while (j >= start && keyMin < bbox[(size_t)order[j] * 6 + 2 * axis]) {
                order[j + 1] = order[j];
                --j;
            }
            order[j + 1] = key;
        }
        const int mid = start + span / 2;
        left = BuildReferenceBvh(s, order, start, mid, bbox);
        right = BuildReferenceBvh(s, order, mid, end, bbox);
    }
    s.nodes[me].left = left;
    s.nodes[me].right = right;
    return me;
}

template <class R> bool LoadScene(const rt_scene_desc* d, Scene<R>& s)
{
    if (!d || d->abi_version != RT_ABI_VERSION || d->n_objects <= 0) return false;
    s.prims.resize(d->n_prims);
    for (int i = 0; i < d->n_prims; ++i) {
        const rt_prim& p = d->prims[i];
        Prim<R>& q = s.prims[i];
        q.type = p.type;
        q.material = p.material;
        q.first_xform = p.first_xform;
        q.xform_count = p.xform_count;
        q.a = FromD<R>(p.a);
        q.b = FromD<R>(p.b);
        q.c = FromD<R>(p.c);
        q.radius = (R)p.radius;
        q.time0 = (R)p.time0;
        q.time1 = (R)p.time1;
        q.D = 0;
        if (p.type == RT_PRIM_QUAD) {
            // Quad.h:31-36
            const V3<R> n = Cross(q.b, q.c);
            q.normal = Unit(n);
            q.D = Dot(q.normal, q.a);
            q.w = n / Dot(n, n);
        }
    }
    s.xforms.resize(d->n_xforms);
    for (int i = 0; i < d->n_xforms; ++i) {
        const rt_xform& x = d->xforms[i];
        s.xforms[i].type = x.type;
        s.xforms[i].offset = FromD<R>(x.v);
        s.xforms[i].sin_t = (R)x.v[0];
        s.xforms[i].cos_t = (R)x.v[1];
    }
    s.objects.assign(d->objects, d->objects + d->n_objects);
    s.materials.assign(d->materials, d->materials + d->n_materials);
    s.textures.assign(d->textures, d->textures + d->n_textures);
    s.perlins = d->perlins;
    s.images = d->images;

    std::vector<double> bbox((size_t)d->n_objects * 6);
    s.object_box.resize(d->n_objects);
    int n_media = 0;
    for (int i = 0; i < d->n_objects; ++i) {
        for (int k = 0; k < 6; ++k) bbox[(size_t)i * 6 + k] = d->objects[i].bbox[k];
        for (int a = 0; a < 3; ++a) {
            s.object_box[i].lo[a] = (R)d->objects[i].bbox[2 * a];
            s.object_box[i].hi[a] = (R)d->objects[i].bbox[2 * a + 1];
        }
        if (d->objects[i].kind == RT_OBJ_MEDIUM) n_media = std::max(n_media, d->objects[i].medium_id + 1);
    }
    std::vector<int> order(d->n_objects);
    for (int i = 0; i < d->n_objects; ++i) order[i] = i;
    s.nodes.clear();
    s.node_box.clear();
    s.root = BuildReferenceBvh(s, order, 0, d->n_objects, bbox);
    s.lights.clear();
    for (int o = 0; o < d->n_objects; ++o) {
        if (d->objects[o].kind == RT_OBJ_MEDIUM) continue;
        for (int k = 0; k < d->objects[o].prim_count; ++k) {
            const int i = d->objects[o].first_prim + k;
            const rt_prim& p = d->prims[i];
            if ((p.type == RT_PRIM_QUAD || p.type == RT_PRIM_SPHERE) && d->materials[p.material].type == RT_MAT_DIFFUSE_LIGHT)
                s.lights.push_back(i);
        }
    }
    std::sort(s.lights.begin(), s.lights.end());
    // trap T2: how often the reference topology references each medium leaf
    s.medium_visits.assign(n_media, 0);
    for (const BvhNodeRef& n : s.nodes) {
        const int kids[2] = {n.left, n.right};
        for (int c = 0; c < 2; ++c)
            if (kids[c] < 0 && s.objects[~kids[c]].kind == RT_OBJ_MEDIUM) ++s.medium_visits[s.objects[~kids[c]].medium_id];
    }
    return true;
}

// ------------------------------------------------------------- primitives
const double kPi = 3.1415926535897932385;

// Sphere.h:73-81
template <class R> void SphereUV(const V3<R>& p, R& u, R& v)
{
    const R pi = (R)kPi;
    const R theta = std::acos(-p[1]);
    const R phi = std::atan2(-p[2], p[0]) + pi;
    u = phi / (R(2) * pi);
    v = theta / pi;
}

// Sphere.h:22-70 with `center` supplied (MovingSphere.h:44-102 lerps it first).
template <class R>
bool HitSphereAt(const V3<R>& center, R radius, int material, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec)
{
    const V3<R> oc = r.o - center;
    const R a = Dot(r.d, r.d);
    const R b = Dot(oc, r.d);
    const R c = Dot(oc, oc) - radius * radius;
    const R disc = b * b - a * c;
    if (disc > R(0)) {
        R t = (-b - std::sqrt(disc)) / a;
        for (int root = 0; root < 2; ++root) {
            if (t < tMax && t > tMin) {
                rec.t = t;
                rec.p = r.At(t);
                const V3<R> outward = (rec.p - center) / radius;
                rec.SetFaceNormal(r, outward);
                SphereUV(outward, rec.u, rec.v);
                rec.material = material;
                return true;
            }
            t = (-b + std::sqrt(disc)) / a;
        }
    }
    return false;
}

// Quad.h:54-99
template <class R> bool HitQuad(const Prim<R>& q, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec)
{
    const R denom = Dot(q.normal, r.d);
    if (std::fabs(denom) < R(1e-8)) return false;
    const R t = (q.D - Dot(q.normal, r.o)) / denom;
    if (t < tMin || t > tMax) return false;
    const V3<R> hit = r.At(t);
    const V3<R> planar = hit - q.a;
    const R alpha = Dot(q.w, Cross(planar, q.c));
    const R beta = Dot(q.w, Cross(q.b, planar));
    // Interval::Contains (Interval.h:37-40): closed [0,1]
    if (!(R(0) <= alpha && alpha <= R(1)) || !(R(0) <= beta && beta <= R(1))) return false;
    rec.u = alpha;
    rec.v = beta;
    rec.t = t;
    rec.p = hit;
    rec.material = q.material;
    rec.SetFaceNormal(r, q.normal);
    return true;
}

template <class R> bool HitBarePrim(const Prim<R>& p, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec, Stats& st)
{
    switch (p.type) {
    case RT_PRIM_SPHERE:
        ++st.sphere_tests;
        return HitSphereAt(p.a, p.radius, p.material, r, tMin, tMax, rec);
    case RT_PRIM_MOVING_SPHERE: {
        ++st.sphere_tests;
        // MovingSphere.h:52-53
        const R frac = (r.time - p.time0) / (p.time1 - p.time0);
        const V3<R> center = p.a + frac * (p.b - p.a);
        return HitSphereAt(center, p.radius, p.material, r, tMin, tMax, rec);
    }
    default:
        ++st.quad_tests;
        return HitQuad(p, r, tMin, tMax, rec);
    }
}

// Instance.h:41-56 (Translate::Hit) and :116-150 (RotateY::Hit): move the ray
// into object space outermost wrapper first, test, move the record back.
template <class R>
bool HitPrim(const Scene<R>& s, const Prim<R>& p, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec, Stats& st)
{
    if (p.xform_count == 0) return HitBarePrim(p, r, tMin, tMax, rec, st);
    Ray<R> lr = r;
    for (int k = 0; k < p.xform_count; ++k) {
        const Xform<R>& x = s.xforms[p.first_xform + k];
        if (x.type == RT_XFORM_TRANSLATE) {
            lr.o = lr.o - x.offset;
        } else {
            const V3<R> o(x.cos_t * lr.o[0] - x.sin_t * lr.o[2], lr.o[1], x.sin_t * lr.o[0] + x.cos_t * lr.o[2]);
            const V3<R> d(x.cos_t * lr.d[0] - x.sin_t * lr.d[2], lr.d[1], x.sin_t * lr.d[0] + x.cos_t * lr.d[2]);
            lr.o = o;
            lr.d = d;
        }
    }
    if (!HitBarePrim(p, lr, tMin, tMax, rec, st)) return false;
    for (int k = p.xform_count - 1; k >= 0; --k) {
        const Xform<R>& x = s.xforms[p.first_xform + k];
        if (x.type == RT_XFORM_TRANSLATE) {
            rec.p = rec.p + x.offset;
        } else {
            rec.p = V3<R>(x.cos_t * rec.p[0] + x.sin_t * rec.p[2], rec.p[1], -x.sin_t * rec.p[0] + x.cos_t * rec.p[2]);
            rec.n = V3<R>(x.cos_t * rec.n[0] + x.sin_t * rec.n[2], rec.n[1], -x.sin_t * rec.n[0] + x.cos_t * rec.n[2]);
        }
    }
    return true;
}

// HittableList.h:39-57: linear closest hit over an object's primitives.
template <class R>
bool HitPrimRange(const Scene<R>& s, int first, int count, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec, Stats& st)
{
    HitRec<R> tmp;
    bool any = false;
    R closest = tMax;
    for (int i = 0; i < count; ++i) {
        if (HitPrim(s, s.prims[first + i], r, tMin, closest, tmp, st)) {
            any = true;
            closest = tmp.t;
            rec = tmp;
        }
    }
    return any;
}

// ConstantMedium.h:52-94.  The uniform comes from the keyed medium stream.
template <class R>
bool HitMedium(const Scene<R>& s, const rt_object& o, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec,
               const Stream& rng, int visit, Stats& st)
{
    ++st.medium_tests;
    const R big = Limits<R>::Max();
    HitRec<R> r1, r2;
    if (!HitPrimRange(s, o.first_prim, o.prim_count, r, -big, big, r1, st)) return false;
    if (!HitPrimRange(s, o.first_prim, o.prim_count, r, r1.t + R(0.0001), big, r2, st)) return false;
    if (r1.t < tMin) r1.t = tMin;
    if (r2.t > tMax) r2.t = tMax;
    if (r1.t >= r2.t) return false;
    if (r1.t < R(0)) r1.t = R(0);
    const R rayLength = r.d.Length();
    const R inside = (r2.t - r1.t) * rayLength;
    const R negInvDensity = R(-1.0) / (R)o.density; // ConstantMedium.h:22
    // `log(curand_uniform(..))` has a float argument: under nvcc that is the
    // float overload (logf), widened afterwards (ConstantMedium.h:79).
    const R hitDistance = negInvDensity * (R)std::log(rng.Medium((uint32_t)o.medium_id, (uint32_t)visit));
    if (hitDistance > inside) return false;
    rec.t = r1.t + hitDistance / rayLength;
    rec.p = r.At(rec.t);
    rec.n = V3<R>(1, 0, 0);
    rec.front = true;
    rec.material = o.phase_material;
    return true;
}

template <class R>
bool HitObject(const Scene<R>& s, int obj, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec, const Stream& rng,
               int* mediumVisit, Stats& st)
{
    const rt_object& o = s.objects[obj];
    if (o.kind == RT_OBJ_MEDIUM) return HitMedium(s, o, r, tMin, tMax, rec, rng, mediumVisit[o.medium_id]++, st);
    if (o.kind == RT_OBJ_PRIM) return HitPrim(s, s.prims[o.first_prim], r, tMin, tMax, rec, st);
    return HitPrimRange(s, o.first_prim, o.prim_count, r, tMin, tMax, rec, st);
}

// BvhNode.h:101-158: explicit 32-entry stack, box test on pop, both children
// examined left then right, leaves tested at once with the shrinking `closest`.
template <class R>
bool HitWorldBvh(const Scene<R>& s, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec, const Stream& rng, Stats& st)
{
    int mediumVisit[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int stack[32];
    int sp = 0;
    int node = s.root;
    bool any = false;
    R closest = tMax;
    while (true) {
        ++st.box_tests;
        if (s.node_box[s.nodes[node].box].Hit(r, tMin, closest)) {
            int next = -1;
            const int kids[2] = {s.nodes[node].left, s.nodes[node].right};
            for (int c = 0; c < 2; ++c) {
                const int kid = kids[c];
                if (kid >= 0) {
                    if (next < 0)
                        next = kid;
                    else if (sp < 32)
                        stack[sp++] = kid;
                } else if (HitObject(s, ~kid, r, tMin, closest, rec, rng, mediumVisit, st)) {
                    any = true;
                    closest = rec.t;
                }
            }
            if (next >= 0) {
                node = next;
                continue;
            }
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
    return any;
}

// The reference's own cross-check (Docs/2권_3장:772): the same world as a plain
// list.  Media get their reference-topology visit count so the two agree.
template <class R>
bool HitWorldList(const Scene<R>& s, const Ray<R>& r, R tMin, R tMax, HitRec<R>& rec, const Stream& rng, Stats& st)
{
    int mediumVisit[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool any = false;
    R closest = tMax;
    for (int i = 0; i < (int)s.objects.size(); ++i) {
        const int reps = s.objects[i].kind == RT_OBJ_MEDIUM ? s.medium_visits[s.objects[i].medium_id] : 1;
        for (int k = 0; k < reps; ++k) {
            HitRec<R> tmp = rec;
            if (HitObject(s, i, r, tMin, closest, tmp, rng, mediumVisit, st)) {
                any = true;
                closest = tmp.t;
                rec = tmp;
            }
        }
    }
    return any;
}

// ------------------------------------------------------------------ textures
// Perlin.h:119-139
template <class R> R PerlinInterp(const V3<R> c[2][2][2], R u, R v, R w)
{
    const R uu = u * u * (R(3) - R(2) * u);
    const R vv = v * v * (R(3) - R(2) * v);
    const R ww = w * w * (R(3) - R(2) * w);
    R accum = 0;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                const V3<R> weight(u - i, v - j, w - k);
                accum += (i * uu + (1 - i) * (1 - uu)) * (j * vv + (1 - j) * (1 - vv)) * (k * ww + (1 - k) * (1 - ww)) *
                         Dot(c[i][j][k], weight);
            }
    return accum;
}

// Perlin.h:38-64
template <class R> R PerlinNoise(const rt_perlin& t, const V3<R>& p)
{
    const R u = p[0] - std::floor(p[0]);
    const R v = p[1] - std::floor(p[1]);
    const R w = p[2] - std::floor(p[2]);
    const int i = int(std::floor(p[0]));
    const int j = int(std::floor(p[1]));
    const int k = int(std::floor(p[2]));
    V3<R> c[2][2][2];
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                const int h = t.perm_x[(i + di) & 255] ^ t.perm_y[(j + dj) & 255] ^ t.perm_z[(k + dk) & 255];
                c[di][dj][dk] = V3<R>((R)t.ranvec[h][0], (R)t.ranvec[h][1], (R)t.ranvec[h][2]);
            }
    return PerlinInterp(c, u, v, w);
}

// Perlin.h:67-80
template <class R> R PerlinTurb(const rt_perlin& t, const V3<R>& p, int depth)
{
    R accum = 0;
    V3<R> q = p;
    R weight = 1;
    for (int i = 0; i < depth; ++i) {
        accum += weight * PerlinNoise(t, q);
        weight *= R(0.5);
        q = R(2) * q;
    }
    return std::fabs(accum);
}

template <class R> V3<R> TextureValue(const Scene<R>& s, int tex, R u, R v, const V3<R>& p)
{
    for (int guard = 0; guard < 64; ++guard) {
        const rt_texture& t = s.textures[tex];
        switch (t.type) {
        case RT_TEX_SOLID: // Texture.h:47-50
            return V3<R>((R)t.color[0], (R)t.color[1], (R)t.color[2]);
        case RT_TEX_CHECKER: { // Texture.h:70-81
            const R inv = R(1.0) / (R)t.scale;
            const int xi = int(std::floor(inv * p[0]));
            const int yi = int(std::floor(inv * p[1]));
            const int zi = int(std::floor(inv * p[2]));
            tex = ((xi + yi + zi) % 2 == 0) ? t.even : t.odd;
            continue;
        }
        case RT_TEX_IMAGE: { // Texture.h:110-133
            if (t.image < 0 || s.images[t.image].height <= 0 || s.images[t.image].rgb == nullptr)
                return V3<R>(0, 1, 1);
            const rt_image& im = s.images[t.image];
            R uu = u < R(0) ? R(0) : (u > R(1) ? R(1) : u);
            R vc = v < R(0) ? R(0) : (v > R(1) ? R(1) : v);
            R vv = R(1.0) - vc;
            int i = int(uu * im.width);
            int j = int(vv * im.height);
            if (i >= im.width) i = im.width - 1;
            if (j >= im.height) j = im.height - 1;
            const uint8_t* px = im.rgb + ((size_t)j * im.width + i) * 3;
            const R cs = R(1.0) / R(255.0);
            return V3<R>(cs * px[0], cs * px[1], cs * px[2]);
        }
        default: { // Texture.h:159-165
            const rt_perlin& pt = s.perlins[t.perlin];
            return V3<R>(R(0.5), R(0.5), R(0.5)) * (R(1.0) + std::sin((R)t.scale * p[2] + R(10.0) * PerlinTurb(pt, p, 7)));
        }
        }
    }
    return V3<R>(0, 0, 0);
}

// ----------------------------------------------------------------- materials
// Material.h:14-24: rejection sampling in the cube, x then y then z.
template <class R> V3<R> RandomInUnitSphere(Stream& rng)
{
    V3<R> p;
    do {
        const R x = (R)rng.Next();
        const R y = (R)rng.Next();
        const R z = (R)rng.Next();
        p = R(2.0) * V3<R>(x, y, z) - V3<R>(1, 1, 1);
    } while (p.LengthSquared() >= R(1.0));
    return p;
}

template <class R> V3<R> Reflect(const V3<R>& v, const V3<R>& n) { return v - R(2.0) * Dot(v, n) * n; } // Vec3.h:122-125

// Vec3.h:127-141
template <class R> V3<R> Refract(const V3<R>& uv, const V3<R>& n, R eta)
{
    const R cosTheta = std::fmin(Dot(-uv, n), R(1.0));
    const V3<R> perp = eta * (uv + cosTheta * n);
    const V3<R> para = -std::sqrt(std::fabs(R(1.0) - perp.LengthSquared())) * n;
    return perp + para;
}

template <class R> bool NearZero(const V3<R>& v)
{
    const R th = R(1e-8);
    return std::fabs(v[0]) < th && std::fabs(v[1]) < th && std::fabs(v[2]) < th;
}

// ------------------------------------------------- importance sampling (f4)
// The machinery of "Ray Tracing: The Rest of Your Life" (P. Shirley et al., v4.0.1, the book the reference's roadmap
// names for its phase 4: PDFs, mixture density, sampling of lights, orthonormal basis) applied to the scattering the
// reference HAS: its Lambertian sends the ray to N + (point in the unit ball) (Material.h:68-86), whose direction
// density is 2 cos^3(theta) / pi (chord of the ball along the direction, cubed, over the ball's volume), and weighs
// it with the albedo alone -- i.e. the scattering function is albedo * 2 cos^3 / pi.  Sampling the mixture
//   1/2 * (that density) + 1/2 * (density of the directions towards the lights)
// and weighing with (scattering density) / (mixture density) estimates the SAME image with less noise wherever
// lights are small.  Isotropic (Material.h:151-162) is uniform on the sphere: 1 / 4 pi.

// Object -> world for a point / a vector of primitive p (Instance.h:136-147 the other way round).
template <class R> V3<R> ToWorld(const Scene<R>& s, const Prim<R>& p, V3<R> v, bool isPoint)
{
    for (int k = p.xform_count - 1; k >= 0; --k) {
        const Xform<R>& x = s.xforms[p.first_xform + k];
        if (x.type == RT_XFORM_TRANSLATE) {
            if (isPoint) v = v + x.offset;
        } else {
            v = V3<R>(x.cos_t * v[0] + x.sin_t * v[2], v[1], -x.sin_t * v[0] + x.cos_t * v[2]);
        }
    }
    return v;
}

// Book 3, quad::pdf_value / sphere::pdf_value: density (per solid angle, seen from `o`) with which light.Random
// produces direction `d`; 0 when the ray (o, d) misses the light.
template <class R> R LightPdf(const Scene<R>& s, int prim, const V3<R>& o, const V3<R>& d, Stats& st)
{
    const Prim<R>& p = s.prims[prim];
    Ray<R> r;
    r.o = o;
    r.d = d;
    HitRec<R> rec;
    if (!HitPrim(s, p, r, R(0.001), Limits<R>::Max(), rec, st)) return R(0);
    if (p.type == RT_PRIM_QUAD) {
        const R area = Cross(p.b, p.c).Length();
        const R dist2 = rec.t * rec.t * d.LengthSquared();
        const R cosine = std::fabs(Dot(d, rec.n) / d.Length());
        return dist2 / (cosine * area);
    }
    const V3<R> c = ToWorld(s, p, p.a, true);
    const R cosMax = std::sqrt(R(1) - p.radius * p.radius / (c - o).LengthSquared());
    return R(1) / (R(2) * (R)kPiValue * (R(1) - cosMax));
}

// Book 3, quad::random / sphere::random (onb + random_to_sphere): a direction from `o` towards the light.
template <class R> V3<R> LightDirection(const Scene<R>& s, int prim, const V3<R>& o, R r1, R r2)
{
    const Prim<R>& p = s.prims[prim];
    if (p.type == RT_PRIM_QUAD) return ToWorld(s, p, p.a + r1 * p.b + r2 * p.c, true) - o;
    const V3<R> dir = ToWorld(s, p, p.a, true) - o;
    const R dist2 = dir.LengthSquared();
    // onb: w = unit(dir), a = |w.x| > 0.9 ? y : x, v = unit(w x a), u = w x v
    const V3<R> w = Unit(dir);
    const V3<R> a = std::fabs(w[0]) > R(0.9) ? V3<R>(0, 1, 0) : V3<R>(1, 0, 0);
    const V3<R> v = Unit(Cross(w, a));
    const V3<R> u = Cross(w, v);
    const R z = R(1) + r2 * (std::sqrt(R(1) - p.radius * p.radius / dist2) - R(1));
    const R phi = R(2) * (R)kPiValue * r1;
    const R rad = std::sqrt(R(1) - z * z);
    return (std::cos(phi) * rad) * u + (std::sin(phi) * rad) * v + z * w;
}

// The scattered direction and the weight (scattering density / mixture density) of a Lambertian or Isotropic hit.
// Returns false when the weight is 0 (the direction carries nothing).
template <class R>
bool ScatterImportance(const Scene<R>& s, bool lambertian, const HitRec<R>& rec, V3<R>& dir, R& weight, Stream& rng, Stats& st)
{
    const int nLights = (int)s.lights.size();
    const R u0 = (R)rng.Importance(0);
    if (nLights > 0 && u0 < R(0.5)) {
        int k = (int)((R)rng.Importance(1) * (R)nLights);
        if (k > nLights - 1) k = nLights - 1;
        const R r1 = (R)rng.Importance(2), r2 = (R)rng.Importance(3);
        dir = LightDirection(s, s.lights[(size_t)k], rec.p, r1, r2);
    } else {
        const V3<R> ball = RandomInUnitSphere<R>(rng);
        if (lambertian) {
            dir = rec.n + ball;
            if (NearZero(dir)) dir = rec.n;
        } else {
            dir = Unit(ball);
        }
    }
    const R len = dir.Length();
    if (!(len > R(0))) return false;
    R pMat;
    if (lambertian) {
        const R c = Dot(dir, rec.n) / len;
        pMat = c > R(0) ? R(2) * c * c * c / (R)kPiValue : R(0);
    } else {
        pMat = R(1) / (R(4) * (R)kPiValue);
    }
    R pdf = pMat;
    if (nLights > 0) {
        R pLight = 0;
        for (int k = 0; k < nLights; ++k) pLight = pLight + LightPdf(s, s.lights[(size_t)k], rec.p, dir, st);
        pdf = R(0.5) * (pLight / (R)nLights) + R(0.5) * pMat;
    }
    if (!(pMat > R(0)) || !(pdf > R(0))) return false;
    weight = pMat / pdf;
    return true;
}

template <class R>
bool Scatter(const Scene<R>& s, const Ray<R>& in, const HitRec<R>& rec, V3<R>& atten, Ray<R>& out, Stream& rng, Stats& st)
{
    const rt_material& m = s.materials[rec.material];
    out.time = in.time;
    out.o = rec.p;
    if (s.importance && (m.type == RT_MAT_LAMBERTIAN || m.type == RT_MAT_ISOTROPIC)) {
        R weight = 0;
        if (!ScatterImportance(s, m.type == RT_MAT_LAMBERTIAN, rec, out.d, weight, rng, st)) return false;
        atten = weight * TextureValue(s, m.texture, rec.u, rec.v, rec.p);
        return true;
    }
    switch (m.type) {
    case RT_MAT_LAMBERTIAN: { // Material.h:68-86
        V3<R> dir = rec.n + RandomInUnitSphere<R>(rng);
        if (NearZero(dir)) dir = rec.n;
        out.d = dir;
        atten = TextureValue(s, m.texture, rec.u, rec.v, rec.p);
        return true;
    }
    case RT_MAT_METAL: { // Metal.h:18-30: draws even when fuzz == 0
        const V3<R> reflected = Reflect(Unit(in.d), rec.n);
        out.d = reflected + (R)m.fuzz * RandomInUnitSphere<R>(rng);
        atten = V3<R>((R)m.albedo[0], (R)m.albedo[1], (R)m.albedo[2]);
        return Dot(out.d, rec.n) > R(0);
    }
    case RT_MAT_DIELECTRIC: { // Dielectric.h:18-54: no draw on total internal reflection
        atten = V3<R>(1, 1, 1);
        const R ratio = rec.front ? (R(1.0) / (R)m.ior) : (R)m.ior;
        const V3<R> unit = Unit(in.d);
        const R cosTheta = std::fmin(Dot(-unit, rec.n), R(1.0));
        const R sinTheta = std::sqrt(R(1.0) - cosTheta * cosTheta);
        const bool cannot = ratio * sinTheta > R(1.0);
        bool reflect = cannot;
        if (!reflect) {
            // Dielectric.h:60-68 Schlick
            R r0 = (R(1.0) - ratio) / (R(1.0) + ratio);
            r0 = r0 * r0;
            const R refl = r0 + (R(1.0) - r0) * std::pow(R(1.0) - cosTheta, R(5.0));
            reflect = refl > (R)rng.Next();
        }
        out.d = reflect ? Reflect(unit, rec.n) : Refract(unit, rec.n, ratio);
        return true;
    }
    case RT_MAT_ISOTROPIC: // Material.h:151-162
        out.d = Unit(RandomInUnitSphere<R>(rng));
        atten = TextureValue(s, m.texture, rec.u, rec.v, rec.p);
        return true;
    default: // DiffuseLight, Material.h:121-128
        return false;
    }
}

// ------------------------------------------------------------------- camera
// Camera.h:36-71
template <class R> struct Cam {
    V3<R> origin, llc, horiz, vert, u, v, w, background;
    R lensRadius, time0, time1;
    int W, H, maxDepth;

    explicit Cam(const rt_camera& c)
    {
        W = c.image_width;
        H = c.image_height;
        maxDepth = c.max_depth;
        const double aspect = double(c.image_width) / double(c.image_height);
        double aperture = c.aperture;
        if (aperture < 0.0)
            aperture = 2.0 * c.focus_dist * std::tan(c.defocus_angle * 3.14159265358979323846 / 360.0);
        background = FromD<R>(c.background);
        time0 = (R)c.time0;
        time1 = (R)c.time1;
        lensRadius = (R)(aperture / 2.0);
        const R theta = (R)(c.vfov * 3.14159265358979323846 / 180.0);
        const R halfHeight = std::tan(theta / R(2.0));
        const R halfWidth = (R)aspect * halfHeight;
        const V3<R> from = FromD<R>(c.lookfrom), at = FromD<R>(c.lookat), vup = FromD<R>(c.vup);
        const R fd = (R)c.focus_dist;
        w = Unit(from - at);
        u = Unit(Cross(vup, w));
        v = Cross(w, u);
        origin = from;
        llc = origin - halfWidth * fd * u - halfHeight * fd * v - fd * w;
        horiz = R(2.0) * halfWidth * fd * u;
        vert = R(2.0) * halfHeight * fd * v;
    }

    // Camera.h:76-85 with RandomInUnitDisk :10-19 (always drawn), then time.
    Ray<R> GetRay(R s, R t, Stream& rng) const
    {
        V3<R> p;
        do {
            const R x = (R)rng.Next();
            const R y = (R)rng.Next();
            p = R(2.0) * V3<R>(x, y, 0) - V3<R>(1, 1, 0);
        } while (Dot(p, p) >= R(1.0));
        const V3<R> rd = lensRadius * p;
        const V3<R> offset = u * rd[0] + v * rd[1];
        Ray<R> r;
        r.time = time0 + (R)rng.Next() * (time1 - time0);
        r.o = origin + offset;
        r.d = llc + s * horiz + t * vert - origin - offset;
        return r;
    }
};

// --------------------------------------------------------------- integrator
// kernel.cu:65-98
template <class R>
V3<R> RayColor(const Scene<R>& s, const Cam<R>& cam, Ray<R> ray, Stream& rng, bool useBvh, Stats& st)
{
    V3<R> throughput(1, 1, 1), accumulated(0, 0, 0);
    for (int bounce = 0; bounce < cam.maxDepth; ++bounce) {
        rng.Begin((uint32_t)bounce + 1u);
        ++st.rays;
        HitRec<R> rec;
        const bool hit = useBvh ? HitWorldBvh(s, ray, R(0.001), Limits<R>::Max(), rec, rng, st)
                                : HitWorldList(s, ray, R(0.001), Limits<R>::Max(), rec, rng, st);
        if (!hit) {
            accumulated = accumulated + throughput * cam.background;
            return accumulated;
        }
        const rt_material& m = s.materials[rec.material];
        if (m.type == RT_MAT_DIFFUSE_LIGHT) { // Material.h:116-119; everything else emits black (:33-36)
            accumulated = accumulated + throughput * TextureValue(s, m.texture, rec.u, rec.v, rec.p);
        } else {
            accumulated = accumulated + throughput * V3<R>(0, 0, 0);
        }
        Ray<R> scattered;
        V3<R> atten;
        if (!Scatter(s, ray, rec, atten, scattered, rng, st)) return accumulated;
        throughput = throughput * atten;
        ray = scattered;
    }
    return accumulated;
}

// kernel.cu:122-154, one row.  Output is the linear SUM over the samples
// rendered (the reference divides by numSamples and takes sqrt afterwards,
// :147-152; parity is defined on linear radiance).
// Window: pixels [x0, x0+w) x [y0, y0+h) of the W x H frame.  Pixel indices (and so the random streams) are the
// GLOBAL ones, pixel = j*W + i; `out` holds only the window, row-major, w*h*3.
struct Window {
    int x0, y0, w, h;
};

template <class R>
void RenderRows(const Scene<R>& s, const Cam<R>& cam, const Window& win, int j0, int j1, int s0, int s1, uint32_t seed,
                bool useBvh, double* out, Stats& st)
{
    Stream rng;
    rng.seed = seed;
    rng.stats = &st;
    for (int j = j0; j < j1; ++j) {
        for (int i = win.x0; i < win.x0 + win.w; ++i) {
            const int pixel = j * cam.W + i;
            rng.pixel = (uint32_t)pixel;
            V3<R> col(0, 0, 0);
            for (int smp = s0; smp < s1; ++smp) {
                rng.sample = (uint32_t)smp;
                rng.Begin(0);
                ++st.paths;
                // `i + curand_uniform()` is an int+float sum, i.e. fp32 (kernel.cu:140-141)
                const float fu = (float)i + rng.Next();
                const float fv = (float)j + rng.Next();
                const R u = (R)((double)fu / double(cam.W));
                const R v = (R)((double)fv / double(cam.H));
                const Ray<R> r = cam.GetRay(u, v, rng);
                col = col + RayColor(s, cam, r, rng, useBvh, st);
            }
            double* o = out + ((size_t)(j - win.y0) * win.w + (size_t)(i - win.x0)) * 3;
            o[0] = (double)col[0];
            o[1] = (double)col[1];
            o[2] = (double)col[2];
        }
    }
}

template <class R>
int RenderT(const rt_scene_desc* d, const rt_camera* c, const Window* window, int s0, int s1, uint32_t seed, int mode,
            int nThreads, double* out, Stats& total)
{
    const bool useBvh = (mode & 1) != 0;
    Scene<R> s;
    if (!LoadScene(d, s)) return -1;
    s.importance = (mode & 2) != 0;
    const Cam<R> cam(*c);
    if (nThreads < 1) nThreads = 1;
    std::vector<Stats> st((size_t)nThreads);
    Window win{0, 0, cam.W, cam.H};
    if (window) win = *window;
    if (win.x0 < 0 || win.y0 < 0 || win.w <= 0 || win.h <= 0 || win.x0 + win.w > cam.W || win.y0 + win.h > cam.H) return -2;
    std::atomic<int> nextRow(win.y0);
    const int rowsPerGrab = 4, rowEnd = win.y0 + win.h;
    auto work = [&](int tid) {
        while (true) {
            const int j0 = nextRow.fetch_add(rowsPerGrab);
            if (j0 >= rowEnd) break;
            RenderRows(s, cam, win, j0, std::min(rowEnd, j0 + rowsPerGrab), s0, s1, seed, useBvh, out, st[(size_t)tid]);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nThreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    for (const Stats& x : st) total += x;
    return 0;
}

} // namespace

extern "C" {

struct oracle_stats {
    uint64_t rays, paths, box_tests, sphere_tests, quad_tests, medium_tests, draws;
    int32_t n_nodes, n_objects;
    int32_t medium_visits[8];
};

// out: W*H*3 doubles, row 0 = bottom row, linear radiance SUM over [s0,s1).
// bvh: bit 0: 1 = reference-topology BVH (BvhNode.h), 0 = linear list; bit 1 (value 2): importance sampling of the
// lights (ScatterImportance above; the default, 0, is the reference's scattering).
// precision: 64 = the oracle; 32 = float study build of the same code.
static int OracleRender(const rt_scene_desc* scene, const rt_camera* cam, const Window* win, int sample_begin,
                        int sample_end, uint32_t seed, int bvh, int precision, int n_threads, double* out,
                        struct oracle_stats* stats);

int oracle_render(const rt_scene_desc* scene, const rt_camera* cam, int sample_begin, int sample_end, uint32_t seed,
                  int bvh, int precision, int n_threads, double* out, oracle_stats* stats)
{
    return OracleRender(scene, cam, nullptr, sample_begin, sample_end, seed, bvh, precision, n_threads, out, stats);
}

// The same for a pixel window of the frame: out = w*h*3 doubles (row 0 = the window's bottom row).  The streams are
// keyed on the GLOBAL pixel index, so the window of a 4K frame is rendered exactly as the full frame would render it,
// at the cost of the window only (full-size parity checks of the BASELINE configs).
int oracle_render_region(const rt_scene_desc* scene, const rt_camera* cam, int x0, int y0, int w, int h,
                         int sample_begin, int sample_end, uint32_t seed, int bvh, int precision, int n_threads,
                         double* out, oracle_stats* stats)
{
    const Window win{x0, y0, w, h};
    return OracleRender(scene, cam, &win, sample_begin, sample_end, seed, bvh, precision, n_threads, out, stats);
}

static int OracleRender(const rt_scene_desc* scene, const rt_camera* cam, const Window* win, int sample_begin,
                        int sample_end, uint32_t seed, int bvh, int precision, int n_threads, double* out,
                        oracle_stats* stats)
{
    if (!scene || !cam || !out || cam->image_width <= 0 || cam->image_height <= 0) return -1;
    Stats st;
    int rc;
    if (precision == 32)
        rc = RenderT<float>(scene, cam, win, sample_begin, sample_end, seed, bvh, n_threads, out, st);
    else
        rc = RenderT<double>(scene, cam, win, sample_begin, sample_end, seed, bvh, n_threads, out, st);
    if (rc != 0) return rc;
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        stats->rays = st.rays;
        stats->paths = st.paths;
        stats->box_tests = st.box_tests;
        stats->sphere_tests = st.sphere_tests;
        stats->quad_tests = st.quad_tests;
        stats->medium_tests = st.medium_tests;
        stats->draws = st.draws;
        Scene<double> s;
        if (LoadScene(scene, s)) {
            stats->n_nodes = (int32_t)s.nodes.size();
            stats->n_objects = (int32_t)s.objects.size();
            for (size_t k = 0; k < s.medium_visits.size() && k < 8; ++k) stats->medium_visits[k] = s.medium_visits[k];
        }
    }
    return 0;
}

float oracle_rng_uniform(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t slot, uint32_t domain, uint32_t dim)
{
    return Stream::At(seed, pixel, sample, slot, domain, dim);
}

// Reference-topology BVH as a flat listing: for node k, out[3k] = left,
// out[3k+1] = right (>=0 node index, <0 = ~object index), out[3k+2] = depth.
int oracle_bvh_topology(const rt_scene_desc* scene, int32_t* out, int32_t capacity_nodes)
{
    Scene<double> s;
    if (!LoadScene(scene, s)) return -1;
    const int n = (int)s.nodes.size();
    if (out) {
        std::vector<int> depth((size_t)n, 0);
        for (int k = 0; k < n && k < capacity_nodes; ++k) {
            out[3 * k] = s.nodes[k].left;
            out[3 * k + 1] = s.nodes[k].right;
            out[3 * k + 2] = depth[k];
            if (s.nodes[k].left >= 0) depth[s.nodes[k].left] = depth[k] + 1;
            if (s.nodes[k].right >= 0) depth[s.nodes[k].right] = depth[k] + 1;
        }
    }
    return n;
}

// Debug aid: the path of (pixel i,j; sample) bounce by bounce, 8 doubles per
// record {hit(1/0), t, material, front, p.x, p.y, p.z, material type}; returns
// the number of records written.  Follows RayColor (kernel.cu:65-98) like above.
int oracle_trace_path(const rt_scene_desc* scene, const rt_camera* cam, int i, int j, int sample, uint32_t seed,
                      double* records, int max_records)
{
    Scene<double> s;
    if (!LoadScene(scene, s) || !cam || !records) return -1;
    const Cam<double> c(*cam);
    Stats st;
    Stream rng;
    rng.seed = seed;
    rng.stats = &st;
    rng.pixel = (uint32_t)(j * c.W + i);
    rng.sample = (uint32_t)sample;
    rng.Begin(0);
    const float fu = (float)i + rng.Next();
    const float fv = (float)j + rng.Next();
    Ray<double> ray = c.GetRay((double)fu / double(c.W), (double)fv / double(c.H), rng);
    int n = 0;
    for (int bounce = 0; bounce < c.maxDepth && n < max_records; ++bounce) {
        rng.Begin((uint32_t)bounce + 1u);
        HitRec<double> rec;
        const bool hit = HitWorldBvh(s, ray, 0.001, Limits<double>::Max(), rec, rng, st);
        double* o = records + 8 * n++;
        o[0] = hit ? 1.0 : 0.0;
        if (!hit) break;
        o[1] = rec.t;
        o[2] = rec.material;
        o[3] = rec.front ? 1.0 : 0.0;
        o[4] = rec.p[0];
        o[5] = rec.p[1];
        o[6] = rec.p[2];
        o[7] = s.materials[rec.material].type;
        Ray<double> scattered;
        V3<double> atten;
        if (!Scatter(s, ray, rec, atten, scattered, rng, st)) break;
        ray = scattered;
    }
    return n;
}

// Texture lookup exactly as the integrator does it (unit tests).
int oracle_texture_value(const rt_scene_desc* scene, int texture, double u, double v, const double* p, double* rgb)
{
    Scene<double> s;
    if (!LoadScene(scene, s) || texture < 0 || texture >= (int)s.textures.size()) return -1;
    const V3<double> c = TextureValue<double>(s, texture, u, v, V3<double>(p[0], p[1], p[2]));
    rgb[0] = c[0];
    rgb[1] = c[1];
    rgb[2] = c[2];
    return 0;
}

// Importance-sampling pieces, for unit tests of their normalisation: n directions -> the light list's density
// (mean over the lights, as ScatterImportance mixes it) seen from `origin`; and n directions drawn towards light
// `light` (index into the light list) from pairs of uniforms.  Return the number of lights, or -1.
int oracle_light_pdf(const rt_scene_desc* scene, const double* origin, const double* dirs, int n, double* pdf)
{
    Scene<double> s;
    if (!LoadScene(scene, s)) return -1;
    Stats st;
    const V3<double> o(origin[0], origin[1], origin[2]);
    for (int k = 0; k < n; ++k) {
        const V3<double> d(dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]);
        double sum = 0.0;
        for (int l : s.lights) sum += LightPdf(s, l, o, d, st);
        pdf[k] = s.lights.empty() ? 0.0 : sum / (double)s.lights.size();
    }
    return (int)s.lights.size();
}

int oracle_light_direction(const rt_scene_desc* scene, int light, const double* origin, const double* r12, int n, double* dirs)
{
    Scene<double> s;
    if (!LoadScene(scene, s) || light < 0 || light >= (int)s.lights.size()) return -1;
    const V3<double> o(origin[0], origin[1], origin[2]);
    for (int k = 0; k < n; ++k) {
        const V3<double> d = LightDirection(s, s.lights[(size_t)light], o, r12[2 * k], r12[2 * k + 1]);
        dirs[3 * k] = d[0];
        dirs[3 * k + 1] = d[1];
        dirs[3 * k + 2] = d[2];
    }
    return (int)s.lights.size();
}

} // extern "C"
