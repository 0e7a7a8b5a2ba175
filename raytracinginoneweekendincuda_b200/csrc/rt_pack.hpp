// rt_pack.hpp -- host side of rt_scene_upload: bake, build the BVH, pack.
//
// Input is the flat FP64 scene description of include/rt_abi.h (the
// reference's object graph as its constructors leave it).  Output is the
// device layout of rt_device_types.h:
//   1. instances are baked: every primitive is moved to world space through
//      its Translate/RotateY chain (reference Instance.h:41-56,116-150 run the
//      other way round: they move the ray into object space per test);
//   2. a BVH is built over the baked primitives -- binned SAH by default, the
//      reference's median-split topology (BvhNode.h:50-90) or a plain list
//      (the reference's BVH==list cross-check) on request;
//   3. primitives are re-ordered so each leaf is a contiguous run of one type.
// Host only, no CUDA.
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rt_abi.h"
#include "rt_device_types.h"

namespace rtpack {

struct Box3 {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
    double hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    void Grow(const double* p)
    {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], p[a]);
            hi[a] = std::max(hi[a], p[a]);
        }
    }
    void Grow(const Box3& b)
    {
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], b.lo[a]);
            hi[a] = std::max(hi[a], b.hi[a]);
        }
    }
    double Area() const
    {
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0 || dy < 0 || dz < 0) return 0.0;
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
};

// A baked primitive, still FP64, with its world box.
struct Baked {
    int type;     // RT_LEAF_SPHERE / MOVING / QUAD
    int material;
    double a[3], b[3], c[3]; // sphere: centre | moving: c0, c1 | quad: Q, u, v
    double radius, time0, time1;
    double yawSin, yawCos; // accumulated RotateY of the instance chain (object -> world); 0, 1 when there is none
    bool rotated;
    Box3 box;
};

// Something the BVH treats as one leaf entry.
struct Item {
    int type;  // RT_LEAF_* (RT_LEAF_BOX: six baked quads that form a closed box, tested as one DevBox)
    int index; // SAH: index into baked[] (or media[] for RT_LEAF_MEDIUM)
    int first, count; // REFERENCE/LIST: run of baked prims (all of `type`)
    Box3 box;
};

struct BuildNode {
    Box3 box;
    int left = -1, right = -1;    // children (BuildNode indices) or -1
    std::vector<int> items;       // leaf: item indices (same type)
};

struct Packed {
    std::vector<DevNode> nodes;
    std::vector<DevSphere> spheres;
    std::vector<int32_t> sphere_material; // parallel to spheres: read once per ray, after traversal
    std::vector<DevMovingSphere> moving;
    std::vector<DevQuad> quads;
    std::vector<DevBox> boxes;
    std::vector<DevMedium> media;
    std::vector<DevMaterial> materials;
    std::vector<double> mat_params;
    std::vector<DevUvFrame> uv_frames;
    std::vector<DevLight> lights; // quads / spheres with a DiffuseLight material, in primitive order
    std::vector<DevTexture> textures;
    std::vector<DevPerlin> perlins;
    std::vector<std::vector<uint8_t>> image_bytes;
    std::vector<int32_t> image_w, image_h;
    uint32_t root_ref = 0;
    uint32_t hoisted[RT_MAX_HOISTED] = {0, 0, 0, 0}; // leaf refs tested before the tree is entered
    int n_hoisted = 0;
    int features = 0;
    int max_depth = 0;
    int medium_visits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n_media = 0;
};

inline void ApplyChainPoint(const rt_scene_desc& d, const rt_prim& p, double* v, bool isPoint)
{
    // object -> world: innermost wrapper first (chain is stored outermost first)
    for (int k = p.xform_count - 1; k >= 0; --k) {
        const rt_xform& x = d.xforms[p.first_xform + k];
        if (x.type == RT_XFORM_TRANSLATE) {
            if (isPoint)
                for (int a = 0; a < 3; ++a) v[a] += x.v[a];
        } else {
            // Instance.h:136-147: x = c x' + s z', z = -s x' + c z'
            const double s = x.v[0], c = x.v[1];
            const double nx = c * v[0] + s * v[2];
            const double nz = -s * v[0] + c * v[2];
            v[0] = nx;
            v[2] = nz;
        }
    }
}

inline Baked Bake(const rt_scene_desc& d, const rt_prim& p)
{
    Baked b;
    std::memset(&b, 0, sizeof b);
    b.box = Box3();
    b.yawSin = 0.0;
    b.yawCos = 1.0;
    for (int k = p.xform_count - 1; k >= 0; --k) { // rotations about one axis compose by adding angles
        const rt_xform& x = d.xforms[p.first_xform + k];
        if (x.type != RT_XFORM_ROTATE_Y) continue;
        const double s = x.v[0], c = x.v[1];
        const double ns = s * b.yawCos + c * b.yawSin, nc = c * b.yawCos - s * b.yawSin;
        b.yawSin = ns;
        b.yawCos = nc;
        b.rotated = true;
    }
    b.material = p.material;
    b.radius = p.radius;
    b.time0 = p.time0;
    b.time1 = p.time1;
    for (int a = 0; a < 3; ++a) {
        b.a[a] = p.a[a];
        b.b[a] = p.b[a];
        b.c[a] = p.c[a];
    }
    if (p.type == RT_PRIM_SPHERE) {
        b.type = RT_LEAF_SPHERE;
        ApplyChainPoint(d, p, b.a, true);
        for (int a = 0; a < 3; ++a) {
            b.box.lo[a] = b.a[a] - p.radius;
            b.box.hi[a] = b.a[a] + p.radius;
        }
    } else if (p.type == RT_PRIM_MOVING_SPHERE) {
        b.type = RT_LEAF_MOVING;
        ApplyChainPoint(d, p, b.a, true);
        ApplyChainPoint(d, p, b.b, true);
        for (int a = 0; a < 3; ++a) {
            b.box.lo[a] = std::min(b.a[a], b.b[a]) - p.radius;
            b.box.hi[a] = std::max(b.a[a], b.b[a]) + p.radius;
        }
    } else if (p.type == RT_PRIM_QUAD) {
        b.type = RT_LEAF_QUAD;
        ApplyChainPoint(d, p, b.a, true);
        ApplyChainPoint(d, p, b.b, false);
        ApplyChainPoint(d, p, b.c, false);
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                double corner[3];
                for (int a = 0; a < 3; ++a) corner[a] = b.a[a] + i * b.b[a] + j * b.c[a];
                b.box.Grow(corner);
            }
    } else {
        throw std::invalid_argument("unknown primitive type");
    }
    return b;
}

// ---------------------------------------------------------------- builders
struct Builder {
    std::vector<Item> items;
    std::vector<BuildNode> nodes;
    int maxLeaf = 2;
    int maxDepth = 0;

    static double IsectCost(int type)
    {
        switch (type) {
        case RT_LEAF_SPHERE: return 1.5;
        case RT_LEAF_MOVING: return 1.8;
        case RT_LEAF_QUAD: return 1.3;
        case RT_LEAF_BOX: return 2.5;
        default: return 12.0;
        }
    }

    // (The cost model is flat around these values: primitive costs scaled by 0.5 / 2 / 4 against the node cost of 1, with
    // leaves of at most 1 / 2 / 4 primitives, move Book 1 and scene 0 by < 0.5 % and the Book 2 final scene by < 2.5 %:
    // profiles/r2_ab_zb.jsonl.)
    static double ItemCost(const Item& it) { return IsectCost(it.type) * std::max(1, it.count); } // (count: prims of a whole list)
    static constexpr int kSurfaceTypes[4] = {RT_LEAF_SPHERE, RT_LEAF_MOVING, RT_LEAF_QUAD, RT_LEAF_BOX};

    int NewLeaf(const std::vector<int>& ids, const Box3& box, int depth = 0)
    {
        // a leaf holds one primitive type, and a medium is always a leaf of its
        // own; mixed sets are chained through internal nodes sharing `box`
        std::vector<std::vector<int>> groups;
        for (int t : kSurfaceTypes) {
            std::vector<int> g;
            for (int id : ids)
                if (items[id].type == t) g.push_back(id);
            if (!g.empty()) groups.push_back(g);
        }
        for (int id : ids)
            if (items[id].type == RT_LEAF_MEDIUM) groups.push_back(std::vector<int>(1, id));
        int made = -1;
        for (const std::vector<int>& g : groups) {
            BuildNode leaf;
            leaf.box = Box3();
            for (int id : g) leaf.box.Grow(items[id].box);
            leaf.items = g;
            nodes.push_back(leaf);
            const int me = (int)nodes.size() - 1;
            if (made < 0) {
                made = me;
            } else {
                BuildNode join;
                join.box = box;
                join.left = made;
                join.right = me;
                nodes.push_back(join);
                made = (int)nodes.size() - 1;
            }
        }
        // a mixed leaf is a chain of joins: each one is a level the traversal stack must hold
        maxDepth = std::max(maxDepth, depth + (int)groups.size() - 1);
        return made;
    }

    // Binned SAH, kBins bins, all three axes.
    int BuildSah(std::vector<int>& ids, int depth)
    {
        maxDepth = std::max(maxDepth, depth);
        Box3 box, cbox;
        for (int id : ids) {
            box.Grow(items[id].box);
            double c[3];
            for (int a = 0; a < 3; ++a) c[a] = 0.5 * (items[id].box.lo[a] + items[id].box.hi[a]);
            cbox.Grow(c);
        }
        const int n = (int)ids.size();
        if (n == 1) return NewLeaf(ids, box, depth);

        double leafCost = 0.0;
        for (int id : ids) leafCost += ItemCost(items[id]);

        const int kBins = 64;
        double bestCost = DBL_MAX;
        int bestAxis = -1, bestBin = -1;
        const double invArea = box.Area() > 0 ? 1.0 / box.Area() : 0.0;
        for (int axis = 0; axis < 3; ++axis) {
            const double ext = cbox.hi[axis] - cbox.lo[axis];
            if (!(ext > 0)) continue;
            Box3 bb[kBins];
            double bc[kBins];
            int bn[kBins];
            for (int k = 0; k < kBins; ++k) {
                bc[k] = 0;
                bn[k] = 0;
            }
            for (int id : ids) {
                const double c = 0.5 * (items[id].box.lo[axis] + items[id].box.hi[axis]);
                int k = (int)(kBins * (c - cbox.lo[axis]) / ext);
                k = std::min(std::max(k, 0), kBins - 1);
                bb[k].Grow(items[id].box);
                bc[k] += ItemCost(items[id]);
                bn[k]++;
            }
            double rightArea[kBins], rightCost[kBins];
            Box3 acc;
            double cacc = 0;
            for (int k = kBins - 1; k > 0; --k) {
                acc.Grow(bb[k]);
                cacc += bc[k];
                rightArea[k] = acc.Area();
                rightCost[k] = cacc;
            }
            acc = Box3();
            cacc = 0;
            int nl = 0;
            for (int k = 0; k < kBins - 1; ++k) {
                acc.Grow(bb[k]);
                cacc += bc[k];
                nl += bn[k];
                if (nl == 0 || nl == n) continue;
                const double cost = 1.0 + invArea * (acc.Area() * cacc + rightArea[k + 1] * rightCost[k + 1]);
                if (cost < bestCost) {
                    bestCost = cost;
                    bestAxis = axis;
                    bestBin = k;
                }
            }
        }
        const bool depthLeft = depth + (int)std::ceil(std::log2((double)std::max(n, 2))) < 26;
        if (n <= maxLeaf && (bestAxis < 0 || leafCost <= bestCost)) return NewLeaf(ids, box, depth);

        std::vector<int> L, R;
        if (bestAxis >= 0 && depthLeft) {
            const double ext = cbox.hi[bestAxis] - cbox.lo[bestAxis];
            for (int id : ids) {
                const double c = 0.5 * (items[id].box.lo[bestAxis] + items[id].box.hi[bestAxis]);
                int k = (int)(kBins * (c - cbox.lo[bestAxis]) / ext);
                k = std::min(std::max(k, 0), kBins - 1);
                (k <= bestBin ? L : R).push_back(id);
            }
        }
        if (L.empty() || R.empty()) {
            // degenerate (coincident centroids) or depth budget: median split on the widest axis
            int axis = 0;
            for (int a = 1; a < 3; ++a)
                if (box.hi[a] - box.lo[a] > box.hi[axis] - box.lo[axis]) axis = a;
            std::vector<int> sorted = ids;
            std::stable_sort(sorted.begin(), sorted.end(), [&](int x, int y) {
                return items[x].box.lo[axis] + items[x].box.hi[axis] < items[y].box.lo[axis] + items[y].box.hi[axis];
            });
            L.assign(sorted.begin(), sorted.begin() + n / 2);
            R.assign(sorted.begin() + n / 2, sorted.end());
        }
        const int l = BuildSah(L, depth + 1);
        const int r = BuildSah(R, depth + 1);
        BuildNode in;
        in.box = box;
        in.left = l;
        in.right = r;
        nodes.push_back(in);
        return (int)nodes.size() - 1;
    }

    // BvhNode.h:50-90 on item order: longest axis of the union box (ties fall
    // towards Z, AABB.h:101-107), stable insertion sort on box-min with strict
    // `<` (BvhNode.h:170-193), midpoint split.  visits[i] counts how often item
    // i is referenced (span-1 nodes reference their leaf twice).
    int BuildReference(std::vector<int>& order, int start, int end, int depth, std::vector<int>& visits)
    {
        maxDepth = std::max(maxDepth, depth);
        Box3 box;
        for (int i = start; i < end; ++i) box.Grow(items[order[i]].box);
        const double sx = box.hi[0] - box.lo[0], sy = box.hi[1] - box.lo[1], sz = box.hi[2] - box.lo[2];
        const int axis = (sx > sy) ? (sx > sz ? 0 : 2) : (sy > sz ? 1 : 2);
        const int span = end - start;
        if (span == 1) {
            visits[order[start]] += 2;
            std::vector<int> one(1, order[start]);
            return NewLeaf(one, box, depth);
        }
        int l, r;
        if (span == 2) {
            visits[order[start]] += 1;
            visits[order[start + 1]] += 1;
            std::vector<int> a(1, order[start]), b(1, order[start + 1]);
            l = NewLeaf(a, items[order[start]].box, depth + 1);
            r = NewLeaf(b, items[order[start + 1]].box, depth + 1);
        } else {
            for (int i = start + 1; i < end; ++i) {
                const int key = order[i];
                const double keyMin = items[key].box.lo[axis];
                int j = i - 1;
                while (j >= start && keyMin < items[order[j]].box.lo[axis]) {
                    order[j + 1] = order[j];
                    --j;
                }
                order[j + 1] = key;
            }
            const int mid = start + span / 2;
            l = BuildReference(order, start, mid, depth + 1, visits);
            r = BuildReference(order, mid, end, depth + 1, visits);
        }
        BuildNode in;
        in.box = box;
        in.left = l;
        in.right = r;
        nodes.push_back(in);
        return (int)nodes.size() - 1;
    }

    // Plain list: a right-deep chain whose boxes are all the scene box.
    int BuildList()
    {
        Box3 box;
        for (const Item& it : items) box.Grow(it.box);
        int chain = -1;
        for (int i = (int)items.size() - 1; i >= 0; --i) {
            std::vector<int> one(1, i);
            const int leaf = NewLeaf(one, box);
            nodes[leaf].box = box;
            if (chain < 0) {
                chain = leaf;
            } else {
                BuildNode in;
                in.box = box;
                in.left = leaf;
                in.right = chain;
                nodes.push_back(in);
                chain = (int)nodes.size() - 1;
            }
        }
        maxDepth = (int)items.size();
        return chain;
    }
};

inline float RoundDown(double v)
{
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -INFINITY);
    return f;
}
inline float RoundUp(double v)
{
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, INFINITY);
    return f;
}

// fp32 box, as centre and half-extent, that contains the FP64 box with a little
// slack for the fp32 slab test (which measures entry/exit as t_centre -+ |e/d|).
inline void PackBox(const Box3& b, DevNode& n)
{
    for (int a = 0; a < 3; ++a) {
        const double mag = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
        const double pad = 8e-7 * mag + 2e-7 * (b.hi[a] - b.lo[a]) + 1e-30;
        const double lo = b.lo[a] - pad, hi = b.hi[a] + pad;
        const float c = (float)(0.5 * (lo + hi));
        n.c[a] = c;
        n.e[a] = RoundUp(std::max(hi - (double)c, (double)c - lo));
    }
}

struct Packer {
    const rt_scene_desc& d;
    const rt_upload_options& opt;
    Packed out;
    std::vector<Baked> baked;         // surfaces first, then medium boundaries
    std::vector<int> bakedOfPrim;     // desc prim index -> baked index
    std::vector<uint32_t> hitOfBaked; // baked index -> RT_HIT_MAKE(type, device index), filled as the runs are emitted

    Packer(const rt_scene_desc& desc, const rt_upload_options& o) : d(desc), opt(o) {}

    // Does a texture lead to an image (directly or through checkers)?  Only then do (u,v) matter.
    bool TextureUsesUv(int tex, int depth = 0) const
    {
        if (tex < 0 || tex >= d.n_textures || depth > 16) return false;
        const rt_texture& t = d.textures[tex];
        if (t.type == RT_TEX_IMAGE) return true;
        if (t.type == RT_TEX_CHECKER) return TextureUsesUv(t.even, depth + 1) || TextureUsesUv(t.odd, depth + 1);
        return false;
    }
    // Material word of a sphere: the index, plus a UV frame for an image-textured sphere baked out of a RotateY chain.
    int32_t SphereMaterialWord(const Baked& b)
    {
        if (b.material > RT_MATERIAL_INDEX_MASK) throw std::invalid_argument("too many materials");
        const rt_material& m = d.materials[b.material];
        const bool textured = m.type != RT_MAT_METAL && m.type != RT_MAT_DIELECTRIC && TextureUsesUv(m.texture);
        if (!b.rotated || !textured) return b.material;
        if (out.uv_frames.size() >= 2047) throw std::invalid_argument("too many rotated image-textured spheres");
        DevUvFrame f;
        f.s = (float)b.yawSin;
        f.c = (float)b.yawCos;
        out.uv_frames.push_back(f);
        return b.material | (int32_t)(out.uv_frames.size() << RT_MATERIAL_INDEX_BITS);
    }

    void PushSphere(const Baked& b)
    {
        DevSphere s;
        s.cx = b.a[0];
        s.cy = b.a[1];
        s.cz = b.a[2];
        s.radius = b.radius;
        out.spheres.push_back(s);
        out.sphere_material.push_back(SphereMaterialWord(b));
    }
    void PushMoving(const Baked& b)
    {
        DevMovingSphere s;
        s.c0x = b.a[0];
        s.c0y = b.a[1];
        s.c0z = b.a[2];
        s.radius = (float)b.radius;
        s.material = SphereMaterialWord(b);
        s.dcx = b.b[0] - b.a[0];
        s.dcy = b.b[1] - b.a[1];
        s.dcz = b.b[2] - b.a[2];
        s.time0 = (float)b.time0;
        s.inv_dt = (float)(1.0 / (b.time1 - b.time0));
        out.moving.push_back(s);
    }
    void PushQuad(const Baked& b)
    {
        // Quad.h:31-36 on the baked Q, u, v
        const double* u = b.b;
        const double* v = b.c;
        const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
        const double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
        const double len = std::sqrt(nn);
        DevQuad q;
        std::memset(&q, 0, sizeof q);
        q.qx = b.a[0];
        q.qy = b.a[1];
        q.qz = b.a[2];
        q.nx = n[0] / len;
        q.ny = n[1] / len;
        q.nz = n[2] / len;
        q.D = q.nx * b.a[0] + q.ny * b.a[1] + q.nz * b.a[2];
        const double w[3] = {n[0] / nn, n[1] / nn, n[2] / nn}; // Quad.h:36
        q.ax = (float)(v[1] * w[2] - v[2] * w[1]); // v x w
        q.ay = (float)(v[2] * w[0] - v[0] * w[2]);
        q.az = (float)(v[0] * w[1] - v[1] * w[0]);
        q.bx = (float)(w[1] * u[2] - w[2] * u[1]); // w x u
        q.by = (float)(w[2] * u[0] - w[0] * u[2]);
        q.bz = (float)(w[0] * u[1] - w[1] * u[0]);
        q.material = b.material;
        out.quads.push_back(q);
    }

    // Six baked quads = a closed box?  (MakeBox, Instance.h:166-184, under any Translate / RotateY chain; any closed
    // parallelepiped of six quads qualifies.)  The faces must pair up into three pairs with parallel normals, the three
    // normals must span space, and the 24 corners must be 8 points shared by three faces each.  Fills the slab record.
    bool MakeDevBox(const int* ids, DevBox& box) const
    {
        double nrm[6][3], corners[24][3];
        double scale = 0.0;
        for (int k = 0; k < 6; ++k) {
            const Baked& q = baked[ids[k]];
            if (q.type != RT_LEAF_QUAD) return false;
            const double* u = q.b;
            const double* v = q.c;
            const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
            const double len = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            if (!(len > 0.0)) return false;
            for (int a = 0; a < 3; ++a) {
                nrm[k][a] = n[a] / len;
                corners[4 * k][a] = q.a[a];
                corners[4 * k + 1][a] = q.a[a] + u[a];
                corners[4 * k + 2][a] = q.a[a] + v[a];
                corners[4 * k + 3][a] = q.a[a] + u[a] + v[a];
                scale = std::max(scale, std::max(std::fabs(u[a]), std::fabs(v[a])));
            }
        }
        const double eps = 1e-9 * scale;
        // 8 distinct corners, each on three faces
        double pts[8][3];
        int uses[8], nPts = 0;
        for (int c = 0; c < 24; ++c) {
            int at = -1;
            for (int k = 0; k < nPts && at < 0; ++k)
                if (std::fabs(pts[k][0] - corners[c][0]) <= eps && std::fabs(pts[k][1] - corners[c][1]) <= eps &&
                    std::fabs(pts[k][2] - corners[c][2]) <= eps)
                    at = k;
            if (at < 0) {
                if (nPts == 8) return false;
                for (int a = 0; a < 3; ++a) pts[nPts][a] = corners[c][a];
                uses[nPts] = 0;
                at = nPts++;
            }
            ++uses[at];
        }
        if (nPts != 8) return false;
        for (int k = 0; k < 8; ++k)
            if (uses[k] != 3) return false;
        // three pairs of parallel faces
        int partner[6] = {-1, -1, -1, -1, -1, -1};
        int pairFirst[3], nPairs = 0;
        for (int k = 0; k < 6; ++k) {
            if (partner[k] >= 0) continue;
            for (int j = k + 1; j < 6 && partner[k] < 0; ++j) {
                if (partner[j] >= 0) continue;
                const double c = nrm[k][0] * nrm[j][0] + nrm[k][1] * nrm[j][1] + nrm[k][2] * nrm[j][2];
                if (std::fabs(c) > 1.0 - 1e-12) {
                    partner[k] = j;
                    partner[j] = k;
                }
            }
            if (partner[k] < 0 || nPairs == 3) return false;
            pairFirst[nPairs++] = k;
        }
        if (nPairs != 3) return false;
        const double* n0 = nrm[pairFirst[0]];
        const double* n1 = nrm[pairFirst[1]];
        const double* n2 = nrm[pairFirst[2]];
        const double det = n0[0] * (n1[1] * n2[2] - n1[2] * n2[1]) - n0[1] * (n1[0] * n2[2] - n1[2] * n2[0]) +
                           n0[2] * (n1[0] * n2[1] - n1[1] * n2[0]);
        if (!(std::fabs(det) > 1e-6)) return false;
        std::memset(&box, 0, sizeof box);
        for (int p = 0; p < 3; ++p) {
            const int k = pairFirst[p], j = partner[k];
            const double* n = nrm[k];
            const double dk = n[0] * baked[ids[k]].a[0] + n[1] * baked[ids[k]].a[1] + n[2] * baked[ids[k]].a[2];
            const double dj = n[0] * baked[ids[j]].a[0] + n[1] * baked[ids[j]].a[1] + n[2] * baked[ids[j]].a[2];
            if (!(std::fabs(dk - dj) > eps)) return false;
            for (int a = 0; a < 3; ++a) box.n[p][a] = n[a];
            const int loFace = dk < dj ? k : j, hiFace = dk < dj ? j : k;
            box.lo[p] = std::min(dk, dj);
            box.hi[p] = std::max(dk, dj);
            box.faces |= (uint32_t)loFace << (6 * p) | (uint32_t)hiFace << (6 * p + 3);
        }
        return true;
    }
    bool IsBoxList(int firstPrim, int count) const
    {
        if (count != 6 || (opt.flags & RT_UPLOAD_NO_BOXES)) return false;
        int ids[6];
        for (int k = 0; k < 6; ++k) ids[k] = firstPrim + k;
        DevBox scratch;
        return MakeDevBox(ids, scratch);
    }

    // Appends baked prims [ids] (one type) to the device arrays; returns a leaf ref.  RT_LEAF_BOX: ids are the quads
    // of whole boxes, six each; the quads go to the quad table (hits are reported against them), one DevBox per six.
    uint32_t EmitRun(int type, const std::vector<int>& ids)
    {
        if (type == RT_LEAF_BOX) {
            if (ids.empty() || ids.size() % 6 != 0 || ids.size() / 6 > (size_t)RT_MAX_LEAF_PRIMS)
                throw std::invalid_argument("box leaf size out of range");
            const size_t firstBox = out.boxes.size();
            for (size_t k = 0; k < ids.size(); k += 6) {
                DevBox box;
                if (!MakeDevBox(&ids[k], box)) throw std::invalid_argument("internal: not a box");
                box.first_quad = (uint32_t)out.quads.size();
                for (int f = 0; f < 6; ++f) {
                    hitOfBaked[ids[k + f]] = RT_HIT_MAKE(RT_LEAF_QUAD, out.quads.size());
                    PushQuad(baked[ids[k + f]]);
                }
                out.boxes.push_back(box);
            }
            if (out.boxes.size() > (size_t)RT_MAX_PRIMS_PER_TYPE || out.quads.size() > (size_t)RT_MAX_PRIMS_PER_TYPE)
                throw std::invalid_argument("too many primitives");
            return RT_REF_MAKE_LEAF(RT_LEAF_BOX, firstBox, ids.size() / 6);
        }
        if (ids.empty() || (int)ids.size() > RT_MAX_LEAF_PRIMS) throw std::invalid_argument("leaf size out of range");
        size_t first = 0;
        if (type == RT_LEAF_SPHERE) {
            first = out.spheres.size();
            for (int i : ids) {
                hitOfBaked[i] = RT_HIT_MAKE(RT_LEAF_SPHERE, out.spheres.size());
                PushSphere(baked[i]);
            }
        } else if (type == RT_LEAF_MOVING) {
            first = out.moving.size();
            for (int i : ids) PushMoving(baked[i]);
        } else {
            first = out.quads.size();
            for (int i : ids) {
                hitOfBaked[i] = RT_HIT_MAKE(RT_LEAF_QUAD, out.quads.size());
                PushQuad(baked[i]);
            }
        }
        if (first + ids.size() > (size_t)RT_MAX_PRIMS_PER_TYPE) throw std::invalid_argument("too many primitives");
        return RT_REF_MAKE_LEAF(type, first, ids.size());
    }

    void PackMaterials()
    {
        for (int i = 0; i < d.n_textures; ++i) {
            const rt_texture& t = d.textures[i];
            DevTexture x;
            std::memset(&x, 0, sizeof x);
            x.type = t.type;
            x.even = t.even;
            x.odd = t.odd;
            x.index = t.type == RT_TEX_IMAGE ? t.image : t.perlin;
            x.r = (float)t.color[0];
            x.g = (float)t.color[1];
            x.b = (float)t.color[2];
            x.scale = (float)t.scale;
            x.inv_scale = t.type == RT_TEX_CHECKER ? 1.0 / t.scale : 0.0;
            if (t.type == RT_TEX_CHECKER && (t.even < 0 || t.even >= d.n_textures || t.odd < 0 || t.odd >= d.n_textures))
                throw std::invalid_argument("checker texture child out of range");
            if (t.type == RT_TEX_NOISE && (t.perlin < 0 || t.perlin >= d.n_perlins))
                throw std::invalid_argument("noise texture perlin out of range");
            if (t.type == RT_TEX_IMAGE && t.image >= d.n_images) throw std::invalid_argument("image index out of range");
            out.textures.push_back(x);
        }
        for (int i = 0; i < d.n_materials; ++i) {
            const rt_material& m = d.materials[i];
            DevMaterial x;
            std::memset(&x, 0, sizeof x);
            if (m.type == RT_MAT_METAL || m.type == RT_MAT_DIELECTRIC) {
                if (m.type == RT_MAT_METAL) {
                    x.r = (float)m.albedo[0];
                    x.g = (float)m.albedo[1];
                    x.b = (float)m.albedo[2];
                }
                x.tt = RT_MAT_TT(m.type, out.mat_params.size());
                out.mat_params.push_back(m.type == RT_MAT_METAL ? m.fuzz : m.ior);
            } else {
                if (m.texture < 0 || m.texture >= d.n_textures) throw std::invalid_argument("material texture out of range");
                const rt_texture& t = d.textures[m.texture];
                x.tt = RT_MAT_TT(m.type, 0);
                if (t.type == RT_TEX_SOLID) {
                    x.r = (float)t.color[0];
                    x.g = (float)t.color[1];
                    x.b = (float)t.color[2];
                } else {
                    x.tt = RT_MAT_TT(m.type, m.texture + 1);
                    out.features |= RT_FEAT_TEXTURE;
                    for (int k = 0; k < d.n_textures; ++k) // a checker may lead to any texture of the scene
                        if (d.textures[k].type == RT_TEX_IMAGE || d.textures[k].type == RT_TEX_NOISE)
                            out.features |= RT_FEAT_TEXTURE_HEAVY;
                }
            }
            out.materials.push_back(x);
        }
        for (int i = 0; i < d.n_perlins; ++i) {
            const rt_perlin& p = d.perlins[i];
            DevPerlin x;
            std::memset(&x, 0, sizeof x);
            for (int k = 0; k < 256; ++k) {
                for (int a = 0; a < 3; ++a) x.ranvec[k][a] = (float)p.ranvec[k][a];
                x.perm_x[k] = (uint8_t)p.perm_x[k];
                x.perm_y[k] = (uint8_t)p.perm_y[k];
                x.perm_z[k] = (uint8_t)p.perm_z[k];
            }
            out.perlins.push_back(x);
        }
        for (int i = 0; i < d.n_images; ++i) {
            const rt_image& im = d.images[i];
            out.image_w.push_back(im.rgb ? im.width : 0);
            out.image_h.push_back(im.rgb ? im.height : 0);
            if (im.rgb && im.width > 0 && im.height > 0)
                out.image_bytes.emplace_back(im.rgb, im.rgb + (size_t)im.width * im.height * 3);
            else
                out.image_bytes.emplace_back();
        }
    }

    void Run()
    {
        if (d.abi_version != RT_ABI_VERSION) throw std::invalid_argument("rt_scene_desc.abi_version mismatch");
        if (d.n_objects <= 0 || d.n_prims <= 0) throw std::invalid_argument("empty scene");
        if (d.n_xforms < 0 || d.n_materials < 0 || d.n_textures < 0 || d.n_perlins < 0 || d.n_images < 0)
            throw std::invalid_argument("negative table size");
        if (!d.objects || !d.prims || (d.n_xforms > 0 && !d.xforms) || (d.n_materials > 0 && !d.materials) ||
            (d.n_textures > 0 && !d.textures) || (d.n_perlins > 0 && !d.perlins) || (d.n_images > 0 && !d.images))
            throw std::invalid_argument("NULL table with a non-zero count");
        for (int i = 0; i < d.n_objects; ++i)
            if (d.objects[i].kind < RT_OBJ_PRIM || d.objects[i].kind > RT_OBJ_MEDIUM)
                throw std::invalid_argument("unknown object kind");
        for (int i = 0; i < d.n_xforms; ++i)
            if (d.xforms[i].type != RT_XFORM_TRANSLATE && d.xforms[i].type != RT_XFORM_ROTATE_Y)
                throw std::invalid_argument("unknown instance transform type");
        for (int i = 0; i < d.n_materials; ++i)
            if (d.materials[i].type < RT_MAT_LAMBERTIAN || d.materials[i].type > RT_MAT_ISOTROPIC)
                throw std::invalid_argument("unknown material type");
        for (int i = 0; i < d.n_textures; ++i)
            if (d.textures[i].type < RT_TEX_SOLID || d.textures[i].type > RT_TEX_NOISE)
                throw std::invalid_argument("unknown texture type");
        for (int i = 0; i < d.n_prims; ++i) {
            const rt_prim& p = d.prims[i];
            if (p.type < RT_PRIM_SPHERE || p.type > RT_PRIM_QUAD) throw std::invalid_argument("unknown primitive type");
            if (p.material < 0 || p.material >= d.n_materials) throw std::invalid_argument("prim material out of range");
            if (p.xform_count < 0 || p.first_xform < 0 || p.first_xform + p.xform_count > d.n_xforms)
                throw std::invalid_argument("prim xform chain out of range");
        }
        PackMaterials();

        // bake everything
        baked.reserve(d.n_prims);
        for (int i = 0; i < d.n_prims; ++i) {
            baked.push_back(Bake(d, d.prims[i]));
            if (baked.back().type == RT_LEAF_MOVING) out.features |= RT_FEAT_MOVING;
            if (baked.back().type == RT_LEAF_QUAD) out.features |= RT_FEAT_QUAD;
        }

        hitOfBaked.assign(baked.size(), RT_HIT_NONE);

        // media: boundary prims go straight to the device arrays (not in the BVH)
        struct Med {
            int object;
            Box3 box;
        };
        std::vector<Med> meds;
        for (int o = 0; o < d.n_objects; ++o) {
            const rt_object& ob = d.objects[o];
            if (ob.first_prim < 0 || ob.prim_count <= 0 || ob.first_prim + ob.prim_count > d.n_prims)
                throw std::invalid_argument("object prim range out of range");
            if (ob.kind != RT_OBJ_MEDIUM) continue;
            if (ob.medium_id < 0 || ob.medium_id >= 8) throw std::invalid_argument("at most 8 media are supported");
            if (ob.phase_material < 0 || ob.phase_material >= d.n_materials)
                throw std::invalid_argument("medium phase material out of range");
            out.features |= RT_FEAT_MEDIUM;
            const int type = baked[ob.first_prim].type;
            std::vector<int> ids;
            Med m;
            m.object = o;
            for (int k = 0; k < ob.prim_count; ++k) {
                if (baked[ob.first_prim + k].type != type)
                    throw std::invalid_argument("medium boundary must be primitives of one type");
                ids.push_back(ob.first_prim + k);
                m.box.Grow(baked[ob.first_prim + k].box);
            }
            DevMedium dm;
            std::memset(&dm, 0, sizeof dm);
            // a MakeBox boundary (Cornell smoke, kernel.cu:424-431): its two boundary queries are the entry and the exit
            // of one slab test instead of two passes over six quads
            const bool boxBoundary = type == RT_LEAF_QUAD && IsBoxList(ob.first_prim, ob.prim_count);
            dm.boundary_ref = EmitRun(boxBoundary ? (int)RT_LEAF_BOX : type, ids);
            dm.phase_material = ob.phase_material;
            dm.neg_inv_density = -1.0 / ob.density;
            dm.medium_id = ob.medium_id;
            dm.visits = 1;
            out.media.push_back(dm);
            meds.push_back(m);
        }
        out.n_media = (int)meds.size();

        // T2: visit multiplicity per medium from the reference topology over objects
        {
            Builder rb;
            for (int o = 0; o < d.n_objects; ++o) {
                Item it;
                it.type = d.objects[o].kind == RT_OBJ_MEDIUM ? RT_LEAF_MEDIUM : RT_LEAF_SPHERE;
                it.index = o;
                it.first = it.count = 0;
                for (int a = 0; a < 3; ++a) {
                    it.box.lo[a] = d.objects[o].bbox[2 * a];
                    it.box.hi[a] = d.objects[o].bbox[2 * a + 1];
                }
                rb.items.push_back(it);
            }
            std::vector<int> order(d.n_objects), visits(d.n_objects, 0);
            for (int i = 0; i < d.n_objects; ++i) order[i] = i;
            rb.BuildReference(order, 0, d.n_objects, 0, visits);
            for (size_t k = 0; k < meds.size(); ++k) {
                out.media[k].visits = visits[meds[k].object];
                out.medium_visits[d.objects[meds[k].object].medium_id] = visits[meds[k].object];
            }
        }

        // BVH items
        const int opt_small_list = (opt.flags & RT_UPLOAD_WHOLE_LISTS) ? 8 : 0;
        Builder b;
        b.maxLeaf = opt.max_leaf_prims > 0 ? std::min(opt.max_leaf_prims, 8) : 2;
        const bool perObject = opt.bvh == RT_BVH_REFERENCE || opt.bvh == RT_BVH_NONE;
        std::vector<std::vector<int>> runOfItem; // perObject: baked ids behind each item
        int medIdx = 0;
        for (int o = 0; o < d.n_objects; ++o) {
            const rt_object& ob = d.objects[o];
            if (ob.kind == RT_OBJ_MEDIUM) {
                Item it;
                it.type = RT_LEAF_MEDIUM;
                it.index = medIdx;
                it.first = it.count = 0;
                it.box = meds[medIdx].box;
                if (perObject)
                    for (int a = 0; a < 3; ++a) {
                        it.box.lo[a] = ob.bbox[2 * a];
                        it.box.hi[a] = ob.bbox[2 * a + 1];
                    }
                b.items.push_back(it);
                runOfItem.emplace_back();
                ++medIdx;
                continue;
            }
            // RT_UPLOAD_WHOLE_LISTS (A/B): a small owning list of one primitive type (MakeBox: six quads,
            // Instance.h:166-184) stays ONE item -- half the nodes (scene 9: 5 062 -> 2 454, the node table then fits
            // in shared memory), but every visit tests all six quads one after the other.  Measured slower than
            // one item per quad: scene 9 4.00 vs 4.19 Grays/s, scene 7 11.6 vs 13.1 (profiles/r2_ab_i.jsonl).
            // A MakeBox list (six quads that close a box) is ONE item, tested by one slab test over its three pairs of
            // faces (DevBox) instead of six plane + interior tests spread over several leaves.
            if (!perObject && ob.kind == RT_OBJ_LIST && IsBoxList(ob.first_prim, ob.prim_count)) {
                Item it;
                it.type = RT_LEAF_BOX;
                it.index = ob.first_prim;
                it.first = it.count = 0;
                it.box = Box3();
                std::vector<int> ids;
                for (int k = 0; k < ob.prim_count; ++k) {
                    it.box.Grow(baked[ob.first_prim + k].box);
                    ids.push_back(ob.first_prim + k);
                }
                b.items.push_back(it);
                runOfItem.push_back(ids);
                continue;
            }
            bool wholeList = !perObject && ob.kind == RT_OBJ_LIST && ob.prim_count >= 2 && ob.prim_count <= opt_small_list;
            for (int k = 1; wholeList && k < ob.prim_count; ++k)
                wholeList = baked[ob.first_prim + k].type == baked[ob.first_prim].type;
            if (wholeList) {
                Item it;
                it.type = baked[ob.first_prim].type;
                it.index = ob.first_prim;
                it.first = 0;
                it.count = ob.prim_count; // weighs the SAH leaf cost
                it.box = Box3();
                std::vector<int> ids;
                for (int k = 0; k < ob.prim_count; ++k) {
                    it.box.Grow(baked[ob.first_prim + k].box);
                    ids.push_back(ob.first_prim + k);
                }
                b.items.push_back(it);
                runOfItem.push_back(ids);
            } else if (!perObject) {
                for (int k = 0; k < ob.prim_count; ++k) {
                    Item it;
                    it.type = baked[ob.first_prim + k].type;
                    it.index = ob.first_prim + k;
                    it.first = it.count = 0;
                    it.box = baked[ob.first_prim + k].box;
                    b.items.push_back(it);
                    runOfItem.emplace_back(1, ob.first_prim + k);
                }
            } else {
                // one item per object; split only where the type changes or the run exceeds a leaf
                int k = 0;
                bool firstPiece = true;
                while (k < ob.prim_count) {
                    const int type = baked[ob.first_prim + k].type;
                    std::vector<int> ids;
                    while (k < ob.prim_count && baked[ob.first_prim + k].type == type && (int)ids.size() < RT_MAX_LEAF_PRIMS)
                        ids.push_back(ob.first_prim + k++);
                    Item it;
                    it.type = type;
                    it.index = o;
                    it.first = it.count = 0;
                    for (int a = 0; a < 3; ++a) {
                        it.box.lo[a] = ob.bbox[2 * a];
                        it.box.hi[a] = ob.bbox[2 * a + 1];
                    }
                    if (!firstPiece && opt.bvh == RT_BVH_REFERENCE)
                        throw std::invalid_argument("reference BVH mode needs single-type objects of <= 1024 primitives");
                    firstPiece = false;
                    b.items.push_back(it);
                    runOfItem.push_back(ids);
                }
            }
        }
        if (b.items.empty()) throw std::invalid_argument("scene has no primitives");

        // Hoisting (SAH mode).  Three kinds of item gain nothing from the hierarchy and lose a lot inside it, because a
        // leaf is visited by the few lanes of a warp that happen to reach it in the same step, while an item tested
        // BEFORE the tree is entered is tested by every lane together, as straight-line code:
        //   * scene-sized items (Book 1's ground sphere, scene 9's r = 5000 mist): nearly every ray enters their box;
        //   * every ConstantMedium: its test is two boundary queries plus a logarithm -- by far the longest leaf --
        //     and ncu put the leaf path of the Cornell-smoke scene on 5.9 of 32 lanes (67 % of its instructions);
        //   * all surfaces of a tiny scene (<= kFlatMax primitives: the Cornell box has 6 walls): a linear pass over a
        //     dozen primitives on a full warp costs less than walking a tree on a third of one.
        // Hoisted items are tested up front and the distance they return then culls the walk through the rest.  The
        // closest hit is a minimum over all primitives (and medium draws are keyed), so the image does not change.
        // Slots (RT_MAX_HOISTED): one per medium, one per primitive type of a flattened scene, one per scene-sized item.
        if (opt.bvh == RT_BVH_SAH && !(opt.flags & RT_UPLOAD_NO_HOIST)) {
            const int kFlatMax = 16;
            std::vector<char> take(b.items.size(), 0);
            int slots = RT_MAX_HOISTED;
            std::vector<uint32_t> refs;
            size_t nSurface = 0, nSurfaceTests = 0;
            bool typePresent[5] = {false, false, false, false, false};
            for (size_t k = 0; k < b.items.size(); ++k)
                if (b.items[k].type != RT_LEAF_MEDIUM) {
                    ++nSurface;
                    nSurfaceTests += b.items[k].type == RT_LEAF_BOX ? 2 : runOfItem[k].size(); // a box: one slab test
                    typePresent[b.items[k].type] = true;
                }
            int nTypes = 0;
            for (int t : Builder::kSurfaceTypes) nTypes += (int)typePresent[t];
            const bool flat = nSurface > 0 && nSurfaceTests <= (size_t)kFlatMax && nTypes <= slots;
            if (flat) {
                for (int t : Builder::kSurfaceTypes) {
                    if (!typePresent[t]) continue;
                    std::vector<int> ids;
                    for (size_t k = 0; k < b.items.size(); ++k)
                        if (b.items[k].type == t) {
                            take[k] = 1;
                            ids.insert(ids.end(), runOfItem[k].begin(), runOfItem[k].end());
                        }
                    refs.push_back(EmitRun(t, ids));
                    --slots;
                }
            }
            for (size_t k = 0; k < b.items.size() && slots > 0; ++k)
                if (b.items[k].type == RT_LEAF_MEDIUM) {
                    take[k] = 1;
                    refs.push_back(RT_REF_MAKE_LEAF(RT_LEAF_MEDIUM, b.items[k].index, 1));
                    --slots;
                }
            if (!flat) {
                Box3 sceneBox;
                for (const Item& it : b.items) sceneBox.Grow(it.box);
                const double sceneArea = sceneBox.Area();
                size_t left = 0;
                for (size_t k = 0; k < b.items.size(); ++k) left += take[k] ? 0 : 1;
                for (size_t k = 0; k < b.items.size() && slots > 0 && left > 2; ++k) {
                    if (take[k] || !(sceneArea > 0.0 && b.items[k].box.Area() >= 0.5 * sceneArea)) continue;
                    take[k] = 1;
                    refs.push_back(EmitRun(b.items[k].type, runOfItem[k]));
                    --slots;
                    --left;
                }
            }
            std::vector<Item> kept;
            std::vector<std::vector<int>> keptRuns;
            for (size_t k = 0; k < b.items.size(); ++k)
                if (!take[k]) {
                    kept.push_back(b.items[k]);
                    keptRuns.push_back(runOfItem[k]);
                }
            for (uint32_t r : refs) out.hoisted[out.n_hoisted++] = r;
            b.items.swap(kept);
            runOfItem.swap(keptRuns);
        }

        if (b.items.empty()) { // everything is tested up front: an empty tree (two zeroed records keep the table non-empty)
            out.nodes.assign(2, DevNode{});
            out.root_ref = 0xffffffffu; // RT_TRAV_DONE: the walk loop does not run
            out.max_depth = 0;
            PackLights();
            return;
        }
        int root;
        if (opt.bvh == RT_BVH_NONE) {
            root = b.BuildList();
        } else if (opt.bvh == RT_BVH_REFERENCE) {
            std::vector<int> order(b.items.size()), visits(b.items.size(), 0);
            for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
            root = b.BuildReference(order, 0, (int)order.size(), 0, visits);
        } else {
            std::vector<int> ids(b.items.size());
            for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int)i;
            root = b.BuildSah(ids, 0);
        }
        out.max_depth = b.maxDepth;
        if (opt.bvh != RT_BVH_NONE && b.maxDepth > 29) throw std::invalid_argument("BVH deeper than the traversal stack");

        // flatten: node 0 = root record, node 1 = pad, children pairs at even indices
        auto leafRef = [&](const BuildNode& n) -> uint32_t {
            const int type = b.items[n.items[0]].type;
            if (type == RT_LEAF_MEDIUM) {
                if (n.items.size() != 1) {
                    // several media in one leaf cannot be expressed as one run unless adjacent; keep it simple
                    throw std::invalid_argument("internal: multi-medium leaf");
                }
                return RT_REF_MAKE_LEAF(RT_LEAF_MEDIUM, b.items[n.items[0]].index, 1);
            }
            std::vector<int> ids;
            for (int it : n.items)
                for (int id : runOfItem[it]) ids.push_back(id);
            return EmitRun(type, ids);
        };
        out.nodes.clear();
        out.nodes.resize(2);
        std::memset(out.nodes.data(), 0, 2 * sizeof(DevNode));
        // iterative DFS; each entry = (build node, slot it occupies)
        struct Todo {
            int node, slot;
        };
        std::vector<Todo> stack;
        stack.push_back(Todo{root, 0});
#if RT_BVH4
        // A/B build: every internal node adopts its grandchildren (largest box first) until it has four children.
        // Stack levels: a visit leaves up to three siblings parked while the walk descends into the fourth.
        while (!stack.empty()) {
            const Todo t = stack.back();
            stack.pop_back();
            const BuildNode& bn = b.nodes[t.node];
            PackBox(bn.box, out.nodes[t.slot]);
            out.nodes[t.slot].aux = 0;
            if (bn.left < 0) {
                out.nodes[t.slot].ref = leafRef(bn);
                continue;
            }
            std::vector<int> kids = {bn.left, bn.right};
            while (kids.size() < 4) {
                int best = -1;
                for (size_t k = 0; k < kids.size(); ++k)
                    if (b.nodes[kids[k]].left >= 0 && (best < 0 || b.nodes[kids[k]].box.Area() > b.nodes[kids[best]].box.Area()))
                        best = (int)k;
                if (best < 0) break;
                const int gone = kids[best];
                kids[best] = b.nodes[gone].left;
                kids.insert(kids.begin() + best + 1, b.nodes[gone].right);
            }
            const int quad = (int)out.nodes.size();
            out.nodes.resize(out.nodes.size() + 4);
            out.nodes[t.slot].ref = (uint32_t)quad;
            for (int k = 0; k < 4; ++k) {
                DevNode& e = out.nodes[quad + k];
                std::memset(&e, 0, sizeof e);
                e.e[0] = e.e[1] = e.e[2] = -1.0f; // empty slot: exit before entry on every axis
                e.ref = RT_REF_MAKE_LEAF(RT_LEAF_SPHERE, 0, 1);
            }
            for (int k = (int)kids.size() - 1; k >= 0; --k) stack.push_back(Todo{kids[k], quad + k});
        }
        {
            // worst-case stack use: depth in collapsed nodes, three parked siblings per level
            struct Walk {
                static int Depth4(const std::vector<DevNode>& n, uint32_t ref)
                {
                    if (ref & RT_REF_LEAF) return 0;
                    int d = 0;
                    for (int k = 0; k < 4; ++k)
                        if (n[ref + k].e[0] >= 0.0f) d = std::max(d, Depth4(n, n[ref + k].ref));
                    return d + 1;
                }
            };
            out.max_depth = 3 * Walk::Depth4(out.nodes, out.nodes[0].ref) + 1;
            if (out.max_depth > 29) throw std::invalid_argument("BVH4 deeper than the traversal stack");
        }
#else
        while (!stack.empty()) {
            const Todo t = stack.back();
            stack.pop_back();
            const BuildNode& bn = b.nodes[t.node];
            PackBox(bn.box, out.nodes[t.slot]);
            out.nodes[t.slot].aux = 0;
            if (bn.left < 0) {
                out.nodes[t.slot].ref = leafRef(bn);
            } else {
                const int pair = (int)out.nodes.size();
                out.nodes.resize(out.nodes.size() + 2);
                out.nodes[t.slot].ref = (uint32_t)pair;
                stack.push_back(Todo{bn.right, pair + 1});
                stack.push_back(Todo{bn.left, pair}); // left is laid out (and its leaves emitted) first
            }
        }
#endif
        out.root_ref = out.nodes[0].ref;
        PackLights();
    }

    // Sampling targets of RT_FLAG_IMPORTANCE: every quad / sphere of a surface object whose material is a DiffuseLight,
    // in primitive order (the order the oracle lists them in).  Called when every run has been emitted.
    void PackLights()
    {
        std::vector<std::pair<int, DevLight>> found;
        for (int o = 0; o < d.n_objects; ++o) {
            const rt_object& ob = d.objects[o];
            if (ob.kind == RT_OBJ_MEDIUM) continue;
            for (int k = 0; k < ob.prim_count; ++k) {
                const int i = ob.first_prim + k;
                const Baked& bk = baked[i];
                if ((bk.type != RT_LEAF_QUAD && bk.type != RT_LEAF_SPHERE) || d.materials[bk.material].type != RT_MAT_DIFFUSE_LIGHT)
                    continue;
                if (hitOfBaked[i] == RT_HIT_NONE) throw std::invalid_argument("internal: light primitive was not emitted");
                DevLight l;
                std::memset(&l, 0, sizeof l);
                for (int a = 0; a < 3; ++a) {
                    l.a[a] = bk.a[a];
                    l.b[a] = bk.type == RT_LEAF_QUAD ? bk.b[a] : (a == 0 ? bk.radius : 0.0);
                    l.c[a] = bk.type == RT_LEAF_QUAD ? bk.c[a] : 0.0;
                }
                l.hit = hitOfBaked[i];
                if (bk.type == RT_LEAF_QUAD) {
                    const double* u = bk.b;
                    const double* v = bk.c;
                    const double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
                    l.area = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                }
                found.emplace_back(i, l);
            }
        }
        std::sort(found.begin(), found.end(), [](const std::pair<int, DevLight>& x, const std::pair<int, DevLight>& y) {
            return x.first < y.first;
        });
        for (const std::pair<int, DevLight>& f : found) out.lights.push_back(f.second);
    }
};

// Camera.h:36-71 in FP64.
inline DevCamera MakeCamera(const rt_camera& c)
{
    DevCamera k;
    std::memset(&k, 0, sizeof k);
    const double aspect = double(c.image_width) / double(c.image_height);
    double aperture = c.aperture;
    if (aperture < 0.0) aperture = 2.0 * c.focus_dist * std::tan(c.defocus_angle * 3.14159265358979323846 / 360.0);
    const double theta = c.vfov * 3.14159265358979323846 / 180.0;
    const double halfHeight = std::tan(theta / 2.0);
    const double halfWidth = aspect * halfHeight;
    double w[3], u[3], v[3];
    double len = 0;
    for (int a = 0; a < 3; ++a) {
        w[a] = c.lookfrom[a] - c.lookat[a];
        len += w[a] * w[a];
    }
    len = std::sqrt(len);
    for (int a = 0; a < 3; ++a) w[a] = (1 / len) * w[a];
    u[0] = c.vup[1] * w[2] - c.vup[2] * w[1];
    u[1] = c.vup[2] * w[0] - c.vup[0] * w[2];
    u[2] = c.vup[0] * w[1] - c.vup[1] * w[0];
    len = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int a = 0; a < 3; ++a) u[a] = (1 / len) * u[a];
    v[0] = w[1] * u[2] - w[2] * u[1];
    v[1] = w[2] * u[0] - w[0] * u[2];
    v[2] = w[0] * u[1] - w[1] * u[0];
    const double fd = c.focus_dist;
    for (int a = 0; a < 3; ++a) {
        k.origin[a] = c.lookfrom[a];
        k.llc[a] = c.lookfrom[a] - halfWidth * fd * u[a] - halfHeight * fd * v[a] - fd * w[a];
        k.horiz[a] = 2.0 * halfWidth * fd * u[a];
        k.vert[a] = 2.0 * halfHeight * fd * v[a];
        k.u[a] = u[a];
        k.v[a] = v[a];
        k.background[a] = (float)c.background[a];
    }
    k.lens_radius = aperture / 2.0;
    k.inv_width = 1.0 / (double)c.image_width;
    k.inv_height = 1.0 / (double)c.image_height;
    k.time0 = (float)c.time0;
    k.time1 = (float)c.time1;
    k.width = c.image_width;
    k.height = c.image_height;
    k.max_depth = c.max_depth;
    return k;
}

} // namespace rtpack
