// rt/scene.hpp -- the reference's scene and camera surface, host side.
//
// Same class names and constructor argument lists as the reference's
// device-side classes, so a scene written for the reference (kernel.cu:199-517)
// reads the same here:
//
//   Hittable / HitRecord-free interface     reference Hittable.h:35-64
//   Sphere, MovingSphere, Quad              Sphere.h, MovingSphere.h, Quad.h
//   Translate, RotateY, MakeBox             Instance.h:28-184
//   HittableList, BvhNode                   HittableList.h, BvhNode.h
//   ConstantMedium                          ConstantMedium.h
//   Lambertian, Metal, Dielectric,
//   DiffuseLight, Isotropic                 Material.h, Metal.h, Dielectric.h
//   SolidColor, CheckerTexture,
//   ImageTexture, NoiseTexture, Perlin      Texture.h, Perlin.h
//   Camera                                  Camera.h:21-102
//
// What differs, on purpose:
//   * These objects live on the HOST.  They hold parameters and compute the
//     construction-time quantities the reference computes (bounding boxes,
//     sin/cos of RotateY, Perlin tables, the BVH topology), all in FP64.  They
//     do not intersect rays -- Flatten() writes the graph into the flat
//     rt_scene_desc of include/rt_abi.h and the CUDA library renders it.
//   * Ownership: the reference leaks/deletes device objects by hand
//     (Instance.h:39,114; Material.h:55-56).  Here every node created with
//     `new` while a SceneScope is alive belongs to that scope and is freed
//     with it, so reference-style `new Sphere(..., new Lambertian(...))` code
//     needs no deletes.
//   * Scene RNG: rt::Xorwow (cuRAND XORWOW restated) replaces curandState*.
#pragma once

#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../rt_abi.h"
#include "math.hpp"
#include "xorwow.hpp"

namespace rt {

// ---------------------------------------------------------------- ownership
class Node;
class SceneScope {
public:
    SceneScope() : prev_(Current()) { Current() = this; }
    ~SceneScope();
    SceneScope(const SceneScope&) = delete;
    SceneScope& operator=(const SceneScope&) = delete;
    static SceneScope*& Current()
    {
        static thread_local SceneScope* cur = nullptr;
        return cur;
    }
    void Adopt(Node* n) { nodes_.push_back(n); }

private:
    SceneScope* prev_;
    std::vector<Node*> nodes_;
};

class Node {
public:
    Node()
    {
        if (SceneScope::Current()) SceneScope::Current()->Adopt(this);
    }
    virtual ~Node() {}
};

inline SceneScope::~SceneScope()
{
    Current() = prev_;
    for (size_t k = nodes_.size(); k-- > 0;) delete nodes_[k];
}

// ------------------------------------------------------------ flat scene out
// Owning storage behind an rt_scene_desc.
struct SceneDesc {
    std::vector<rt_object> objects;
    std::vector<rt_prim> prims;
    std::vector<rt_xform> xforms;
    std::vector<rt_material> materials;
    std::vector<rt_texture> textures;
    std::vector<rt_perlin> perlins;
    std::vector<rt_image> images;
    std::vector<std::vector<uint8_t>> image_bytes;

    rt_scene_desc View() const
    {
        rt_scene_desc d;
        std::memset(&d, 0, sizeof d);
        d.abi_version = RT_ABI_VERSION;
        d.n_objects = (int32_t)objects.size();
        d.n_prims = (int32_t)prims.size();
        d.n_xforms = (int32_t)xforms.size();
        d.n_materials = (int32_t)materials.size();
        d.n_textures = (int32_t)textures.size();
        d.n_perlins = (int32_t)perlins.size();
        d.n_images = (int32_t)images.size();
        d.objects = objects.data();
        d.prims = prims.data();
        d.xforms = xforms.data();
        d.materials = materials.data();
        d.textures = textures.data();
        d.perlins = perlins.data();
        d.images = images.data();
        return d;
    }
};

using XformChain = std::vector<rt_xform>; // outermost first

class Texture;
class Material;
class Perlin;

// Collects the graph into a SceneDesc, de-duplicating shared textures,
// materials and instance chains by identity/content.
class Emitter {
public:
    explicit Emitter(SceneDesc& out) : out_(out) {}
    int TextureIndex(const Texture* t);
    int MaterialIndex(const Material* m);
    int ChainIndex(const XformChain& chain)
    {
        if (chain.empty()) return 0;
        for (const auto& kv : chains_) {
            if (kv.second.size() == chain.size() &&
                std::memcmp(kv.second.data(), chain.data(), chain.size() * sizeof(rt_xform)) == 0)
                return kv.first;
        }
        const int first = (int)out_.xforms.size();
        out_.xforms.insert(out_.xforms.end(), chain.begin(), chain.end());
        chains_[first] = chain;
        return first;
    }
    SceneDesc& Out() { return out_; }

private:
    SceneDesc& out_;
    std::map<const Texture*, int> tex_;
    std::map<const Material*, int> mat_;
    std::map<int, XformChain> chains_;
};

// ----------------------------------------------------------------- textures
class Texture : public Node {
public:
    virtual int Emit(Emitter& em) const = 0;
};

// Texture.h:35-56
class SolidColor : public Texture {
public:
    explicit SolidColor(const Color& albedo) : mAlbedo(albedo) {}
    SolidColor(double r, double g, double b) : mAlbedo(r, g, b) {}
    int Emit(Emitter& em) const override
    {
        rt_texture t;
        std::memset(&t, 0, sizeof t);
        t.type = RT_TEX_SOLID;
        t.even = t.odd = t.image = t.perlin = -1;
        for (int k = 0; k < 3; ++k) t.color[k] = mAlbedo[k];
        em.Out().textures.push_back(t);
        return (int)em.Out().textures.size() - 1;
    }

private:
    Color mAlbedo;
};

// Texture.h:58-88.  `scale` is stored; consumers apply 1.0/scale (Texture.h:65).
class CheckerTexture : public Texture {
public:
    CheckerTexture(double scale, Texture* even, Texture* odd) : mScale(scale), mEven(even), mOdd(odd) {}
    int Emit(Emitter& em) const override
    {
        rt_texture t;
        std::memset(&t, 0, sizeof t);
        t.type = RT_TEX_CHECKER;
        t.image = t.perlin = -1;
        t.scale = mScale;
        t.even = em.TextureIndex(mEven);
        t.odd = em.TextureIndex(mOdd);
        em.Out().textures.push_back(t);
        return (int)em.Out().textures.size() - 1;
    }

private:
    double mScale;
    Texture* mEven;
    Texture* mOdd;
};

// Texture.h:98-139.  `data` = RGB8 texels as RtwImage produces them
// (RtwImage.h:51-105), NULL or height<=0 renders cyan (Texture.h:113-114).
class ImageTexture : public Texture {
public:
    ImageTexture(const unsigned char* data, int width, int height) : mData(data), mWidth(width), mHeight(height) {}
    int Emit(Emitter& em) const override
    {
        rt_texture t;
        std::memset(&t, 0, sizeof t);
        t.type = RT_TEX_IMAGE;
        t.even = t.odd = t.perlin = -1;
        t.image = -1;
        if (mData != nullptr && mHeight > 0 && mWidth > 0) {
            SceneDesc& o = em.Out();
            o.image_bytes.emplace_back(mData, mData + (size_t)mWidth * mHeight * 3);
            rt_image im;
            im.width = mWidth;
            im.height = mHeight;
            im.rgb = nullptr; // patched in Flatten() once the vectors stop moving
            o.images.push_back(im);
            t.image = (int)o.images.size() - 1;
        }
        em.Out().textures.push_back(t);
        return (int)em.Out().textures.size() - 1;
    }

private:
    const unsigned char* mData;
    int mWidth, mHeight;
};

// Perlin.h:22-34,84-117.  Tables are drawn from the *scene* stream at the
// point of construction (SURVEY.md trap T8): 256 x (x,y,z) in that order,
// then the X, Y, Z permutations.
class Perlin {
public:
    explicit Perlin(Xorwow* rng)
    {
        for (int i = 0; i < 256; ++i) {
            // Perlin.h:88-94: min + range*U with U a float widened to double;
            // the device evaluates the three draws left to right (trap T1).
            const double x = -1.0 + 2.0 * (double)rng->Uniform();
            const double y = -1.0 + 2.0 * (double)rng->Uniform();
            const double z = -1.0 + 2.0 * (double)rng->Uniform();
            const Vector3 u = UnitVector(Vector3(x, y, z));
            for (int k = 0; k < 3; ++k) mTables.ranvec[i][k] = u[k];
        }
        GeneratePerm(mTables.perm_x, rng);
        GeneratePerm(mTables.perm_y, rng);
        GeneratePerm(mTables.perm_z, rng);
    }
    const rt_perlin& Tables() const { return mTables; }

private:
    rt_perlin mTables;
    // Perlin.h:96-117.  `curand_uniform(s) * (i + 1)` is a float*int product,
    // i.e. evaluated in fp32 before the int() truncation.
    static void GeneratePerm(int32_t* p, Xorwow* rng)
    {
        for (int i = 0; i < 256; ++i) p[i] = i;
        for (int i = 255; i > 0; --i) {
            int target = int(rng->Uniform() * (float)(i + 1));
            if (target > i) target = i;
            const int32_t tmp = p[i];
            p[i] = p[target];
            p[target] = tmp;
        }
    }
};

// Texture.h:149-176
class NoiseTexture : public Texture {
public:
    NoiseTexture(double scale, Xorwow* rng) : mNoise(rng), mScale(scale) {}
    int Emit(Emitter& em) const override
    {
        rt_texture t;
        std::memset(&t, 0, sizeof t);
        t.type = RT_TEX_NOISE;
        t.even = t.odd = t.image = -1;
        t.scale = mScale;
        em.Out().perlins.push_back(mNoise.Tables());
        t.perlin = (int)em.Out().perlins.size() - 1;
        em.Out().textures.push_back(t);
        return (int)em.Out().textures.size() - 1;
    }

private:
    Perlin mNoise;
    double mScale;
};

// ---------------------------------------------------------------- materials
class Material : public Node {
public:
    virtual rt_material Describe(Emitter& em) const = 0;
};

inline rt_material BlankMaterial(int type)
{
    rt_material m;
    std::memset(&m, 0, sizeof m);
    m.type = type;
    m.texture = -1;
    return m;
}

// Material.h:45-86
class Lambertian : public Material {
public:
    explicit Lambertian(const Color& albedo) : mTexture(new SolidColor(albedo)) {}
    explicit Lambertian(Texture* texture) : mTexture(texture) {}
    rt_material Describe(Emitter& em) const override
    {
        rt_material m = BlankMaterial(RT_MAT_LAMBERTIAN);
        m.texture = em.TextureIndex(mTexture);
        return m;
    }

private:
    Texture* mTexture;
};

// Metal.h:9-35: fuzz is clamped to 1 at construction.
class Metal : public Material {
public:
    Metal(const Color& albedo, double fuzz) : mAlbedo(albedo), mFuzz(fuzz < 1.0 ? fuzz : 1.0) {}
    rt_material Describe(Emitter&) const override
    {
        rt_material m = BlankMaterial(RT_MAT_METAL);
        for (int k = 0; k < 3; ++k) m.albedo[k] = mAlbedo[k];
        m.fuzz = mFuzz;
        return m;
    }

private:
    Color mAlbedo;
    double mFuzz;
};

// Dielectric.h:10-69
class Dielectric : public Material {
public:
    explicit Dielectric(double refractionIndex) : mRefractionIndex(refractionIndex) {}
    rt_material Describe(Emitter&) const override
    {
        rt_material m = BlankMaterial(RT_MAT_DIELECTRIC);
        m.ior = mRefractionIndex;
        return m;
    }

private:
    double mRefractionIndex;
};

// Material.h:100-131
class DiffuseLight : public Material {
public:
    explicit DiffuseLight(Texture* texture) : mTexture(texture) {}
    explicit DiffuseLight(const Color& emit) : mTexture(new SolidColor(emit)) {}
    rt_material Describe(Emitter& em) const override
    {
        rt_material m = BlankMaterial(RT_MAT_DIFFUSE_LIGHT);
        m.texture = em.TextureIndex(mTexture);
        return m;
    }

private:
    Texture* mTexture;
};

// Material.h:139-167
class Isotropic : public Material {
public:
    explicit Isotropic(const Color& albedo) : mTexture(new SolidColor(albedo)) {}
    explicit Isotropic(Texture* texture) : mTexture(texture) {}
    rt_material Describe(Emitter& em) const override
    {
        rt_material m = BlankMaterial(RT_MAT_ISOTROPIC);
        m.texture = em.TextureIndex(mTexture);
        return m;
    }

private:
    Texture* mTexture;
};

inline int Emitter::TextureIndex(const Texture* t)
{
    if (t == nullptr) throw std::invalid_argument("rt: null Texture");
    auto it = tex_.find(t);
    if (it != tex_.end()) return it->second;
    const int idx = t->Emit(*this);
    tex_[t] = idx;
    return idx;
}

inline int Emitter::MaterialIndex(const Material* m)
{
    if (m == nullptr) throw std::invalid_argument("rt: null Material");
    auto it = mat_.find(m);
    if (it != mat_.end()) return it->second;
    const rt_material d = m->Describe(*this);
    out_.materials.push_back(d);
    const int idx = (int)out_.materials.size() - 1;
    mat_[m] = idx;
    return idx;
}

// ----------------------------------------------------------------- geometry
class Hittable : public Node {
public:
    virtual Aabb BoundingBox() const = 0;
    virtual bool IsBvhNode() const { return false; }
    // Appends this object's primitives, each carrying `chain` (the instance
    // wrappers between it and the top-level list), to the emitter.
    virtual void CollectPrims(Emitter& em, const XformChain& chain) const = 0;
    // Top-level role of this object in list[] (RT_OBJ_*).
    virtual int ObjectKind() const { return RT_OBJ_PRIM; }
    virtual const class ConstantMedium* AsMedium() const { return nullptr; }
};

inline rt_prim BlankPrim(int type, int material, Emitter& em, const XformChain& chain)
{
    rt_prim p;
    std::memset(&p, 0, sizeof p);
    p.type = type;
    p.material = material;
    p.first_xform = em.ChainIndex(chain);
    p.xform_count = (int32_t)chain.size();
    return p;
}

// Sphere.h:8-21
class Sphere : public Hittable {
public:
    Sphere(const Point3& center, double radius, Material* material)
        : mCenter(center), mRadius(radius), mMaterial(material)
    {
        const Vector3 rvec(radius, radius, radius);
        mBBox = Aabb(center - rvec, center + rvec);
    }
    Aabb BoundingBox() const override { return mBBox; }
    void CollectPrims(Emitter& em, const XformChain& chain) const override
    {
        rt_prim p = BlankPrim(RT_PRIM_SPHERE, em.MaterialIndex(mMaterial), em, chain);
        for (int k = 0; k < 3; ++k) p.a[k] = mCenter[k];
        p.radius = mRadius;
        em.Out().prims.push_back(p);
    }

private:
    Point3 mCenter;
    double mRadius;
    Material* mMaterial;
    Aabb mBBox;
};

// MovingSphere.h:20-36: box over the whole shutter interval.
class MovingSphere : public Hittable {
public:
    MovingSphere(Point3 center0, Point3 center1, double time0, double time1, double radius, Material* material)
        : mCenter0(center0), mCenter1(center1), mTime0(time0), mTime1(time1), mRadius(radius), mMaterial(material)
    {
        const Vector3 rvec(radius, radius, radius);
        const Aabb box0(center0 - rvec, center0 + rvec);
        const Aabb box1(center1 - rvec, center1 + rvec);
        mBBox = Aabb(box0, box1);
    }
    Aabb BoundingBox() const override { return mBBox; }
    void CollectPrims(Emitter& em, const XformChain& chain) const override
    {
        rt_prim p = BlankPrim(RT_PRIM_MOVING_SPHERE, em.MaterialIndex(mMaterial), em, chain);
        for (int k = 0; k < 3; ++k) {
            p.a[k] = mCenter0[k];
            p.b[k] = mCenter1[k];
        }
        p.radius = mRadius;
        p.time0 = mTime0;
        p.time1 = mTime1;
        em.Out().prims.push_back(p);
    }

private:
    Point3 mCenter0, mCenter1;
    double mTime0, mTime1, mRadius;
    Material* mMaterial;
    Aabb mBBox;
};

// Quad.h:25-52: box = union of the boxes of the two diagonals.
class Quad : public Hittable {
public:
    Quad(const Point3& q, const Vector3& u, const Vector3& v, Material* material)
        : mQ(q), mU(u), mV(v), mMaterial(material)
    {
        const Aabb diag1(mQ, mQ + mU + mV);
        const Aabb diag2(mQ + mU, mQ + mV);
        mBBox = Aabb(diag1, diag2);
    }
    Aabb BoundingBox() const override { return mBBox; }
    void CollectPrims(Emitter& em, const XformChain& chain) const override
    {
        rt_prim p = BlankPrim(RT_PRIM_QUAD, em.MaterialIndex(mMaterial), em, chain);
        for (int k = 0; k < 3; ++k) {
            p.a[k] = mQ[k];
            p.b[k] = mU[k];
            p.c[k] = mV[k];
        }
        em.Out().prims.push_back(p);
    }

private:
    Point3 mQ;
    Vector3 mU, mV;
    Material* mMaterial;
    Aabb mBBox;
};

// HittableList.h:21-57
class HittableList : public Hittable {
public:
    HittableList(Hittable** list, int count, bool /*bOwns*/ = false) : mItems(list, list + count)
    {
        for (int i = 0; i < count; ++i) mBBox = Aabb(mBBox, list[i]->BoundingBox());
    }
    Aabb BoundingBox() const override { return mBBox; }
    int ObjectKind() const override { return RT_OBJ_LIST; }
    void CollectPrims(Emitter& em, const XformChain& chain) const override
    {
        for (const Hittable* h : mItems) {
            if (h->AsMedium()) throw std::invalid_argument("rt: ConstantMedium inside a list is not supported");
            h->CollectPrims(em, chain);
        }
    }
    const std::vector<Hittable*>& Items() const { return mItems; }

private:
    std::vector<Hittable*> mItems;
    Aabb mBBox;
};

// Instance.h:28-64
class Translate : public Hittable {
public:
    Translate(Hittable* object, const Vector3& offset) : mObject(object), mOffset(offset)
    {
        mBBox = object->BoundingBox() + offset;
    }
    Aabb BoundingBox() const override { return mBBox; }
    int ObjectKind() const override { return mObject->ObjectKind(); }
    void CollectPrims(Emitter& em, const XformChain& chain) const override
    {
        if (mObject->AsMedium()) throw std::invalid_argument("rt: instanced ConstantMedium is not supported");
        XformChain c = chain;
        rt_xform x;
        std::memset(&x, 0, sizeof x);
        x.type = RT_XFORM_TRANSLATE;
        for (int k = 0; k < 3; ++k) x.v[k] = mOffset[k];
        c.push_back(x);
        mObject->CollectPrims(em, c);
    }

private:
    Hittable* mObject;
    Vector3 mOffset;
    Aabb mBBox;
};

// Instance.h:71-159: sin/cos from degrees, box from the 8 rotated corners.
class RotateY : public Hittable {
public:
    RotateY(Hittable* object, double angle) : mObject(object), mAngle(angle)
    {
        const double radians = angle * 3.1415926535897932385 / 180.0;
        mSinTheta = std::sin(radians);
        mCosTheta = std::cos(radians);
        const Aabb in = object->BoundingBox();
        double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
        double hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 2; ++k) {
                    const double x = i * in.X.Max + (1 - i) * in.X.Min;
                    const double y = j * in.Y.Max + (1 - j) * in.Y.Min;
                    const double z = k * in.Z.Max + (1 - k) * in.Z.Min;
                    const double corner[3] = {mCosTheta * x + mSinTheta * z, y, -mSinTheta * x + mCosTheta * z};
                    for (int c = 0; c < 3; ++c) {
                        lo[c] = std::fmin(lo[c], corner[c]);
                        hi[c] = std::fmax(hi[c], corner[c]);
                    }
                }
        mBBox = Aabb(Point3(lo[0], lo[1], lo[2]), Point3(hi[0], hi[1], hi[2]));
    }
    Aabb BoundingBox() const override { return mBBox; }
    int ObjectKind() const override { return mObject->ObjectKind(); }
    void CollectPrims(Emitter& em, const XformChain& chain) const override
    {
        if (mObject->AsMedium()) throw std::invalid_argument("rt: instanced ConstantMedium is not supported");
        XformChain c = chain;
        rt_xform x;
        std::memset(&x, 0, sizeof x);
        x.type = RT_XFORM_ROTATE_Y;
        x.v[0] = mSinTheta;
        x.v[1] = mCosTheta;
        x.v[2] = mAngle;
        c.push_back(x);
        mObject->CollectPrims(em, c);
    }

private:
    Hittable* mObject;
    double mAngle, mSinTheta, mCosTheta;
    Aabb mBBox;
};

// Instance.h:166-184: six quads, front/right/back/left/top/bottom.
inline Hittable* MakeBox(const Point3& a, const Point3& b, Material* mat)
{
    const Point3 lo(std::fmin(a.X(), b.X()), std::fmin(a.Y(), b.Y()), std::fmin(a.Z(), b.Z()));
    const Point3 hi(std::fmax(a.X(), b.X()), std::fmax(a.Y(), b.Y()), std::fmax(a.Z(), b.Z()));
    const Vector3 dx(hi.X() - lo.X(), 0, 0);
    const Vector3 dy(0, hi.Y() - lo.Y(), 0);
    const Vector3 dz(0, 0, hi.Z() - lo.Z());
    Hittable* sides[6] = {
        new Quad(Point3(lo.X(), lo.Y(), hi.Z()), dx, dy, mat),  // front
        new Quad(Point3(hi.X(), lo.Y(), hi.Z()), -dz, dy, mat), // right
        new Quad(Point3(hi.X(), lo.Y(), lo.Z()), -dx, dy, mat), // back
        new Quad(Point3(lo.X(), lo.Y(), lo.Z()), dz, dy, mat),  // left
        new Quad(Point3(lo.X(), hi.Y(), hi.Z()), dx, -dz, mat), // top
        new Quad(Point3(lo.X(), lo.Y(), lo.Z()), dx, dz, mat),  // bottom
    };
    return new HittableList(sides, 6, true);
}

// ConstantMedium.h:18-104.  The medium id is its construction order; it keys
// the medium's random draws (include/rt_rng.h, trap T3).
class ConstantMedium : public Hittable {
public:
    ConstantMedium(Hittable* boundary, double density, Texture* tex)
        : mBoundary(boundary), mDensity(density), mPhaseFunction(new Isotropic(tex)), mSerial(NextSerial())
    {
    }
    ConstantMedium(Hittable* boundary, double density, const Color& albedo)
        : mBoundary(boundary), mDensity(density), mPhaseFunction(new Isotropic(albedo)), mSerial(NextSerial())
    {
    }
    Aabb BoundingBox() const override { return mBoundary->BoundingBox(); }
    int ObjectKind() const override { return RT_OBJ_MEDIUM; }
    const ConstantMedium* AsMedium() const override { return this; }
    void CollectPrims(Emitter& em, const XformChain& chain) const override
    {
        if (mBoundary->AsMedium()) throw std::invalid_argument("rt: nested ConstantMedium is not supported");
        mBoundary->CollectPrims(em, chain);
    }
    double Density() const { return mDensity; }
    const Material* Phase() const { return mPhaseFunction; }
    unsigned long long Serial() const { return mSerial; }

private:
    Hittable* mBoundary;
    double mDensity;
    Material* mPhaseFunction;
    unsigned long long mSerial;
    static unsigned long long NextSerial()
    {
        static unsigned long long counter = 0;
        return counter++;
    }
};

// BvhNode.h:50-90,170-193.  Host build of the reference's exact topology:
// node box = union of the range, split axis = LongestAxis of that box, stable
// insertion sort on box-min (strict <), midpoint split, span 2 -> two leaves,
// span 1 -> the same leaf twice.  Like the reference constructor this sorts
// `objects[start,end)` in place; the order the caller passed is remembered so
// Flatten() can hand the objects over in construction order.
class BvhNode : public Hittable {
public:
    BvhNode(Hittable** objects, int start, int end) : mOriginal(objects + start, objects + end)
    {
        for (int i = start; i < end; ++i) mBBox = Aabb(mBBox, objects[i]->BoundingBox());
        const int axis = mBBox.LongestAxis();
        const int span = end - start;
        if (span == 1) {
            mLeft = mRight = objects[start];
        } else if (span == 2) {
            mLeft = objects[start];
            mRight = objects[start + 1];
        } else {
            for (int i = start + 1; i < end; ++i) {
                Hittable* key = objects[i];
                const double keyMin = key->BoundingBox().AxisInterval(axis).Min;
                int j = i - 1;
                while (j >= start && keyMin < objects[j]->BoundingBox().AxisInterval(axis).Min) {
                    objects[j + 1] = objects[j];
                    --j;
                }
                objects[j + 1] = key;
            }
            const int mid = start + span / 2;
            mLeft = new BvhNode(objects, start, mid);
            mRight = new BvhNode(objects, mid, end);
        }
    }
    // Book-style convenience: build over a whole list.
    BvhNode(Hittable** objects, int count) : BvhNode(objects, 0, count) {}

    Aabb BoundingBox() const override { return mBBox; }
    bool IsBvhNode() const override { return true; }
    int ObjectKind() const override { return RT_OBJ_LIST; }
    void CollectPrims(Emitter& em, const XformChain& chain) const override
    {
        for (const Hittable* h : mOriginal) {
            if (h->AsMedium()) throw std::invalid_argument("rt: ConstantMedium inside a nested BVH is not supported");
            h->CollectPrims(em, chain);
        }
    }
    const Hittable* Left() const { return mLeft; }
    const Hittable* Right() const { return mRight; }
    const std::vector<Hittable*>& OriginalOrder() const { return mOriginal; }
    int NodeCount() const
    {
        int n = 1;
        if (mLeft->IsBvhNode()) n += static_cast<const BvhNode*>(mLeft)->NodeCount();
        if (mRight->IsBvhNode()) n += static_cast<const BvhNode*>(mRight)->NodeCount();
        return n;
    }

private:
    std::vector<Hittable*> mOriginal;
    Hittable* mLeft;
    Hittable* mRight;
    Aabb mBBox;
};

// ------------------------------------------------------------------- camera
// Camera.h:36-46 argument list; ToAbi() maps it onto rt_camera.
class Camera {
public:
    Camera(Point3 lookfrom, Point3 lookat, Vector3 vup, double vfov, double aspect, double aperture, double focusDist,
           double time0 = 0.0, double time1 = 0.0, Color background = Color(0.70, 0.80, 1.00))
        : mLookfrom(lookfrom), mLookat(lookat), mVup(vup), mVfov(vfov), mAspect(aspect), mAperture(aperture),
          mFocusDist(focusDist), mTime0(time0), mTime1(time1), mBackground(background)
    {
    }
    rt_camera ToAbi(int imageWidth, int imageHeight, int samplesPerPixel, int maxDepth = 50) const
    {
        rt_camera c;
        std::memset(&c, 0, sizeof c);
        c.image_width = imageWidth;
        c.image_height = imageHeight;
        c.samples_per_pixel = samplesPerPixel;
        c.max_depth = maxDepth;
        c.vfov = mVfov;
        for (int k = 0; k < 3; ++k) {
            c.lookfrom[k] = mLookfrom[k];
            c.lookat[k] = mLookat[k];
            c.vup[k] = mVup[k];
            c.background[k] = mBackground[k];
        }
        c.aperture = mAperture;
        c.focus_dist = mFocusDist;
        // aperture = 2*focus_dist*tan(defocus_angle/2)
        c.defocus_angle = 2.0 * std::atan(mAperture / (2.0 * mFocusDist)) * 180.0 / 3.14159265358979323846;
        c.time0 = mTime0;
        c.time1 = mTime1;
        return c;
    }
    double Aspect() const { return mAspect; }

private:
    Point3 mLookfrom, mLookat;
    Vector3 mVup;
    double mVfov, mAspect, mAperture, mFocusDist, mTime0, mTime1;
    Color mBackground;
};

// ------------------------------------------------------------------ flatten
// Writes list[0..count) -- the array the reference hands to
// `new BvhNode(list, 0, i, ...)` (kernel.cu:525) -- into `out`, in the order
// given.  Call it BEFORE building an rt::BvhNode over the same array (the
// build sorts it), or pass the root to the overload below.
inline void Flatten(Hittable* const* list, int count, SceneDesc& out)
{
    Emitter em(out);
    // medium ids = construction order among the media present
    std::vector<const ConstantMedium*> media;
    for (int i = 0; i < count; ++i)
        if (const ConstantMedium* m = list[i]->AsMedium()) media.push_back(m);
    std::vector<const ConstantMedium*> byAge = media;
    for (size_t a = 1; a < byAge.size(); ++a)
        for (size_t b = a; b > 0 && byAge[b]->Serial() < byAge[b - 1]->Serial(); --b) std::swap(byAge[b], byAge[b - 1]);

    for (int i = 0; i < count; ++i) {
        const Hittable* h = list[i];
        rt_object o;
        std::memset(&o, 0, sizeof o);
        o.kind = h->ObjectKind();
        o.first_prim = (int32_t)out.prims.size();
        o.phase_material = -1;
        o.medium_id = -1;
        h->CollectPrims(em, XformChain());
        o.prim_count = (int32_t)out.prims.size() - o.first_prim;
        if (o.prim_count <= 0) throw std::invalid_argument("rt: object without primitives");
        if (const ConstantMedium* m = h->AsMedium()) {
            o.density = m->Density();
            o.phase_material = em.MaterialIndex(m->Phase());
            for (size_t a = 0; a < byAge.size(); ++a)
                if (byAge[a] == m) o.medium_id = (int32_t)a;
        }
        const Aabb bb = h->BoundingBox();
        o.bbox[0] = bb.X.Min;
        o.bbox[1] = bb.X.Max;
        o.bbox[2] = bb.Y.Min;
        o.bbox[3] = bb.Y.Max;
        o.bbox[4] = bb.Z.Min;
        o.bbox[5] = bb.Z.Max;
        out.objects.push_back(o);
    }
    for (size_t k = 0; k < out.images.size(); ++k) out.images[k].rgb = out.image_bytes[k].data();
}

// Flatten from a world root, as the reference keeps it (`*world = root`,
// kernel.cu:528): a BvhNode (objects taken in the order its constructor
// received them) or a plain HittableList.
inline void Flatten(const Hittable* world, SceneDesc& out)
{
    if (world->IsBvhNode()) {
        const auto& v = static_cast<const BvhNode*>(world)->OriginalOrder();
        Flatten(v.data(), (int)v.size(), out);
    } else if (const HittableList* l = dynamic_cast<const HittableList*>(world)) {
        Flatten(l->Items().data(), (int)l->Items().size(), out);
    } else {
        Hittable* one[1] = {const_cast<Hittable*>(world)};
        Flatten(one, 1, out);
    }
}

} // namespace rt
