// ref_stream_main.cpp -- drives the REFERENCE'S OWN scene classes on the host.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/build_ref.py into
// oracle/_ref/libref_stream.so (git-ignored), from the reference headers where
// they lie under /root/reference plus the shims in oracle/shim/.  It is what
// pins oracle/rt_oracle.cpp: same scenes, same random stream, the reference's
// Hit/Scatter/Value code doing the work.
//
// What is restated here (it lives in the reference's kernel.cu next to main()
// and the kernel launches, so it cannot be included):
//   RayColor   reference kernel.cu:65-98    (max depth is a parameter, not 50)
//   Render     reference kernel.cu:122-154  (loop over pixels instead of a grid;
//                                            returns the linear sum, before the
//                                            divide/sqrt at :147-152)
// What is the reference's, verbatim: CreateWorld (kernel.cu:157-545) is pulled
// in as text by build_ref.py ("create_world.inc"), and every class header.
// The sequencing patches build_ref.py applies (trap T1) call the helpers
// declared just below.
#include <atomic>
#include <cfloat>
#include <cstdio>
#include <thread>
#include <type_traits>
#include <vector>

#include "cuda_runtime.h"
#include <curand_kernel.h>

#include "Vec3.h"
#include "Ray.h"
#include "Hittable.h"
#include "HittableList.h"
#include "BvhNode.h"
#include "Sphere.h"
#include "MovingSphere.h"
#include "Quad.h"
#include "Instance.h"
#include "ConstantMedium.h"
#include "Texture.h"
#include "Material.h"
#include "Metal.h"
#include "Dielectric.h"
#include "Camera.h"

static_assert(std::is_same<decltype(log(1.0f)), float>::value,
              "log(float) must resolve to the float overload, as it does under nvcc");

static int g_nextMediumId = 0;
int rtshim_next_medium_id() { return g_nextMediumId++; }

// Left-to-right sequencing of the multi-draw expressions in CreateWorld
// (kernel.cu:216,229,237-238,502); nvcc device code evaluates them in this order.
static Color RtShimColorProducts(curandState* s)
{
    const float r0 = curand_uniform(s), r1 = curand_uniform(s);
    const float g0 = curand_uniform(s), g1 = curand_uniform(s);
    const float b0 = curand_uniform(s), b1 = curand_uniform(s);
    return Color(r0 * r1, g0 * g1, b0 * b1);
}
static Material* RtShimNewMetal(curandState* s)
{
    const double r = 0.5 * (1.0 + curand_uniform(s));
    const double g = 0.5 * (1.0 + curand_uniform(s));
    const double b = 0.5 * (1.0 + curand_uniform(s));
    const double fuzz = 0.5 * curand_uniform(s);
    return new Metal(Color(r, g, b), fuzz);
}
static Point3 RtShimPoint(curandState* s, double scale)
{
    const double x = scale * curand_uniform(s);
    const double y = scale * curand_uniform(s);
    const double z = scale * curand_uniform(s);
    return Point3(x, y, z);
}
static Vector3 RtShimCenter(curandState* s, int a, int b)
{
    const double x = a + 0.9 * curand_uniform(s);
    const double z = b + 0.9 * curand_uniform(s);
    return Vector3(x, 0.2, z);
}

#include "create_world.inc"

namespace {

// kernel.cu:65-98
Color RayColorRestated(const Ray& r, const Color& background, Hittable** world, curandState* rs, int maxDepth,
                       unsigned long long& rays)
{
    Ray current = r;
    Color throughput(1.0, 1.0, 1.0);
    Color accumulated(0.0, 0.0, 0.0);
    for (int i = 0; i < maxDepth; i++) {
        rtshim_slot(rs, (uint32_t)i + 1u);
        ++rays;
        HitRecord rec;
        rec.U = rec.V = 0.0;
        if (!(*world)->Hit(current, 0.001, DBL_MAX, rec, rs)) {
            accumulated += throughput * background;
            return accumulated;
        }
        Color emission = rec.MaterialPtr->Emitted(rec.U, rec.V, rec.P);
        accumulated += throughput * emission;
        Ray scattered;
        Color attenuation;
        if (!rec.MaterialPtr->Scatter(current, rec, attenuation, scattered, rs)) return accumulated;
        throughput = throughput * attenuation;
        current = scattered;
    }
    return accumulated;
}

struct World {
    std::vector<Hittable*> list;
    std::vector<Hittable*> nodes;
    Hittable* root = nullptr;
    Camera* camera = nullptr;
    int count = 0, nodeCount = 0;
    std::vector<double> bboxInConstructionOrder;
    unsigned long long sceneDraws = 0;
};

// Runs the reference's CreateWorld (scene + BVH + camera) on the host.
void BuildWorld(World& w, int sceneId, int W, int H, const unsigned char* earth, int ew, int eh)
{
    g_nextMediumId = 0;
    curandState sceneRng;
    curand_init(1984, 0, 0, &sceneRng); // kernel.cu:105
    w.list.assign(4096, nullptr);
    w.nodes.assign(8192, nullptr);
    CreateWorld(w.list.data(), &w.root, &w.camera, W, H, &sceneRng, &w.count, w.nodes.data(), &w.nodeCount, sceneId,
                earth, ew, eh);
    w.sceneDraws = sceneRng.draws;
}

} // namespace

extern "C" {

struct ref_stream_stats {
    unsigned long long rays, paths, draws, scene_draws;
    int n_objects, n_nodes;
};

// out: W*H*3 doubles, row 0 = bottom, linear SUM over samples [s0,s1).
int ref_stream_render(int sceneId, int W, int H, int s0, int s1, int maxDepth, unsigned seed,
                      const unsigned char* earth, int earthW, int earthH, int nThreads, double* out,
                      ref_stream_stats* stats)
{
    World w;
    BuildWorld(w, sceneId, W, H, earth, earthW, earthH);
    Hittable* worldPtr = w.root;
    Camera* cam = w.camera;
    const Color background = cam->Background();
    if (nThreads < 1) nThreads = 1;
    std::atomic<int> nextRow(0);
    std::vector<unsigned long long> rays((size_t)nThreads, 0), draws((size_t)nThreads, 0);
    auto work = [&](int tid) {
        curandState rs;
        while (true) {
            const int j = nextRow.fetch_add(1);
            if (j >= H) break;
            for (int i = 0; i < W; ++i) {
                const int pixelIndex = j * W + i; // kernel.cu:131
                Color col(0.0, 0.0, 0.0);
                for (int s = s0; s < s1; ++s) { // kernel.cu:138-144
                    rtshim_key(&rs, seed, (uint32_t)pixelIndex, (uint32_t)s);
                    rtshim_slot(&rs, 0);
                    const float fu = i + curand_uniform(&rs);
                    const float fv = j + curand_uniform(&rs);
                    const double u = double(fu) / double(W);
                    const double v = double(fv) / double(H);
                    Ray r = cam->GetRay(u, v, &rs);
                    col += RayColorRestated(r, background, &worldPtr, &rs, maxDepth, rays[(size_t)tid]);
                }
                out[(size_t)pixelIndex * 3 + 0] = col.X();
                out[(size_t)pixelIndex * 3 + 1] = col.Y();
                out[(size_t)pixelIndex * 3 + 2] = col.Z();
            }
        }
        draws[(size_t)tid] = rs.draws;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nThreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    if (stats) {
        stats->rays = 0;
        stats->draws = 0;
        for (int t = 0; t < nThreads; ++t) {
            stats->rays += rays[(size_t)t];
            stats->draws += draws[(size_t)t];
        }
        stats->paths = (unsigned long long)W * H * (unsigned long long)(s1 - s0);
        stats->scene_draws = w.sceneDraws;
        stats->n_objects = w.count;
        stats->n_nodes = w.nodeCount;
    }
    return 0; // the world is leaked on purpose: FreeWorld's delete graph is not part of the path
}

// Bounding boxes of list[0..n) AFTER the BVH build sorted it (6 doubles each:
// xmin,xmax,ymin,ymax,zmin,zmax); returns n.  As a set this pins the host scene
// builders of include/rt/scenes.hpp against the reference's CreateWorld.
int ref_stream_scene_boxes(int sceneId, int W, int H, const unsigned char* earth, int earthW, int earthH, double* out,
                           int capacity, unsigned long long* sceneDraws)
{
    World w;
    BuildWorld(w, sceneId, W, H, earth, earthW, earthH);
    for (int i = 0; i < w.count && i < capacity; ++i) {
        const Aabb b = w.list[(size_t)i]->BoundingBox();
        out[6 * i + 0] = b.X.Min;
        out[6 * i + 1] = b.X.Max;
        out[6 * i + 2] = b.Y.Min;
        out[6 * i + 3] = b.Y.Max;
        out[6 * i + 4] = b.Z.Min;
        out[6 * i + 5] = b.Z.Max;
    }
    if (sceneDraws) *sceneDraws = w.sceneDraws;
    return w.count;
}

} // extern "C"
