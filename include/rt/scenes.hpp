// rt/scenes.hpp -- the reference's ten scenes, built on the host.
//
// Restates CreateWorld (reference kernel.cu:176-543), which the reference runs
// as a <<<1,1>>> kernel with device-side `new`, against the host classes of
// rt/scene.hpp.  Scene ids 0..9 are the reference's `sceneId` switch
// (kernel.cu:163-172); id 10 is "Book 1 final", which the reference does not
// carry as such: it is scene 0 with static spheres, a plain grey ground and a
// closed shutter -- the RNG draws, and therefore the sphere
// positions and colours, are those of scene 0.
//
// Random numbers: one rt::Xorwow(1984) stream consumed in source order.  The
// reference writes several draws in one expression, e.g.
//   Vector3 center(a + 0.9 * RND, 0.2, b + 0.9 * RND);        (kernel.cu:216)
// whose evaluation order C++ leaves open; nvcc device code evaluates left to
// right (SURVEY.md trap T1), so every draw below is sequenced explicitly in
// that order.  `RND * RND` is a float*float product (curand_uniform returns
// float), kept as such.
#pragma once

#include <vector>

#include "scene.hpp"

namespace rt {

enum SceneId {
    kSceneBouncingSpheres = 0,
    kSceneCheckeredSpheres = 1,
    kSceneEarth = 2,
    kScenePerlinSpheres = 3,
    kSceneQuads = 4,
    kSceneSimpleLight = 5,
    kSceneCornellEmpty = 6,
    kSceneCornellBoxes = 7,
    kSceneCornellSmoke = 8,
    kSceneFinal = 9,
    kSceneBook1Final = 10,
    kSceneCount = 11
};

struct SceneCamera {
    Vector3 lookfrom{13.0, 2.0, 3.0}; // kernel.cu:189-197 defaults
    Vector3 lookat{0.0, 0.0, 0.0};
    double vfov = 20.0;
    double aperture = 0.0;
    double distToFocus = 10.0;
    double shutterOpen = 0.0;
    double shutterClose = 0.0;
    Color background{0.70, 0.80, 1.00};

    Camera Make(int imageWidth, int imageHeight) const
    {
        // kernel.cu:531-541
        return Camera(lookfrom, lookat, Vector3(0.0, 1.0, 0.0), vfov, double(imageWidth) / double(imageHeight),
                      aperture, distToFocus, shutterOpen, shutterClose, background);
    }
};

namespace detail {

inline void CornellWalls(std::vector<Hittable*>& list, Material* red, Material* white, Material* green,
                         Material* light, bool smokeLayout)
{
    list.push_back(new Quad(Vector3(555, 0, 0), Vector3(0, 555, 0), Vector3(0, 0, 555), green));
    list.push_back(new Quad(Vector3(0, 0, 0), Vector3(0, 555, 0), Vector3(0, 0, 555), red));
    if (!smokeLayout) {
        // kernel.cu:350-355 / 374-379
        list.push_back(new Quad(Vector3(343, 554, 332), Vector3(-130, 0, 0), Vector3(0, 0, -105), light));
        list.push_back(new Quad(Vector3(0, 0, 0), Vector3(555, 0, 0), Vector3(0, 0, 555), white));
        list.push_back(new Quad(Vector3(555, 555, 555), Vector3(-555, 0, 0), Vector3(0, 0, -555), white));
        list.push_back(new Quad(Vector3(0, 0, 555), Vector3(555, 0, 0), Vector3(0, 555, 0), white));
    } else {
        // kernel.cu:411-416: bigger, dimmer light; ceiling/floor written differently
        list.push_back(new Quad(Vector3(113, 554, 127), Vector3(330, 0, 0), Vector3(0, 0, 305), light));
        list.push_back(new Quad(Vector3(0, 555, 0), Vector3(555, 0, 0), Vector3(0, 0, 555), white));
        list.push_back(new Quad(Vector3(0, 0, 0), Vector3(555, 0, 0), Vector3(0, 0, 555), white));
        list.push_back(new Quad(Vector3(0, 0, 555), Vector3(555, 0, 0), Vector3(0, 555, 0), white));
    }
}

inline Hittable* CornellBlock(const Point3& hi, double degrees, const Vector3& offset, Material* mat)
{
    Hittable* box = MakeBox(Point3(0, 0, 0), hi, mat);
    box = new RotateY(box, degrees);
    return new Translate(box, offset);
}

// kernel.cu:199-258 (scene 0) and its Book-1 variant (scene 10).
inline void RandomSpheres(bool book1, Xorwow& rng, std::vector<Hittable*>& list, SceneCamera& cam)
{
    if (book1) {
        list.push_back(new Sphere(Vector3(0.0, -1000.0, -1.0), 1000.0, new Lambertian(Color(0.5, 0.5, 0.5))));
    } else {
        Texture* checker =
            new CheckerTexture(0.32, new SolidColor(Color(0.2, 0.3, 0.1)), new SolidColor(Color(0.9, 0.9, 0.9)));
        list.push_back(new Sphere(Vector3(0.0, -1000.0, -1.0), 1000.0, new Lambertian(checker)));
    }
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            const double chooseMat = rng.Uniform();
            const double cx = a + 0.9 * rng.Uniform();
            const double cz = b + 0.9 * rng.Uniform();
            const Vector3 center(cx, 0.2, cz);
            const Vector3 diff = center - Vector3(4.0, 0.2, 0.0);
            if (diff.Length() <= 0.9) continue;

            if (chooseMat < 0.8) {
                const Vector3 center2 = center + Vector3(0.0, 0.5 * rng.Uniform(), 0.0);
                const float r0 = rng.Uniform(), r1 = rng.Uniform();
                const float g0 = rng.Uniform(), g1 = rng.Uniform();
                const float b0 = rng.Uniform(), b1 = rng.Uniform();
                Material* mat = new Lambertian(Color(r0 * r1, g0 * g1, b0 * b1));
                if (book1)
                    list.push_back(new Sphere(center, 0.2, mat));
                else
                    list.push_back(new MovingSphere(center, center2, 0.0, 1.0, 0.2, mat));
            } else if (chooseMat < 0.95) {
                const double mr = 0.5 * (1.0 + rng.Uniform());
                const double mg = 0.5 * (1.0 + rng.Uniform());
                const double mb = 0.5 * (1.0 + rng.Uniform());
                const double fuzz = 0.5 * rng.Uniform();
                list.push_back(new Sphere(center, 0.2, new Metal(Color(mr, mg, mb), fuzz)));
            } else {
                list.push_back(new Sphere(center, 0.2, new Dielectric(1.5)));
            }
        }
    }
    list.push_back(new Sphere(Vector3(0.0, 1.0, 0.0), 1.0, new Dielectric(1.5)));
    list.push_back(new Sphere(Vector3(-4.0, 1.0, 0.0), 1.0, new Lambertian(Color(0.4, 0.2, 0.1))));
    list.push_back(new Sphere(Vector3(4.0, 1.0, 0.0), 1.0, new Metal(Color(0.7, 0.6, 0.5), 0.0)));

    cam.lookfrom = Vector3(13.0, 2.0, 3.0);
    cam.vfov = 30.0;
    cam.aperture = 0.1;
    cam.shutterOpen = 0.0;
    cam.shutterClose = book1 ? 0.0 : 1.0;
}

// kernel.cu:436-517
inline void FinalScene(Xorwow& rng, const unsigned char* earth, int earthW, int earthH, std::vector<Hittable*>& list,
                       SceneCamera& cam)
{
    Material* ground = new Lambertian(Color(0.48, 0.83, 0.53));
    const int boxesPerSide = 20;
    for (int bi = 0; bi < boxesPerSide; bi++) {
        for (int bj = 0; bj < boxesPerSide; bj++) {
            const double w = 100.0;
            const double x0 = -1000.0 + bi * w;
            const double z0 = -1000.0 + bj * w;
            const double x1 = x0 + w;
            const double y1 = 1.0 + 100.0 * rng.Uniform();
            const double z1 = z0 + w;
            list.push_back(MakeBox(Point3(x0, 0.0, z0), Point3(x1, y1, z1), ground));
        }
    }
    Material* light = new DiffuseLight(Color(7.0, 7.0, 7.0));
    list.push_back(new Quad(Vector3(123, 554, 147), Vector3(300, 0, 0), Vector3(0, 0, 265), light));

    Material* sphereMaterial = new Lambertian(Color(0.7, 0.3, 0.1));
    list.push_back(new MovingSphere(Point3(400, 400, 200), Point3(430, 400, 200), 0.0, 1.0, 50.0, sphereMaterial));

    list.push_back(new Sphere(Point3(260, 150, 45), 50.0, new Dielectric(1.5)));
    list.push_back(new Sphere(Point3(0, 150, 145), 50.0, new Metal(Color(0.8, 0.8, 0.9), 1.0)));

    // glass shell + the blue medium it bounds (two identical spheres, kernel.cu:476-478)
    list.push_back(new Sphere(Point3(360, 150, 145), 70.0, new Dielectric(1.5)));
    Hittable* blueBoundary = new Sphere(Point3(360, 150, 145), 70.0, new Dielectric(1.5));
    list.push_back(new ConstantMedium(blueBoundary, 0.2, Color(0.2, 0.4, 0.9)));

    Hittable* mistBoundary = new Sphere(Point3(0, 0, 0), 5000.0, new Dielectric(1.5));
    list.push_back(new ConstantMedium(mistBoundary, 0.0001, Color(1.0, 1.0, 1.0)));

    Texture* earthTex = new ImageTexture(earth, earthW, earthH);
    list.push_back(new Sphere(Point3(400, 200, 400), 100.0, new Lambertian(earthTex)));

    Texture* pertext = new NoiseTexture(0.2, &rng); // tables drawn here (trap T8)
    list.push_back(new Sphere(Point3(220, 280, 300), 80.0, new Lambertian(pertext)));

    Material* white = new Lambertian(Color(0.73, 0.73, 0.73));
    const int ns = 1000;
    std::vector<Hittable*> cluster(ns);
    for (int s = 0; s < ns; s++) {
        const double x = 165.0 * rng.Uniform();
        const double y = 165.0 * rng.Uniform();
        const double z = 165.0 * rng.Uniform();
        cluster[s] = new Sphere(Point3(x, y, z), 10.0, white);
    }
    Hittable* group = new HittableList(cluster.data(), ns, true);
    group = new RotateY(group, 15.0);
    group = new Translate(group, Vector3(-100, 270, 395));
    list.push_back(group);

    cam.background = Color(0.0, 0.0, 0.0);
    cam.lookfrom = Vector3(478.0, 278.0, -600.0);
    cam.lookat = Vector3(278.0, 278.0, 0.0);
    cam.vfov = 40.0;
    cam.aperture = 0.0;
    cam.shutterOpen = 0.0;
    cam.shutterClose = 1.0;
}

} // namespace detail

// Fills `list` (the reference's list[0..i)) and `cam` for one scene id.
// Must run inside a SceneScope.  `earth` may be null (cyan fallback).
inline void BuildScene(int sceneId, Xorwow& rng, const unsigned char* earth, int earthW, int earthH,
                       std::vector<Hittable*>& list, SceneCamera& cam)
{
    using namespace detail;
    cam = SceneCamera();
    switch (sceneId) {
    case kSceneBouncingSpheres:
        RandomSpheres(false, rng, list, cam);
        break;
    case kSceneBook1Final:
        RandomSpheres(true, rng, list, cam);
        break;
    case kSceneCheckeredSpheres: { // kernel.cu:259-274
        Texture* checker =
            new CheckerTexture(0.32, new SolidColor(Color(0.2, 0.3, 0.1)), new SolidColor(Color(0.9, 0.9, 0.9)));
        list.push_back(new Sphere(Vector3(0.0, -10.0, 0.0), 10.0, new Lambertian(checker)));
        list.push_back(new Sphere(Vector3(0.0, 10.0, 0.0), 10.0, new Lambertian(checker)));
        break;
    }
    case kSceneEarth: { // kernel.cu:275-286
        Texture* earthTex = new ImageTexture(earth, earthW, earthH);
        list.push_back(new Sphere(Vector3(0.0, 0.0, 0.0), 2.0, new Lambertian(earthTex)));
        cam.lookfrom = Vector3(0.0, 0.0, 12.0);
        break;
    }
    case kScenePerlinSpheres: { // kernel.cu:287-299
        Texture* pertext = new NoiseTexture(4.0, &rng);
        list.push_back(new Sphere(Vector3(0.0, -1000.0, 0.0), 1000.0, new Lambertian(pertext)));
        list.push_back(new Sphere(Vector3(0.0, 2.0, 0.0), 2.0, new Lambertian(pertext)));
        break;
    }
    case kSceneQuads: { // kernel.cu:300-320
        list.push_back(new Quad(Vector3(-3, -2, 5), Vector3(0, 0, -4), Vector3(0, 4, 0), new Lambertian(Color(1.0, 0.2, 0.2))));
        list.push_back(new Quad(Vector3(-2, -2, 0), Vector3(4, 0, 0), Vector3(0, 4, 0), new Lambertian(Color(0.2, 1.0, 0.2))));
        list.push_back(new Quad(Vector3(3, -2, 1), Vector3(0, 0, 4), Vector3(0, 4, 0), new Lambertian(Color(0.2, 0.2, 1.0))));
        list.push_back(new Quad(Vector3(-2, 3, 1), Vector3(4, 0, 0), Vector3(0, 0, 4), new Lambertian(Color(1.0, 0.5, 0.0))));
        list.push_back(new Quad(Vector3(-2, -3, 5), Vector3(4, 0, 0), Vector3(0, 0, -4), new Lambertian(Color(0.2, 0.8, 0.8))));
        cam.lookfrom = Vector3(0.0, 0.0, 9.0);
        cam.vfov = 80.0;
        break;
    }
    case kSceneSimpleLight: { // kernel.cu:321-340
        Texture* pertext = new NoiseTexture(4.0, &rng);
        list.push_back(new Sphere(Vector3(0.0, -1000.0, 0.0), 1000.0, new Lambertian(pertext)));
        list.push_back(new Sphere(Vector3(0.0, 2.0, 0.0), 2.0, new Lambertian(pertext)));
        Material* diffLight = new DiffuseLight(Color(4.0, 4.0, 4.0));
        list.push_back(new Sphere(Vector3(0.0, 7.0, 0.0), 2.0, diffLight));
        list.push_back(new Quad(Vector3(3.0, 1.0, -2.0), Vector3(2.0, 0.0, 0.0), Vector3(0.0, 2.0, 0.0), diffLight));
        cam.background = Color(0.0, 0.0, 0.0);
        cam.lookfrom = Vector3(26.0, 3.0, 6.0);
        cam.lookat = Vector3(0.0, 2.0, 0.0);
        break;
    }
    case kSceneCornellEmpty:   // kernel.cu:341-362
    case kSceneCornellBoxes:   // kernel.cu:363-398
    case kSceneCornellSmoke: { // kernel.cu:399-435
        Material* red = new Lambertian(Color(0.65, 0.05, 0.05));
        Material* white = new Lambertian(Color(0.73, 0.73, 0.73));
        Material* green = new Lambertian(Color(0.12, 0.45, 0.15));
        const bool smoke = sceneId == kSceneCornellSmoke;
        Material* light = new DiffuseLight(smoke ? Color(7.0, 7.0, 7.0) : Color(15.0, 15.0, 15.0));
        CornellWalls(list, red, white, green, light, smoke);
        if (sceneId == kSceneCornellBoxes) {
            list.push_back(CornellBlock(Point3(165, 330, 165), 15.0, Vector3(265, 0, 295), white));
            list.push_back(CornellBlock(Point3(165, 165, 165), -18.0, Vector3(130, 0, 65), white));
        } else if (smoke) {
            Hittable* box1 = CornellBlock(Point3(165, 330, 165), 15.0, Vector3(265, 0, 295), white);
            list.push_back(new ConstantMedium(box1, 0.01, Color(0.0, 0.0, 0.0)));
            Hittable* box2 = CornellBlock(Point3(165, 165, 165), -18.0, Vector3(130, 0, 65), white);
            list.push_back(new ConstantMedium(box2, 0.01, Color(1.0, 1.0, 1.0)));
        }
        cam.background = Color(0.0, 0.0, 0.0);
        cam.lookfrom = Vector3(278.0, 278.0, -800.0);
        cam.lookat = Vector3(278.0, 278.0, 0.0);
        cam.vfov = 40.0;
        break;
    }
    case kSceneFinal:
        FinalScene(rng, earth, earthW, earthH, list, cam);
        break;
    default:
        throw std::invalid_argument("rt: unknown scene id");
    }
}

} // namespace rt
