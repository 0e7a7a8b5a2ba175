// rt_cli.cpp -- command-line renderer: the reference's main() (reference
// kernel.cu:570-742) with its hard-coded constants (image size :572-573, scene
// :589, samples :593, depth :71) turned into flags.  Builds the scene on the
// host with the reference-style classes, renders through the C ABI, writes the
// same P3 PPM.
//
//   rt_cli --scene 10 --width 1200 --height 675 --spp 10 --depth 50 --out out.ppm
//          [--seed 1984] [--device 0] [--earth earthmap.rgb8 W H] [--bvh sah|reference|list]
//          [--variant megakernel|wavefront] [--p6]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rt/scenes.hpp"
#include "../../include/rt_abi.h"

static int Fail(const char* what)
{
    std::fprintf(stderr, "rt_cli: %s: %s\n", what, rt_last_error());
    return 1;
}

int main(int argc, char** argv)
{
    int sceneId = 9, width = 1440, height = 720, spp = -1, depth = 50, device = 0, bvh = RT_BVH_SAH;
    unsigned seed = 1984;
    bool binary = false;
    int variant = RT_VARIANT_AUTO;
    std::string out = "output.ppm", earthPath;
    int earthW = 0, earthH = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char* name) -> const char* {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "rt_cli: %s needs a value\n", name);
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--scene") sceneId = std::atoi(next("--scene"));
        else if (a == "--width") width = std::atoi(next("--width"));
        else if (a == "--height") height = std::atoi(next("--height"));
        else if (a == "--spp") spp = std::atoi(next("--spp"));
        else if (a == "--depth") depth = std::atoi(next("--depth"));
        else if (a == "--seed") seed = (unsigned)std::strtoul(next("--seed"), nullptr, 10);
        else if (a == "--device") device = std::atoi(next("--device"));
        else if (a == "--out") out = next("--out");
        else if (a == "--earth") {
            earthPath = next("--earth");
            earthW = std::atoi(next("--earth W"));
            earthH = std::atoi(next("--earth H"));
        } else if (a == "--p6") {
            binary = true;
        } else if (a == "--variant") {
            const std::string v = next("--variant");
            variant = v == "wavefront" ? RT_VARIANT_WAVEFRONT : (v == "megakernel" ? RT_VARIANT_MEGAKERNEL : RT_VARIANT_AUTO);
        } else if (a == "--bvh") {
            const std::string v = next("--bvh");
            bvh = v == "reference" ? RT_BVH_REFERENCE : (v == "list" ? RT_BVH_NONE : RT_BVH_SAH);
        } else {
            std::fprintf(stderr, "rt_cli: unknown flag %s\n", a.c_str());
            return 2;
        }
    }
    // kernel.cu:593: the reference's per-scene sample counts
    if (spp < 0) spp = (sceneId == 9) ? 100 : ((sceneId >= 5 && sceneId <= 8) ? 200 : 10);

    std::vector<unsigned char> earth;
    if (!earthPath.empty()) {
        FILE* f = std::fopen(earthPath.c_str(), "rb");
        if (f) {
            earth.resize((size_t)earthW * earthH * 3);
            if (std::fread(earth.data(), 1, earth.size(), f) != earth.size()) earth.clear();
            std::fclose(f);
        }
        if (earth.empty()) std::fprintf(stderr, "rt_cli: could not read %s; the image texture renders cyan\n", earthPath.c_str());
    }

    rt::SceneDesc desc;
    rt_camera cam;
    {
        rt::SceneScope scope;
        rt::Xorwow rng(1984ULL);
        std::vector<rt::Hittable*> list;
        rt::SceneCamera sc;
        rt::BuildScene(sceneId, rng, earth.empty() ? nullptr : earth.data(), earthW, earthH, list, sc);
        rt::Flatten(list.data(), (int)list.size(), desc);
        cam = sc.Make(width, height).ToAbi(width, height, spp, depth);
    }
    std::fprintf(stderr, "Rendering a %dx%d image with %d samples per pixel (scene %d, %zu objects, %zu primitives).\n",
                 width, height, spp, sceneId, desc.objects.size(), desc.prims.size());

    const rt_scene_desc view = desc.View();
    rt_upload_options opt{};
    opt.device = device;
    opt.bvh = bvh;
    rt_scene_handle scene = nullptr;
    if (rt_scene_upload(&view, &opt, &scene) != RT_OK) return Fail("rt_scene_upload");

    rt_render_params p{};
    p.sample_begin = 0;
    p.sample_end = spp;
    p.seed = seed;
    p.clear = 1;
    p.variant = variant;
    const auto t0 = std::chrono::steady_clock::now();
    if (rt_render(scene, &cam, &p) != RT_OK) return Fail("rt_render");
    if (rt_sync(scene) != RT_OK) return Fail("rt_sync");
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<uint8_t> srgb((size_t)width * height * 3);
    rt_stats st{};
    if (rt_readback(scene, nullptr, nullptr, srgb.data(), &st) != RT_OK) return Fail("rt_readback");
    const double sec = std::chrono::duration<double>(t1 - t0).count();
    std::fprintf(stderr, "took %.4f seconds: %llu rays, %.1f Mrays/s.\n", sec, (unsigned long long)st.rays,
                 (double)st.rays / sec * 1e-6);
    if ((binary ? rt_write_ppm_binary : rt_write_ppm)(out.c_str(), srgb.data(), width, height) != RT_OK)
        return Fail("rt_write_ppm");
    std::fprintf(stderr, "Done. Saved to %s\n", out.c_str());
    rt_scene_free(scene);
    return 0;
}
