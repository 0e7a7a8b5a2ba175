// rt_cli.cpp -- command-line renderer: the reference's main() (reference
// kernel.cu:570-742) with its hard-coded constants (image size :572-573, scene
// :589, samples :593, depth :71) turned into flags.  Builds the scene on the
// host with the reference-style classes, renders through the C ABI, writes the
// same P3 PPM.
//
//   rt_cli --scene 10 --width 1200 --height 675 --spp 10 --depth 50 --out out.ppm
//          [--seed 1984] [--device 0 | --gpus N] [--earth earthmap.jpg] [--bvh sah|reference|list]
//          [--variant auto|hitqueue|headtail|megakernel|wavefront] [--p6] [--progressive BATCH] [--nccl] [--importance]
//   --gpus N         one process, N devices: samples split across them, accumulators reduced on device 0
//   --earth FILE     the image texture of scenes 2 and 9: a baseline JPEG, decoded by the library as RtwImage does
//   --progressive B  re-write the output file after every B samples while the next batch renders
//   --importance     RT_FLAG_IMPORTANCE: diffuse bounces sample the lights (same image in expectation, less noise)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rt/scenes.hpp"
#include "../../include/rt_abi.h"
#include "../../include/rt_scenes_c.h"

static int Fail(const char* what)
{
    std::fprintf(stderr, "rt_cli: %s: %s\n", what, rt_last_error());
    return 1;
}

struct Progress {
    std::string out;
    int width, height;
    bool binary;
    std::chrono::steady_clock::time_point t0;
};

static void OnFrame(void* user, const float*, const uint8_t* srgb8, int32_t done, int32_t total)
{
    Progress* pr = static_cast<Progress*>(user);
    const std::string tmp = pr->out + ".part";
    if ((pr->binary ? rt_write_ppm_binary : rt_write_ppm)(tmp.c_str(), srgb8, pr->width, pr->height) == RT_OK)
        std::rename(tmp.c_str(), pr->out.c_str());
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - pr->t0).count();
    std::fprintf(stderr, "  %d / %d samples after %.3f s -> %s\n", done, total, sec, pr->out.c_str());
}

int main(int argc, char** argv)
{
    int sceneId = 9, width = 1440, height = 720, spp = -1, depth = 50, device = 0, bvh = RT_BVH_SAH, gpus = 1;
    int progressive = 0, uploadFlags = 0, renderFlags = 0;
    unsigned seed = 1984;
    bool binary = false;
    int variant = RT_VARIANT_AUTO;
    std::string out = "output.ppm", earthPath;
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
        else if (a == "--gpus") gpus = std::atoi(next("--gpus"));
        else if (a == "--out") out = next("--out");
        else if (a == "--earth") earthPath = next("--earth");
        else if (a == "--progressive") progressive = std::atoi(next("--progressive"));
        else if (a == "--nccl") uploadFlags |= RT_UPLOAD_REDUCE_NCCL;
        else if (a == "--importance") renderFlags |= RT_FLAG_IMPORTANCE;
        else if (a == "--p6") binary = true;
        else if (a == "--variant") {
            const std::string v = next("--variant");
            if (v == "auto") variant = RT_VARIANT_AUTO;
            else if (v == "hitqueue") variant = RT_VARIANT_HITQUEUE;
            else if (v == "headtail") variant = RT_VARIANT_HEADTAIL;
            else if (v == "megakernel") variant = RT_VARIANT_MEGAKERNEL;
            else if (v == "wavefront") variant = RT_VARIANT_WAVEFRONT;
            else {
                std::fprintf(stderr, "rt_cli: unknown variant %s\n", v.c_str());
                return 2;
            }
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
    if (gpus < 1 || gpus > 16 || spp <= 0 || width <= 0 || height <= 0) {
        std::fprintf(stderr, "rt_cli: bad --gpus / --spp / image size\n");
        return 2;
    }

    // RtwImage("earthmap.jpg") (kernel.cu:474): decoded by the library; a missing file renders cyan like the reference
    std::vector<unsigned char> earth;
    int earthW = 0, earthH = 0;
    if (!earthPath.empty()) {
        if (rt_image_load(earthPath.c_str(), &earthW, &earthH, nullptr, 0) == RT_OK) {
            earth.resize((size_t)earthW * earthH * 3);
            if (rt_image_load(earthPath.c_str(), &earthW, &earthH, earth.data(), earth.size()) != RT_OK) earth.clear();
        }
        if (earth.empty())
            std::fprintf(stderr, "rt_cli: could not load %s (%s); the image texture renders cyan\n", earthPath.c_str(), rt_last_error());
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
    std::fprintf(stderr, "Rendering a %dx%d image with %d samples per pixel (scene %d, %zu objects, %zu primitives) on %d GPU(s).\n",
                 width, height, spp, sceneId, desc.objects.size(), desc.prims.size(), gpus);

    const rt_scene_desc view = desc.View();
    rt_upload_options opt{};
    opt.device = device;
    opt.bvh = bvh;
    opt.flags = uploadFlags;
    opt.n_devices = gpus > 1 ? gpus : 0; // devices 0 .. gpus-1
    rt_scene_handle scene = nullptr;
    if (rt_scene_upload(&view, &opt, &scene) != RT_OK) return Fail("rt_scene_upload");

    rt_render_params p{};
    p.sample_begin = 0;
    p.sample_end = spp;
    p.seed = seed;
    p.clear = 1;
    p.variant = variant;
    p.flags = renderFlags;
    const auto t0 = std::chrono::steady_clock::now();
    if (progressive > 0) {
        Progress pr{out, width, height, binary, t0};
        if (rt_render_progressive(scene, &cam, &p, progressive, 0, 1, OnFrame, &pr) != RT_OK) return Fail("rt_render_progressive");
    } else {
        if (rt_render(scene, &cam, &p) != RT_OK) return Fail("rt_render");
        if (rt_sync(scene) != RT_OK) return Fail("rt_sync");
    }
    const auto t1 = std::chrono::steady_clock::now();
    rt_timing tm{};
    if (progressive <= 0 && rt_get_timing(scene, &tm) == RT_OK && tm.n_devices > 1) {
        std::fprintf(stderr, "per-device kernel ms:");
        for (int k = 0; k < tm.n_devices; ++k) std::fprintf(stderr, " %.2f", tm.render_ms[k]);
        std::fprintf(stderr, "\n");
    }
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
