// rt_host.cpp -- C entry points over the host scene surface (include/rt_scenes_c.h).
//
// Builds the reference's scenes (restated in include/rt/scenes.hpp from
// reference kernel.cu:176-543) with the host classes of include/rt/scene.hpp
// and hands out the flat FP64 description.  No CUDA here.
#include <cmath>
#include <cstdio>
#include <new>
#include <string>
#include <vector>

#include "../../include/rt/scenes.hpp"
#include "../../include/rt_scenes_c.h"

void rt_set_error(const char* fmt, ...); // rt_error.cpp

struct rt_host_scene_s {
    rt::SceneDesc desc;
    rt_scene_desc view;
    rt::SceneCamera camera;
    uint64_t draws = 0;
    int32_t ref_nodes = 0;
};

extern "C" {

int rt_host_scene_builtin(int32_t scene_id, const uint8_t* earth_rgb, int32_t earth_w, int32_t earth_h,
                          rt_host_scene* out)
{
    if (!out) {
        rt_set_error("rt_host_scene_builtin: out is NULL");
        return RT_ERR_INVALID;
    }
    *out = nullptr;
    try {
        rt_host_scene_s* hs = new rt_host_scene_s();
        {
            rt::SceneScope scope;
            rt::Xorwow rng(1984ULL); // kernel.cu:105
            std::vector<rt::Hittable*> list;
            rt::BuildScene(scene_id, rng, earth_rgb, earth_w, earth_h, list, hs->camera);
            rt::Flatten(list.data(), (int)list.size(), hs->desc);
            hs->draws = rng.Count();
            // kernel.cu:525: the world is a BvhNode over list[0..i)
            rt::BvhNode* root = new rt::BvhNode(list.data(), 0, (int)list.size());
            hs->ref_nodes = root->NodeCount();
        }
        hs->view = hs->desc.View();
        *out = hs;
        return RT_OK;
    } catch (const std::exception& e) {
        rt_set_error("rt_host_scene_builtin(%d): %s", scene_id, e.what());
        return RT_ERR_INVALID;
    }
}

const rt_scene_desc* rt_host_scene_desc(rt_host_scene s) { return s ? &s->view : nullptr; }

int rt_host_scene_camera(rt_host_scene s, int32_t width, int32_t height, int32_t samples_per_pixel,
                         int32_t max_depth, rt_camera* out)
{
    if (!s || !out || width <= 0 || height <= 0) {
        rt_set_error("rt_host_scene_camera: bad argument");
        return RT_ERR_INVALID;
    }
    *out = s->camera.Make(width, height).ToAbi(width, height, samples_per_pixel, max_depth);
    return RT_OK;
}

uint64_t rt_host_scene_rng_draws(rt_host_scene s) { return s ? s->draws : 0; }
int32_t rt_host_scene_reference_bvh_nodes(rt_host_scene s) { return s ? s->ref_nodes : 0; }

int rt_host_scene_free(rt_host_scene s)
{
    delete s;
    return RT_OK;
}

void rt_image_linearize_rgb8(const uint8_t* srgb, uint8_t* out, uint64_t n)
{
    // stb_image v2.30 stbi__ldr_to_hdr: (float)(pow(x/255.0f, 2.2f) * 1.0f), then
    // RtwImage::FloatToByte (RtwImage.h:100-105).
    uint8_t lut[256];
    for (int x = 0; x < 256; ++x) {
        const float f = (float)(std::pow(x / 255.0f, 2.2f) * 1.0f);
        lut[x] = f <= 0.0f ? 0 : (1.0f <= f ? 255 : (uint8_t)(256.0f * f));
    }
    for (uint64_t k = 0; k < n; ++k) out[k] = lut[srgb[k]];
}

// P3 text: the reference's format (kernel.cu:696-723), byte for byte; digits come from a
// 256-entry table instead of a printf per pixel (4K: 95 MB of text in ~0.1 s instead of ~2 s).
static int WritePpm(const char* path, const uint8_t* srgb8, int32_t width, int32_t height, bool binary, const char* who)
{
    if (!path || !srgb8 || width <= 0 || height <= 0) {
        rt_set_error("%s: bad argument", who);
        return RT_ERR_INVALID;
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) {
        rt_set_error("%s: cannot open %s", who, path);
        return RT_ERR_INVALID;
    }
    const size_t n = (size_t)width * height;
    bool ok = std::fprintf(f, "%s\n%d %d\n255\n", binary ? "P6" : "P3", width, height) > 0;
    if (binary) {
        ok = ok && std::fwrite(srgb8, 1, n * 3, f) == n * 3;
    } else {
        char digits[256][4];
        uint8_t len[256];
        for (int v = 0; v < 256; ++v) len[v] = (uint8_t)std::snprintf(digits[v], sizeof digits[v], "%d", v);
        std::vector<char> buf;
        buf.reserve((1u << 20) + 16);
        for (size_t k = 0; k < n && ok; ++k) {
            for (int c = 0; c < 3; ++c) {
                const uint8_t v = srgb8[3 * k + c];
                buf.insert(buf.end(), digits[v], digits[v] + len[v]);
                buf.push_back(c == 2 ? '\n' : ' ');
            }
            if (buf.size() >= (1u << 20)) {
                ok = std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
                buf.clear();
            }
        }
        ok = ok && std::fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) {
        rt_set_error("%s: write to %s failed", who, path);
        return RT_ERR_INVALID;
    }
    return RT_OK;
}

int rt_write_ppm(const char* path, const uint8_t* srgb8, int32_t width, int32_t height)
{
    return WritePpm(path, srgb8, width, height, false, "rt_write_ppm");
}

int rt_write_ppm_binary(const char* path, const uint8_t* srgb8, int32_t width, int32_t height)
{
    return WritePpm(path, srgb8, width, height, true, "rt_write_ppm_binary");
}

} // extern "C"
