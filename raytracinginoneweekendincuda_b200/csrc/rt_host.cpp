// rt_host.cpp -- C entry points over the host scene surface (include/rt_scenes_c.h).
//
// Builds the reference's scenes (restated in include/rt/scenes.hpp from
// reference kernel.cu:176-543) with the host classes of include/rt/scene.hpp
// and hands out the flat FP64 description.  No CUDA here.
#include <cmath>
#include <cstdio>
#include <new>
#include <string>

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

} // extern "C"
