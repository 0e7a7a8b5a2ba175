/*
 * rt_scenes_c.h -- C access to the host-side scene surface.
 *
 * The scene classes (include/rt/scene.hpp) and the reference's ten scenes
 * (include/rt/scenes.hpp, restating CreateWorld, reference kernel.cu:176-543)
 * are C++; these entry points let a non-C++ host (the ctypes binding the tests
 * use, a cgo/JNI stub) build one of them and get the flat rt_scene_desc that
 * rt_scene_upload consumes.  Host only: no CUDA call is made here.
 */
#ifndef RT_SCENES_C_H
#define RT_SCENES_C_H

#include "rt_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_host_scene_s* rt_host_scene;

/* scene_id 0..9 = the reference's sceneId (kernel.cu:163-172); 10 = Book 1 final.
 * earth_rgb: texels for ImageTexture (scenes 2 and 9) as RtwImage produces them,
 * or NULL (cyan fallback, Texture.h:113-114). Scene RNG = XORWOW(1984). */
int rt_host_scene_builtin(int32_t scene_id, const uint8_t* earth_rgb, int32_t earth_w, int32_t earth_h,
                          rt_host_scene* out);
const rt_scene_desc* rt_host_scene_desc(rt_host_scene s);
/* Camera of the scene for an image size (aspect = width/height, kernel.cu:536). */
int rt_host_scene_camera(rt_host_scene s, int32_t width, int32_t height, int32_t samples_per_pixel,
                         int32_t max_depth, rt_camera* out);
/* Uniforms the scene build consumed from the scene stream. */
uint64_t rt_host_scene_rng_draws(rt_host_scene s);
/* Node count of the reference-topology BVH (rt::BvhNode) over the scene. */
int32_t rt_host_scene_reference_bvh_nodes(rt_host_scene s);
int rt_host_scene_free(rt_host_scene s);

/* RtwImage's texel pipeline on decoded 8-bit sRGB input (RtwImage.h:48-50,
 * 100-105 on top of stb_image's stbi_loadf): u8 -> pow(x/255, 2.2) -> (uchar)(256*f).
 * n = number of bytes. */
void rt_image_linearize_rgb8(const uint8_t* srgb, uint8_t* out, uint64_t n);

/* The image half of RtwImage::Load (RtwImage.h:51-87 over stb_image's stbi_loadf, StbImageImpl.cpp:19-21): decodes
 * a baseline JPEG with stb_image v2.30's arithmetic (integer IDCT, tent-filter chroma upsampling, fixed-point YCbCr:
 * the texels come out byte for byte as the reference's) into RGB8, row 0 = top.  linearize != 0 applies
 * rt_image_linearize_rgb8, i.e. returns exactly what RtwImage hands ImageTexture.  Call with rgb_out = NULL to get
 * the size first.  Progressive JPEG and other formats: RT_ERR_UNSUPPORTED. */
int rt_image_decode_jpeg(const uint8_t* bytes, uint64_t n_bytes, int32_t* width, int32_t* height, uint8_t* rgb_out,
                         uint64_t capacity, int32_t linearize);
/* RtwImage::Load(path): reads the file and decodes + linearises it as above. */
int rt_image_load(const char* path, int32_t* width, int32_t* height, uint8_t* rgb_out, uint64_t capacity);

#ifdef __cplusplus
}
#endif
#endif
