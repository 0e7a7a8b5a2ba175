"""ctypes mirror of include/rt_abi.h and include/rt_scenes_c.h.

Field order and types follow the C headers one-for-one; tests/test_abi.py
checks every sizeof against the values the library itself reports.
"""
from __future__ import annotations

import ctypes as C

RT_ABI_VERSION = 2

RT_OK = 0
RT_ERR_INVALID = -1
RT_ERR_CUDA = -2
RT_ERR_NO_DEVICE = -3
RT_ERR_UNSUPPORTED = -4
RT_ERR_STATE = -5

RT_PRIM_SPHERE, RT_PRIM_MOVING_SPHERE, RT_PRIM_QUAD = 0, 1, 2
RT_XFORM_TRANSLATE, RT_XFORM_ROTATE_Y = 0, 1
RT_OBJ_PRIM, RT_OBJ_LIST, RT_OBJ_MEDIUM = 0, 1, 2
RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT, RT_MAT_ISOTROPIC = 0, 1, 2, 3, 4
RT_TEX_SOLID, RT_TEX_CHECKER, RT_TEX_IMAGE, RT_TEX_NOISE = 0, 1, 2, 3
RT_VARIANT_AUTO, RT_VARIANT_MEGAKERNEL, RT_VARIANT_WAVEFRONT, RT_VARIANT_HEADTAIL, RT_VARIANT_HITQUEUE = 0, 1, 2, 3, 4
RT_BVH_SAH, RT_BVH_REFERENCE, RT_BVH_NONE = 0, 1, 2
RT_FLAG_STATS = 0x100
RT_FLAG_SCENE_IN_GLOBAL = 0x200
RT_FLAG_NODES_IN_GLOBAL = 0x400
RT_FLAG_IMPORTANCE = 0x800
RT_UPLOAD_NO_HOIST = 1
RT_UPLOAD_REDUCE_NCCL = 2
RT_UPLOAD_WHOLE_LISTS = 4
RT_UPLOAD_NO_BOXES = 8

D3 = C.c_double * 3


class rt_prim(C.Structure):
    _fields_ = [("type", C.c_int32), ("material", C.c_int32), ("first_xform", C.c_int32), ("xform_count", C.c_int32),
                ("a", D3), ("b", D3), ("c", D3), ("radius", C.c_double), ("time0", C.c_double),
                ("time1", C.c_double)]


class rt_xform(C.Structure):
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32), ("v", D3)]


class rt_object(C.Structure):
    _fields_ = [("kind", C.c_int32), ("first_prim", C.c_int32), ("prim_count", C.c_int32),
                ("phase_material", C.c_int32), ("medium_id", C.c_int32), ("_pad", C.c_int32),
                ("density", C.c_double), ("bbox", C.c_double * 6)]


class rt_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("texture", C.c_int32), ("albedo", D3), ("fuzz", C.c_double),
                ("ior", C.c_double)]


class rt_texture(C.Structure):
    _fields_ = [("type", C.c_int32), ("even", C.c_int32), ("odd", C.c_int32), ("image", C.c_int32),
                ("perlin", C.c_int32), ("_pad", C.c_int32), ("color", D3), ("scale", C.c_double)]


class rt_perlin(C.Structure):
    _fields_ = [("ranvec", (C.c_double * 3) * 256), ("perm_x", C.c_int32 * 256), ("perm_y", C.c_int32 * 256),
                ("perm_z", C.c_int32 * 256)]


class rt_image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_uint8))]


class rt_scene_desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_objects", C.c_int32), ("n_prims", C.c_int32),
                ("n_xforms", C.c_int32), ("n_materials", C.c_int32), ("n_textures", C.c_int32),
                ("n_perlins", C.c_int32), ("n_images", C.c_int32),
                ("objects", C.POINTER(rt_object)), ("prims", C.POINTER(rt_prim)), ("xforms", C.POINTER(rt_xform)),
                ("materials", C.POINTER(rt_material)), ("textures", C.POINTER(rt_texture)),
                ("perlins", C.POINTER(rt_perlin)), ("images", C.POINTER(rt_image))]


class rt_camera(C.Structure):
    _fields_ = [("image_width", C.c_int32), ("image_height", C.c_int32), ("samples_per_pixel", C.c_int32),
                ("max_depth", C.c_int32), ("vfov", C.c_double), ("lookfrom", D3), ("lookat", D3), ("vup", D3),
                ("defocus_angle", C.c_double), ("focus_dist", C.c_double), ("aperture", C.c_double),
                ("time0", C.c_double), ("time1", C.c_double), ("background", D3)]


class rt_upload_options(C.Structure):
    _fields_ = [("device", C.c_int32), ("bvh", C.c_int32), ("max_leaf_prims", C.c_int32), ("flags", C.c_int32),
                ("n_devices", C.c_int32), ("_pad", C.c_int32), ("device_ids", C.POINTER(C.c_int32))]


class rt_render_params(C.Structure):
    _fields_ = [("sample_begin", C.c_int32), ("sample_end", C.c_int32), ("seed", C.c_uint32),
                ("variant", C.c_int32), ("clear", C.c_int32), ("block_threads", C.c_int32),
                ("blocks_per_sm", C.c_int32), ("flags", C.c_int32), ("stream", C.c_void_p),
                ("accum", C.c_void_p)]


class rt_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("paths", C.c_uint64), ("node_tests", C.c_uint64),
                ("prim_tests", C.c_uint64)]


class rt_scene_info(C.Structure):
    _fields_ = [("n_prims_baked", C.c_int32), ("n_nodes", C.c_int32), ("n_media", C.c_int32),
                ("max_depth_bvh", C.c_int32), ("features", C.c_int32), ("scene_in_smem", C.c_int32),
                ("variant", C.c_int32), ("n_devices", C.c_int32), ("reduce_path", C.c_int32),
                ("block_threads", C.c_int32), ("registers", C.c_int32), ("nodes_in_smem", C.c_int32),
                ("_pad", C.c_int32), ("device_bytes", C.c_uint64),
                ("medium_visits", C.c_int32 * 8)]


class rt_timing(C.Structure):
    _fields_ = [("n_devices", C.c_int32), ("_pad", C.c_int32), ("render_ms", C.c_float * 16),
                ("reduce_ms", C.c_float), ("resolve_ms", C.c_float)]


# void fn(void* user, const float* linear_rgb, const uint8_t* srgb8, int32 samples_done, int32 samples_total)
rt_progress_fn = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.c_int32, C.c_int32)


class rt_pack_info(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_spheres", C.c_int32), ("n_moving", C.c_int32), ("n_quads", C.c_int32),
                ("n_media", C.c_int32), ("n_materials", C.c_int32), ("n_mat_params", C.c_int32),
                ("max_depth_bvh", C.c_int32), ("features", C.c_int32), ("n_hoisted", C.c_int32),
                ("hoisted", C.c_uint32 * 4), ("staged_bytes", C.c_uint64), ("n_boxes", C.c_int32), ("n_lights", C.c_int32)]


def declare_host(lib: C.CDLL) -> None:
    """Prototypes of the host-only entry points (no CUDA call behind them)."""
    lib.rt_last_error.restype = C.c_char_p
    lib.rt_last_error.argtypes = []
    lib.rt_host_scene_builtin.restype = C.c_int
    lib.rt_host_scene_builtin.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
    lib.rt_host_scene_desc.restype = C.POINTER(rt_scene_desc)
    lib.rt_host_scene_desc.argtypes = [C.c_void_p]
    lib.rt_host_scene_camera.restype = C.c_int
    lib.rt_host_scene_camera.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(rt_camera)]
    lib.rt_host_scene_rng_draws.restype = C.c_uint64
    lib.rt_host_scene_rng_draws.argtypes = [C.c_void_p]
    lib.rt_host_scene_reference_bvh_nodes.restype = C.c_int32
    lib.rt_host_scene_reference_bvh_nodes.argtypes = [C.c_void_p]
    lib.rt_host_scene_free.restype = C.c_int
    lib.rt_host_scene_free.argtypes = [C.c_void_p]
    lib.rt_scene_pack_info.restype = C.c_int
    lib.rt_scene_pack_info.argtypes = [C.POINTER(rt_scene_desc), C.POINTER(rt_upload_options), C.POINTER(rt_pack_info)]
    lib.rt_image_decode_jpeg.restype = C.c_int
    lib.rt_image_decode_jpeg.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p,
                                         C.c_uint64, C.c_int32]
    lib.rt_image_load.restype = C.c_int
    lib.rt_image_load.argtypes = [C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_uint64]
    lib.rt_image_linearize_rgb8.restype = None
    lib.rt_image_linearize_rgb8.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]


def declare_device(lib: C.CDLL) -> None:
    """Prototypes of the render path proper (include/rt_abi.h)."""
    lib.rt_scene_upload.restype = C.c_int
    lib.rt_scene_upload.argtypes = [C.POINTER(rt_scene_desc), C.POINTER(rt_upload_options), C.POINTER(C.c_void_p)]
    lib.rt_render.restype = C.c_int
    lib.rt_render.argtypes = [C.c_void_p, C.POINTER(rt_camera), C.POINTER(rt_render_params)]
    lib.rt_render_progressive.restype = C.c_int
    lib.rt_render_progressive.argtypes = [C.c_void_p, C.POINTER(rt_camera), C.POINTER(rt_render_params), C.c_int32,
                                          C.c_int32, C.c_int32, rt_progress_fn, C.c_void_p]
    lib.rt_get_timing.restype = C.c_int
    lib.rt_get_timing.argtypes = [C.c_void_p, C.POINTER(rt_timing)]
    lib.rt_accum_ptr.restype = C.c_int
    lib.rt_accum_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.rt_sync.restype = C.c_int
    lib.rt_sync.argtypes = [C.c_void_p]
    lib.rt_readback.restype = C.c_int
    lib.rt_readback.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(rt_stats)]
    lib.rt_scene_free.restype = C.c_int
    lib.rt_scene_free.argtypes = [C.c_void_p]
    lib.rt_scene_get_info.restype = C.c_int
    lib.rt_scene_get_info.argtypes = [C.c_void_p, C.POINTER(rt_scene_info)]
    lib.rt_rng_uniform.restype = C.c_float
    lib.rt_rng_uniform.argtypes = [C.c_uint32] * 6
    lib.rt_write_ppm.restype = C.c_int
    lib.rt_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.rt_write_ppm_binary.restype = C.c_int
    lib.rt_write_ppm_binary.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.rt_debug_trace_path.restype = C.c_int
    lib.rt_debug_trace_path.argtypes = [C.c_void_p, C.POINTER(rt_camera), C.POINTER(rt_render_params), C.c_int32,
                                        C.c_int32, C.c_void_p, C.c_int32]
    lib.rt_measure_fp32_peak.restype = C.c_int
    lib.rt_measure_fp32_peak.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.rt_release_cached_memory.restype = C.c_int
    lib.rt_release_cached_memory.argtypes = []
    lib.rt_abi_sizeof.restype = C.c_int
    lib.rt_abi_sizeof.argtypes = [C.c_char_p]
