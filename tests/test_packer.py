"""The host packer behind rt_scene_upload (csrc/rt_pack.hpp), through the host-only rt_scene_pack_info:
hoisting of scene-sized items, stack-depth accounting, validation of malformed input.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from conftest import A, BoxScene, HandScene
from raytracinginoneweekendincuda_b200 import BuiltinScene


def pack(lib, desc, bvh=A.RT_BVH_SAH, flags=0, max_leaf=0):
    info = A.rt_pack_info()
    opt = A.rt_upload_options(device=0, bvh=bvh, max_leaf_prims=max_leaf, flags=flags)
    rc = lib.rt_scene_pack_info(desc, C.byref(opt), C.byref(info))
    return rc, info


def test_hoisting_policy(lib, earth):
    """What is tested before the tree instead of inside it (csrc/rt_pack.hpp): scene-sized surfaces (Book 1 / scene 0:
    the r = 1000 ground sphere), every ConstantMedium (scene 9: the blue-glass medium and the r = 5000 mist whose box
    contains the whole scene, reference kernel.cu:476-482; scene 8: the two smoke boxes), and all surfaces of a tiny
    scene (scene 8: the six Cornell walls as one run of quads -- its tree is then empty)."""
    expect = {10: [0], 0: [0], 9: [3, 3], 8: [2, 3, 3], 7: [2, 4]}  # leaf-ref types: 0 sphere, 2 quad run, 3 medium, 4 box run
    for sid, types in expect.items():
        sc = BuiltinScene(sid, earth if sid == 9 else None)
        rc, i = pack(lib, sc.desc)
        assert rc == 0 and i.n_hoisted == len(types), (sid, i.n_hoisted)
        for k, t in enumerate(types):
            assert (i.hoisted[k] >> 31) == 1 and ((i.hoisted[k] >> 28) & 7) == t, (sid, k, hex(i.hoisted[k]))
        rc, j = pack(lib, sc.desc, flags=A.RT_UPLOAD_NO_HOIST)
        assert rc == 0 and j.n_hoisted == 0
        assert j.n_nodes >= i.n_nodes
        assert (i.n_spheres, i.n_moving, i.n_quads, i.n_media) == (j.n_spheres, j.n_moving, j.n_quads, j.n_media)
    sc = BuiltinScene(8)
    rc, i = pack(lib, sc.desc)
    assert i.n_nodes == 2 and i.max_depth_bvh == 0 and ((i.hoisted[0] >> 18) & 1023) + 1 == 6  # empty tree, run of 6 quads


def test_closed_six_quad_lists_become_slab_tested_boxes(lib, earth):
    """rt_pack.hpp MakeDevBox: MakeBox lists (Instance.h:166-184), with or without RotateY / Translate around them,
    are one BVH item each, tested by one slab test; their six quads stay in the quad table (hits are reported against
    them).  Scene 7: the two Cornell boxes; scene 8: the two smoke boundaries; scene 9: the 400 ground boxes -- half
    the nodes, and the node table then fits in shared memory.  Reference / list topologies keep the quads."""
    for sid, boxes in ((7, 2), (8, 2), (9, 400)):
        sc = BuiltinScene(sid, earth if sid == 9 else None)
        rc, i = pack(lib, sc.desc)
        rc2, j = pack(lib, sc.desc, flags=A.RT_UPLOAD_NO_BOXES)
        assert rc == 0 and rc2 == 0 and i.n_boxes == boxes and j.n_boxes == 0, (sid, i.n_boxes, j.n_boxes)
        assert i.n_quads == j.n_quads and i.n_nodes <= j.n_nodes
        for bvh in (A.RT_BVH_REFERENCE, A.RT_BVH_NONE):
            rc, k = pack(lib, sc.desc, bvh=bvh)
            assert rc == 0 and k.n_boxes == (boxes if sid == 8 else 0)  # (a medium's boundary is not a BVH item)
    sc = BuiltinScene(9, earth)
    rc, i = pack(lib, sc.desc)
    assert i.n_nodes * 32 < 100 * 1024


def test_boxes_under_instances_and_as_medium_boundaries(lib):
    """conftest.BoxScene: a rotated + translated glass box, a plain metal box, a medium bounded by a rotated box: three
    DevBox records; the scene is small enough to be tested flat (ground sphere, the two surface boxes, the medium)."""
    sc = BoxScene()
    rc, i = pack(lib, sc.desc)
    assert rc == 0, lib.rt_last_error()
    assert (i.n_boxes, i.n_quads, i.n_spheres, i.n_media) == (3, 18, 1, 1)
    assert i.n_hoisted == 3 and [(i.hoisted[k] >> 28) & 7 for k in range(3)] == [0, 4, 3]
    rc, j = pack(lib, sc.desc, flags=A.RT_UPLOAD_NO_BOXES)
    assert rc == 0 and j.n_boxes == 0 and j.n_quads == 18


def _six_quad_scene(shift_top=0.0, tilt=0.0):
    """One owning list of six quads laid out like MakeBox((0,0,0),(1,2,3)) (Instance.h:166-184) plus a sphere."""
    lo, hi = (0.0, 0.0, 0.0), (1.0, 2.0, 3.0)
    dx, dy, dz = (1.0, 0.0, 0.0), (0.0, 2.0, 0.0), (0.0, 0.0, 3.0)

    def neg(v):
        return tuple(-c for c in v)

    quads = [((lo[0], lo[1], hi[2]), dx, dy), ((hi[0], lo[1], hi[2]), neg(dz), dy), ((hi[0], lo[1], lo[2]), neg(dx), dy),
             ((lo[0], lo[1], lo[2]), dz, dy), ((lo[0], hi[1] + shift_top, hi[2]), dx, (0.0, tilt, -3.0)),
             ((lo[0], lo[1], lo[2]), dx, dz)]
    prims = (A.rt_prim * 7)()
    for k, (q, u, v) in enumerate(quads):
        prims[k].type = A.RT_PRIM_QUAD
        prims[k].a[:] = list(q)
        prims[k].b[:] = list(u)
        prims[k].c[:] = list(v)
    prims[6].type = A.RT_PRIM_SPHERE
    prims[6].a[:] = [5.0, 5.0, 5.0]
    prims[6].radius = 1.0
    objs = (A.rt_object * 2)()
    objs[0].kind = A.RT_OBJ_LIST
    objs[0].first_prim, objs[0].prim_count = 0, 6
    objs[0].bbox[:] = [0, 1, 0, 2.5, 0, 3]
    objs[1].kind = A.RT_OBJ_PRIM
    objs[1].first_prim, objs[1].prim_count = 6, 1
    objs[1].bbox[:] = [4, 6, 4, 6, 4, 6]
    mats = (A.rt_material * 1)()
    mats[0].type = A.RT_MAT_LAMBERTIAN
    mats[0].texture = 0
    tex = (A.rt_texture * 1)()
    tex[0].type = A.RT_TEX_SOLID
    d = A.rt_scene_desc(abi_version=A.RT_ABI_VERSION, n_objects=2, n_prims=7, n_materials=1, n_textures=1,
                        objects=objs, prims=prims, materials=mats, textures=tex)
    d._keep = (prims, objs, mats, tex)
    return d


def test_an_open_or_skewed_six_quad_list_is_not_a_box(lib):
    """Six quads that do not close a box (one face lifted off; one face tilted) stay six quads."""
    for kw, want in ((dict(), 1), (dict(shift_top=0.25), 0), (dict(tilt=0.1), 0)):
        d = _six_quad_scene(**kw)
        rc, i = pack(lib, C.byref(d))
        assert rc == 0, lib.rt_last_error()
        assert i.n_boxes == want and i.n_quads == 6, (kw, i.n_boxes)


def test_reference_and_list_modes_never_hoist(lib):
    sc = BuiltinScene(10)
    for bvh in (A.RT_BVH_REFERENCE, A.RT_BVH_NONE):
        rc, i = pack(lib, sc.desc, bvh=bvh)
        assert rc == 0 and i.n_hoisted == 0


def test_book1_fits_shared_memory_with_24_warps(lib):
    """staged scene + traversal stacks + hit queues of 24 warps within the 227 KB a CTA may use."""
    sc = BuiltinScene(10)
    rc, i = pack(lib, sc.desc)
    assert rc == 0
    assert i.n_materials == 485 and i.n_mat_params > 50  # metals + dielectrics carry an FP64 parameter
    stack = 768 * 4 * (i.max_depth_bvh + 3)
    queues = 24 * (64 * 68 + 384)
    assert i.staged_bytes + stack + queues + 16 <= 232448


def _mixed_leaf_scene(n_types):
    """Coincident primitives of several types: SAH cannot separate them, so they end in ONE mixed leaf, which the
    packer chains through join nodes (one stack level each)."""
    prims = (A.rt_prim * 8)()
    objs = (A.rt_object * 8)()
    n = 0
    for rep in range(2):
        for t in range(n_types):
            p = prims[n]
            p.type = [A.RT_PRIM_SPHERE, A.RT_PRIM_MOVING_SPHERE, A.RT_PRIM_QUAD][t]
            p.material = 0
            p.a[:] = [0.0, 0.0, 0.0]
            p.b[:] = [0.0, 0.0, 0.0] if t == 1 else [1.0, 0.0, 0.0]
            p.c[:] = [0.0, 1.0, 0.0]
            p.radius = 1.0
            p.time0, p.time1 = 0.0, 1.0
            o = objs[n]
            o.kind = A.RT_OBJ_PRIM
            o.first_prim, o.prim_count = n, 1
            o.bbox[:] = [-1, 1, -1, 1, -1, 1]
            n += 1
    mats = (A.rt_material * 1)()
    mats[0].type = A.RT_MAT_LAMBERTIAN
    mats[0].texture = 0
    tex = (A.rt_texture * 1)()
    tex[0].type = A.RT_TEX_SOLID
    d = A.rt_scene_desc(abi_version=A.RT_ABI_VERSION, n_objects=n, n_prims=n, n_materials=1, n_textures=1,
                        objects=objs, prims=prims, materials=mats, textures=tex)
    d._keep = (prims, objs, mats, tex)
    return d


def test_mixed_leaves_count_their_join_levels(lib):
    """a leaf holding three primitive types becomes two joins above three typed leaves; the stack depth
    reported to the kernel must include them."""
    d1 = _mixed_leaf_scene(1)
    d3 = _mixed_leaf_scene(3)
    rc1, i1 = pack(lib, C.byref(d1), max_leaf=8, flags=A.RT_UPLOAD_NO_HOIST)  # (coincident = all scene-sized)
    rc3, i3 = pack(lib, C.byref(d3), max_leaf=8, flags=A.RT_UPLOAD_NO_HOIST)
    assert rc1 == 0 and rc3 == 0
    assert i3.max_depth_bvh >= i1.max_depth_bvh + 2
    assert i3.n_nodes >= 2 + 2 * 2  # root pair + two joins


def test_validation_rejects_unknown_enums_and_null_tables(lib):
    sc = BuiltinScene(9, np.zeros((4, 4, 3), np.uint8))
    d = sc.desc.contents

    def expect_invalid(mutate, needle):
        bad = A.rt_scene_desc.from_buffer_copy(d)
        keep = mutate(bad)
        rc, _ = pack(lib, C.byref(bad))
        assert rc == A.RT_ERR_INVALID, needle
        assert needle in lib.rt_last_error(), lib.rt_last_error()
        return keep

    def bad_texture(b):
        t = (A.rt_texture * d.n_textures)(*[d.textures[i] for i in range(d.n_textures)])
        t[0].type = 7
        b.textures = t
        return t

    def bad_material(b):
        m = (A.rt_material * d.n_materials)(*[d.materials[i] for i in range(d.n_materials)])
        m[1].type = -1
        b.materials = m
        return m

    def bad_object(b):
        o = (A.rt_object * d.n_objects)(*[d.objects[i] for i in range(d.n_objects)])
        o[0].kind = 3
        b.objects = o
        return o

    def bad_prim(b):
        p = (A.rt_prim * d.n_prims)(*[d.prims[i] for i in range(d.n_prims)])
        p[5].type = 9
        b.prims = p
        return p

    def null_table(b):
        b.materials = None

    def negative_count(b):
        b.n_perlins = -1

    expect_invalid(bad_texture, b"texture type")
    expect_invalid(bad_material, b"material type")
    expect_invalid(bad_object, b"object kind")
    expect_invalid(bad_prim, b"primitive type")
    expect_invalid(null_table, b"NULL table")
    expect_invalid(negative_count, b"negative")
    rc, _ = pack(lib, None)
    assert rc == A.RT_ERR_INVALID


def test_rotated_image_textured_sphere_gets_a_uv_frame(lib, earth):
    """the reference computes sphere (u,v) in OBJECT space (Sphere.h:42-44 inside RotateY::Hit,
    Instance.h:116-150), so baking a RotateY chain into a sphere must keep its yaw for the texture lookup.  The packer
    accepts such a scene; rendering parity with the oracle (which moves the ray into object space like the reference)
    is checked on the GPU in test_parity_gpu.py."""
    for kw in (dict(degrees=70.0, offset=(0.3, 0.0, 0.0)), dict(degrees=70.0, offset=(0.3, 0.0, 0.0), checker=True),
               dict()):
        sc = HandScene(earth, **kw)
        rc, i = pack(lib, sc.desc)
        assert rc == 0, lib.rt_last_error()
        assert i.n_spheres == 1 and i.features & 16  # RT_FEAT_TEXTURE_HEAVY
