"""Host scene surface: the flat description the C ABI consumes."""
import numpy as np

from conftest import A
from raytracinginoneweekendincuda_b200 import BuiltinScene


def test_book1_final_scene_shape():
    sc = BuiltinScene(10)
    d = sc.desc.contents
    assert d.n_objects == d.n_prims == 485  # 22*22 grid minus the skipped ones + ground + 3 big
    kinds = {d.objects[i].kind for i in range(d.n_objects)}
    assert kinds == {A.RT_OBJ_PRIM}
    types = [d.prims[i].type for i in range(d.n_prims)]
    assert set(types) == {A.RT_PRIM_SPHERE}
    mats = [d.materials[d.prims[i].material].type for i in range(d.n_prims)]
    assert mats.count(A.RT_MAT_DIELECTRIC) > 5 and mats.count(A.RT_MAT_METAL) > 30
    cam = sc.camera(1200, 675, 10)
    assert cam.vfov == 30.0 and abs(cam.aperture - 0.1) < 1e-15 and cam.focus_dist == 10.0
    assert abs(cam.defocus_angle - np.degrees(2 * np.arctan(0.005))) < 1e-12
    assert cam.time0 == 0.0 and cam.time1 == 0.0


def test_scene0_shares_geometry_with_book1():
    sa, sb = BuiltinScene(0), BuiltinScene(10)
    a, b = sa.desc.contents, sb.desc.contents
    assert a.n_prims == b.n_prims
    moving = 0
    for i in range(a.n_prims):
        assert list(a.prims[i].a) == list(b.prims[i].a)
        moving += a.prims[i].type == A.RT_PRIM_MOVING_SPHERE
    assert moving > 300
    assert a.textures[a.materials[a.prims[0].material].texture].type == A.RT_TEX_CHECKER


def test_cornell_boxes_are_instanced_lists():
    sc = BuiltinScene(7)
    d = sc.desc.contents
    assert d.n_objects == 8 and d.n_prims == 18
    box = d.objects[6]
    assert box.kind == A.RT_OBJ_LIST and box.prim_count == 6
    p = d.prims[box.first_prim]
    assert p.xform_count == 2
    assert d.xforms[p.first_xform].type == A.RT_XFORM_TRANSLATE  # outermost first
    rot = d.xforms[p.first_xform + 1]
    assert rot.type == A.RT_XFORM_ROTATE_Y and rot.v[2] == 15.0
    assert abs(rot.v[0] - np.sin(np.radians(15.0))) < 1e-15


def test_final_scene_inventory(earth):
    sc = BuiltinScene(9, earth)
    d = sc.desc.contents
    assert d.n_objects == 410  # 400 boxes + 10 others (reference kernel.cu:626)
    assert d.n_prims == 400 * 6 + 1 + 1 + 2 + 1 + 1 + 1 + 1 + 1 + 1000
    media = [d.objects[i] for i in range(d.n_objects) if d.objects[i].kind == A.RT_OBJ_MEDIUM]
    assert [m.medium_id for m in media] == [0, 1]
    assert [m.density for m in media] == [0.2, 0.0001]
    assert d.n_images == 1 and d.images[0].width == 1024 and d.images[0].height == 512
    assert d.n_perlins == 1
    cluster = d.objects[d.n_objects - 1]
    assert cluster.kind == A.RT_OBJ_LIST and cluster.prim_count == 1000


def test_image_linearize_lut(lib):
    src = np.arange(256, dtype=np.uint8)
    out = np.zeros(256, np.uint8)
    lib.rt_image_linearize_rgb8(src.ctypes.data, out.ctypes.data, 256)
    assert out[0] == 0 and out[255] == 255
    assert np.all(np.diff(out.astype(int)) >= 0)
    assert out[128] == int(256 * np.float32((128 / 255) ** 2.2))
