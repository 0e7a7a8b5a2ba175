"""Image-texture pipeline of the host library: JPEG -> RtwImage texels, in the product.

The reference reads earthmap.jpg through stb_image's stbi_loadf and re-quantises (RtwImage.h:48-105); nearest-texel
lookups make the decoded bytes part of the parity contract (trap T8).  rt_image_load / rt_image_decode_jpeg restate
that decoder's arithmetic; here they are pinned against the texels the reference's own path produced (committed
golden) and, where the reference-derived checker is built, against stb_image itself on other kinds of JPEG."""
import ctypes as C
import io
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, A, O



def _asset():
    """The reference's texture asset: beside the reference-derived checker (oracle/_ref, which travels to the GPU
    box), or in the reference tree itself."""
    for path in (os.path.join(ROOT, "oracle", "_ref", "earthmap.jpg"),
                 os.path.join(os.environ.get("RT_REFERENCE_DIR", "/root/reference"), "RayTracinginOneWeekend", "earthmap.jpg")):
        if os.path.exists(path):
            return path
    pytest.skip("earthmap.jpg not available (needs /root/reference or a built oracle/_ref)")


def decode(lib, data: bytes, linearize=1):
    w, h = C.c_int32(), C.c_int32()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    rc = lib.rt_image_decode_jpeg(buf, len(data), C.byref(w), C.byref(h), None, 0, linearize)
    if rc != 0:
        return rc, None
    out = np.zeros((h.value, w.value, 3), np.uint8)
    rc = lib.rt_image_decode_jpeg(buf, len(data), C.byref(w), C.byref(h), out.ctypes.data, out.size, linearize)
    return rc, out


def test_earthmap_texels_equal_the_reference_pipeline_byte_for_byte(lib):
    """The shipped asset, decoded + linearised by the product, equals tests/golden/earthmap_rgb8.npz (made by the
    reference's RtwImage over its vendored stb_image, tests/golden/make_golden.py)."""
    ASSET = _asset()
    want = np.load(os.path.join(GOLDEN, "earthmap_rgb8.npz"))["rgb"]
    w, h = C.c_int32(), C.c_int32()
    assert lib.rt_image_load(ASSET.encode(), C.byref(w), C.byref(h), None, 0) == 0
    assert (w.value, h.value) == (1024, 512)
    got = np.zeros((h.value, w.value, 3), np.uint8)
    assert lib.rt_image_load(ASSET.encode(), C.byref(w), C.byref(h), got.ctypes.data, got.size) == 0
    assert np.array_equal(got, want)
    # and through the Python surface: a path given to BuiltinScene is decoded by the host library
    from raytracinginoneweekendincuda_b200 import BuiltinScene, load_image
    assert np.array_equal(load_image(ASSET), want)
    sc = BuiltinScene(9, ASSET)
    im = sc.desc.contents.images[0]
    assert (im.width, im.height) == (1024, 512)
    assert np.array_equal(np.ctypeslib.as_array(im.rgb, shape=(512, 1024, 3)), want)


def test_decode_errors(lib, tmp_path):
    w, h = C.c_int32(), C.c_int32()
    assert lib.rt_image_load(str(tmp_path / "missing.jpg").encode(), C.byref(w), C.byref(h), None, 0) == A.RT_ERR_INVALID
    rc, _ = decode(lib, b"P6\n2 2\n255\n" + bytes(12))
    assert rc == A.RT_ERR_INVALID and b"SOI" in lib.rt_last_error()
    data = open(_asset(), "rb").read()
    rc, _ = decode(lib, data[:4000])  # truncated inside the scan: the rows decoded so far are kept, like stb does
    assert rc in (0, A.RT_ERR_INVALID)
    small = (C.c_uint8 * 16)()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    assert lib.rt_image_decode_jpeg(buf, len(data), C.byref(w), C.byref(h), small, 16, 1) == A.RT_ERR_INVALID
    assert lib.rt_image_decode_jpeg(None, 0, C.byref(w), C.byref(h), None, 0, 1) == A.RT_ERR_INVALID


def _pil():
    try:
        from PIL import Image
        return Image
    except Exception:  # noqa: BLE001
        pytest.skip("Pillow not available")


def _test_picture(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 255 // max(1, w - 1)), (y * 255 // max(1, h - 1)), ((x * y) % 256)], axis=2).astype(np.float64)
    img += rng.normal(0, 25, img.shape)
    img[h // 3:h // 2, w // 4:w // 2] = (250, 10, 200)  # a hard chroma edge: exercises the tent filters
    return np.clip(img, 0, 255).astype(np.uint8)


def test_progressive_jpeg_is_refused_not_misread(lib):
    Image = _pil()
    bio = io.BytesIO()
    Image.fromarray(_test_picture(40, 30, 1)).save(bio, "JPEG", progressive=True)
    rc, _ = decode(lib, bio.getvalue())
    assert rc == A.RT_ERR_UNSUPPORTED and b"progressive" in lib.rt_last_error()


@pytest.mark.parametrize("w,h,kw", [
    (64, 48, dict(subsampling=0, quality=90)),            # 4:4:4
    (64, 48, dict(subsampling=1, quality=85)),            # 4:2:2 -> "h_2"
    (64, 48, dict(subsampling=2, quality=75)),            # 4:2:0 -> "hv_2"
    (37, 23, dict(subsampling=2, quality=95)),            # not a multiple of the MCU
    (17, 1, dict(subsampling=2, quality=80)),             # one row
    (1, 9, dict(subsampling=1, quality=80)),              # one column (w == 1 branches of the filters)
    (100, 60, dict(subsampling=2, quality=50, restart_marker_blocks=3)),  # DRI + RSTn
    (50, 50, dict(subsampling=0, quality=100, optimize=True)),            # optimised Huffman tables, 16 steps of quantiser 1
    (48, 32, dict(grey=True, quality=88)),                # one component
])
def test_decoder_equals_stb_image_on_other_jpeg_kinds(lib, tmp_path, w, h, kw):
    """Byte equality with the reference's decoder (its vendored stb_image, reached through
    oracle/_ref/libref_stream.so: ref_load_image_rgb8 = RtwImage::Load) on subsampled, odd-sized, restart-interval,
    optimised-table and greyscale files made with Pillow."""
    Image = _pil()
    ref = O.load_ref_stream()
    if ref is None:
        pytest.skip("oracle/_ref/libref_stream.so not built (needs /root/reference)")
    kw = dict(kw)
    pic = _test_picture(w, h, w * 1000 + h)
    im = Image.fromarray(pic[:, :, 0] if kw.pop("grey", False) else pic)
    path = tmp_path / "t.jpg"
    try:
        im.save(path, "JPEG", **kw)
    except TypeError:
        pytest.skip("this Pillow lacks an encoder option used here")
    rw, rh = C.c_int(), C.c_int()
    assert ref.ref_load_image_rgb8(str(path).encode(), C.byref(rw), C.byref(rh), None, 0) == 0
    want = np.zeros((rh.value, rw.value, 3), np.uint8)
    assert ref.ref_load_image_rgb8(str(path).encode(), C.byref(rw), C.byref(rh), want.ctypes.data, want.size) == 0
    rc, got = decode(lib, path.read_bytes())
    assert rc == 0, lib.rt_last_error()
    assert got.shape == want.shape == (h, w, 3)
    assert np.array_equal(got, want), f"{(got != want).sum()} of {got.size} bytes differ"
