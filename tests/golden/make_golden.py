#!/usr/bin/env python3
"""Generates the committed fixtures from the REFERENCE's own code.

Needs oracle/_ref/libref_stream.so, i.e. /root/reference and
`python oracle/build_ref.py stream`.  Writes next to this file:

  earthmap_rgb8.npz      the reference's earthmap.jpg after its image load path
                         (stb_image stbi_loadf + RtwImage::FloatToByte,
                         reference RtwImage.h:51-105): the texels ImageTexture sees.
  ref_stream_s<id>.npz   for every scene id 0..10: a 48x27, 2-spp, depth-50 render
                         by the reference's classes on the counter-based stream
                         (float64 linear sums, row 0 = bottom), its ray count, the
                         scene-stream draw count and the sorted top-level boxes.

The oracle must reproduce every render bit for bit (tests/test_oracle_pin.py).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import bindings as O  # noqa: E402

W, H, SPP, DEPTH, SEED = 48, 27, 2, 50, 1984


def main():
    ref = O.load_ref_stream()
    jpg = os.path.join(os.environ.get("RT_REFERENCE_DIR", "/root/reference"), "earthmap.jpg").encode()
    w, h = C.c_int(), C.c_int()
    assert ref.ref_load_image_rgb8(jpg, C.byref(w), C.byref(h), None, 0) == 0
    earth = np.zeros((h.value, w.value, 3), np.uint8)
    assert ref.ref_load_image_rgb8(jpg, C.byref(w), C.byref(h), earth.ctypes.data, earth.size) == 0
    np.savez_compressed(os.path.join(HERE, "earthmap_rgb8.npz"), rgb=earth)
    for sid in range(11):
        out = np.zeros((H, W, 3), np.float64)
        st = O.ref_stream_stats()
        ref.ref_stream_render(sid, W, H, 0, SPP, DEPTH, SEED, earth.ctypes.data, w.value, h.value, 4,
                              out.ctypes.data, C.byref(st))
        boxes = np.zeros((4096, 6), np.float64)
        draws = C.c_ulonglong()
        n = ref.ref_stream_scene_boxes(sid, W, H, earth.ctypes.data, w.value, h.value, boxes.ctypes.data, 4096,
                                       C.byref(draws))
        b = boxes[:n]
        b = b[np.lexsort(b.T[::-1])]
        np.savez_compressed(os.path.join(HERE, f"ref_stream_s{sid}.npz"), image=out, rays=np.uint64(st.rays),
                            scene_draws=np.uint64(st.scene_draws), n_objects=np.int32(st.n_objects),
                            n_nodes=np.int32(st.n_nodes), boxes=b,
                            params=np.array([W, H, SPP, DEPTH, SEED], np.int64))
        print(f"scene {sid}: rays {st.rays} objects {st.n_objects} nodes {st.n_nodes} mean {out.mean():.6f}")


if __name__ == "__main__":
    main()
