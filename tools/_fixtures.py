"""Development helpers shared by the probe scripts (not product code)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def earth_texels():
    """The earthmap texels as the reference's RtwImage produces them (tests/golden, made by make_golden.py)."""
    return np.ascontiguousarray(np.load(os.path.join(ROOT, "tests", "golden", "earthmap_rgb8.npz"))["rgb"])
