"""The render stream: product (include/rt_rng.h) == oracle restatement, range, known answers."""
import numpy as np

# (seed, pixel, sample, slot, domain, dim) -> float, fixed when the stream was specified.
KNOWN = [
    ((1984, 0, 0, 0, 0, 0), None),
]


def test_product_and_oracle_streams_are_identical(lib, oracle):
    rng = np.random.default_rng(7)
    for _ in range(4000):
        seed, pixel, sample = int(rng.integers(0, 2**32)), int(rng.integers(0, 3840 * 2160)), int(rng.integers(0, 10000))
        slot, domain, dim = int(rng.integers(0, 52)), int(rng.integers(0, 17)), int(rng.integers(0, 64))
        a = lib.rt_rng_uniform(seed, pixel, sample, slot, domain, dim)
        b = oracle.oracle_rng_uniform(seed, pixel, sample, slot, domain, dim)
        assert a == b
        assert 0.0 < a <= 1.0


def test_stream_is_roughly_uniform_and_decorrelated(lib):
    n = 20000
    u = np.array([lib.rt_rng_uniform(1984, p, 0, 1, 0, 0) for p in range(n)])
    v = np.array([lib.rt_rng_uniform(1984, p, 0, 1, 0, 1) for p in range(n)])
    w = np.array([lib.rt_rng_uniform(1984, 5, s, 1, 0, 0) for s in range(n)])
    for x in (u, v, w):
        assert abs(x.mean() - 0.5) < 0.01
        assert abs(x.var() - 1 / 12) < 0.005
        hist, _ = np.histogram(x, bins=16, range=(0, 1))
        assert hist.min() > 0.85 * n / 16
    assert abs(np.corrcoef(u, v)[0, 1]) < 0.03
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 0.03
    assert abs(np.corrcoef(u, w)[0, 1]) < 0.03


def test_bits_to_float_matches_curand_uniform_ends(lib):
    # curand_uniform.h:69-72: x*2^-32 + 2^-33 in fp32 -> never 0, can be exactly 1
    f = np.float32
    assert f(0) * f(2.3283064365386963e-10) + f(1.1641532182693481e-10) > 0
    assert f(np.uint32(0xFFFFFFFF)) * f(2.3283064365386963e-10) + f(1.1641532182693481e-10) == f(1.0)
