"""CPU: the pinned transcendental functions (include/pbr_pinned_math.h) against numpy float64,
through the oracle's scalar probes.  Tolerance: 1 float32 ulp (they are computed in binary64 and
rounded once, so <= 0.5 ulp + double-rounding slack is expected)."""
import numpy as np


def _ulp_err(got, ref64):
    ref32 = ref64.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    ulp = np.maximum(ulp, np.finfo(np.float32).tiny)
    return np.abs(got.astype(np.float64) - ref64) / ulp


def _apply(fn, xs):
    return np.array([fn(float(x)) for x in xs], np.float32)


def test_sin_cos_tan(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-10, 10, 4000), rng.uniform(-3000, 3000, 4000),
                         [0.0, 1e-8, -1e-8, np.pi / 2, np.pi, 0.0333, 1.0333]]).astype(np.float32)
    x64 = xs.astype(np.float64)
    assert _ulp_err(_apply(L.oracle_pm_sin, xs), np.sin(x64)).max() <= 1.0
    assert _ulp_err(_apply(L.oracle_pm_cos, xs), np.cos(x64)).max() <= 1.0
    t = _apply(L.oracle_pm_tan, xs)
    ok = np.abs(np.cos(x64)) > 1e-3
    assert _ulp_err(t[ok], np.tan(x64[ok])).max() <= 1.0
    assert np.isnan(L.oracle_pm_sin(float("inf")))


def test_acos_atan(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(-1, 1, 6000), [-1.0, 1.0, 0.0, 0.999999, -0.999999]]).astype(np.float32)
    assert _ulp_err(_apply(L.oracle_pm_acos, xs), np.arccos(xs.astype(np.float64))).max() <= 1.0
    assert np.isnan(L.oracle_pm_acos(1.5)) and np.isnan(L.oracle_pm_acos(-1.5))
    ys = np.concatenate([rng.uniform(-50, 50, 6000), rng.uniform(-1, 1, 2000), [0.0, 1e30, -1e30]]).astype(np.float32)
    assert _ulp_err(_apply(L.oracle_pm_atan, ys), np.arctan(ys.astype(np.float64))).max() <= 1.0
    assert abs(L.oracle_pm_atan(float("inf")) - np.float32(np.pi / 2)) < 1e-7


def test_pow_cbrt(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(3)
    x = rng.uniform(0.0, 1.0, 6000).astype(np.float32)
    y = np.concatenate([rng.uniform(0, 4, 2000), rng.uniform(0, 2000, 2000), rng.uniform(0, 200000, 2000)]).astype(np.float32)
    got = np.array([L.oracle_pm_pow(float(a), float(b)) for a, b in zip(x, y)], np.float32)
    ref = np.power(x.astype(np.float64), y.astype(np.float64))
    ok = ref > 1e-37
    assert _ulp_err(got[ok], ref[ok]).max() <= 1.0
    assert np.all(got[~ok] < 2e-37)
    # special cases used by the BRDF code
    assert L.oracle_pm_pow(-0.5, 0.0) == 1.0
    assert L.oracle_pm_pow(0.0, 2.0) == 0.0
    assert np.isnan(L.oracle_pm_pow(-0.5, 1.5))
    assert L.oracle_pm_pow(-2.0, 3.0) == -8.0
    assert L.oracle_pm_pow(2.0, 0.5) == np.float32(np.sqrt(2.0))
    c = rng.uniform(-1000, 1000, 3000).astype(np.float32)
    assert _ulp_err(_apply(L.oracle_pm_cbrt, c), np.cbrt(c.astype(np.float64))).max() <= 1.0


def test_rand_stream(oracle):
    """rand() = fract(sin(++seed) * 43758.5453123f) (pt_utils.cl:39-44): values in [0,1) and the
    seed advances by exactly 1."""
    import ctypes
    L = oracle.lib()
    seed = ctypes.c_float(0.0333)
    vals = []
    for i in range(2000):
        before = seed.value
        vals.append(L.oracle_rand(ctypes.byref(seed)))
        assert seed.value == np.float32(np.float32(before) + np.float32(1.0))
    vals = np.array(vals)
    assert vals.min() >= 0.0 and vals.max() < 1.0
    assert 0.4 < vals.mean() < 0.6
