"""CPU: the oracle's deterministic expf / acos / acosf / sin against libm (<= 1 ulp), so that the
oracle's scores stay within the 1e-4 budget of a glibc build of the reference."""
import math

import numpy as np


def ulp_diff32(a, b):
    ia = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    ib = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def test_expf_within_one_ulp_of_libm(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-30, 0, 20000), rng.uniform(-1, 1, 5000), rng.uniform(-87, 88, 5000)]).astype(np.float32)
    got = np.array([L.orc_kat_expf(float(v)) for v in x], dtype=np.float32)
    ref = np.exp(x.astype(np.float64)).astype(np.float32)
    d = ulp_diff32(got, ref)
    assert d.max() <= 1, d.max()
    assert (d == 0).mean() > 0.999          # correctly rounded almost everywhere
    assert L.orc_kat_expf(0.0) == 1.0
    assert L.orc_kat_expf(-200.0) == 0.0
    assert math.isinf(L.orc_kat_expf(100.0))


def test_expf_monotone_at_the_truncation_threshold(oracle):
    """The early exits of the CUDA kernels rely on exp(x) <= 0.5 for x < -0.70."""
    L = oracle.lib()
    x = np.linspace(-0.72, -0.69, 20001).astype(np.float32)
    y = np.array([L.orc_kat_expf(float(v)) for v in x], dtype=np.float32)
    assert (np.diff(y) >= 0).all()
    assert (y[x < -0.70] < 0.4966).all()


def test_acos_and_sin_accuracy(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-1, 1, 20000), [1.0, -1.0, 0.0, 0.5, -0.5, 0.4999999, 0.5000001]])
    got = np.array([L.orc_kat_acos(float(v)) for v in x])
    ref = np.arccos(x)
    assert np.max(np.abs(got - ref)) < 4e-16 * math.pi
    xf = x.astype(np.float32)
    gotf = np.array([L.orc_kat_acosf(float(v)) for v in xf], dtype=np.float32)
    reff = np.arccos(xf.astype(np.float64)).astype(np.float32)
    assert ulp_diff32(gotf, reff).max() <= 1
    a = rng.uniform(0, math.pi, 20000)
    s = np.array([L.orc_kat_sin(float(v)) for v in a])
    assert np.max(np.abs(s - np.sin(a))) < 3e-16
