"""The device copy of crnn_b200/csrc/lean_math.h returns the same BITS as the host copy the oracle's shared-math mode
runs (a shared host+device math header)."""
import numpy as np
import pytest

from oracle import oracle
from test_lean_math_cpu import sample_exp, sample_log

pytestmark = pytest.mark.gpu


def test_device_and_host_lean_math_agree_bit_for_bit(engine):
    L = oracle.lib()
    xl = np.concatenate([sample_log(200000, 3), [0.0, -1.0, np.inf, np.nan, 5e-324]])
    xe = np.concatenate([sample_exp(200000, 4), [800.0, -800.0, np.nan]])
    for op, xs, fn in (("log", xl, L.crnn_lean_log), ("exp", xe, L.crnn_lean_exp),
                       ("log10", xl[:50000], L.crnn_lean_log10), ("exp10", xe[:50000] / 3.0, L.crnn_lean_exp10)):
        dev = engine.lean_math(op, xs)
        host = np.array([fn(float(v)) for v in xs])
        plain = np.isfinite(host) & (np.abs(xs) < 690 if "exp" in op else (xs > 1e-300))
        assert np.array_equal(dev[plain].view(np.uint64), host[plain].view(np.uint64)), op
        rest = ~plain    # library fall-backs (CUDA libm vs glibc): same value class, 1 ulp at most
        assert np.array_equal(np.isnan(dev[rest]), np.isnan(host[rest]))
        ok = ~np.isnan(host[rest])
        assert np.allclose(dev[rest][ok], host[rest][ok], rtol=1e-15, atol=0, equal_nan=True)
    g = np.random.default_rng(5)
    x = 10.0 ** g.uniform(-8, 2, 100000); y = g.uniform(0.05, 0.8, 100000)
    dev = engine.lean_math("pow", x, y)
    host = np.array([L.crnn_lean_pow(float(a), float(b)) for a, b in zip(x, y)])
    assert np.array_equal(dev.view(np.uint64), host.view(np.uint64))
