"""crnn_b200/csrc/lean_math.h, host instantiation (oracle/lean_math_host.c): accuracy against mpmath.  The same header
compiles into the CUDA kernels; tests/test_lean_math_gpu.py shows the device copy returns the same bits."""
import mpmath as mp
import numpy as np

from oracle import oracle


def _ulps(got, exact):
    return float(abs(mp.mpf(got) - exact) / mp.mpf(float(np.spacing(abs(float(exact))))))


def sample_log(n=4000, seed=0):
    g = np.random.default_rng(seed)
    return np.concatenate([10.0 ** g.uniform(-300, 300, n), g.uniform(0.5, 2.0, n), 1.0 + g.uniform(-1e-3, 1e-3, n // 2),
                           [1e-8, 1e-6, 1e-5, 10.0, 100.0, np.sqrt(0.5), np.sqrt(2.0), 2.0, 0.5]])


def sample_exp(n=4000, seed=1):
    g = np.random.default_rng(seed)
    return np.concatenate([g.uniform(-700, 700, n), g.uniform(-1, 1, n), [0.0, -0.0, 1.0, -1.0, 30.0, -30.0]])


def test_lean_log_exp_are_one_ulp_functions():
    mp.mp.prec = 200
    L = oracle.lib()
    wl = max(_ulps(L.crnn_lean_log(float(x)), mp.log(mp.mpf(float(x)))) for x in sample_log() if x != 1.0)
    we = max(_ulps(L.crnn_lean_exp(float(x)), mp.exp(mp.mpf(float(x)))) for x in sample_exp())
    assert wl < 1.0 and we < 1.0, (wl, we)
    assert L.crnn_lean_log(1.0) == 0.0 and L.crnn_lean_exp(0.0) == 1.0


def test_out_of_range_arguments_take_the_library_path():
    L = oracle.lib()
    assert L.crnn_lean_log(0.0) == -np.inf and np.isnan(L.crnn_lean_log(-1.0)) and L.crnn_lean_log(np.inf) == np.inf
    assert L.crnn_lean_log(5e-324) == np.log(5e-324)
    assert L.crnn_lean_exp(800.0) == np.inf and L.crnn_lean_exp(-800.0) == 0.0 and np.isnan(L.crnn_lean_exp(np.nan))


def test_pow_and_base10_helpers():
    L = oracle.lib()
    g = np.random.default_rng(2)
    for x, y in zip(10.0 ** g.uniform(-8, 2, 500), g.uniform(0.05, 0.8, 500)):   # controller exponents on EEst
        assert abs(L.crnn_lean_pow(float(x), float(y)) / float(x) ** float(y) - 1.0) < 4e-15
    for x in 10.0 ** g.uniform(-12, 6, 500):
        assert abs(L.crnn_lean_log10(float(x)) - np.log10(x)) < 4e-15 * max(1.0, abs(np.log10(x)))
        assert abs(L.crnn_lean_exp10(float(np.log10(x))) / x - 1.0) < 1e-14


def test_shared_math_switch_changes_rounding_only(golden=None):
    import json, os
    from problems import make_problem
    with open(os.path.join(os.path.dirname(__file__), "golden", "checkpoints.json")) as f:
        golden = json.load(f)
    pb = make_problem("case2", golden, 16)
    a = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
    assert oracle.lib().crnn_oracle_get_shared_math() == 0          # default: the C library
    with oracle.shared_math():
        b = oracle.solve_batch(pb["model"], pb["opts"], pb["u0"])
    assert np.array_equal(a["stats"]["n_accept"], b["stats"]["n_accept"])
    assert not np.array_equal(a["pred"], b["pred"])
    np.testing.assert_allclose(b["pred"], a["pred"], rtol=1e-9, atol=1e-13)
