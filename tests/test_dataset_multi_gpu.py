"""Device-resident datasets (crnn_dataset_create / crnn_loss_grad_indexed) and the single-process multi-GPU handle
(crnn_create_multi: contiguous shards + ncclAllReduce of [sum loss, n, grad]) against the plain batched call."""
import numpy as np
import pytest

from crnn_b200 import _abi, cases, synth
from crnn_b200.engine import Engine
from oracle import oracle
from problems import make_problem

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def test_indexed_dataset_call_equals_the_batched_call(engine, golden):
    pb = make_problem("case2", golden, 700)
    ds = engine.dataset(pb["u0"], pb["data"])
    args = (pb["model"], pb["opts"], pb["seed"])
    whole = engine.loss_grad_batch(*args, pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    r = engine.loss_grad_indexed(*args, ds, pb["yscale"], pb["loss_kind"], want_loss=True, want_stats=True)
    assert np.array_equal(r["loss"], whole["loss"]) and np.array_equal(r["grad_sum"], whole["grad_sum"])
    assert np.array_equal(r["stats"]["n_accept"], whole["stats"]["n_accept"]) and r["n_ok"] == 700
    assert r["loss_sum"] == pytest.approx(whole["loss"].sum(), rel=1e-13)
    # a shuffled mini-batch with a random time truncation per picked row (rober_crnn.jl:218)
    g = np.random.default_rng(0)
    idx = g.permutation(700)[:123]
    nsu = g.integers(1, 51, size=idx.size).astype(np.int32)
    part = engine.loss_grad_batch(*args, pb["u0"][idx], pb["data"][idx], pb["yscale"], pb["loss_kind"], n_save_used=nsu)
    ri = engine.loss_grad_indexed(*args, ds, pb["yscale"], pb["loss_kind"], idx=idx, n_save_used=nsu, want_loss=True)
    assert np.array_equal(ri["loss"], part["loss"]) and np.array_equal(ri["n_saved"], nsu)
    assert np.array_equal(ri["grad_sum"], part["grad_sum"])
    ref = oracle.loss_grad_batch(*args, pb["u0"][idx], pb["data"][idx], pb["yscale"], pb["loss_kind"], n_save_used=nsu)
    np.testing.assert_allclose(ri["loss"], ref["loss"], rtol=1e-10)
    np.testing.assert_allclose(ri["grad_sum"], ref["grad_sum"], rtol=1e-8, atol=1e-10 * np.abs(ref["grad_sum"]).max())
    # empty pick, bad index, dataset / handle mismatch
    e0 = engine.loss_grad_indexed(*args, ds, pb["yscale"], pb["loss_kind"], idx=np.zeros(0, dtype=np.int64))
    assert e0["n_ok"] == 0 and np.all(e0["grad_sum"] == 0)
    with pytest.raises(Exception):
        engine.loss_grad_indexed(*args, ds, pb["yscale"], pb["loss_kind"], idx=np.array([700]))
    ds.close()


@pytest.mark.parametrize("mode", ["adjoint", "rosenbrock"])
def test_indexed_call_on_the_other_gradient_kernels(engine, golden, mode):
    if mode == "adjoint":
        pb = make_problem("case2", golden, 200)
        o = pb["case"].opts(obs_idx=np.arange(6), sens_mode=_abi.SENS_DISCRETE_ADJOINT)
    else:
        pb = make_problem("robertson", golden, 200)
        o = pb["opts"]
    ds = engine.dataset(pb["u0"], pb["data"])
    idx = np.random.default_rng(1).permutation(200)[:77]
    a = engine.loss_grad_batch(pb["model"], o, pb["seed"], pb["u0"][idx], pb["data"][idx], pb["yscale"], pb["loss_kind"])
    b = engine.loss_grad_indexed(pb["model"], o, pb["seed"], ds, pb["yscale"], pb["loss_kind"], idx=idx, want_loss=True)
    assert np.array_equal(a["loss"], b["loss"]) and np.array_equal(a["grad_sum"], b["grad_sum"])


def test_multi_device_handle_two_gpu_gradient_equals_one_gpu(engine, golden):
    """SURVEY §7.5 "multi-GPU" row: N-GPU grad == 1-GPU grad up to summation order (1e-12 relative)."""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    pb = make_problem("case2", golden, 4099)           # uneven split
    args = (pb["model"], pb["opts"], pb["seed"])
    one = engine.loss_grad_batch(*args, pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    for nd in sorted({2, min(_n_gpus(), 8)}):
        multi = Engine(devices=nd)
        assert multi.n_devices == nd
        ds = multi.dataset(pb["u0"], pb["data"])
        r = multi.loss_grad_indexed(*args, ds, pb["yscale"], pb["loss_kind"], want_loss=True, want_stats=True)
        assert np.array_equal(r["loss"], one["loss"])                       # per-trajectory results are bit-identical
        assert np.array_equal(r["stats"]["n_accept"], one["stats"]["n_accept"])
        gmax = np.abs(one["grad_sum"]).max()
        assert np.abs(r["grad_sum"] - one["grad_sum"]).max() <= 1e-12 * gmax
        assert r["loss_sum"] == pytest.approx(one["loss"].sum(), rel=1e-12) and r["n_ok"] == 4099
        # shuffled pick across the shards, with truncation
        g = np.random.default_rng(2)
        idx = g.permutation(4099)[:1000]; nsu = g.integers(1, 51, size=1000).astype(np.int32)
        a = engine.loss_grad_batch(*args, pb["u0"][idx], pb["data"][idx], pb["yscale"], pb["loss_kind"], n_save_used=nsu)
        b = multi.loss_grad_indexed(*args, ds, pb["yscale"], pb["loss_kind"], idx=idx, n_save_used=nsu, want_loss=True)
        assert np.array_equal(a["loss"], b["loss"]) and np.array_equal(b["n_saved"], nsu)
        assert np.abs(a["grad_sum"] - b["grad_sum"]).max() <= 1e-12 * np.abs(a["grad_sum"]).max()
        # host-buffer entry points on the multi-device handle
        hb = multi.loss_grad_batch(*args, pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
        assert np.array_equal(hb["loss"], one["loss"])
        assert np.abs(hb["grad_sum"] - one["grad_sum"]).max() <= 1e-12 * gmax
        sv = multi.solve_batch(pb["model"], pb["opts"], pb["u0"])
        sv1 = engine.solve_batch(pb["model"], pb["opts"], pb["u0"])
        assert np.array_equal(sv["pred"], sv1["pred"])
        ds.close(); multi.close()


def test_multi_device_handle_with_one_device_needs_no_nccl(engine, golden):
    pb = make_problem("case2", golden, 64)
    m1 = Engine(devices=[0])
    ds = m1.dataset(pb["u0"], pb["data"])
    a = m1.loss_grad_indexed(pb["model"], pb["opts"], pb["seed"], ds, pb["yscale"], pb["loss_kind"], want_loss=True)
    b = engine.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"], pb["loss_kind"])
    assert np.array_equal(a["loss"], b["loss"]) and np.array_equal(a["grad_sum"], b["grad_sum"])
    ds.close(); m1.close()
