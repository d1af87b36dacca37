"""End-to-end: the reference's case2 training loop (case2/case2.jl:62-88,192-207) running on the engine — its data
generation (true mechanism + 5 % multiplicative noise, 20 training + 10 validation experiments), its parameter
initialisation, its optimiser chain, one optimiser step per experiment.  The reference's committed loss curve
(case2/figs/loss.png, checkpoint loss history) falls from ~0.14 to ~0.014 over 3 700 epochs; a short run here must
show the same descent and move the Arrhenius parameters towards the generating mechanism."""
import numpy as np
import pytest

from crnn_b200 import _abi, cases, optim, synth
from crnn_b200.frontend import CRNNProblem

pytestmark = pytest.mark.gpu


def test_case2_training_loop_descends_like_the_reference(engine, golden):
    c = cases.CASES["case2"]
    n_exp_train, n_exp_val = 20, 10
    n_exp = n_exp_train + n_exp_val
    u0 = synth.make_u0("case2", n_exp, seed=1234)                      # TG, ROH ~ U(0.2, 2.2), T ~ U(323, 343) (:62-65)
    truth = engine.solve_batch(cases.true_model_case2(), c.opts(obs_idx=np.arange(c.ns), pred_clamp=(-np.inf, np.inf)), u0)
    assert (truth["retcode"] == 1).all()
    data = synth.noisy_targets(truth["pred"], 0.05)                    # ode_data += randn .* ode_data .* noise (:79)
    yscale = synth.yscale_from(data, c.lb)                             # (:81-83)
    prob = CRNNProblem("case2", u0, data, yscale, engine=engine)
    g = np.random.default_rng(1234)
    ns, nr = c.ns, c.nr
    p = g.standard_normal(c.n_p) * 0.1                                 # (:85-88)
    p[:nr] += 0.8
    p[nr * (ns + 1):nr * (ns + 2)] += 0.8
    p[-1] = 0.1
    opt = optim.Optimiser(optim.ExpDecay(5e-3, 0.5, 500 * n_exp_train, 1e-4), *optim.ADAMW(0.005, (0.9, 0.999), 1e-6).chain)
    # NB: as in the script, ExpDecay sits BEFORE ADAM and only rescales a gradient ADAM then normalises
    p_end, hist = prob.train(p, opt, n_epoch=60, n_exp_train=n_exp_train, batch=1, rng=g)
    l0_train, l0_val = hist[0][0], hist[0][1]
    l_train, l_val = hist[-1][0], hist[-1][1]
    assert np.isfinite([h[0] for h in hist]).all()
    assert l0_train > 0.08                                             # starts where the reference starts (~0.14)
    assert l_train < 0.5 * l0_train and l_val < 0.5 * l0_val, (l0_train, l_train, l0_val, l_val)
    best = min(h[0] for h in hist)
    assert l_train < 1.5 * best                                        # no blow-up at the end
    # one mini-batch step with the whole training set gives the same descent direction as the mean of the per-experiment ones
    loss_b, grad_b = prob.loss_grad(p_end, np.arange(n_exp_train))
    grads = [prob.loss_grad(p_end, i)[1] for i in range(n_exp_train)]
    np.testing.assert_allclose(grad_b, np.mean(grads, axis=0), rtol=1e-9, atol=1e-12)


def _case2_problem(engine, n_exp=30, seed=1234):
    c = cases.CASES["case2"]
    u0 = synth.make_u0("case2", n_exp, seed=seed)
    truth = engine.solve_batch(cases.true_model_case2(), c.opts(obs_idx=np.arange(c.ns), pred_clamp=(-np.inf, np.inf)), u0)
    data = synth.noisy_targets(truth["pred"], 0.05)
    prob = CRNNProblem("case2", u0, data, synth.yscale_from(data, c.lb), engine=engine)
    g = np.random.default_rng(seed)
    p = g.standard_normal(c.n_p) * 0.1
    p[:c.nr] += 0.8
    p[c.nr * (c.ns + 1):c.nr * (c.ns + 2)] += 0.8
    p[-1] = 0.1
    return prob, p, g


@pytest.mark.parametrize("variant", ["expdecay_adamw_batch1", "nadam_clip_batch4"])
def test_on_device_training_loop_equals_the_host_loop(engine, variant):
    """crnn_train_steps (p2vec kernel -> sensitivity kernel -> reduction -> Flux chain, nothing returning to the host
    between steps) against the same steps taken one C call at a time with the host mirror of Flux's optimisers
    (case2/case2.jl:31-32,192-198; rober_crnn.jl:220-223 for the clip; case3.jl:20 for NADAM)."""
    prob, p0, g = _case2_problem(engine)
    n_train = 20
    if variant == "expdecay_adamw_batch1":
        batch, n_steps, grad_max = 1, 60, None
        opt = optim.Optimiser(optim.ExpDecay(5e-3, 0.5, 25, 1e-4), optim.ADAMW(0.005, (0.9, 0.999), 1e-6))
        kw = dict(optimiser="adam", eta=0.005, beta=(0.9, 0.999), weight_decay=1e-6, expdecay=(5e-3, 0.5, 25, 1e-4))
    else:
        batch, n_steps, grad_max = 4, 15, 0.05
        opt = optim.Optimiser(optim.NADAM(0.002, (0.9, 0.999)))
        kw = dict(optimiser="nadam", eta=0.002, beta=(0.9, 0.999), grad_max=grad_max)
    order = np.concatenate([g.permutation(n_train) for _ in range(-(-n_steps * batch // n_train))])[:n_steps * batch]
    model, _ = prob.case.model(p0)
    adam = opt.chain[-1].chain[0] if variant == "expdecay_adamw_batch1" else opt.chain[0]
    # (1) every device step against the host mirror taking the same step FROM THE SAME p AND STATE (a free-running
    # comparison is not meaningful beyond ~20 steps: the adaptive solver's accept / reject decisions make loss(p)
    # piecewise smooth, a last-bit difference in p becomes 1e-8 in the gradient at a flipped decision and ADAM amplifies it)
    p, st = p0.copy(), None
    gn_seen = []
    for s in range(n_steps):
        idx = order[s * batch:(s + 1) * batch]
        loss, grad = prob.loss_grad(p, idx)
        grad, gn = optim.clip_by_norm(grad, grad_max) if grad_max is not None else (grad, float(np.linalg.norm(grad)))
        gn_seen.append(gn)
        p_host = p.copy()
        opt.update(p_host, grad)
        r = engine.train_steps(model, prob.opts, prob.dataset, idx, prob.yscale, p, st, prob.case.loss_kind, batch=batch, **kw)
        np.testing.assert_allclose(r["step_loss"][0], loss, rtol=1e-13)
        np.testing.assert_allclose(r["step_gnorm"][0], gn, rtol=1e-13)
        np.testing.assert_allclose(r["p"], p_host, rtol=1e-12, atol=1e-15, err_msg=f"step {s}")
        p, st = r["p"], r["opt_state"]
        n_p = p.size                                          # the host mirror continues from the device's state
        np.testing.assert_allclose(st[:n_p], adam.m, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(st[n_p:2 * n_p], adam.v, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(st[2 * n_p:2 * n_p + 2], adam.bp, rtol=1e-13)
        adam.m, adam.v = st[:n_p].copy(), st[n_p:2 * n_p].copy()
        if variant == "expdecay_adamw_batch1":
            assert st[2 * n_p + 2] == opt.chain[0].eta and st[2 * n_p + 3] == opt.chain[0].count
    assert not np.allclose(p, p0, rtol=1e-3)
    if variant == "expdecay_adamw_batch1":
        assert opt.chain[0].eta == 5e-3 * 0.25               # two decays happened
    else:
        assert max(gn_seen) > grad_max                       # the clip was active
    # (2) all steps in ONE call (and in two calls carrying the state) = the step-by-step run, bit for bit
    r_all = engine.train_steps(model, prob.opts, prob.dataset, order, prob.yscale, p0, None, prob.case.loss_kind, batch=batch, **kw)
    h = (n_steps // 2) * batch
    r1 = engine.train_steps(model, prob.opts, prob.dataset, order[:h], prob.yscale, p0, None, prob.case.loss_kind, batch=batch, **kw)
    r2 = engine.train_steps(model, prob.opts, prob.dataset, order[h:], prob.yscale, r1["p"], r1["opt_state"], prob.case.loss_kind,
                            batch=batch, **kw)
    assert np.array_equal(r_all["p"], p) and np.array_equal(r2["p"], p) and np.array_equal(r_all["opt_state"], st)
    assert np.array_equal(np.concatenate([r1["step_loss"], r2["step_loss"]]), r_all["step_loss"])


def test_on_device_loop_case1_p2vec(engine, golden):
    """p2vec_kind = 1: case1/case1.jl:70-78 on the device (w_b = p + b0, w_in = clamp(-w_out, 0, 2.5)); every step against the host
    mirror from the same (p, state), ADAMW(0.001, (0.9, 0.999), 1e-8) as in case1.jl:18"""
    from problems import make_problem
    pb = make_problem("case1", golden, 12)
    prob = CRNNProblem("case1", pb["u0"], pb["data"], pb["yscale"], engine=engine)
    g = np.random.default_rng(3)
    p = 0.1 * g.standard_normal(24)                                        # case1.jl:86
    opt = optim.ADAMW(0.001, (0.9, 0.999), 1e-8)
    kw = dict(p2vec_kind=1, optimiser="adam", eta=0.001, beta=(0.9, 0.999), weight_decay=1e-8)
    model, _ = prob.case.model(p)
    st = None
    order = np.concatenate([g.permutation(12) for _ in range(3)])
    for s in range(30):
        idx = order[s:s + 1]
        loss, grad = prob.loss_grad(p, idx)
        p_host = p.copy(); opt.update(p_host, grad)
        r = engine.train_steps(model, prob.opts, prob.dataset, idx, prob.yscale, p, st, prob.case.loss_kind, **kw)
        np.testing.assert_allclose(r["step_loss"][0], loss, rtol=1e-13)
        np.testing.assert_allclose(r["p"], p_host, rtol=1e-12, atol=1e-15, err_msg=f"step {s}")
        p, st = r["p"], r["opt_state"]
        adam = opt.chain[0]
        adam.m, adam.v = st[:24].copy(), st[24:48].copy()
    # a model the device has no p2vec for is refused, as is a mismatching kind
    from crnn_b200.engine import EngineError
    with pytest.raises(EngineError):
        engine.train_steps(model, prob.opts, prob.dataset, order[:2], prob.yscale, p, None, prob.case.loss_kind, p2vec_kind=2)
    pbr = make_problem("robertson", golden, 4)                                 # the robertson loop integrates with Rosenbrock23
    probr = CRNNProblem("robertson", pbr["u0"], pbr["data"], pbr["yscale"], engine=engine, alg=_abi.ALG_TSIT5)
    with pytest.raises(EngineError):
        engine.train_steps(pbr["model"], probr.opts, probr.dataset, np.arange(2), probr.yscale, np.zeros(43), None, probr.case.loss_kind, p2vec_kind=4)
    with pytest.raises(EngineError):                                           # dataset row out of range
        engine.train_steps(model, prob.opts, prob.dataset, np.array([99]), prob.yscale, p, None, prob.case.loss_kind, **kw)


def test_on_device_loop_case3_p2vec(engine, golden):
    """p2vec_kind = 3: case3/case3.jl:42-53 on the device (w_out = -w_in .* |w_out_raw|, w_in = clamp(w_in, 0, 4), dy_std folded in),
    153 parameters = five warps per trajectory in the forward-sensitivity kernel reading its weights and seed columns from device
    memory, log-MAE loss, NADAM(0.001) as in case3.jl:20; every step against the host mirror from the same (p, state)"""
    from problems import make_problem, trained_p
    pb = make_problem("case3", golden, 10)
    prob = CRNNProblem("case3", pb["u0"], np.abs(pb["data"]) + 1e-6, pb["yscale"], out_scale=pb["model"].out_scale, engine=engine)
    g = np.random.default_rng(5)
    p = trained_p("case3", golden)                                         # the script's Xavier initialisation (case3.jl:35-36)
    assert p.size == 153 and prob.opts.sens_mode == _abi.SENS_FORWARD
    opt = optim.Optimiser(optim.NADAM(0.001, (0.9, 0.999)))
    kw = dict(p2vec_kind=3, optimiser="nadam", eta=0.001, beta=(0.9, 0.999))
    model, _ = prob.case.model(p, prob.out_scale)
    st = None
    order = np.concatenate([g.permutation(10) for _ in range(2)])
    for s in range(16):
        idx = order[s:s + 1]
        loss, grad = prob.loss_grad(p, idx)
        p_host = p.copy(); opt.update(p_host, grad)
        r = engine.train_steps(model, prob.opts, prob.dataset, idx, prob.yscale, p, st, prob.case.loss_kind, **kw)
        np.testing.assert_allclose(r["step_loss"][0], loss, rtol=1e-13)
        np.testing.assert_allclose(r["step_gnorm"][0], np.linalg.norm(grad), rtol=1e-11)
        np.testing.assert_allclose(r["p"], p_host, rtol=1e-11, atol=1e-14, err_msg=f"step {s}")
        p, st = r["p"], r["opt_state"]
        nadam = opt.chain[0]
        nadam.m, nadam.v = st[:153].copy(), st[153:306].copy()
    # a mini-batch of four, two optimiser steps in one call = the same steps taken one call at a time
    r_a = engine.train_steps(model, prob.opts, prob.dataset, order[:8], prob.yscale, p, st, prob.case.loss_kind, batch=4, **kw)
    r_1 = engine.train_steps(model, prob.opts, prob.dataset, order[:4], prob.yscale, p, st, prob.case.loss_kind, batch=4, **kw)
    r_2 = engine.train_steps(model, prob.opts, prob.dataset, order[4:8], prob.yscale, r_1["p"], r_1["opt_state"], prob.case.loss_kind, batch=4, **kw)
    assert np.array_equal(r_a["p"], r_2["p"]) and np.array_equal(r_a["opt_state"], r_2["opt_state"])
    l4, g4 = prob.loss_grad(p, order[:4])
    np.testing.assert_allclose(r_1["step_loss"][0], l4, rtol=1e-12)
    # and the frontend's epoch loop picks the kernel by the case name
    p_end, hist = prob.train_on_device(p, n_epoch=2, n_exp_train=8, rng=g, optimiser="nadam", eta=0.001)
    assert np.isfinite([h[0] for h in hist]).all() and p_end.shape == (153,)


def test_on_device_loop_robertson_p2vec(engine, golden):
    """p2vec_kind = 4: robertson/rober_crnn.jl:85-96 on the device (slope = |p[end]|, w_b = p .* 10 slope, w_out = -w_in .* 10 .^ w_out,
    w_in = clamp(w_in, 0, 2.5), dydt_scale folded in), the Rosenbrock23 sensitivity kernel reading weights and seed columns from
    device memory, the per-visit truncation `sample = rand(batchsize:datasize)` (:218), the 2-norm clip (:220-223) and
    ADAMW(0.005, (0.9, 0.999), 1e-6) (:19); every step against the host mirror from the same (p, state), from the reference's
    committed checkpoint"""
    from problems import make_problem
    pb = make_problem("robertson", golden, 12)
    prob = CRNNProblem("robertson", pb["u0"], pb["data"], pb["yscale"], out_scale=pb["model"].out_scale, engine=engine)
    assert prob.opts.alg == _abi.ALG_ROSENBROCK23
    g = np.random.default_rng(8)
    p = np.array(golden["robertson"]["p"], dtype=np.float64)
    opt = optim.ADAMW(0.005, (0.9, 0.999), 1e-6)
    grad_max = 10.0                                                        # rober_crnn.jl:29
    kw = dict(p2vec_kind=4, optimiser="adam", eta=0.005, beta=(0.9, 0.999), weight_decay=1e-6, grad_max=grad_max)
    model, _ = prob.case.model(p, prob.out_scale)
    st = None
    order = np.concatenate([g.permutation(12) for _ in range(2)])
    sample = g.integers(32, 41, size=order.size)                           # batchsize = 32, datasize = 40 (:21,218)
    for s in range(20):
        idx = order[s:s + 1]
        loss, grad = prob.loss_grad(p, idx, sample=sample[s:s + 1])
        gn = np.linalg.norm(grad)
        gc, _ = optim.clip_by_norm(grad, grad_max)
        p_host = p.copy(); opt.update(p_host, gc)
        r = engine.train_steps(model, prob.opts, prob.dataset, idx, prob.yscale, p, st, prob.case.loss_kind, n_save_used=sample[s:s + 1], **kw)
        np.testing.assert_allclose(r["step_loss"][0], loss, rtol=1e-10)
        np.testing.assert_allclose(r["step_gnorm"][0], gn, rtol=1e-8)
        np.testing.assert_allclose(r["p"], p_host, rtol=1e-9, atol=1e-12, err_msg=f"step {s}")
        p, st = r["p"], r["opt_state"]
        adam = opt.chain[0]
        adam.m, adam.v = st[:43].copy(), st[43:86].copy()
    # K steps in one call = K calls of one step; the frontend's epoch loop draws the truncation itself
    kw["grad_max"] = 1e-3                                                  # every step clipped
    r_a = engine.train_steps(model, prob.opts, prob.dataset, order[:6], prob.yscale, p, st, prob.case.loss_kind, batch=2, n_save_used=sample[:6], **kw)
    assert (r_a["step_gnorm"] > 1e-3).all()
    q, stq = p, st
    for k in range(3):
        r_k = engine.train_steps(model, prob.opts, prob.dataset, order[2 * k:2 * k + 2], prob.yscale, q, stq, prob.case.loss_kind, batch=2,
                                 n_save_used=sample[2 * k:2 * k + 2], **kw)
        q, stq = r_k["p"], r_k["opt_state"]
    assert np.array_equal(r_a["p"], q) and np.array_equal(r_a["opt_state"], stq)
    with pytest.raises(Exception):
        engine.train_steps(model, prob.opts, prob.dataset, order[:2], prob.yscale, p, st, prob.case.loss_kind, n_save_used=np.array([0, 41]), **kw)
    p_end, hist = prob.train_on_device(p, n_epoch=2, n_exp_train=10, rng=g, sample_range=(32, 40), optimiser="adam", eta=0.005,
                                       weight_decay=1e-6, grad_max=grad_max)
    assert np.isfinite([h[0] for h in hist]).all() and p_end.shape == (43,)


def test_on_device_epochs_descend(engine):
    prob, p0, g = _case2_problem(engine)
    p_end, hist = prob.train_on_device(p0, n_epoch=30, n_exp_train=20, rng=g, optimiser="adam", eta=0.005, weight_decay=1e-6,
                                       expdecay=(5e-3, 0.5, 500 * 20, 1e-4))
    assert np.isfinite([h[0] for h in hist]).all()
    assert hist[-1][0] < 0.6 * hist[0][0] and hist[-1][1] < 0.6 * hist[0][1], (hist[0], hist[-1])


def test_hychem_training_loop_with_the_adjoint_gradient(engine):
    """HyChem/crnn_pyrolysis_mass.jl's loop (:197-212) on the engine: F2 RHS, 211 parameters, random time truncation
    `sample = rand(batch_size:ntotal)`, gradient clipping at 10, ADAMW(5e-3) — with the gradient from the discrete
    adjoint kernel instead of 18 chunked ForwardDiff re-solves.  Synthetic targets from a second random CRNN of the same
    shape (the script's data file is not in the reference tree); non-stiff variant, see DESIGN.md §8.1."""
    ys = np.array([0.05, 0.01, 0.01, 0.01, 0.02, 0.9, 0.01, 1e-4, 1e-3])
    c = cases.hychem_case()
    n_exp = 8
    u0 = cases.hychem_u0(n_exp)
    m_true, _ = c.model(cases.hychem_p(1, lnA_shift=-2.0), out_scale=ys / 0.01)
    data = engine.solve_batch(m_true, c.opts(alg=1), u0)["pred"]                  # targets by Rosenbrock23
    yscale = np.maximum(data.max(axis=(0, 1)) - data.min(axis=(0, 1)), 1e-8)      # clamp.(ymax - ymin, lb, Inf) (:71)
    prob = CRNNProblem(c, u0, data, yscale, out_scale=ys / 0.01, engine=engine)
    p0 = cases.hychem_p(0, lnA_shift=-2.0)
    g = np.random.default_rng(0)
    opt = optim.ADAMW(0.005, (0.9, 0.999), 1e-6)
    p_end, hist = prob.train(p0, opt, n_epoch=40, n_exp_train=n_exp, batch=n_exp, grad_max=10.0, rng=g, sample_range=(32, 40))
    losses = [h[0] for h in hist]
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.6 * losses[0], (losses[0], losses[-1])
    assert min(losses) > 0.0


def test_readme_python_example_runs(engine):
    """the code block under 'Using it from Python' in README.md, executed as written"""
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "README.md")).read()
    block = re.search(r"## Using it from Python.*?```python\n(.*?)```", txt, flags=re.S).group(1)
    cwd = os.getcwd()
    os.chdir(root)
    try:
        ns = {}
        exec(compile(block, "README.md", "exec"), ns)
    finally:
        os.chdir(cwd)
    assert ns["pred"].shape == (6, 50) and np.isfinite(ns["loss"]) and ns["grad"].shape == (25,)
    assert len(ns["hist"]) == 10 and ns["hist"][-1][0] < 0.05
    ns["eng"].close()


def test_yeast_epoch_loop_by_the_adjoint_gradient(engine, golden):
    """yeast_glycolysis.jl:243-249 on the engine: batch-1 steps over randperm(n_exp_train), the gradient of all 294 parameters
    (CRNN + Flux chain) from the discrete-adjoint kernel, the script's ExpDecay -> ADAMW chain (:39-40) and its random save-count
    `batch = rand(batch_min:ntotal)` (:245); from the reference's committed checkpoint knocked off by 3 % the loss comes back down"""
    from scipy.integrate import solve_ivp
    c = cases.CASES["yeast"]
    rng = np.random.default_rng(7)
    n_exp, n_train = 12, 10
    u0 = cases.YEAST_IC_LB + rng.random((n_exp, 7)) * (cases.YEAST_IC_UB - cases.YEAST_IC_LB)
    ts = c.saveat()
    data = np.array([solve_ivp(cases.yeast_true_rhs, (0, 5), u, method="Radau", rtol=1e-9, atol=1e-12, t_eval=ts).y.T for u in u0])
    prob = CRNNProblem("yeast", u0, data, data.std(axis=1).max(axis=0) + 1e-5, engine=engine)
    pc = np.array(golden["yeast"]["p"])
    p0 = pc * (1.0 + 0.03 * rng.standard_normal(pc.size))
    l_ckpt = float(np.mean(prob.loss_neuralode(pc, np.arange(n_train))))
    l0 = float(np.mean(prob.loss_neuralode(p0, np.arange(n_train))))
    assert l0 > 1.15 * l_ckpt
    # the frontend's gradient is the C call's: mean of the per-experiment gradients, all 294 entries live
    loss_b, grad_b = prob.loss_grad(p0, np.arange(n_train))
    grads = [prob.loss_grad(p0, i)[1] for i in range(n_train)]
    np.testing.assert_allclose(grad_b, np.mean(grads, axis=0), rtol=1e-9, atol=1e-12)
    assert grad_b.shape == (294,) and np.count_nonzero(grad_b[164:]) == 130 and np.count_nonzero(grad_b[:164]) > 100   # clamp(-w_out, 0, 4) zeroes some
    opt = optim.Optimiser(optim.ExpDecay(5e-3, 0.5, 100 * n_train, 1e-5), optim.ADAMW(0.005, (0.9, 0.999), 1e-6))
    p_end, hist = prob.train(p0, opt, 6, n_train, batch=1, rng=rng, sample_range=(32, 300))
    l_train = [h[0] for h in hist]
    assert np.isfinite(l_train).all() and min(l_train) < 0.75 * l0, (l0, l_train)
