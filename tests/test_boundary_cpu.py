"""CPU tests of the drop-in boundary and the host logic: the C-ABI library loads and exports
every symbol include/crnn_b200.h declares (no compute calls: there is no GPU here), the ctypes
mirror matches the header, the product path fails loudly without CUDA, and the host-side
pieces (optimisers, sharding, synthetic inputs, the world_size-2 exchange over gloo) work."""
import ctypes
import json
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from crnn_b200 import _abi, cases, dist as cdist, optim, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "crnn_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crnn_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    declared = _declared_functions()
    assert set(declared) == set(_abi.EXPORTS), (declared, _abi.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.crnn_version() == 200


def test_ctypes_mirror_matches_header_layout():
    """Compile a probe against the real header and compare sizeof/offsetof with the ctypes structs."""
    probe = textwrap.dedent("""
        #include <stdio.h>
        #include <stddef.h>
        #include "crnn_b200.h"
        int main(void) {
          printf("%zu %zu %zu ", sizeof(crnn_model), sizeof(crnn_opts), sizeof(crnn_stats));
          printf("%zu %zu %zu %zu ", offsetof(crnn_model, lb), offsetof(crnn_model, out_scale), offsetof(crnn_model, w_out),
                 offsetof(crnn_opts, maxiters));
          printf("%zu %zu %zu %zu\\n", offsetof(crnn_opts, pred_clamp_lo), offsetof(crnn_opts, obs_idx), offsetof(crnn_opts, qmin),
                 offsetof(crnn_opts, stream));
          return 0;
        }""")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c"); exe = os.path.join(d, "p")
        open(c, "w").write(probe)
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.run([cc, "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    M, O, S = _abi.CModel, _abi.COpts, _abi.CStats
    want = [ctypes.sizeof(M), ctypes.sizeof(O), ctypes.sizeof(S), M.lb.offset, M.out_scale.offset, M.w_out.offset,
            O.maxiters.offset, O.pred_clamp_lo.offset, O.obs_idx.offset, O.qmin.offset, O.stream.offset]
    assert got == want


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to exist; without the .so it refuses to load."""
    import torch
    from crnn_b200.engine import Engine, EngineError
    if not torch.cuda.is_available():
        with pytest.raises(EngineError):
            Engine(0)
        lib = _abi.load_library()
        h = ctypes.c_void_p()
        assert lib.crnn_create(ctypes.byref(h), 0) == _abi.ERR_NO_DEVICE
    with pytest.raises(RuntimeError):
        _abi.load_library(os.path.join(ROOT, "does_not_exist.so"))


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under crnn_b200/ may import, include, link or load it
    (comments may cite it as the specification a kernel mirrors)."""
    bad = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|#\s*include\s*[\"<][^\">]*oracle|liboracle|oracle\.(lib|build)\(", re.M)
    for base, _, files in os.walk(os.path.join(ROOT, "crnn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(base, f)).read()
                assert not bad.search(txt), f"{f} uses the oracle"


def test_model_and_opts_validation():
    with pytest.raises(ValueError):
        cases.CRNNModel(w_in=np.ones((3, 2)), w_b=np.zeros(3), w_out=np.ones((3, 2)))
    with pytest.raises(ValueError):
        cases.CRNNModel(w_in=np.ones((3, 2)), w_b=np.zeros(2), w_out=np.ones((3, 2)), rhs_kind=_abi.RHS_F1)
    m = cases.true_model_case2()
    assert (m.n_state, m.n_species, m.n_in, m.n_reac, m.n_w) == (7, 6, 7, 3, 42)
    with pytest.raises(ValueError):
        cases.SolveOpts(saveat=np.array([0.0, 60.0]), t0=0.0, t1=50.0).to_c(7)
    with pytest.raises(ValueError):
        cases.SolveOpts(saveat=np.array([0.0, 5.0]), t0=0.0, t1=50.0, obs_idx=np.array([9])).to_c(7)
    o, keep = cases.CASES["robertson"].opts().to_c(3)
    assert (o.n_abstol, o.n_reltol, o.n_save, o.alg) == (3, 3, 40, _abi.ALG_ROSENBROCK23)
    assert abs(o.saveat[39] - 1e5) < 1e-6 and o.saveat[0] == 1.0


def test_synth_inputs_follow_the_scripts_and_shard():
    u = synth.make_u0("case2", 5000)
    assert u[:, 0:2].min() >= 0.2 and u[:, 0:2].max() <= 2.2 and np.all(u[:, 2:6] == 0)
    assert u[:, 6].min() >= 323 and u[:, 6].max() <= 343                  # case2.jl:62-65
    r = synth.make_u0("robertson", 3000)
    assert np.all(r[:, 1] == 1e-8) and r[:, [0, 2]].min() >= 0.5 and r[:, [0, 2]].max() <= 1.5
    c3 = synth.make_u0("case3", 2000)
    assert c3.min() >= 1e-3 and c3.max() <= 1.0                           # 10^(-3 U)
    assert np.array_equal(synth.make_u0("case2", 700, start=900), u[900:1600])   # shard == slice of the whole
    x = np.ones((1500, 4, 3))
    assert np.array_equal(synth.noisy_targets(x[:200], 0.05, start=1100), synth.noisy_targets(x, 0.05)[1100:1300])
    assert np.allclose(synth.yscale_from(np.arange(24.0).reshape(2, 4, 3), 1.0), [10, 10, 10])


def test_shard_bounds_cover_exactly():
    for N in (0, 1, 7, 65536, 1000003):
        for w in (1, 2, 3, 8):
            b = [cdist.shard_bounds(N, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_optimiser_chains_nest_like_flux():
    """case2/case2.jl:31-32 written literally: Flux.Optimiser(ExpDecay(...), ADAMW(...)) - ADAMW is itself an Optimiser"""
    from crnn_b200 import optim
    g = np.random.default_rng(0).standard_normal(7)
    p1, p2 = np.ones(7), np.ones(7)
    nested = optim.Optimiser(optim.ExpDecay(5e-3, 0.5, 500, 1e-4), optim.ADAMW(5e-3, (0.9, 0.999), 1e-6))
    flat = optim.Optimiser(optim.ExpDecay(5e-3, 0.5, 500, 1e-4), *optim.ADAMW(5e-3, (0.9, 0.999), 1e-6).chain)
    for _ in range(5):
        nested.update(p1, g); flat.update(p2, g)
    assert np.array_equal(p1, p2) and not np.array_equal(p1, np.ones(7))


def test_optimisers_match_flux_formulas():
    """ADAMW = ADAM then decoupled weight decay."""
    p = np.array([1.0, -2.0]); g = np.array([0.5, 0.25])
    opt = optim.ADAMW(5e-3, (0.9, 0.999), 1e-6)
    p1 = opt.update(p.copy(), g)
    m = 0.1 * g; v = 0.001 * g * g
    d = m / (1 - 0.9) / (np.sqrt(v / (1 - 0.999)) + 1e-8) * 5e-3 + 1e-6 * p
    np.testing.assert_allclose(p1, p - d, rtol=1e-14)
    # second step uses beta powers ^2
    p2 = opt.update(p1.copy(), g)
    m = 0.9 * m + 0.1 * g; v = 0.999 * v + 0.001 * g * g
    d = m / (1 - 0.81) / (np.sqrt(v / (1 - 0.999**2)) + 1e-8) * 5e-3 + 1e-6 * p1
    np.testing.assert_allclose(p2, p1 - d, rtol=1e-13)
    e = optim.ExpDecay(5e-3, 0.5, 2, 1e-4)
    assert np.allclose(e.apply(p, g), 5e-3 * g) and np.allclose(e.apply(p, g), 2.5e-3 * g)   # decays on the 2nd call
    gc, n = optim.clip_by_norm(np.array([30.0, 40.0]), 10.0)
    assert n == 50.0 and np.allclose(gc, [6.0, 8.0])
    n1 = optim.Optimiser(optim.NADAM(1e-3)).update(p.copy(), g)
    assert np.all(np.abs(n1 - p) < 2e-3) and np.all(np.sign(p - n1) == np.sign(g))


WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
from crnn_b200 import dist as cdist, synth
from oracle import oracle
from problems import make_problem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
golden = json.load(open(os.path.join({root!r}, "tests", "golden", "checkpoints.json")))
N = 37                                  # uneven split on purpose
pb = make_problem("case2", golden, N)
lo, hi = cdist.shard_bounds(N, world, rank)
# each rank regenerates ITS shard of the inputs from the counter-based stream (no scatter)
u0 = synth.make_u0("case2", hi - lo, start=lo)
assert np.array_equal(u0, pb["u0"][lo:hi])
r = oracle.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], u0, pb["data"][lo:hi], pb["yscale"])
loss, grad = cdist.allreduce_loss_grad(r["loss"].sum(), hi - lo, r["grad_sum"])
if rank == 0:
    whole = oracle.loss_grad_batch(pb["model"], pb["opts"], pb["seed"], pb["u0"], pb["data"], pb["yscale"])
    np.testing.assert_allclose(loss, whole["loss"].mean(), rtol=1e-13)
    np.testing.assert_allclose(grad, whole["grad_sum"] / N, rtol=1e-11, atol=1e-15)
    print("DIST_OK", world)
dist.destroy_process_group()
'''


def test_two_rank_shard_and_allreduce_over_gloo(tmp_path):
    """world_size 2 on CPU: shards + the one exchange ([loss, n, grad_sum] all-reduce) reproduce the
    whole-batch result (the oracle stands in for the per-rank engine: there is no GPU here)."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK 2" in r.stdout


# ---------------------------------------------------------------- BSON.jl checkpoints (host-side utility)
def test_checkpoint_round_trip(tmp_path):
    from crnn_b200 import checkpoint as ck
    p = np.random.default_rng(0).standard_normal(25)
    f = str(tmp_path / "mymodel.bson")
    ck.save(f, p, 3700, l_loss_train=[0.139, 0.0165], l_loss_val=[0.1235, 0.01396])
    c = ck.load(f)
    assert np.array_equal(c["p"], p) and c["iter"] == 3700 and c["opt"] is None
    np.testing.assert_allclose(c["l_loss_train"], [0.139, 0.0165], rtol=1e-7)      # Float32 on disk, like the reference's
    np.testing.assert_allclose(c["l_loss_val"], [0.1235, 0.01396], rtol=1e-7)
    raw = open(f, "rb").read()
    doc, end = ck.parse(raw)
    assert end == len(raw) and ck.encode(doc) == raw
    assert list(doc.keys()) == ["iter", "p", "l_loss_train", "l_loss_val"] and doc["p"]["type"]["name"] == ["Core", "Float64"]
    # matrices are stored column-major, like Julia
    m = np.arange(6.0).reshape(2, 3)
    node = ck.lower_array(m)
    assert node["size"] == [2, 3] and np.array_equal(ck.resolve({}, node), m)
    assert np.frombuffer(bytes(node["data"])).tolist() == [0.0, 3.0, 1.0, 4.0, 2.0, 5.0]


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the authoring container")
def test_checkpoint_reader_writer_speak_the_reference_dialect(golden):
    """the four checkpoints the reference commits: parse -> encode is byte-exact, our array lowering equals theirs node for
    node, and the decoded parameter vectors are the golden fixtures"""
    from crnn_b200 import checkpoint as ck
    for name, key in (("case2", "case2"), ("robertson", "robertson"), ("gene-regulatory-network", "gene"), ("yeast-glycolysis", None)):
        path = f"/root/reference/{name}/checkpoint/mymodel.bson"
        raw = open(path, "rb").read()
        doc, end = ck.parse(raw)
        assert end == len(raw) and ck.encode(doc) == raw, name
        c = ck.load(path)
        node = doc["p"]
        if isinstance(node, dict) and node.get("tag") == "backref":
            node = doc["_backrefs"][node["ref"] - 1]
        assert ck.encode(ck.lower_array(c["p"])) == ck.encode(node), name
        if key:
            assert np.array_equal(c["p"], np.array(golden[key]["p"])) and c["iter"] == golden[key]["iter"]
        hist = [k for k in c if k.startswith("l_loss") or k.startswith("list_loss")]
        assert len(hist) >= 1 and all(len(c[k]) == c["iter"] for k in hist), name
    # carrying the optimiser state and the shared-object table over: still a well-formed document with the same p
    import tempfile
    c = ck.load("/root/reference/case2/checkpoint/mymodel.bson")
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "resaved.bson")
        pnew = c["p"] + 1e-3
        assert c["p_ref"] == 4        # in the reference's file p is a backref: the object that keys the ADAM state's IdDict
        ck.save(f, pnew, c["iter"] + 1, opt=c["opt"], backrefs=c["doc"]["_backrefs"], p_ref=c["p_ref"],
                l_loss_train=c["l_loss_train"], l_loss_val=c["l_loss_val"])
        c2 = ck.load(f)
        assert np.array_equal(c2["p"], pnew) and c2["iter"] == c["iter"] + 1 and c2["opt"] == c["opt"]
        # ... and it still is one after re-saving with new values: the optimiser state stays attached to the live p
        assert c2["doc"]["p"] == {"tag": "backref", "ref": 4}
        assert np.array_equal(np.asarray(ck.resolve(c2["doc"], c2["doc"]["_backrefs"][3])).reshape(-1), pnew)
        np.testing.assert_allclose(c2["l_loss_train"], c["l_loss_train"], rtol=0, atol=0)


def test_pruned_p2vec_variants(golden):
    """the evaluate-only pruning scripts: hard thresholds inside p2vec (case1_hardthreshhold.jl:76-77, case2_pruning.jl:105-106,
    case3_pruning.jl:243-247) — the engine serves them unchanged, the map zeroes weights and their seed rows"""
    p = np.array(golden["case2"]["p"])
    w_in, w_b, w_out, seed = cases.p2vec_case2(p, p_cutoff=0.2)
    w_in0, w_b0, w_out0, _ = cases.p2vec_case2(p)
    raw = p[3:21].reshape(3, 6).T
    assert np.array_equal(w_out == 0.0, np.abs(raw) < 0.2) and (np.abs(raw) < 0.2).sum() >= 5
    assert np.array_equal(w_out[w_out != 0], w_out0[w_out != 0]) and np.array_equal(w_b, w_b0)
    assert np.all(w_in[:6][w_out == 0.0] == 0.0) and np.array_equal(w_in[6], w_in0[6])       # Ea row untouched
    off_out = 7 * 3 + 3
    assert np.all(seed[off_out:][(w_out == 0.0).reshape(-1, order="F")] == 0.0)              # pruned weights do not train
    # the pruned trained model still follows the generating mechanism (the trained w_out is near-integer)
    from oracle import oracle
    c = cases.CASES["case2"]
    u0 = synth.make_u0("case2", 16)
    m = cases.CRNNModel(w_in=w_in, w_b=w_b, w_out=w_out, rhs_kind=_abi.RHS_F1, lb=c.lb, ub=c.ub)
    m0, _ = c.model(p)
    a = oracle.solve_batch(m, c.opts(obs_idx=np.arange(6)), u0)["pred"]
    b = oracle.solve_batch(m0, c.opts(obs_idx=np.arange(6)), u0)["pred"]
    assert np.abs(a - b).max() < 0.05
    g = np.random.default_rng(0)
    p1 = g.standard_normal(24)
    w_in1, _, w_out1, _ = cases.p2vec_case1(p1, p_cutoff=0.5)
    assert np.array_equal(w_out1 == 0.0, np.abs(p1[4:].reshape(4, 5).T) < 0.5) and np.all(w_in1[w_out1 == 0.0] == 0.0)
    p3 = (g.random(153) - 0.5) * 2 * np.sqrt(6 / 17)
    w_in3, _, w_out3, seed3 = cases.p2vec_case3(p3, p_cutoff=0.1, dy_std=np.linspace(0.5, 1.5, 9))
    w_in3f, _, w_out3f, _ = cases.p2vec_case3(p3)
    assert (w_out3 == 0).sum() > (w_out3f == 0).sum() and np.all(np.abs(w_in3[w_in3 != 0]) >= 0.1)
    assert np.array_equal(w_out3[w_out3 != 0], w_out3f[w_out3 != 0])


def test_julia_shim_structs_follow_the_header():
    """julia/CRNNB200.jl cannot be executed here (no Julia): at least its struct mirrors must list the header's fields in the
    header's order with matching widths, and every ccall must name an exported symbol."""
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    jl = open(os.path.join(ROOT, "julia", "CRNNB200.jl")).read()
    jl_nocomment = re.sub(r"#.*", "", jl)
    ctype = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64"}

    def c_fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), hdr, flags=re.S).group(1)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s*(\*?)\s*(.*)", decl)
            base, ptr, names = m.group(2), m.group(3), m.group(4)
            for nm in names.split(","):
                nm = nm.strip()
                is_ptr = bool(ptr) or nm.startswith("*")
                nm = nm.lstrip("* ")
                if base == "void":
                    jt = "Ptr{Cvoid}"
                else:
                    jt = f"Ptr{{{ctype[base]}}}" if is_ptr else ctype[base]
                out.append((nm, jt))
        return out

    def jl_fields(struct):
        body = re.search(r"struct %s\n(.*?)\nend" % struct, jl_nocomment, flags=re.S).group(1)
        return [(n, t) for n, t in re.findall(r"(\w+)::([\w{}]+)", body)]

    assert jl_fields("CModel") == c_fields("crnn_model")
    assert jl_fields("COpts") == c_fields("crnn_opts")
    assert jl_fields("CTrainOpts") == c_fields("crnn_train_opts")
    assert [n for n, _ in _abi.CTrainOpts._fields_] == [n for n, _ in c_fields("crnn_train_opts")]
    called = set(re.findall(r"ccall\(\(:(\w+), LIB\)", jl))
    assert called and called <= set(_abi.EXPORTS), called - set(_abi.EXPORTS)
    assert {"crnn_create", "crnn_destroy", "crnn_solve_batch", "crnn_loss_grad_batch", "crnn_last_error"} <= called
    # every ccall passes as many arguments as the C prototype declares
    flat = re.sub(r"\s+", " ", hdr)
    for fn in called:
        proto = re.search(r"\b%s\s*\((.*?)\)\s*;" % fn, flat).group(1)
        n_c = 0 if proto.strip() in ("", "void") else proto.count(",") + 1
        m = re.search(r"ccall\(\(:%s, LIB\), \w+,\s*\((.*?)\),\s*\n?\s*(.*?)\)\)?\s*(?:\n|$)" % fn, jl, flags=re.S)
        types = m.group(1)
        depth = 0; n_j = 1 if types.strip() else 0
        for ch in types:
            depth += ch == "{"; depth -= ch == "}"
            n_j += (ch == "," and depth == 0)
        if types.rstrip().endswith(","):
            n_j -= 1
        assert n_j == n_c, (fn, n_j, n_c)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a small sample: exactly one JSON line with the
    contract's keys; needs no GPU"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "512"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "trajectories/s" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] and "workload" in d["config"]
    # a non-zero rank of a torchrun launch exits silently
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                        text=True, cwd=ROOT, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
