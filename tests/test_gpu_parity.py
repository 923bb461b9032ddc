"""Parity of the CUDA path (through the nn.Module API -> autograd.Function -> C ABI) against the
numpy oracle and the reference-generated golden fixtures. Run on a B200: pytest -m gpu.

Tolerances (rel-Frobenius), from BASELINE.json's north star: 1e-3 for the fp32 modes, 2e-2 for
the bf16 mode on outputs. Token / graph gradients in the single-pass bf16 mode are ~10x the
output error and are gated at 6e-2; the fp32 modes hold 1e-3 on gradients too.
"""
import numpy as np
import pytest
import torch

from conftest import golden, make_inputs, params_of, rel_err
from oracle import moment_oracle as O

pytestmark = pytest.mark.gpu

TOL_OUT = {"fp32_simt": 5e-5, "fp32": 1e-3, "bf16": 2e-2}
TOL_GRAD = {"fp32_simt": 5e-5, "fp32": 1e-3, "bf16": 6e-2}
MODES = ["fp32_simt", "fp32", "bf16"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda")


def npy(t):
    return t.detach().cpu().double().numpy()


# ------------------------------------------------------------------ stage parity vs oracle
@pytest.fixture(scope="module")
def stage_case():
    B, N, D, P, Q, K = 2, 197, 768, 3, 3, 5
    anchor, positive = make_inputs(B, N, D)
    torch.manual_seed(0)
    alpha = torch.rand(P + 1, Q + 1) * 0.1
    gd = torch.Generator().manual_seed(4321)
    dvec = torch.randn(B, D * (D + 1) // 2, generator=gd)
    S = 2048
    hashes = torch.randint(0, S, (3, D), generator=gd)
    signs = torch.randint(0, 2, (3, D), generator=gd) * 2 - 1
    dsk = torch.randn(B, S, generator=gd)
    a64, p64, al64 = npy(anchor), npy(positive), npy(alpha)
    fw = O.gpf_forward(a64, p64, al64)
    third = {"hashes": hashes.numpy(), "signs": signs.numpy(), "sketch_dim": S}
    st = O.moment_forward(a64, fw["G"], K, 1e-5, third)
    dZ, dG = O.moment_backward(a64, fw["G"], K, npy(dvec), 1e-5, third, npy(dsk))
    da, dp, dal = O.gpf_backward(a64, p64, al64, dG)
    return dict(anchor=anchor, positive=positive, alpha=alpha, dvec=dvec, dsk=dsk, hashes=hashes,
                signs=signs, S=S, K=K, fw=fw, st=st, dG=dG, da=da + dZ, dp=dp, dal=dal)


@pytest.mark.parametrize("mode", MODES)
def test_every_stage_and_gradient_matches_the_oracle(pkg, dev, stage_case, mode):
    c = stage_case
    EF = pkg.functional
    with EF.precision(mode):
        a = c["anchor"].to(dev).requires_grad_(True)
        p = c["positive"].to(dev).requires_grad_(True)
        al = c["alpha"].to(dev).requires_grad_(True)
        G = EF.gpf_fused_graph(a, p, torch.nn.functional.softplus(al))
        G.retain_grad()
        M2, u = EF.graph_weighted_pool(a, G, eps=1e-5, third_order=True)
        isq = EF.newton_schulz(M2, c["K"], 1e-5)
        vec = EF.half_vectorize(isq)
        h, s = c["hashes"].to(dev), c["signs"].to(dev)
        sk = EF.tensor_sketch(u, h, s, EF.build_sketch_csr(h, s, c["S"]), c["S"])
        ((vec * c["dvec"].to(dev)).sum() + (sk * c["dsk"].to(dev)).sum()).backward()
    to, tg = TOL_OUT[mode], TOL_GRAD[mode]
    assert rel_err(npy(G), c["fw"]["G"]) < to
    assert rel_err(npy(M2), c["st"]["M2"]) < to
    assert rel_err(npy(u), c["st"]["u"]) < to
    assert rel_err(npy(isq), c["st"]["isqrt"]) < to
    assert rel_err(npy(vec), c["st"]["vec"]) < to
    assert rel_err(npy(sk), c["st"]["sketch"]) < to
    assert rel_err(npy(G.grad), c["dG"]) < tg
    assert rel_err(npy(a.grad), c["da"]) < tg
    assert rel_err(npy(p.grad), c["dp"]) < tg
    assert rel_err(npy(al.grad), c["dal"]) < tg
    # the graph is exactly symmetric and non-negative (gpf_kernel.py:153-157)
    assert torch.equal(G, G.transpose(-2, -1)) and float(G.detach().min()) >= 0.0


@pytest.mark.parametrize("mode", MODES)
def test_lowrank_newton_schulz_is_the_same_function(pkg, dev, stage_case, mode):
    """functional.moment_isqrt(algorithm='lowrank'): pooling + iSQRT-COV with every Newton-Schulz
    product on N x N matrices. Same outputs and gradients as the oracle's dense
    restatement of moment_head.py:279-296."""
    c = stage_case
    EF = pkg.functional
    with EF.precision(mode):
        a = c["anchor"].to(dev).requires_grad_(True)
        G = torch.from_numpy(c["fw"]["G"]).float().to(dev).requires_grad_(True)
        isq, u = EF.moment_isqrt(a, G, c["K"], eps=1e-5, third_order=True, algorithm="lowrank")
        vec = EF.half_vectorize(isq)
        h, s = c["hashes"].to(dev), c["signs"].to(dev)
        sk = EF.tensor_sketch(u, h, s, EF.build_sketch_csr(h, s, c["S"]), c["S"])
        ((vec * c["dvec"].to(dev)).sum() + (sk * c["dsk"].to(dev)).sum()).backward()
    to, tg = TOL_OUT[mode], TOL_GRAD[mode]
    assert rel_err(npy(isq), c["st"]["isqrt"]) < to
    assert rel_err(npy(u), c["st"]["u"]) < to
    assert rel_err(npy(G.grad), c["dG"]) < tg
    dZ, _ = O.moment_backward(npy(c["anchor"]), c["fw"]["G"], c["K"], npy(c["dvec"]), 1e-5,
                              {"hashes": c["hashes"].numpy(), "signs": c["signs"].numpy(), "sketch_dim": c["S"]},
                              npy(c["dsk"]))
    assert rel_err(npy(a.grad), dZ) < tg


@pytest.mark.parametrize("K", [1, 2, 3])
def test_lowrank_on_a_nonsymmetric_graph_and_few_iterations(pkg, dev, K):
    EF = pkg.functional
    g = torch.Generator().manual_seed(K)
    B, N, D = 3, 21, 40
    Z = torch.randn(B, N, D, generator=g)
    graph = torch.rand(B, N, N, generator=g) + 0.05        # not symmetric
    dvec = torch.randn(B, D * (D + 1) // 2, generator=g)
    st = O.moment_forward(npy(Z), npy(graph), K, 1e-5)
    dZ, dG = O.moment_backward(npy(Z), npy(graph), K, npy(dvec), 1e-5)
    for mode, tol in (("fp32_simt", 2e-4), ("fp32", 2e-3)):
        with EF.precision(mode):
            z = Z.to(dev).requires_grad_(True)
            gr = graph.to(dev).requires_grad_(True)
            vec = EF.half_vectorize(EF.moment_isqrt(z, gr, K, eps=1e-5, algorithm="lowrank"))
            (vec * dvec.to(dev)).sum().backward()
        assert rel_err(npy(vec), st["vec"]) < tol
        assert rel_err(npy(z.grad), dZ) < 5 * tol
        assert rel_err(npy(gr.grad), dG) < 5 * tol


def test_moment_head_lowrank_switch(pkg, dev):
    """set_ns_algorithm('lowrank') changes how MomentHead evaluates, not what."""
    EF = pkg.functional
    torch.manual_seed(3)
    head = pkg.MomentHead(96, 24, use_third_order=True, isqrt_iterations=5, sketch_dim=128).to(dev).eval()
    gpf = pkg.GraphPolynomialFusion(2, 2).to(dev)
    a = torch.randn(4, 30, 96, device=dev)
    p = a + 0.5 * torch.randn_like(a)
    outs = {}
    for algo in ("dense", "lowrank"):
        EF.set_ns_algorithm(algo)
        try:
            x = a.clone().requires_grad_(True)
            out = head(x, gpf(x, p))
            out.square().sum().backward()
            outs[algo] = (out.detach(), x.grad.clone())
        finally:
            EF.set_ns_algorithm("dense")
    assert rel_err(npy(outs["lowrank"][0]), npy(outs["dense"][0])) < 1e-3
    assert rel_err(npy(outs["lowrank"][1]), npy(outs["dense"][1])) < 2e-3
    with pytest.raises(ValueError):
        EF.set_ns_algorithm("magic")


@pytest.mark.parametrize("algo,K", [("dense", 2), ("dense", 3), ("dense", 5), ("lowrank", 1), ("lowrank", 4)])
@pytest.mark.parametrize("shape", [(3, 21, 40), (2, 50, 136), (2, 70, 64)])
def test_fused_moment_head_linear_matches_the_oracle(pkg, dev, algo, K, shape):
    """functional.moment_head_linear: pool -> iSQRT-COV -> packed half-vector -> Linear in one
    operator (the last Newton-Schulz product writes the Linear's operand planes; trace(M2), <dO,O>
    and <dA,A> come out of GEMM epilogues). Same outputs and gradients as the oracle's restatement
    of moment_head.py:279-300 followed by a Linear, incl. the third-order u branch and a
    non-symmetric graph."""
    EF = pkg.functional
    B, N, D = shape
    g = torch.Generator().manual_seed(7 * K + N)
    Z = torch.randn(B, N, D, generator=g)
    graph = torch.rand(B, N, N, generator=g) + 0.05        # not symmetric
    L, n_out = D * (D + 1) // 2, 24
    Wt = torch.randn(n_out, L, generator=g) / L ** 0.5
    bias = torch.randn(n_out, generator=g)
    dy = torch.randn(B, n_out, generator=g)
    du = torch.randn(B, D, generator=g)
    st = O.moment_forward(npy(Z), npy(graph), K, 1e-5, {"hashes": np.zeros((3, D), np.int64),
                                                        "signs": np.ones((3, D), np.int64), "sketch_dim": 4})
    y_ref = st["vec"] @ npy(Wt).T + npy(bias)
    dvec = npy(dy) @ npy(Wt)
    for mode, tol in (("fp32", 1e-3), ("bf16", 6e-2)):
        with EF.precision(mode):
            z = Z.to(dev).requires_grad_(True)
            gr = graph.to(dev).requires_grad_(True)
            w = Wt.to(dev).requires_grad_(True)
            b = bias.to(dev).requires_grad_(True)
            res = EF.moment_head_linear(z, gr, w, b, K, eps=1e-5, third_order=True, algorithm=algo)
            if algo == "lowrank" and N >= D:
                # N >= D: the low-rank form does not apply, the dense fused chain needs K >= 2
                assert (res is None) == (K < 2)
                if res is None:
                    continue
            y, u = res
            ((y * dy.to(dev)).sum() + (u * du.to(dev)).sum()).backward()
        assert rel_err(npy(y), y_ref) < tol
        assert rel_err(npy(u), st["u"]) < tol
        # gradients: the oracle's backward takes d vec and (through the sketch hook) d u
        dZ, dG = O.moment_backward(npy(Z), npy(graph), K, dvec, 1e-5, None, None, du=npy(du))
        assert rel_err(npy(z.grad), dZ) < 3 * tol
        assert rel_err(npy(gr.grad), dG) < 3 * tol
        assert rel_err(npy(w.grad), npy(dy).T @ st["vec"]) < tol
        assert rel_err(npy(b.grad), npy(dy).sum(0)) < 1e-5
    # strict fp32 mode and K < 2 fall back to the composed operators
    with EF.precision("fp32_simt"):
        assert EF.moment_head_linear(Z.to(dev), graph.to(dev), Wt.to(dev), None, K) is None


@pytest.mark.parametrize("K", [2, 3, 5])
@pytest.mark.parametrize("shape", [(3, 21, 40), (2, 50, 136), (2, 70, 300), (1, 40, 520)])
def test_symmetric_graph_fast_path_matches_the_oracle(pkg, dev, K, shape):
    """EGM_MHD_SYMMETRIC_GRAPH: with an exactly symmetric graph every Newton-Schulz matrix is symmetric;
    the engine evaluates / stores upper 256-blocks only and the backward is the symmetric tangent chain.
    y, u, dZ, dW, db equal the oracle; dG equals the oracle's up to a skew-symmetric part."""
    EF = pkg.functional
    B, N, D = shape
    g = torch.Generator().manual_seed(11 * K + D)
    Z = torch.randn(B, N, D, generator=g)
    graph = torch.rand(B, N, N, generator=g) + 0.05
    graph = 0.5 * (graph + graph.transpose(-2, -1))          # exactly symmetric in fp32
    assert torch.equal(graph, graph.transpose(-2, -1))
    L, n_out = D * (D + 1) // 2, 16
    Wt = torch.randn(n_out, L, generator=g) / L ** 0.5
    bias = torch.randn(n_out, generator=g)
    dy = torch.randn(B, n_out, generator=g)
    du = torch.randn(B, D, generator=g)
    st = O.moment_forward(npy(Z), npy(graph), K, 1e-5, {"hashes": np.zeros((3, D), np.int64),
                                                        "signs": np.ones((3, D), np.int64), "sketch_dim": 4})
    y_ref = st["vec"] @ npy(Wt).T + npy(bias)
    dZ, dG = O.moment_backward(npy(Z), npy(graph), K, npy(dy) @ npy(Wt), 1e-5, None, None, du=npy(du))
    sym = lambda x: 0.5 * (x + np.swapaxes(x, -1, -2))
    for mode, tol in (("fp32", 1e-3), ("bf16", 6e-2)):
        with EF.precision(mode):
            z = Z.to(dev).requires_grad_(True)
            gr = EF.mark_symmetric(graph.to(dev).requires_grad_(True))
            assert EF.graph_is_symmetric(gr)
            w = Wt.to(dev).requires_grad_(True)
            b = bias.to(dev).requires_grad_(True)
            y, u = EF.moment_head_linear(z, gr, w, b, K, eps=1e-5, third_order=True, algorithm="dense")
            ((y * dy.to(dev)).sum() + (u * du.to(dev)).sum()).backward()
        assert rel_err(npy(y), y_ref) < tol
        assert rel_err(npy(u), st["u"]) < tol
        assert rel_err(npy(z.grad), dZ) < 3 * tol
        assert rel_err(npy(gr.grad), sym(dG)) < 3 * tol          # the symmetric part, itself symmetric
        assert rel_err(npy(w.grad), npy(dy).T @ st["vec"]) < tol
        assert rel_err(npy(b.grad), npy(dy).sum(0)) < 1e-5


def test_symmetric_tag_follows_the_tensor_version(pkg, dev):
    """GraphPolynomialFusion tags its (bit-exactly symmetric) output; an in-place edit or
    symmetric_enforce=False drops the tag, and the switch turns the fast path off."""
    EF = pkg.functional
    a, p = (t.to(dev) for t in make_inputs(2, 30, 64))
    gpf = pkg.GraphPolynomialFusion(2, 2).to(dev)
    G = gpf(a, p)
    assert torch.equal(G, G.transpose(-2, -1)) and EF.graph_is_symmetric(G)
    EF.set_symmetric_fast_path(False)
    try:
        assert not EF.graph_is_symmetric(G)
    finally:
        EF.set_symmetric_fast_path(True)
    with torch.no_grad():
        G[:, 0, 1] += 1.0
    assert not EF.graph_is_symmetric(G)
    assert not EF.graph_is_symmetric(pkg.GraphPolynomialFusion(2, 2, symmetric_enforce=False).to(dev)(a, p))
    assert not EF.graph_is_symmetric(gpf(a, p).detach().clone())


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_fast_path_and_general_path_give_the_same_training_gradients(pkg, dev, mode):
    """GPF -> MomentHead (BASELINE config 1 shape, B=4): outputs and every gradient the optimiser sees
    agree between the symmetric fast path and the general chain."""
    EF = pkg.functional
    B, N, D = 4, 197, 768
    a0, p0 = make_inputs(B, N, D)
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(3, 3).to(dev)
    head = pkg.MomentHead(D, 256, isqrt_iterations=5).to(dev).eval()
    dout = torch.randn(B, 256, generator=torch.Generator().manual_seed(4321)).to(dev)
    res = {}
    for fast in (True, False):
        EF.set_symmetric_fast_path(fast)
        try:
            with EF.precision(mode):
                a = a0.to(dev).requires_grad_(True)
                p = p0.to(dev).requires_grad_(True)
                gpf.zero_grad(); head.zero_grad()
                out = head(a, gpf(a, p))
                (out * dout).sum().backward()
        finally:
            EF.set_symmetric_fast_path(True)
        res[fast] = [npy(out), npy(a.grad), npy(p.grad), npy(gpf.alpha_coeffs.grad),
                     npy(head.second_net[0].weight.grad)]
    # each path sits within ~5e-4 (fp32 mode) of the fp64 oracle on token gradients (tests/gpu_fused_report.py)
    tol = 1e-3 if mode == "fp32" else 6e-2
    for x, y in zip(res[True], res[False]):
        assert rel_err(x, y) < tol


# ----------------------------------------------------------- golden fixtures (reference)
def _build_from_golden(pkg, rec, dev):
    B, N, D, P, Q, K, d_out, third, S, sym, train = [int(v) for v in rec["cfg"]]
    gpf = pkg.GraphPolynomialFusion(P, Q, similarity=str(rec["similarity"]), symmetric_enforce=bool(sym))
    head = pkg.MomentHead(D, d_out, use_third_order=bool(third), isqrt_iterations=K, sketch_dim=S)
    with torch.no_grad():
        gpf.alpha_coeffs.copy_(torch.from_numpy(rec["alpha"]).float())
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in params_of(rec).items()}
    sd = {k: (v.float() if v.dtype == torch.float64 else v) for k, v in sd.items()}
    head.load_state_dict(sd)
    for m in head.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    head.train(bool(train))
    return gpf.to(dev), head.to(dev)


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("mode", ["fp32_simt", "fp32"])
@pytest.mark.parametrize("name", ["small_p2q2", "small_p3q3_third", "small_dot_nosym", "small_trainbn"])
def test_small_goldens_forward_and_backward(pkg, dev, name, mode, fast):
    rec = golden(name)
    gpf, head = _build_from_golden(pkg, rec, dev)
    EF = pkg.functional
    a = torch.from_numpy(rec["anchor"]).float().to(dev).requires_grad_(True)
    p = torch.from_numpy(rec["positive"]).float().to(dev).requires_grad_(True)
    EF.set_symmetric_fast_path(fast)
    try:
        with EF.precision(mode):
            G = gpf(a, p)
            tagged = EF.graph_is_symmetric(G)
            out = head(a, G)
            # torch.autograd.grad captures the gradient that flows into G without observing the tensor
            # (no hook, no retain_grad), so a tagged graph stays on the symmetric fast path
            dG, da, dp, dal = torch.autograd.grad((out * torch.from_numpy(rec["dOut"]).float().to(dev)).sum(),
                                                  [G, a, p, gpf.alpha_coeffs])
    finally:
        EF.set_symmetric_fast_path(True)
    assert tagged == (fast and bool(gpf.symmetric_enforce))
    tol = 2e-4 if mode == "fp32_simt" else 1e-3
    if name == "small_trainbn":
        tol *= 20   # B=4 train-mode BatchNorm amplifies
    assert rel_err(npy(G), rec["G"]) < tol
    assert rel_err(npy(out), rec["out"]) < tol
    assert rel_err(npy(dal), rec["d_alpha"]) < 5 * tol
    assert rel_err(npy(da), rec["d_anchor"]) < 5 * tol
    assert rel_err(npy(dp), rec["d_positive"]) < 5 * tol
    # the fused tensor-core head on a graph tagged symmetric returns the symmetric part of the
    # reference's graph gradient (include/egm_b200.h, EGM_MHD_SYMMETRIC_GRAPH); every other
    # combination returns the reference's gradient itself
    d_G = rec["d_G"]
    if tagged and mode == "fp32" and head.isqrt_cov.num_iterations >= 2:
        d_G = 0.5 * (d_G + np.swapaxes(d_G, -1, -2))
        assert rel_err(npy(dG), npy(dG.transpose(-2, -1))) < 1e-4   # dW = V Zc^T rounds per element
    assert rel_err(npy(dG), d_G) < 5 * tol


@pytest.mark.parametrize("name", ["small_p2q2", "small_trainbn"])
def test_observed_graph_gets_the_reference_gradient(pkg, dev, name):
    """`G.retain_grad()` (or a tensor hook) on the tagged graph means somebody looks at dG
    itself - such a graph leaves the symmetric fast path and G.grad is the reference's dG, not its
    symmetric part, with the fast path switched on."""
    rec = golden(name)
    gpf, head = _build_from_golden(pkg, rec, dev)
    EF = pkg.functional
    a = torch.from_numpy(rec["anchor"]).float().to(dev).requires_grad_(True)
    p = torch.from_numpy(rec["positive"]).float().to(dev).requires_grad_(True)
    with EF.precision("fp32"):
        G = gpf(a, p)
        assert EF.graph_is_symmetric(G)
        G.retain_grad()
        assert not EF.graph_is_symmetric(G)
        out = head(a, G)
        (out * torch.from_numpy(rec["dOut"]).float().to(dev)).sum().backward()
    tol = 5e-3 * (20 if name == "small_trainbn" else 1)
    assert rel_err(npy(G.grad), rec["d_G"]) < tol
    assert rel_err(npy(a.grad), rec["d_anchor"]) < tol
    G2 = gpf(a, p)
    G2.register_hook(lambda g: g)
    assert not EF.graph_is_symmetric(G2)


@pytest.mark.parametrize("mode", MODES)
def test_config1_golden(pkg, dev, mode):
    """BASELINE config 1 (B=8, N=197, D=768 -> d_out=256, degree 3, 5 NS iterations): the reference's
    fp32 CPU output and gradients, modules initialised from the same seed as the reference."""
    rec = golden("cfg1_b8_n197_d768")
    B, N, D, P, Q, K, d_out = [int(v) for v in rec["cfg"][:7]]
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(P, Q)
    head = pkg.MomentHead(D, d_out, use_third_order=False, isqrt_iterations=K)
    head.eval()
    gpf, head = gpf.to(dev), head.to(dev)
    anchor, positive = make_inputs(B, N, D)
    assert abs(float(anchor.double().sum()) - rec["in_sum"][0]) < 1e-3   # same synthetic inputs
    a = anchor.to(dev).requires_grad_(True)
    p = positive.to(dev).requires_grad_(True)
    with pkg.functional.precision(mode):
        G = gpf(a, p)
        out = head(a, G)
        (out * torch.from_numpy(rec["dOut"]).to(dev)).sum().backward()
    to, tg = TOL_OUT[mode], TOL_GRAD[mode]
    to = max(to, 2e-5)   # the fixture itself is an fp32 computation
    tg = max(tg, 2e-4)
    assert rel_err(npy(G)[:, :6, :6], rec["G_probe"]) < to
    assert rel_err(npy(G).mean(axis=(1, 2)), rec["G_mean"]) < to
    assert rel_err(npy(out), rec["out"]) < to
    assert rel_err(npy(gpf.alpha_coeffs.grad), rec["d_alpha"]) < tg
    assert rel_err(npy(a.grad)[:, :4, :16], rec["d_anchor_probe"]) < tg
    assert rel_err(npy(p.grad)[:, :4, :16], rec["d_positive_probe"]) < tg
    assert rel_err(np.sqrt((npy(a.grad) ** 2).sum(axis=(1, 2))), rec["d_anchor_fro"]) < tg
    assert rel_err(npy(head.second_net[0].weight.grad)[:4, :32], rec["d_w_probe"]) < tg


def test_config1_train_mode_batchnorm(pkg, dev):
    """Same as above through train-mode BatchNorm1d, which amplifies upstream error ~50x on iid
    tokens: the fp32 mode still holds 1e-3 here... measured, gate at 5e-3."""
    rec = golden("cfg1_trainbn")
    B, N, D, P, Q, K, d_out = [int(v) for v in rec["cfg"][:7]]
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(P, Q).to(dev)
    head = pkg.MomentHead(D, d_out, use_third_order=False, isqrt_iterations=K)
    for m in head.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    head = head.to(dev).train()
    anchor, positive = make_inputs(B, N, D)
    with torch.no_grad(), pkg.functional.precision("fp32"):
        out = head(anchor.to(dev), gpf(anchor.to(dev), positive.to(dev)))
    assert rel_err(npy(out), rec["out"]) < 5e-3


def _run_config_golden(pkg, dev, name, mode, adaptive=None, patch_sketch=False):
    rec = golden(name)
    B, N, D, P, Q, K, d_out, third, S = [int(v) for v in rec["cfg"][:9]]
    torch.manual_seed(0)
    if adaptive:
        gpf = pkg.AdaptiveGraphPolynomialFusion(P, Q, adaptive_type=adaptive)
    else:
        gpf = pkg.GraphPolynomialFusion(P, Q)
    head = pkg.MomentHead(D, d_out, use_third_order=bool(third), isqrt_iterations=K, sketch_dim=S)
    if patch_sketch:
        with pytest.raises(RuntimeError, match="out of bounds"):     # the unpatched reference bug
            head.to(dev).eval()(torch.randn(1, N, D, device=dev), torch.rand(1, N, N, device=dev))
        head.tensor_sketch.sketch_dim = S        # same instance patch as the oracle
    head.eval()
    gpf, head = gpf.to(dev), head.to(dev)
    w = head.second_net[0].weight.detach().cpu().numpy()
    assert np.array_equal(w[:2, :8], rec["w_head"])                  # identical initialisation
    anchor, positive = make_inputs(B, N, D)
    a = anchor.to(dev).requires_grad_(True)
    p = positive.to(dev).requires_grad_(True)
    with pkg.functional.precision(mode):
        out = head(a, gpf(a, p))
        (out * torch.from_numpy(rec["dOut"]).to(dev)).sum().backward()
    to, tg = max(TOL_OUT[mode], 2e-5), max(TOL_GRAD[mode], 2e-4)
    assert out.shape == (B, d_out)
    assert rel_err(npy(out), rec["out"]) < to
    assert rel_err(npy(gpf.alpha_coeffs.grad), rec["d_alpha"]) < tg
    assert rel_err(npy(a.grad)[:, :4, :16], rec["d_anchor_probe"]) < tg
    assert rel_err(npy(p.grad)[:, :4, :16], rec["d_positive_probe"]) < tg
    assert rel_err(np.sqrt((npy(p.grad) ** 2).sum(axis=(1, 2))), rec["d_positive_fro"]) < tg
    assert rel_err(npy(head.second_net[0].weight.grad)[:4, :32], rec["d_w_probe"]) < tg
    return gpf, head


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_config3_third_order_sketch_8192(pkg, dev, mode):
    """BASELINE config 3: 3rd-order Tensor-Sketch (sketch_dim 8192 > 4*768) + 2nd-order iSQRT-COV."""
    _run_config_golden(pkg, dev, "cfg3_third_s8192", mode, patch_sketch=True)


@pytest.mark.parametrize("name,adaptive", [("cfg5_swin_p3q2", "global"), ("cfg5_swin_p1q1", "attention")])
def test_config5_swin_shape_adaptive_gpf(pkg, dev, name, adaptive):
    """BASELINE config 5: Swin-B 384px final-stage tokens (N=144, D=1024 -> d=256), adaptive GPF."""
    gpf, _ = _run_config_golden(pkg, dev, name, "fp32", adaptive=adaptive)
    if adaptive == "attention":   # registered but unused, exactly like the reference
        assert all(q.grad is None for q in gpf.coeff_attention.parameters())


def test_external_graph_golden_with_third_order(pkg, dev):
    rec = golden("extgraph")
    B, N, D, K, d_out, S = [int(v) for v in rec["cfg"]]
    head = pkg.MomentHead(D, d_out, use_third_order=True, isqrt_iterations=K, sketch_dim=S)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in params_of(rec).items()}
    head.load_state_dict({k: (v.float() if v.dtype == torch.float64 else v) for k, v in sd.items()})
    head = head.to(dev).eval()
    tokens = torch.from_numpy(rec["tokens"]).float().to(dev).requires_grad_(True)
    graph = torch.from_numpy(rec["graph"]).float().to(dev).requires_grad_(True)
    g_before = graph.detach().clone()
    with pkg.functional.precision("fp32_simt"):
        out = head(tokens, graph)
        (out * torch.from_numpy(rec["dOut"]).float().to(dev)).sum().backward()
    assert torch.equal(graph.detach(), g_before)            # inputs are never mutated
    assert rel_err(npy(out), rec["out"]) < 2e-4
    assert rel_err(npy(tokens.grad), rec["d_tokens"]) < 1e-3
    assert rel_err(npy(graph.grad), rec["d_graph"]) < 1e-3


def test_ops_helpers_golden(pkg, dev):
    import importlib
    ops = importlib.import_module("ego-moment-cle-vit_b200.utils.ops")
    rec = golden("ops")
    M = torch.from_numpy(rec["M"]).float().to(dev)
    graph = torch.from_numpy(rec["graph"]).float().to(dev)
    feats = torch.from_numpy(rec["feats"]).float().to(dev)
    with pkg.functional.precision("fp32_simt"):
        assert rel_err(npy(ops.matrix_sqrt_newton_schulz(M, 5, 1e-5)), rec["sqrt_ns"]) < 1e-4
        assert rel_err(npy(ops.cosine_similarity_matrix(feats)), rec["cos"]) < 1e-5
        assert rel_err(npy(ops.cosine_similarity_matrix(feats[0])), rec["cos2d"]) < 1e-5
    with pkg.functional.precision("fp32"):
        assert rel_err(npy(ops.matrix_sqrt_newton_schulz(M, 5, 1e-5)), rec["sqrt_ns"]) < 1e-3
        assert rel_err(npy(ops.cosine_similarity_matrix(feats)), rec["cos"]) < 1e-3
    assert np.array_equal(npy(ops.half_vectorize_symmetric(M)), rec["halfvec"].astype(np.float32).astype(np.float64))
    assert rel_err(npy(ops.normalize_graph(graph, "symmetric")), rec["norm_sym"]) < 1e-6
    assert rel_err(npy(ops.normalize_graph(graph, "random_walk")), rec["norm_rw"]) < 1e-6
    assert rel_err(npy(ops.batch_trace(M)), rec["trace"]) < 1e-6
    # the helpers stay differentiable, like the torch ops they replace
    g = graph.clone().requires_grad_(True)
    ops.normalize_graph(g, "symmetric").square().sum().backward()
    g64 = torch.from_numpy(rec["graph"]).requires_grad_(True)
    deg = g64.sum(-1).clamp(min=1e-8)
    (g64 * (1 / deg.sqrt()).unsqueeze(-1) * (1 / deg.sqrt()).unsqueeze(-2)).square().sum().backward()
    assert rel_err(npy(g.grad), g64.grad.numpy()) < 1e-5
    m = M.clone().requires_grad_(True)
    ops.batch_trace(m).sum().backward()
    assert torch.equal(m.grad, torch.eye(M.shape[-1], device=dev).expand_as(M))


# ----------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("B,N,D,P,Q,K,kind,sym", [
    (1, 3, 8, 0, 0, 1, "cosine", True),      # tiny, degree 0, one iteration
    (2, 7, 20, 1, 3, 2, "cosine", True),     # D not a multiple of 8 (padded planes)
    (3, 33, 50, 2, 2, 0, "dot", False),      # zero iterations, dot similarity, no symmetrise
    (2, 130, 260, 3, 1, 3, "cosine", True),  # ragged across the 128/256 tile edges
    (2, 9, 16, 5, 4, 4, "dot", True),        # higher degrees
])
@pytest.mark.parametrize("mode", ["fp32_simt", "fp32"])
def test_ragged_shapes_and_degenerate_settings(pkg, dev, B, N, D, P, Q, K, kind, sym, mode):
    EF = pkg.functional
    g = torch.Generator().manual_seed(B * 1000 + N)
    anchor = torch.randn(B, N, D, generator=g)
    positive = anchor + 0.5 * torch.randn(B, N, D, generator=g)
    if kind == "dot":
        anchor, positive = anchor / D ** 0.5, positive / D ** 0.5
    alpha = torch.rand(P + 1, Q + 1, generator=g) * 0.1
    dvec = torch.randn(B, D * (D + 1) // 2, generator=g)
    a64, p64, al64 = npy(anchor), npy(positive), npy(alpha)
    fw = O.gpf_forward(a64, p64, al64, kind, 1e-6, sym)
    st = O.moment_forward(a64, fw["G"], K, 1e-5)
    dZ, dG = O.moment_backward(a64, fw["G"], K, npy(dvec), 1e-5)
    da, dp, dal = O.gpf_backward(a64, p64, al64, dG, kind, 1e-6, sym)
    with EF.precision(mode):
        a = anchor.to(dev).requires_grad_(True)
        p = positive.to(dev).requires_grad_(True)
        al = alpha.to(dev).requires_grad_(True)
        G = EF.gpf_fused_graph(a, p, torch.nn.functional.softplus(al), cosine=(kind == "cosine"),
                               symmetric=sym)
        M2 = EF.graph_weighted_pool(a, G, eps=1e-5)
        vec = EF.half_vectorize(EF.newton_schulz(M2, K, 1e-5))
        (vec * dvec.to(dev)).sum().backward()
    tol = 2e-4 if mode == "fp32_simt" else 2e-3
    assert rel_err(npy(G), fw["G"]) < tol
    assert rel_err(npy(vec), st["vec"]) < tol
    assert rel_err(npy(a.grad), da + dZ) < 5 * tol
    assert rel_err(npy(p.grad), dp) < 5 * tol
    # degree (0,0) makes G a constant and d alpha exactly zero: compare with an absolute floor
    assert np.linalg.norm(npy(al.grad) - dal) < 5 * tol * (np.linalg.norm(dal) + 1e-5)


def test_zero_token_row_and_zero_degree_row_hit_the_clamps(pkg, dev):
    """F.normalize's eps clamp (gpf_kernel.py:87) and the degree clamp (moment_head.py:261)."""
    EF = pkg.functional
    B, N, D = 2, 6, 8
    g = torch.Generator().manual_seed(3)
    anchor = torch.randn(B, N, D, generator=g)
    anchor[0, 2] = 0.0                                   # ||x|| = 0 < eps
    positive = torch.randn(B, N, D, generator=g)
    alpha = torch.rand(3, 3, generator=g)
    graph = torch.rand(B, N, N, generator=g)
    graph[1, 4] = 0.0                                    # zero degree -> clamp at eps
    graph[0, 2, 3] = -0.05                               # a negative entry is legal input
    dG_up = torch.randn(B, N, N, generator=g)
    dvec = torch.randn(B, D * (D + 1) // 2, generator=g)
    fw = O.gpf_forward(npy(anchor), npy(positive), npy(alpha))
    da, dp, dal = O.gpf_backward(npy(anchor), npy(positive), npy(alpha), npy(dG_up))
    st = O.moment_forward(npy(anchor), npy(graph), 3, 1e-5)
    dZ, dGm = O.moment_backward(npy(anchor), npy(graph), 3, npy(dvec), 1e-5)
    with EF.precision("fp32_simt"):
        a = anchor.to(dev).requires_grad_(True)
        p = positive.to(dev).requires_grad_(True)
        G = EF.gpf_fused_graph(a, p, torch.nn.functional.softplus(alpha.to(dev)))
        (G * dG_up.to(dev)).sum().backward()
        assert rel_err(npy(G), fw["G"]) < 1e-5
        assert rel_err(npy(a.grad), da) < 1e-4 and rel_err(npy(p.grad), dp) < 1e-4
        z = anchor.to(dev).requires_grad_(True)
        gr = graph.to(dev).requires_grad_(True)
        vec = EF.half_vectorize(EF.newton_schulz(EF.graph_weighted_pool(z, gr, eps=1e-5), 3, 1e-5))
        (vec * dvec.to(dev)).sum().backward()
        assert rel_err(npy(vec), st["vec"]) < 1e-4
        assert rel_err(npy(z.grad), dZ) < 1e-3 and rel_err(npy(gr.grad), dGm) < 1e-3


def test_module_api_semantics(pkg, dev):
    """no_grad, eval/train, non-contiguous inputs, fresh outputs, dtype / shape errors."""
    torch.manual_seed(1)
    gpf = pkg.GraphPolynomialFusion(2, 2).to(dev)
    head = pkg.MomentHead(32, 16, use_third_order=True, isqrt_iterations=3, sketch_dim=64).to(dev)
    x = torch.randn(4, 10, 64, device=dev)
    a, p = x[:, :, :32], x[:, :, 32:]                      # non-contiguous views
    with torch.no_grad():
        G = gpf(a, p)
        out_eval = head.eval()(a, G)
        assert not G.requires_grad and G.shape == (4, 10, 10) and out_eval.shape == (4, 16)
        assert torch.equal(head(a, G), out_eval)            # deterministic (no atomics anywhere)
    assert torch.equal(gpf(a.contiguous(), p.contiguous()), G)
    head.train()
    o1 = head(a, G)
    assert o1.requires_grad and not torch.equal(o1, out_eval)   # BN batch stats + dropout
    with pytest.raises(RuntimeError, match="float32"):
        gpf(a.double(), p.double())
    with pytest.raises(RuntimeError, match="shapes differ"):
        gpf(a, p[:, :5])
    with pytest.raises(RuntimeError, match="does not match"):
        head(a, G[:, :5, :5])
    with pytest.raises(RuntimeError, match="out of bounds"):   # reference bug, kept
        pkg.MomentHead(8, 4, use_third_order=True, sketch_dim=64).to(dev)(
            torch.randn(2, 5, 8, device=dev), torch.rand(2, 5, 5, device=dev))
    # inference(): both views identical (ego_moment_clevit.py:318-331)
    Gs = gpf(a, a)
    assert torch.equal(Gs, Gs.transpose(-2, -1))


def test_training_step_updates_parameters(pkg, dev):
    """The wiring of ego_moment_clevit.py:156-159 + an optimiser step."""
    torch.manual_seed(2)
    gpf = pkg.GraphPolynomialFusion(2, 2).to(dev)
    head = pkg.MomentHead(64, 32, use_third_order=True, isqrt_iterations=5, sketch_dim=256).to(dev)
    clf = torch.nn.Linear(32, 5).to(dev)
    params = list(gpf.parameters()) + list(head.parameters()) + list(clf.parameters())
    opt = torch.optim.SGD(params, lr=0.1)
    a = torch.randn(8, 20, 64, device=dev)
    p = a + 0.5 * torch.randn_like(a)
    y = torch.randint(0, 5, (8,), device=dev)
    before = [q.detach().clone() for q in params]
    loss = torch.nn.functional.cross_entropy(clf(head(a, gpf(a, p))), y) + gpf.get_sparsity_loss()
    loss.backward()
    assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in params)
    opt.step()
    assert all(not torch.equal(b, q.detach()) for b, q in zip(before, params))


@pytest.mark.parametrize("M,K,N", [(256, 768 * 769 // 2, 256), (8, 768 * 769 // 2, 128), (5, 100, 7),
                                   (33, 1001, 40), (300, 4096, 300)])
def test_linear_on_the_tensor_core_engine(pkg, dev, M, K, N):
    """functional.linear == F.linear (second_net / third_net Linear, moment_head.py:187,196):
    forward with split-K, dx, dW, dbias."""
    EF = pkg.functional
    g = torch.Generator(device="cuda").manual_seed(M + K)
    x = torch.randn(M, K, device=dev, generator=g) * 0.02
    w = torch.randn(N, K, device=dev, generator=g) * (1.0 / K ** 0.5)
    b = torch.randn(N, device=dev, generator=g)
    dy = torch.randn(M, N, device=dev, generator=g)
    ref_in = [t.double().requires_grad_(True) for t in (x, w, b)]
    torch.nn.functional.linear(*ref_in).backward(dy.double())
    for mode, tol in (("fp32", 1e-4), ("bf16", 1e-2)):
        xs, ws_, bs = (t.clone().requires_grad_(True) for t in (x, w, b))
        with EF.precision(mode):
            y = EF.linear(xs, ws_, bs)
            y.backward(dy)
        yr = torch.nn.functional.linear(*[t.detach() for t in ref_in])
        assert rel_err(npy(y), npy(yr)) < tol
        assert rel_err(npy(xs.grad), npy(ref_in[0].grad)) < tol
        assert rel_err(npy(ws_.grad), npy(ref_in[1].grad)) < tol
        assert rel_err(npy(bs.grad), npy(ref_in[2].grad)) < 1e-5
    with EF.precision("fp32"):                      # no bias, no grad for x
        y2 = EF.linear(x, w.clone().requires_grad_(True))
        assert rel_err(npy(y2), npy(torch.nn.functional.linear(x.double(), w.double()))) < 1e-4


# ------------------------------------------------------- full-size, size-independent checks
def test_full_size_properties_config2(pkg, dev):
    """B=256, N=197, D=768 (BASELINE config 2): properties that need no oracle."""
    EF = pkg.functional
    B, N, D, K = 256, 197, 768, 5
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(B, N, D, device=dev, generator=g)
    p = a + 0.5 * torch.randn(B, N, D, device=dev, generator=g)
    coef = torch.nn.functional.softplus(torch.rand(4, 4, device=dev, generator=g) * 0.1)
    with torch.no_grad(), EF.precision("fp32"):
        G = EF.gpf_fused_graph(a, p, coef)
        assert torch.equal(G, G.transpose(-2, -1)) and float(G.min()) >= 0 and torch.isfinite(G).all()
        # shard independence: any slice of the batch gives bit-identical rows (what data
        # parallelism relies on)
        G2 = EF.gpf_fused_graph(a[100:132].contiguous(), p[100:132].contiguous(), coef)
        assert torch.equal(G2, G[100:132])
        M2 = EF.graph_weighted_pool(a, G, eps=1e-5)
        assert rel_err(npy(M2[:4]), npy(M2[:4].transpose(-2, -1))) < 1e-5   # W symmetric -> M2 symmetric
        isq = EF.newton_schulz(M2, K, 1e-5)
        assert torch.isfinite(isq).all()
        isq2 = EF.newton_schulz(M2[17:18].contiguous(), K, 1e-5)
        assert torch.equal(isq2, isq[17:18])
        v = EF.half_vectorize(isq)
        assert v.shape == (B, D * (D + 1) // 2)
        assert torch.equal(v[:, :D], isq[:, 0, :]) and torch.equal(v[:, -1], isq[:, -1, -1])
        # scaling law of the trace normalisation: isqrt(c M) = isqrt(M) / sqrt(c)   (eps-small)
        isq_c = EF.newton_schulz(4.0 * M2[:8].contiguous(), K, 1e-5)
        assert rel_err(npy(isq_c), 0.5 * npy(isq[:8])) < 1e-4


def test_newton_schulz_converges_to_inverse_square_root(pkg, dev):
    """With enough iterations on a well-conditioned SPD input, O M O = I."""
    EF = pkg.functional
    B, D = 4, 256
    g = torch.Generator().manual_seed(5)
    Qm, _ = torch.linalg.qr(torch.randn(B, D, D, generator=g))
    lam = 0.5 + torch.rand(B, D, generator=g)
    M = (Qm * lam.unsqueeze(1)) @ Qm.transpose(-2, -1)
    M = (0.5 * (M + M.transpose(-2, -1))).to(dev)
    eye = torch.eye(D, device=dev)
    for mode, tol in (("fp32_simt", 1e-4), ("fp32", 1e-3), ("bf16", 5e-2)):
        with torch.no_grad(), EF.precision(mode):
            Om = EF.newton_schulz(M, 14, 1e-7)
        res = Om @ M @ Om - eye
        assert float(res.norm() / eye.norm() / B ** 0.5) < tol, mode


def test_sketch_is_trilinear_and_triu_round_trips(pkg, dev):
    EF = pkg.functional
    ts = pkg.TensorSketch(96, 256).to(dev)
    x = torch.randn(5, 96, device=dev)
    with torch.no_grad():
        assert rel_err(npy(ts(2.0 * x)), 8.0 * npy(ts(x))) < 1e-5
        ref, _ = O.tensor_sketch(npy(x), [npy(ts.hash1).astype(int), npy(ts.hash2).astype(int), npy(ts.hash3).astype(int)],
                                 [npy(ts.sign1), npy(ts.sign2), npy(ts.sign3)], 256)
        assert rel_err(npy(ts(x)), ref) < 1e-5
    M = torch.randn(3, 37, 37, device=dev, requires_grad=True)
    v = EF.half_vectorize(M)
    assert np.array_equal(npy(v), O.half_vectorize(npy(M)))
    v.backward(torch.ones_like(v))
    assert torch.equal(M.grad, torch.triu(torch.ones(37, 37, device=dev)).expand(3, 37, 37))


# ------------------------------------------------------------------ graph alignment loss
def test_graph_alignment_loss_matches_reference_and_oracle(pkg, dev):
    """functional.graph_alignment_loss vs the reference's own _graph_alignment_loss (golden) and the
    oracle at the benchmark's batch size; patch_alignment_loss routes the reference method to it."""
    EF = pkg.functional
    rec = golden("align")
    for tag in ("a", "b"):
        G = torch.from_numpy(rec[f"{tag}_G"]).float().to(dev).requires_grad_(True)
        labels = torch.from_numpy(rec[f"{tag}_labels"]).to(dev)
        loss = EF.graph_alignment_loss(G, labels)
        (loss * 3.0).backward()
        assert abs(float(loss.detach()) - float(rec[f"{tag}_loss"])) < 2e-6
        assert rel_err(npy(G.grad), rec[f"{tag}_dG_x3"]) < 1e-4
    g = torch.Generator().manual_seed(3)
    B, N = 256, 197
    G = torch.rand(B, N, N, generator=g)
    labels = torch.randint(0, 80, (B,), generator=g)
    ref_loss, _, _ = O.graph_alignment_loss(npy(G), labels.numpy())
    ref_dG = O.graph_alignment_loss_backward(npy(G), labels.numpy())
    Gd = G.to(dev).requires_grad_(True)
    loss = EF.graph_alignment_loss(Gd, labels.to(dev))
    loss.backward()
    assert abs(float(loss.detach()) - ref_loss) < 1e-5 * max(1.0, ref_loss)
    assert rel_err(npy(Gd.grad), ref_dG) < 1e-4

    class Model:                      # stands in for the reference class: only the method is replaced
        def _graph_alignment_loss(self, fused_graph, labels):
            raise AssertionError("not patched")
    pkg.patch_alignment_loss(Model)
    assert abs(float(Model()._graph_alignment_loss(Gd.detach(), labels.to(dev))) - ref_loss) < 1e-5
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        EF.graph_alignment_loss(G, labels)


# ------------------------------------------------------------------------ CUDA graph capture
def test_whole_step_is_cuda_graph_capturable(pkg, dev):
    """The library never allocates or synchronises and takes the caller's stream, so a full
    GPF -> MomentHead forward + backward can be captured once and replayed on new inputs
    (what a launch-bound small-batch loop needs)."""
    EF = pkg.functional
    B, N, D = 4, 50, 128
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(3, 3).to(dev)
    head = pkg.MomentHead(D, 32, use_third_order=True, isqrt_iterations=5, sketch_dim=256).to(dev).eval()
    params = list(gpf.parameters()) + list(head.parameters())
    a = torch.zeros(B, N, D, device=dev, requires_grad=True)
    p = torch.zeros(B, N, D, device=dev, requires_grad=True)
    dout = torch.randn(B, 32, device=dev)
    xs = [tuple(t.to(dev) for t in make_inputs(B, N, D, seed=s)) for s in (1, 2, 3)]

    def step():
        for t in params + [a, p]:
            t.grad = None
        out = head(a, gpf(a, p))
        (out * dout).sum().backward()
        return out

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                 # warm-up off the capture (lazy init, CSR tables)
        with torch.no_grad():
            a.copy_(xs[0][0]); p.copy_(xs[0][1])
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_g = step()
    grads_g = [a.grad, p.grad, gpf.alpha_coeffs.grad, head.second_net[0].weight.grad]
    for xa, xp in xs[1:]:
        with torch.no_grad():
            a.copy_(xa); p.copy_(xp)
        graph.replay()
        torch.cuda.synchronize()
        got = [out_g.clone()] + [g.clone() for g in grads_g]
        out_e = step()                              # eager, same inputs
        want = [out_e, a.grad, p.grad, gpf.alpha_coeffs.grad, head.second_net[0].weight.grad]
        for x, y in zip(got, want):
            assert torch.equal(x, y)                # deterministic kernels: bit-identical


# ------------------------------------------------------------------- uninitialised-memory check
def _every_path(pkg, dev, B, N, D):
    """Outputs and gradients of every code path of the package on one seeded problem."""
    EF = pkg.functional
    res = []
    a0, p0 = make_inputs(B, N, D, seed=5)
    dout = torch.randn(B, 24, generator=torch.Generator().manual_seed(6)).to(dev)
    labels = torch.arange(B) % 2
    for mode, algo, fast in (("fp32", "dense", True), ("fp32", "dense", False), ("bf16", "dense", True),
                             ("fp32", "lowrank", True), ("fp32_simt", "dense", True)):
        torch.manual_seed(0)
        gpf = pkg.GraphPolynomialFusion(3, 2).to(dev)
        head = pkg.MomentHead(D, 24, use_third_order=True, isqrt_iterations=4, sketch_dim=64).to(dev).eval()
        EF.set_symmetric_fast_path(fast)
        EF.set_ns_algorithm(algo)
        try:
            with EF.precision(mode):
                a = a0.to(dev).requires_grad_(True)
                p = p0.to(dev).requires_grad_(True)
                G = gpf(a, p)
                out = head(a, G)
                loss = (out * dout).sum() + EF.graph_alignment_loss(G, labels.to(dev))
                loss.backward()
        finally:
            EF.set_symmetric_fast_path(True)
            EF.set_ns_algorithm("dense")
        res += [out.detach(), a.grad, p.grad, gpf.alpha_coeffs.grad, head.second_net[0].weight.grad,
                head.third_net[0].weight.grad]
    with torch.no_grad():
        M = torch.randn(B, D, D, generator=torch.Generator().manual_seed(7)).to(dev)
        M = M @ M.transpose(1, 2)
        res += [pkg.NewtonSchulzSqrtm(3)(M), EF.half_vectorize(M), EF.similarity_matrix(a0.to(dev))]
    return res


@pytest.mark.parametrize("shape", [(2, 50, 136), (1, 40, 520)])
def test_no_kernel_reads_memory_nobody_wrote(pkg, dev, shape):
    """EGM_POISON: every buffer the package allocates (outputs, saved state, scratch, the absent blocks
    of the symmetric storage) starts as 0xFF bytes = NaN. All kernels are deterministic, so the results
    must be bit-identical to the unpoisoned run - and finite."""
    EF = pkg.functional
    clean = _every_path(pkg, dev, *shape)
    EF.set_poison(True)
    try:
        poisoned = _every_path(pkg, dev, *shape)
    finally:
        EF.set_poison(False)
    for x, y in zip(clean, poisoned):
        assert torch.isfinite(y).all()
        assert torch.equal(x, y)


# ----------------------------------------------------------------------- the GEMM engine alone
def test_native_gemm_engine_cases(dev):
    """tests/native/test_gemm_tc: 28 cases of the tcgen05 engine against a double-precision host
    product (every operand major-ness, ragged edges, two-term products, addends, secondary output,
    <C,F>, packed triangle, symmetric block storage with NaN-poisoned absent blocks)."""
    import os
    import subprocess
    from conftest import ROOT
    native = os.path.join(ROOT, "tests", "native")
    exe = os.path.join(native, "test_gemm_tc")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", native], check=True, capture_output=True)
    r = subprocess.run([exe, "quick"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "PASSED: 0 failing" in r.stdout, r.stdout[-4000:]
    assert r.stdout.count("[ ok ]") >= 28
