"""Pin oracle/moment_oracle.py against fixtures produced by the reference itself
(tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest

from conftest import golden, params_of, rel_err
from oracle import moment_oracle as O

SMALL = ["small_p2q2", "small_p3q3_third", "small_dot_nosym", "small_trainbn"]


def _cfg(rec):
    B, N, D, P, Q, K, d_out, third, S, sym, train = [int(v) for v in rec["cfg"]]
    return dict(B=B, N=N, D=D, P=P, Q=Q, K=K, d_out=d_out, third=bool(third), S=S, sym=bool(sym),
                train=bool(train), kind=str(rec["similarity"]))


@pytest.mark.parametrize("name", SMALL)
def test_forward_taps_match_reference(name):
    rec = golden(name)
    c = _cfg(rec)
    fw = O.gpf_forward(rec["anchor"], rec["positive"], rec["alpha"], c["kind"], 1e-6, c["sym"])
    assert rel_err(fw["G"], rec["G"]) < 1e-12
    st = O.head_forward(rec["anchor"], fw["G"], params_of(rec), c["K"], c["third"], c["S"], c["train"])
    for key in ("W", "mu", "M2", "isqrt", "vec", "pre_bn", "out"):
        assert rel_err(st[key], rec[key]) < 1e-9, key
    if c["third"]:
        assert rel_err(st["u"], rec["u"]) < 1e-9
        assert rel_err(st["sketch"], rec["sketch"]) < 1e-9


@pytest.mark.parametrize("name", SMALL)
def test_backward_matches_reference_autograd(name):
    rec = golden(name)
    c = _cfg(rec)
    fw = O.gpf_forward(rec["anchor"], rec["positive"], rec["alpha"], c["kind"], 1e-6, c["sym"])
    # d vec -> d M2 through iSQRT-COV
    dO = O.half_vectorize_backward(rec["d_vec"], c["D"])
    dM = O.newton_schulz_backward(rec["M2"], dO, c["K"], 1e-5)
    assert rel_err(dM, rec["d_M2"]) < 1e-8
    # graph / token gradients of the pooling stage, driven by the reference's own d G
    # (d_G in the fixture is the total gradient of the real forward w.r.t. G)
    da, dp, dalpha = O.gpf_backward(rec["anchor"], rec["positive"], rec["alpha"], rec["d_G"], c["kind"],
                                    1e-6, c["sym"])
    assert rel_err(dalpha, rec["d_alpha"]) < 1e-8
    assert rel_err(dp, rec["d_positive"]) < 1e-8
    if not c["train"] and not c["third"]:
        # eval-mode BN is an affine map: d vec of the real forward equals the tapped d vec
        dZ, dG = O.moment_backward(rec["anchor"], fw["G"], c["K"], rec["d_vec"], 1e-5)
        assert rel_err(dG, rec["d_G"]) < 1e-8
        assert rel_err(dZ + da, rec["d_anchor"]) < 1e-8


def test_external_graph_with_third_order():
    rec = golden("extgraph")
    B, N, D, K, d_out, S = [int(v) for v in rec["cfg"]]
    prm = params_of(rec)
    st = O.head_forward(rec["tokens"], rec["graph"], prm, K, True, S, False)
    assert rel_err(st["out"], rec["out"]) < 1e-9
    # full backward incl. the sketch branch: build d vec / d sketch through the (affine, eval-mode)
    # feature nets analytically
    d2 = prm["second_net.0.weight"].shape[0]
    third = {"hashes": [prm[f"tensor_sketch.hash{k}"] for k in (1, 2, 3)],
             "signs": [prm[f"tensor_sketch.sign{k}"] for k in (1, 2, 3)], "sketch_dim": S}

    def net_bwd(x, prefix, dout):
        W = prm[f"{prefix}.0.weight"]
        pre = x @ W.T + prm[f"{prefix}.0.bias"]
        scale = prm[f"{prefix}.1.weight"] / np.sqrt(prm[f"{prefix}.1.running_var"] + 1e-5)
        bn = (pre - prm[f"{prefix}.1.running_mean"]) * scale + prm[f"{prefix}.1.bias"]
        from scipy.special import erf
        dgelu = 0.5 * (1 + erf(bn / np.sqrt(2))) + bn * np.exp(-bn * bn / 2) / np.sqrt(2 * np.pi)
        return (dout * dgelu * scale) @ W

    dvec = net_bwd(st["vec"], "second_net", rec["dOut"][:, :d2])
    dsk = net_bwd(st["sketch"], "third_net", rec["dOut"][:, d2:])
    dZ, dG = O.moment_backward(rec["tokens"], rec["graph"], K, dvec, 1e-5, third, dsk)
    assert rel_err(dZ, rec["d_tokens"]) < 1e-8
    assert rel_err(dG, rec["d_graph"]) < 1e-8


def test_ops_helpers():
    rec = golden("ops")
    assert rel_err(O.matrix_sqrt_newton_schulz(rec["M"], 5, 1e-5), rec["sqrt_ns"]) < 1e-10
    assert rel_err(O.half_vectorize(rec["M"]), rec["halfvec"]) == 0.0
    assert rel_err(O.normalize_graph(rec["graph"], "symmetric"), rec["norm_sym"]) < 1e-12
    assert rel_err(O.normalize_graph(rec["graph"], "random_walk"), rec["norm_rw"]) < 1e-12
    assert rel_err(O.batch_trace(rec["M"]), rec["trace"]) < 1e-12
    assert rel_err(O.cosine_similarity_matrix(rec["feats"]), rec["cos"]) < 1e-12
    assert rel_err(O.cosine_similarity_matrix(rec["feats"][0][None])[0], rec["cos2d"]) < 1e-12
    with pytest.raises(ValueError):
        O.normalize_graph(rec["graph"], "bogus")
    with pytest.raises(ValueError):
        O.similarity(rec["feats"], "bogus")


@pytest.mark.parametrize("tag,tol", [("a", 2e-6), ("b", 2e-6)])
def test_graph_alignment_loss_matches_reference(tag, tol):
    """ego_moment_clevit.py:278-316 run by the reference itself. The reference builds its B x B matrix
    with zeros_like(label_sim.float()), i.e. in float32 whatever the graph's dtype: fp32 tolerances."""
    rec = golden("align")
    G, labels = rec[f"{tag}_G"], rec[f"{tag}_labels"]
    loss, g, S = O.graph_alignment_loss(G, labels)
    assert abs(loss - float(rec[f"{tag}_loss"])) < tol * max(1.0, abs(loss))
    dG = O.graph_alignment_loss_backward(G, labels, dloss=3.0)
    assert rel_err(dG, rec[f"{tag}_dG_x3"]) < max(tol, 1e-9) * 10


@pytest.mark.parametrize("K", [2, 3, 5])
def test_symmetric_tangent_chain_is_the_symmetric_part_of_the_reverse_mode_gradient(K):
    """The identity the symmetric-graph fast path rests on (csrc/egm_api.cu, ns_tangent_bwd): for a
    symmetric M, Y_K is a fixed polynomial in A = M/(tr M + eps), its Frechet derivative is self-adjoint
    and commutes with transposition, so the symmetric part of the reference's reverse-mode dM equals the
    FORWARD tangent of the commuting chain along sym(dO) - every tangent being symmetric itself."""
    rng = np.random.default_rng(K)
    B, N, D, eps = 2, 10, 24, 1e-5
    Z = rng.standard_normal((B, N, D))
    M = np.einsum("bnd,bne->bde", Z, Z)                      # symmetric, rank N < D like Zc^T W Zc
    dO = O.half_vectorize_backward(rng.standard_normal((B, D * (D + 1) // 2)), D)   # upper-triangular
    dM_ref = O.newton_schulz_backward(M, dO, K, eps)         # the reference's gradient (not symmetric)
    tau = np.trace(M, axis1=1, axis2=2)
    inv, post = 1.0 / (tau + eps), (tau + eps) ** -0.5
    A = M * inv[:, None, None]
    I = np.eye(D)[None]
    # the commuting forward chain of the CUDA path: Y_1 = T_0, Z_1 = T_0 A, ...
    T = [1.5 * I - 0.5 * A]
    Y, Zs = {1: T[0]}, {1: T[0] @ A}
    for k in range(1, K):
        T.append(1.5 * I - 0.5 * Zs[k] @ Y[k])
        Y[k + 1] = Y[k] @ T[k]
        Zs[k + 1] = T[k] @ Zs[k]
    assert rel_err(post[:, None, None] * Y[K], O.newton_schulz(M, K, eps)) < 1e-12
    # the tangent chain along E = post * sym(dO)
    E = post[:, None, None] * 0.5 * (dO + np.swapaxes(dO, 1, 2))
    Yd = -0.5 * E
    Zd = -3.0 * Yd + Yd @ A + A @ Yd
    for k in range(1, K):
        Td = -0.5 * (Zd @ Y[k] + Zs[k] @ Yd)
        Yd, Zd = Yd @ T[k] + Y[k] @ Td, Td @ Zs[k] + T[k] @ Zd
        assert rel_err(Yd, np.swapaxes(Yd, 1, 2)) < 1e-12     # tangents stay symmetric
    dA = Yd
    dotO = np.einsum("bij,bij->b", dO, post[:, None, None] * Y[K])
    dotA = np.einsum("bij,bij->b", dA, A)
    dtau = (-0.5 * dotO - dotA) * inv
    dM = inv[:, None, None] * dA + dtau[:, None, None] * I
    assert rel_err(dM, 0.5 * (dM_ref + np.swapaxes(dM_ref, 1, 2))) < 1e-10


def test_torch_eager_comparator_matches_the_oracle():
    """baseline/torch_eager.py (bench.py's "PyTorch on B200" comparator) is the same function as the
    numpy oracle: forward taps and autograd gradients in float64 on the CPU."""
    import torch
    from baseline import torch_eager as TE
    rec = golden("small_p2q2")
    c = _cfg(rec)
    a = torch.from_numpy(rec["anchor"]).requires_grad_(True)
    p = torch.from_numpy(rec["positive"]).requires_grad_(True)
    alpha = torch.from_numpy(rec["alpha"]).requires_grad_(True)
    G = TE.gpf_forward(a, p, alpha)
    assert rel_err(G.detach().numpy(), rec["G"]) < 1e-12
    vec = TE.moment_vector(a, G, c["K"])
    assert rel_err(vec.detach().numpy(), rec["vec"]) < 1e-10
    (vec * torch.from_numpy(rec["d_vec"])).sum().backward()
    fw = O.gpf_forward(rec["anchor"], rec["positive"], rec["alpha"])
    dZ, dG = O.moment_backward(rec["anchor"], fw["G"], c["K"], rec["d_vec"], 1e-5)
    da, dp, dal = O.gpf_backward(rec["anchor"], rec["positive"], rec["alpha"], dG)
    assert rel_err(a.grad.numpy(), da + dZ) < 1e-8
    assert rel_err(p.grad.numpy(), dp) < 1e-8
    assert rel_err(alpha.grad.numpy(), dal) < 1e-8


def test_normalize_backward_folds_into_the_pair_gradient_matrix():
    """The identity behind the folded GPF backward (csrc/egm_kernels.cu, gpf_poly3_bwd_kernel FOLD): with
    xh = x / m, m = max(||x||, eps) (F.normalize, gpf_kernel.py:87), E = dR + dR^T and dxh = E xh,
        dx_i = (dxh_i - [n_i >= eps] xh_i <xh_i, dxh_i>) / m_i      and   <xh_i, dxh_i> = sum_j E_ij R_ij = s_i,
    hence dx = E'' X on the RAW tokens with E''_ij = (E_ij - [n_i >= eps] s_i delta_ij) / (m_i m_j) - no pass
    over [N, D] for the normalisation's backward. Checked against the oracle's own similarity backward
    (which follows the reference's autograd), including a row below the eps clamp."""
    rng = np.random.default_rng(3)
    B, N, D, eps = 2, 9, 14, 1e-6
    X = rng.standard_normal((B, N, D))
    X[0, 2] = 0.0                                   # ||x|| < eps: x / eps, no projection term
    dR = rng.standard_normal((B, N, N))
    n = np.linalg.norm(X, axis=-1)
    m = np.maximum(n, eps)
    Xh = X / m[..., None]
    R = Xh @ np.swapaxes(Xh, 1, 2)
    assert rel_err(R, O.similarity(X, "cosine", eps)) < 1e-12
    E = dR + np.swapaxes(dR, 1, 2)
    # two-step form (what rownorm_bwd used to do)
    dXh = E @ Xh
    s_ref = np.einsum("bnd,bnd->bn", Xh, dXh)
    gate = (n >= eps)[..., None]
    dX_two_step = (dXh - gate * Xh * s_ref[..., None]) / m[..., None]
    # folded form
    s = np.einsum("bij,bij->bi", E, R)
    assert rel_err(s, s_ref) < 1e-12
    E2 = E / (m[:, :, None] * m[:, None, :])
    idx = np.arange(N)
    E2[:, idx, idx] -= np.where(n >= eps, s / (m * m), 0.0)
    dX_folded = E2 @ X
    assert rel_err(dX_folded, dX_two_step) < 1e-12
    # and the two-step form IS the reference's gradient: drive the oracle's GPF backward with a polynomial
    # that is the identity in R_a (alpha so that only c_10 matters) ... simpler: autograd of the same graph
    import torch
    xt = torch.from_numpy(X).requires_grad_(True)
    xh = torch.nn.functional.normalize(xt, p=2, dim=-1, eps=eps)
    (torch.bmm(xh, xh.transpose(1, 2)) * torch.from_numpy(dR)).sum().backward()
    assert rel_err(dX_folded, xt.grad.numpy()) < 1e-10
