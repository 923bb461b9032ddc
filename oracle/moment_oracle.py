"""CPU oracle for the moment-pooling hot path - TEST INFRASTRUCTURE ONLY.

A plain-numpy restatement of the reference's algorithm (hibana2077/EGO-Moment-CLE-ViT), each
function citing the reference file:line it follows. It is the checker for the CUDA path: only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it. Nothing in the product package (`ego-moment-cle-vit_b200/`) imports
it, and the product has no CPU fallback.

Pinning: the reference's own tests hold no golden vectors, so the oracle
is pinned against outputs of the reference itself, generated in the build container by
`tests/golden/make_golden.py` (which imports the reference modules from /root/reference by file
path) and committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every stage
tap and every gradient of this file against those fixtures.

The forward functions follow the reference's operation order literally (including the 4K-GEMM
Newton-Schulz loop and the `bmm(W, ones)` broadcast of the third-order branch); the backward
functions are the reverse-mode derivatives of exactly those operations.
`dtype` defaults to float64 so that the oracle is a tighter reference than either fp32 side.
"""
from __future__ import annotations

import math

import numpy as np

try:  # exact (erf) GELU, as torch.nn.GELU() default
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover - scipy is in the image
    _erf = np.vectorize(math.erf)


def softplus(x):
    """torch.nn.functional.softplus (beta=1, threshold=20), used at gpf_kernel.py:133."""
    x = np.asarray(x)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-np.asarray(x)))


def _bT(x):
    return np.swapaxes(x, -1, -2)


_MM_BACKEND = "numpy"


def set_matmul_backend(name):
    """'numpy' (default) or 'torch': which CPU BLAS evaluates the batched products. bench.py's
    CPU legs select 'torch' so the port runs on the same threaded torch.bmm the reference itself
    calls (numpy loops over the batch single-matrix at a time)."""
    global _MM_BACKEND
    if name not in ("numpy", "torch"):
        raise ValueError(name)
    _MM_BACKEND = name


def _mm(a, b):
    if _MM_BACKEND == "torch" and a.ndim == 3 and b.ndim == 3 and a.dtype == b.dtype:
        import torch
        return torch.bmm(torch.from_numpy(a), torch.from_numpy(b)).numpy()   # strided views, no copy
    return a @ b


# ------------------------------------------------------------------------------ GPF
def l2_normalize(x, eps):
    """F.normalize(tokens, p=2, dim=-1, eps) = x / max(||x||, eps)  (gpf_kernel.py:87)."""
    n = np.sqrt(np.sum(x * x, axis=-1, keepdims=True))
    return x / np.maximum(n, eps), n


def similarity(tokens, kind="cosine", eps=1e-6):
    """GraphPolynomialFusion._compute_similarity  (gpf_kernel.py:75-94)."""
    if kind == "cosine":
        xn, _ = l2_normalize(tokens, eps)
        return _mm(xn, _bT(xn))
    if kind == "dot":
        return _mm(tokens, _bT(tokens))
    raise ValueError(f"Unknown similarity function: {kind}")


def hadamard_power(R, power):
    """_hadamard_power (gpf_kernel.py:96-115): 0 -> ones, 1 -> R (no clamp), k -> clamp(R,0)^k."""
    if power == 0:
        return np.ones_like(R)
    if power == 1:
        return R
    return np.maximum(R, 0.0) ** power


def _hadamard_dpower(R, power):
    if power == 0:
        return np.zeros_like(R)
    if power == 1:
        return np.ones_like(R)
    return power * np.maximum(R, 0.0) ** (power - 1)


def gpf_forward(tokens_anchor, tokens_positive, alpha, kind="cosine", eps=1e-6, symmetric=True,
                dtype=np.float64):
    """GraphPolynomialFusion.forward (gpf_kernel.py:117-159). Returns a dict of stage taps."""
    a = np.asarray(tokens_anchor, dtype)
    p = np.asarray(tokens_positive, dtype)
    alpha = np.asarray(alpha, dtype)
    P, Q = alpha.shape[0] - 1, alpha.shape[1] - 1
    Ra, Rp = similarity(a, kind, eps), similarity(p, kind, eps)
    c = softplus(alpha)
    F = np.zeros_like(Ra)
    for pp in range(P + 1):
        for qq in range(Q + 1):
            F = F + c[pp, qq] * (hadamard_power(Ra, pp) * hadamard_power(Rp, qq))
    S = 0.5 * (F + _bT(F)) if symmetric else F
    G = np.maximum(S, 0.0)
    return {"Ra": Ra, "Rp": Rp, "coef": c, "F": F, "S": S, "G": G}


def gpf_backward(tokens_anchor, tokens_positive, alpha, dG, kind="cosine", eps=1e-6, symmetric=True,
                 dtype=np.float64, fwd=None):
    """Reverse mode of gpf_forward. Returns (d tokens_anchor, d tokens_positive, d alpha).
    `fwd` = the gpf_forward output for the same inputs (reuse instead of recomputing)."""
    a = np.asarray(tokens_anchor, dtype)
    p = np.asarray(tokens_positive, dtype)
    alpha = np.asarray(alpha, dtype)
    dG = np.asarray(dG, dtype)
    P, Q = alpha.shape[0] - 1, alpha.shape[1] - 1
    fw = fwd if fwd is not None else gpf_forward(a, p, alpha, kind, eps, symmetric, dtype)
    Ra, Rp, c, S = fw["Ra"], fw["Rp"], fw["coef"], fw["S"]
    dS = dG * (S >= 0.0)                       # torch.clamp(min=0) passes the gradient at equality
    dF = 0.5 * (dS + _bT(dS)) if symmetric else dS
    dc = np.zeros_like(c)
    dRa = np.zeros_like(Ra)
    dRp = np.zeros_like(Rp)
    for pp in range(P + 1):
        for qq in range(Q + 1):
            fa, fb = hadamard_power(Ra, pp), hadamard_power(Rp, qq)
            dc[pp, qq] = np.sum(dF * fa * fb)
            dRa += dF * c[pp, qq] * _hadamard_dpower(Ra, pp) * fb
            dRp += dF * c[pp, qq] * fa * _hadamard_dpower(Rp, qq)
    dalpha = dc * sigmoid(alpha)

    def sim_bwd(x, dR):
        if kind == "dot":
            return _mm(dR + _bT(dR), x)
        xn, n = l2_normalize(x, eps)
        dxn = _mm(dR + _bT(dR), xn)
        live = n >= eps                         # clamp_min passes the gradient when ||x|| >= eps
        proj = np.sum(xn * dxn, axis=-1, keepdims=True)
        return np.where(live, (dxn - xn * proj) / np.maximum(n, eps), dxn / eps)

    return sim_bwd(a, dRa), sim_bwd(p, dRp), dalpha


# ---------------------------------------------------------------------- moment head
def normalize_weight_matrix(graph, eps):
    """MomentHead._normalize_weight_matrix (moment_head.py:246-266)."""
    deg = graph.sum(axis=-1)
    s = 1.0 / np.sqrt(np.maximum(deg, eps))
    return graph * s[..., :, None] * s[..., None, :], deg, s


def graph_weighted_mean(tokens, W, eps):
    """MomentHead._graph_weighted_mean (moment_head.py:222-244): normaliser is tr(W)."""
    w = W.sum(axis=-1)
    t = np.trace(W, axis1=-2, axis2=-1)
    mu = np.einsum("bnd,bn->bd", tokens, w) / (t[:, None] + eps)
    return mu, w, t


def newton_schulz(M, iters, eps, post="divide", keep=False):
    """NewtonSchulzSqrtm.forward (moment_head.py:28-70); post='multiply' is
    utils/ops.py:matrix_sqrt_newton_schulz (ops.py:122-165)."""
    B, D, _ = M.shape
    tr = np.trace(M, axis1=-2, axis2=-1)[:, None, None]
    A = M / (tr + eps)
    I = np.eye(D, dtype=M.dtype)
    Y = np.broadcast_to(I, M.shape).copy()
    Z = A
    Ys, Zs = [], []
    for _ in range(iters):
        if keep:
            Ys.append(Y)
            Zs.append(Z)
        ZY = _mm(Z, Y)
        YZ = _mm(Y, Z)
        I3 = 3.0 * I
        Y_new = 0.5 * _mm(Y, I3 - ZY)
        Z = 0.5 * _mm(I3 - YZ, Z)
        Y = Y_new
    r = np.sqrt(tr + eps)
    O = Y / r if post == "divide" else Y * r
    if keep:
        return O, (tr, A, Y, Ys, Zs)
    return O


def newton_schulz_backward(M, dO, iters, eps, post="divide", cache=None):
    """Reverse mode of newton_schulz. `cache` = the second return value of
    newton_schulz(..., keep=True), to reuse the forward's Y_k / Z_k as autograd would."""
    B, D, _ = M.shape
    if cache is None:
        _, cache = newton_schulz(M, iters, eps, post, keep=True)
    tr, A, YK, Ys, Zs = cache
    I = np.eye(D, dtype=M.dtype)
    te = tr + eps
    if post == "divide":
        dY = dO / np.sqrt(te)
        dtr = -0.5 * te ** -1.5 * np.sum(dO * YK, axis=(-1, -2), keepdims=True)
    else:
        dY = dO * np.sqrt(te)
        dtr = 0.5 * te ** -0.5 * np.sum(dO * YK, axis=(-1, -2), keepdims=True)
    dZ = np.zeros_like(M)
    for k in range(iters - 1, -1, -1):
        Y, Z = Ys[k], Zs[k]
        T1 = 3.0 * I - _mm(Z, Y)
        T2 = 3.0 * I - _mm(Y, Z)
        dT1 = 0.5 * _mm(_bT(Y), dY)
        dT2 = 0.5 * _mm(dZ, _bT(Z))
        dY_new = 0.5 * _mm(dY, _bT(T1)) - _mm(_bT(Z), dT1) - _mm(dT2, _bT(Z))
        dZ_new = 0.5 * _mm(_bT(T2), dZ) - _mm(dT1, _bT(Y)) - _mm(_bT(Y), dT2)
        dY, dZ = dY_new, dZ_new
    dA = dZ
    return dA / te + (dtr - np.sum(dA * M, axis=(-1, -2), keepdims=True) / te ** 2) * I


def triu_indices(D):
    """torch.triu_indices(D, D, offset=0) ordering: row-major (moment_head.py:215)."""
    return np.triu_indices(D)


def half_vectorize(M):
    """MomentHead._half_vectorize (moment_head.py:202-220) / ops.half_vectorize_symmetric."""
    r, c = triu_indices(M.shape[-1])
    return M[:, r, c]


def half_vectorize_backward(dv, D):
    r, c = triu_indices(D)
    dM = np.zeros((dv.shape[0], D, D), dtype=dv.dtype)
    dM[:, r, c] = dv
    return dM


def count_sketch(x, hash_idx, signs, sketch_dim):
    """TensorSketch._count_sketch (moment_head.py:100-112): scatter_add of sign*x."""
    out = np.zeros((x.shape[0], sketch_dim), dtype=x.dtype)
    np.add.at(out, (slice(None), np.asarray(hash_idx)), x * np.asarray(signs, dtype=x.dtype)[None, :])
    return out


def tensor_sketch(x, hashes, signs, sketch_dim):
    """TensorSketch.forward (moment_head.py:114-133): product of three count sketches (no FFT)."""
    cs = [count_sketch(x, hashes[k], signs[k], sketch_dim) for k in range(3)]
    return cs[0] * cs[1] * cs[2], cs


def tensor_sketch_backward(x, hashes, signs, sketch_dim, dout):
    _, cs = tensor_sketch(x, hashes, signs, sketch_dim)
    dx = np.zeros_like(x)
    for k in range(3):
        others = np.ones_like(cs[0])
        for j in range(3):
            if j != k:
                others = others * cs[j]
        dcs = dout * others
        dx += dcs[:, np.asarray(hashes[k])] * np.asarray(signs[k], dtype=x.dtype)[None, :]
    return dx


def moment_forward(tokens, graph, iters, eps=1e-5, third=None, dtype=np.float64, keep=False):
    """MomentHead.forward up to (and excluding) second_net / third_net (moment_head.py:268-317).
    `third` = dict(hashes[3,D], signs[3,D], sketch_dim) enables the third-order branch.
    keep=True also returns what autograd would save (for moment_backward(cache=...))."""
    Z = np.asarray(tokens, dtype)
    G = np.asarray(graph, dtype)
    W, deg, s = normalize_weight_matrix(G, eps)
    mu, w, t = graph_weighted_mean(Z, W, eps)
    Zc = Z - mu[:, None, :]
    U = _mm(W, Zc)
    M2 = _mm(_bT(Zc), U)
    if keep:
        isqrt, ns_cache = newton_schulz(M2, iters, eps, "divide", keep=True)
    else:
        isqrt, ns_cache = newton_schulz(M2, iters, eps, "divide"), None
    vec = half_vectorize(isqrt)
    out = {"W": W, "mu": mu, "Zc": Zc, "M2": M2, "isqrt": isqrt, "vec": vec}
    if keep:
        out["cache"] = {"deg": deg, "s": s, "w": w, "t": t, "ns": ns_cache}
    if third is not None:
        tw = W @ np.ones_like(Zc)                                    # moment_head.py:310
        u = (Zc * tw).sum(axis=1) / (t[:, None] + eps)                # moment_head.py:311
        sk, _ = tensor_sketch(u, third["hashes"], third["signs"], third["sketch_dim"])
        out.update({"u": u, "sketch": sk})
    return out


def moment_backward(tokens, graph, iters, dvec, eps=1e-5, third=None, dsketch=None, dtype=np.float64,
                    fwd=None, du=None):
    """Reverse mode of moment_forward: returns (d tokens, d graph) given d vec (and d sketch, or
    directly the gradient `du` of the third-order weighted mean u, moment_head.py:305-311).
    `fwd` = moment_forward(..., keep=True) output, to reuse saved tensors as autograd would."""
    Z = np.asarray(tokens, dtype)
    G = np.asarray(graph, dtype)
    B, N, D = Z.shape
    if fwd is not None:
        W, mu, Zc, M2 = fwd["W"], fwd["mu"], fwd["Zc"], fwd["M2"]
        deg, s, w, t, ns_cache = (fwd["cache"][k] for k in ("deg", "s", "w", "t", "ns"))
    else:
        W, deg, s = normalize_weight_matrix(G, eps)
        mu, w, t = graph_weighted_mean(Z, W, eps)
        Zc = Z - mu[:, None, :]
        M2 = _bT(Zc) @ (W @ Zc)
        ns_cache = None
    te = (t + eps)[:, None]
    dO = half_vectorize_backward(np.asarray(dvec, dtype), D)
    dM = newton_schulz_backward(M2, dO, iters, eps, "divide", cache=ns_cache)
    dZc = _mm(_mm(W, Zc), _bT(dM)) + _mm(_mm(_bT(W), Zc), dM)
    dW = _mm(_mm(Zc, dM), _bT(Zc))
    dw = np.zeros((B, N), dtype)
    dt = np.zeros((B,), dtype)
    if (third is not None and dsketch is not None) or du is not None:
        u = np.einsum("bnd,bn->bd", Zc, w) / te
        if du is None:
            du = tensor_sketch_backward(u, third["hashes"], third["signs"], third["sketch_dim"],
                                        np.asarray(dsketch, dtype))
        du = np.asarray(du, dtype)
        dZc = dZc + w[:, :, None] * du[:, None, :] / te[:, :, None]
        dw = dw + np.einsum("bnd,bd->bn", Zc, du) / te
        dt = dt - np.sum(u * du, axis=-1) / te[:, 0]
    dZ = dZc.copy()
    dmu = -dZc.sum(axis=1)
    dZ = dZ + w[:, :, None] * dmu[:, None, :] / te[:, :, None]
    dw = dw + np.einsum("bnd,bd->bn", Z, dmu) / te
    dt = dt - np.sum(mu * dmu, axis=-1) / te[:, 0]
    dW = dW + dw[:, :, None]
    idx = np.arange(N)
    dW[:, idx, idx] += dt[:, None]
    dG = s[:, :, None] * dW * s[:, None, :]
    ds = np.sum(dW * G * s[:, None, :], axis=2) + np.sum(dW * G * s[:, :, None], axis=1)
    dtil = np.maximum(deg, eps)
    ddeg = -0.5 * dtil ** -1.5 * ds * (deg >= eps)
    dG = dG + ddeg[:, :, None]
    return dZ, dG


# ---------------------------------------------------------------- second_net / third_net
def gelu(x):
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def feature_net(v, params, prefix, train, bn_eps=1e-5):
    """nn.Sequential(Linear, BatchNorm1d, GELU, Dropout) of moment_head.py:186-200 with dropout
    inactive (p = 0 or eval). `params` is a state_dict-like mapping of numpy arrays."""
    Wt = np.asarray(params[f"{prefix}.0.weight"])
    pre = v @ Wt.T.astype(v.dtype) + np.asarray(params[f"{prefix}.0.bias"], v.dtype)
    if train:
        mean, var = pre.mean(axis=0), pre.var(axis=0)
    else:
        mean = np.asarray(params[f"{prefix}.1.running_mean"], v.dtype)
        var = np.asarray(params[f"{prefix}.1.running_var"], v.dtype)
    bn = (pre - mean) / np.sqrt(var + bn_eps) * np.asarray(params[f"{prefix}.1.weight"], v.dtype) \
        + np.asarray(params[f"{prefix}.1.bias"], v.dtype)
    return gelu(bn), pre


def head_forward(tokens, graph, params, iters, use_third=False, sketch_dim=None, train=False,
                 eps=1e-5, dtype=np.float64):
    """Full MomentHead.forward (moment_head.py:268-322), dropout inactive."""
    third = None
    if use_third:
        third = {
            "hashes": [np.asarray(params[f"tensor_sketch.hash{k}"]) for k in (1, 2, 3)],
            "signs": [np.asarray(params[f"tensor_sketch.sign{k}"]) for k in (1, 2, 3)],
            "sketch_dim": sketch_dim,
        }
    st = moment_forward(tokens, graph, iters, eps, third, dtype)
    second, pre2 = feature_net(st["vec"], params, "second_net", train)
    st["pre_bn"] = pre2
    feats = [second]
    if use_third:
        thirdf, _ = feature_net(st["sketch"], params, "third_net", train)
        feats.append(thirdf)
    st["out"] = np.concatenate(feats, axis=-1)
    return st


# ----------------------------------------------------------------------- utils/ops.py
def matrix_sqrt_newton_schulz(M, iters=5, eps=1e-5, dtype=np.float64):
    """utils/ops.py:122-165 (post-multiplies by sqrt(tr+eps))."""
    return newton_schulz(np.asarray(M, dtype), iters, eps, "multiply")


def normalize_graph(graph, method="symmetric", dtype=np.float64):
    """utils/ops.py:238-271 (eps = 1e-8, 1/sqrt)."""
    G = np.asarray(graph, dtype)
    if method == "none":
        return G
    deg = np.maximum(G.sum(axis=-1), 1e-8)
    if method == "symmetric":
        s = 1.0 / np.sqrt(deg)
        return G * s[..., :, None] * s[..., None, :]
    if method == "random_walk":
        return G * (1.0 / deg)[..., :, None]
    raise ValueError(f"Unknown normalization method: {method}")


def cosine_similarity_matrix(features, eps=1e-8, dtype=np.float64):
    """utils/ops.py:355-381."""
    return similarity(np.asarray(features, dtype), "cosine", eps)


def batch_trace(M):
    """utils/ops.py:316-326."""
    return np.trace(M, axis1=-2, axis2=-1)


# ------------------------------------------------------------- graph alignment loss
def graph_alignment_loss(fused_graph, labels, dtype=np.float64):
    """EGOMomentCLEViT._graph_alignment_loss (ego_moment_clevit.py:278-316): g = mean(G) per image,
    S_ij = sigmoid(g_i g_j) (the reference fills it with a B x B Python loop), loss =
    mse_loss(S, [labels_i == labels_j]). Returns (loss, g, S)."""
    G = np.asarray(fused_graph, dtype)
    labels = np.asarray(labels)
    label_sim = (labels[None, :] == labels[:, None]).astype(dtype)
    g = G.mean(axis=(1, 2))
    S = sigmoid(g[:, None] * g[None, :])
    return ((S - label_sim) ** 2).mean(), g, S


def graph_alignment_loss_backward(fused_graph, labels, dloss=1.0, dtype=np.float64):
    """d loss / d fused_graph: dS = 2 (S - L)/B^2, d(g g^T) = dS S (1 - S), dg = (X + X^T) g,
    dG[b] = dg_b / N^2 broadcast over the image's N x N entries."""
    G = np.asarray(fused_graph, dtype)
    B, N = G.shape[0], G.shape[1]
    labels = np.asarray(labels)
    label_sim = (labels[None, :] == labels[:, None]).astype(dtype)
    _, g, S = graph_alignment_loss(G, labels, dtype)
    X = dloss * 2.0 * (S - label_sim) / (B * B) * S * (1.0 - S)
    dg = (X + X.T) @ g
    return np.broadcast_to((dg / (N * G.shape[2]))[:, None, None], G.shape).copy()


# ------------------------------------------------- whole-path step (bench.py cpu baseline)
def feature_net_backward(v, params, prefix, dout, train, bn_eps=1e-5):
    """d/d v and d/d weight of Linear -> BatchNorm1d -> GELU (dropout inactive)."""
    Wt = np.asarray(params[f"{prefix}.0.weight"]).astype(v.dtype, copy=False)
    gamma = np.asarray(params[f"{prefix}.1.weight"], v.dtype)
    pre = v @ Wt.T + np.asarray(params[f"{prefix}.0.bias"], v.dtype)
    if train:
        mean, var = pre.mean(axis=0), pre.var(axis=0)
    else:
        mean = np.asarray(params[f"{prefix}.1.running_mean"], v.dtype)
        var = np.asarray(params[f"{prefix}.1.running_var"], v.dtype)
    rstd = 1.0 / np.sqrt(var + bn_eps)
    xhat = (pre - mean) * rstd
    bn = xhat * gamma + np.asarray(params[f"{prefix}.1.bias"], v.dtype)
    dbn = dout * (0.5 * (1.0 + _erf(bn / math.sqrt(2.0))) + bn * np.exp(-0.5 * bn * bn) / math.sqrt(2.0 * math.pi))
    dxhat = dbn * gamma
    if train:
        n = pre.shape[0]
        dpre = rstd / n * (n * dxhat - dxhat.sum(axis=0) - xhat * (dxhat * xhat).sum(axis=0))
    else:
        dpre = dxhat * rstd
    return dpre @ Wt, dpre.T @ v


def path_step(anchor, positive, alpha, params, iters, dout, train=True, dtype=np.float32):
    """One forward+backward of GPF -> MomentHead (2nd order) with the reference's operation
    count: forward intermediates are kept and reused by the backward exactly as autograd would
    (20 + 40 Newton-Schulz products). Returns (out, d anchor, d positive, d alpha, d weight)."""
    a = np.asarray(anchor, dtype)
    p = np.asarray(positive, dtype)
    al = np.asarray(alpha, dtype)
    fw = gpf_forward(a, p, al, dtype=dtype)
    st = moment_forward(a, fw["G"], iters, 1e-5, None, dtype, keep=True)
    out, _ = feature_net(st["vec"], params, "second_net", train)
    dvec, dWt = feature_net_backward(st["vec"], params, "second_net", np.asarray(dout, dtype), train)
    dZ, dG = moment_backward(a, fw["G"], iters, dvec, 1e-5, dtype=dtype, fwd=st)
    da, dp, dal = gpf_backward(a, p, al, dG, dtype=dtype, fwd=fw)
    return out, da + dZ, dp, dal, dWt
