"""PyTorch-eager restatement of the reference algorithm - BENCHMARK BASELINE ONLY (not the oracle, not the product).

The "PyTorch on B200" comparator of SURVEY.md 8(d): the same arithmetic as the reference modules
(gpf_kernel.py:75-159, moment_head.py:28-70 and :222-300) written as plain torch ops with autograd,
so `bench.py` can time what the stock framework does with this algorithm on the same GPU (cuBLAS
`torch.bmm`, ~120 launches for GPF, 4K matrix products forward and 8K backward for Newton-Schulz).
The reference tree itself is not present on the GPU box. Pinned against `oracle/moment_oracle.py` (which is
pinned against the reference's own goldens) by tests/test_oracle_golden.py. Nothing in the product
package imports this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gpf_forward(tokens_anchor, tokens_positive, alpha, eps: float = 1e-6, symmetric: bool = True):
    """GraphPolynomialFusion.forward, cosine similarity (gpf_kernel.py:75-159)."""
    def sim(x):
        xn = F.normalize(x, p=2, dim=-1, eps=eps)
        return torch.bmm(xn, xn.transpose(-2, -1))

    def power(R, k):                      # gpf_kernel.py:96-115
        if k == 0:
            return torch.ones_like(R)
        if k == 1:
            return R
        return torch.pow(torch.clamp(R, min=0.0), k)

    Ra, Rp = sim(tokens_anchor), sim(tokens_positive)
    coef = F.softplus(alpha)
    fused = torch.zeros_like(Ra)
    for p in range(alpha.shape[0]):
        for q in range(alpha.shape[1]):
            fused = fused + coef[p, q] * (power(Ra, p) * power(Rp, q))
    if symmetric:
        fused = 0.5 * (fused + fused.transpose(-2, -1))
    return torch.clamp(fused, min=0.0)


def newton_schulz(M, iters: int, eps: float = 1e-5):
    """NewtonSchulzSqrtm.forward (moment_head.py:28-70): 4 matrix products per iteration."""
    D = M.shape[-1]
    tr = torch.diagonal(M, dim1=-2, dim2=-1).sum(-1, keepdim=True).unsqueeze(-1)
    A = M / (tr + eps)
    I = torch.eye(D, device=M.device, dtype=M.dtype).unsqueeze(0).expand_as(M)
    Y, Z = I, A
    for _ in range(iters):
        ZY = torch.bmm(Z, Y)
        YZ = torch.bmm(Y, Z)
        Y_new = 0.5 * torch.bmm(Y, 3.0 * I - ZY)
        Z = 0.5 * torch.bmm(3.0 * I - YZ, Z)
        Y = Y_new
    return Y / torch.sqrt(tr + eps)


def moment_vector(tokens, graph, iters: int, eps: float = 1e-5):
    """MomentHead.forward up to the half-vector (moment_head.py:222-300, 2nd-order branch)."""
    deg = graph.sum(dim=-1)
    s = torch.rsqrt(torch.clamp(deg, min=eps))
    W = graph * s.unsqueeze(-1) * s.unsqueeze(-2)
    w = W.sum(dim=-1)
    t = torch.diagonal(W, dim1=-2, dim2=-1).sum(-1, keepdim=True)
    mu = torch.bmm(w.unsqueeze(1), tokens).squeeze(1) / (t + eps)
    Zc = tokens - mu.unsqueeze(1)
    M2 = torch.bmm(Zc.transpose(-2, -1), torch.bmm(W, Zc))
    O = newton_schulz(M2, iters, eps)
    D = O.shape[-1]
    iu = torch.triu_indices(D, D, device=O.device)
    return O[:, iu[0], iu[1]]


def head_forward(tokens, graph, second_net: torch.nn.Module, iters: int, eps: float = 1e-5):
    """... followed by second_net = Linear -> BatchNorm1d -> GELU -> Dropout (moment_head.py:186-200)."""
    return second_net(moment_vector(tokens, graph, iters, eps))
