"""Graph Polynomial Fusion - drop-in for the reference's `src/models/gpf_kernel.py`.

Same classes, constructor arguments, attributes, parameter names and forward signatures
(reference: gpf_kernel.py:15-217); the arithmetic runs in the sm_100a library: token
normalisation -> two Gram matrices on tcgen05 -> one fused pass for the Hadamard-power polynomial,
symmetrisation and clamp, with a hand-written backward. There is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as EF


class GraphPolynomialFusion(nn.Module):
    """G = clamp(sym(sum_{p,q} softplus(alpha)[p,q] * R_a^{o p} * R_p^{o q}), min=0).

    Args mirror the reference (gpf_kernel.py:32-40): degree_p, degree_q, similarity in
    {'cosine','dot'}, eps, symmetric_enforce, coeff_init in {'uniform','xavier','identity'}.
    Hadamard powers follow gpf_kernel.py:96-115: power 0 -> ones, power 1 -> R (not clamped),
    power k>=2 -> clamp(R, 0)^k.
    """

    def __init__(self, degree_p: int = 2, degree_q: int = 2, similarity: str = 'cosine',
                 eps: float = 1e-6, symmetric_enforce: bool = True, coeff_init: str = 'uniform'):
        super().__init__()
        self.degree_p = degree_p
        self.degree_q = degree_q
        self.similarity = similarity
        self.eps = eps
        self.symmetric_enforce = symmetric_enforce
        self.num_terms = (degree_p + 1) * (degree_q + 1)
        # A_pq = softplus(alpha_pq) >= 0 keeps the fused kernel PSD
        self.alpha_coeffs = nn.Parameter(torch.zeros(degree_p + 1, degree_q + 1))
        self._init_coefficients(coeff_init)

    def _init_coefficients(self, init_method: str):
        if init_method == 'uniform':
            nn.init.uniform_(self.alpha_coeffs, 0.0, 0.1)
        elif init_method == 'xavier':
            nn.init.xavier_uniform_(self.alpha_coeffs)
        elif init_method == 'identity':
            self.alpha_coeffs.data.fill_(0.01)
            if self.degree_p >= 0 and self.degree_q >= 0:
                self.alpha_coeffs.data[0, 0] = 0.5
            if self.degree_p >= 1 and self.degree_q >= 1:
                self.alpha_coeffs.data[1, 1] = 0.5
        else:
            raise ValueError(f"Unknown initialization method: {init_method}")

    def _check_similarity(self) -> bool:
        if self.similarity == 'cosine':
            return True
        if self.similarity == 'dot':
            return False
        raise ValueError(f"Unknown similarity function: {self.similarity}")

    def _compute_similarity(self, tokens: torch.Tensor) -> torch.Tensor:
        """[B,N,D] -> [B,N,N] cosine or dot Gram matrix (gpf_kernel.py:75-94)."""
        return EF.similarity_matrix(tokens, cosine=self._check_similarity(), eps=self.eps)

    def _hadamard_power(self, matrix: torch.Tensor, power: int) -> torch.Tensor:
        """Element-wise power with the reference's clamping rule (gpf_kernel.py:96-115)."""
        if power == 0:
            return torch.ones_like(matrix)
        if power == 1:
            return matrix
        return torch.pow(torch.clamp(matrix, min=0.0), power)

    def forward(self, tokens_anchor: torch.Tensor, tokens_positive: torch.Tensor) -> torch.Tensor:
        """tokens_* [B,N,D] -> fused relation graph [B,N,N] (fp32, freshly allocated)."""
        cosine = self._check_similarity()
        coef = F.softplus(self.alpha_coeffs)
        return EF.gpf_fused_graph(tokens_anchor, tokens_positive, coef, cosine=cosine, eps=self.eps,
                                  symmetric=bool(self.symmetric_enforce))

    def get_coefficient_matrix(self) -> torch.Tensor:
        """A_pq = softplus(alpha_pq)  (gpf_kernel.py:161-163)."""
        return F.softplus(self.alpha_coeffs)

    def get_sparsity_loss(self, lambda_sparse: float = 0.01) -> torch.Tensor:
        """L1 penalty on the non-negative coefficients (gpf_kernel.py:165-176)."""
        coeffs = F.softplus(self.alpha_coeffs)
        return lambda_sparse * torch.sum(torch.abs(coeffs))


# the north star's name for the same module
GPFKernel = GraphPolynomialFusion


class AdaptiveGraphPolynomialFusion(GraphPolynomialFusion):
    """Reference subclass (gpf_kernel.py:179-217): extra `adaptive_type`; for 'attention' an
    (unused) `coeff_attention` MultiheadAttention is registered so state_dicts line up; every
    adaptive_type evaluates the base forward, exactly as the reference does."""

    def __init__(self, degree_p: int = 2, degree_q: int = 2, similarity: str = 'cosine',
                 eps: float = 1e-6, symmetric_enforce: bool = True, coeff_init: str = 'uniform',
                 adaptive_type: str = 'global'):
        super().__init__(degree_p, degree_q, similarity, eps, symmetric_enforce, coeff_init)
        self.adaptive_type = adaptive_type
        if adaptive_type == 'attention':
            self.coeff_attention = nn.MultiheadAttention(embed_dim=self.num_terms, num_heads=1,
                                                         batch_first=True)

    def forward(self, tokens_anchor: torch.Tensor, tokens_positive: torch.Tensor) -> torch.Tensor:
        return super().forward(tokens_anchor, tokens_positive)


def test_gpf():
    """Smoke test mirroring the reference's module-level test (needs a B200)."""
    dev = torch.device('cuda')
    tokens_anchor = torch.randn(2, 196, 768, device=dev)
    tokens_positive = torch.randn(2, 196, 768, device=dev)
    gpf = GraphPolynomialFusion(degree_p=2, degree_q=2, similarity='cosine').to(dev)
    with torch.no_grad():
        g = gpf(tokens_anchor, tokens_positive)
    print(f"Fused graph shape: {tuple(g.shape)} range [{g.min().item():.4f}, {g.max().item():.4f}]")
    print(f"Is symmetric: {torch.allclose(g, g.transpose(-2, -1), atol=1e-6)}")
    print(f"Coefficients:\n{gpf.get_coefficient_matrix()}")


if __name__ == "__main__":
    test_gpf()
