"""Moment head - drop-in for the reference's `src/models/moment_head.py`.

`NewtonSchulzSqrtm`, `TensorSketch` and `MomentHead` keep the reference's constructor arguments,
attributes, submodule order and state_dict keys (moment_head.py:15-322); the pooling, iSQRT-COV,
half-vectorisation and count-sketch arithmetic run in the sm_100a library with hand-written
backward passes. `second_net` / `third_net` stay `nn.Sequential` (state_dict compatibility).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as EF


class NewtonSchulzSqrtm(nn.Module):
    """Trace-normalised coupled Newton-Schulz iteration (iSQRT-COV), moment_head.py:15-70.

    Y0 = I, Z0 = M/(tr M + eps); K x {Y <- 1/2 Y(3I - ZY); Z <- 1/2 (3I - YZ) Z};
    returns Y_K / sqrt(tr M + eps). Does not mutate its input."""

    def __init__(self, num_iterations: int = 3, eps: float = 1e-5):
        super().__init__()
        self.num_iterations = num_iterations
        self.eps = eps

    def forward(self, matrix: torch.Tensor) -> torch.Tensor:
        return EF.newton_schulz(matrix, self.num_iterations, self.eps, post="divide")


class TensorSketch(nn.Module):
    """Product of three count-sketches (moment_head.py:73-133).

    Keeps the reference's observable behaviour, including `torch.manual_seed(seed)` in the
    constructor (global RNG side effect, moment_head.py:88), the int64 `hash1..3` / `sign1..3`
    buffers drawn in the reference's order, the `min(sketch_dim, 4*input_dim)` cap on
    `self.sketch_dim` and the resulting out-of-bounds failure when sketch_dim > 4*input_dim."""

    def __init__(self, input_dim: int, sketch_dim: int = 2048, seed: int = 42):
        super().__init__()
        self.input_dim = input_dim
        self.sketch_dim = min(sketch_dim, input_dim * 4)
        torch.manual_seed(seed)
        self.register_buffer('hash1', torch.randint(0, sketch_dim, (input_dim,)))
        self.register_buffer('hash2', torch.randint(0, sketch_dim, (input_dim,)))
        self.register_buffer('hash3', torch.randint(0, sketch_dim, (input_dim,)))
        self.register_buffer('sign1', torch.randint(0, 2, (input_dim,)) * 2 - 1)
        self.register_buffer('sign2', torch.randint(0, 2, (input_dim,)) * 2 - 1)
        self.register_buffer('sign3', torch.randint(0, 2, (input_dim,)) * 2 - 1)
        self._csr_key = None
        self._csr = None

    def _tables(self):
        """Stacked [3,D] hash/sign buffers and the CSR inverse map, rebuilt when buffers change."""
        bufs = (self.hash1, self.hash2, self.hash3, self.sign1, self.sign2, self.sign3)
        key = tuple((b.data_ptr(), b._version, str(b.device)) for b in bufs) + (self.sketch_dim,)
        if key != self._csr_key:
            hashes = torch.stack(bufs[:3]).to(torch.int64).contiguous()
            signs = torch.stack(bufs[3:]).to(torch.int64).contiguous()
            csr = EF.build_sketch_csr(hashes, signs, self.sketch_dim)
            self._csr = (hashes, signs, csr)
            self._csr_key = key
        return self._csr

    def _count_sketch(self, x: torch.Tensor, hash_idx: torch.Tensor, signs: torch.Tensor) -> torch.Tensor:
        """Single count sketch [B,D] -> [B,sketch_dim] (moment_head.py:100-112)."""
        # private helper nobody on the path calls (forward uses the fused 3-sketch kernel)
        sketched = torch.zeros(x.shape[0], self.sketch_dim, device=x.device, dtype=x.dtype)
        sketched.scatter_add_(1, hash_idx.unsqueeze(0).expand(x.shape[0], -1), x * signs.unsqueeze(0))
        return sketched

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        hashes, signs, csr = self._tables()
        return EF.tensor_sketch(x, hashes, signs, csr, self.sketch_dim)


class MomentHead(nn.Module):
    """Graph-weighted 2nd (+ optional 3rd) order moment pooling head (moment_head.py:136-322).

    forward(tokens [B,N,D], graph [B,N,N]) -> [B,d_out]. `graph` may be any real matrix.

    Gradient contract. Outputs, token gradients and parameter gradients are the reference's in every
    case. When `graph` is the (exactly symmetric, tagged) output of GraphPolynomialFusion with
    symmetric_enforce=True the head takes the symmetric fast path (upper tiles only, tangent-form
    backward) and the gradient returned FOR THE GRAPH is the symmetric part (dG + dG^T)/2 of the
    reference's dG - the gradient with respect to a symmetric matrix, and exactly what the
    symmetrisation in GraphPolynomialFusion's backward lets through. A graph that is untagged,
    edited in place, produced with symmetric_enforce=False, or observed (`graph.retain_grad()`, a
    tensor hook) takes the general path and receives the reference's dG itself;
    `functional.set_symmetric_fast_path(False)` turns the fast path off globally."""

    def __init__(self, d_in: int, d_out: int = 512, use_third_order: bool = False,
                 isqrt_iterations: int = 3, sketch_dim: int = 2048, eps: float = 1e-5):
        super().__init__()
        self.d_in = d_in
        self.d_out = d_out
        self.use_third_order = use_third_order
        self.eps = eps
        self.isqrt_cov = NewtonSchulzSqrtm(num_iterations=isqrt_iterations, eps=eps)
        if use_third_order:
            self.tensor_sketch = TensorSketch(d_in, sketch_dim)
        if use_third_order:
            self.d_second = d_out // 2
            self.d_third = d_out - self.d_second
        else:
            self.d_second = d_out
            self.d_third = 0
        second_input_dim = (d_in * (d_in + 1)) // 2
        self.second_net = nn.Sequential(
            nn.Linear(second_input_dim, self.d_second),
            nn.BatchNorm1d(self.d_second),
            nn.GELU(),
            nn.Dropout(0.1),
        )
        if use_third_order:
            self.third_net = nn.Sequential(
                nn.Linear(sketch_dim, self.d_third),
                nn.BatchNorm1d(self.d_third),
                nn.GELU(),
                nn.Dropout(0.1),
            )

    def _half_vectorize(self, matrix: torch.Tensor) -> torch.Tensor:
        return EF.half_vectorize(matrix)

    def _normalize_weight_matrix(self, graph: torch.Tensor) -> torch.Tensor:
        """W = D^-1/2 G D^-1/2 with rsqrt(clamp(deg, eps))  (moment_head.py:246-266)."""
        from ..utils.ops import _normalize_graph_rsqrt
        return _normalize_graph_rsqrt(graph, self.eps)

    def _graph_weighted_mean(self, tokens: torch.Tensor, weight_matrix: torch.Tensor) -> torch.Tensor:
        """mu = Z^T W 1 / (tr W + eps)  (moment_head.py:222-244); small helper, torch ops."""
        w = weight_matrix.sum(dim=-1)
        tr = torch.diagonal(weight_matrix, dim1=-2, dim2=-1).sum(-1, keepdim=True)
        return torch.einsum('bn,bnd->bd', w, tokens) / (tr + self.eps)

    def forward(self, tokens: torch.Tensor, graph: torch.Tensor) -> torch.Tensor:
        # fused operator: pool -> iSQRT-COV -> half-vector -> Linear of second_net, no fp32 round
        # trips between the stages (tensor-core precision modes)
        lin = self.second_net[0]
        fused = EF.moment_head_linear(tokens, graph, lin.weight, lin.bias, self.isqrt_cov.num_iterations,
                                      eps=self.eps, third_order=self.use_third_order)
        if fused is not None:
            y, u = fused if self.use_third_order else (fused, None)
            features = [EF.feature_tail(y, list(self.second_net)[1:])]
            if self.use_third_order:
                features.append(self._feature_net(self.third_net, self.tensor_sketch(u)))
            return torch.cat(features, dim=-1)
        if EF.get_ns_algorithm() == "lowrank":
            # opt-in: same function, Newton-Schulz on N x N matrices (functional.moment_isqrt)
            res = EF.moment_isqrt(tokens, graph, self.isqrt_cov.num_iterations, eps=self.eps,
                                  third_order=self.use_third_order)
            M2_normalized, u = res if self.use_third_order else (res, None)
        else:
            if self.use_third_order:
                M2, u = EF.graph_weighted_pool(tokens, graph, eps=self.eps, third_order=True)
            else:
                M2 = EF.graph_weighted_pool(tokens, graph, eps=self.eps, third_order=False)
            M2_normalized = self.isqrt_cov(M2)
        M2_vec = EF.half_vectorize(M2_normalized)
        features = [self._feature_net(self.second_net, M2_vec)]
        if self.use_third_order:
            features.append(self._feature_net(self.third_net, self.tensor_sketch(u)))
        return torch.cat(features, dim=-1)

    @staticmethod
    def _feature_net(net: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
        """Linear -> BatchNorm1d -> GELU -> Dropout (moment_head.py:186-200). The parameters stay
        in the reference's nn.Sequential; the Linear's GEMMs (fwd, dx, dW) run on the library's
        GEMM engines and BatchNorm + GELU + Dropout as one kernel each way (functional.feature_tail)."""
        lin = net[0]
        return EF.feature_tail(EF.linear(x, lin.weight, lin.bias), list(net)[1:])


def test_moment_head():
    """Smoke test mirroring the reference's module-level test (needs a B200)."""
    dev = torch.device('cuda')
    tokens = torch.randn(2, 196, 768, device=dev)
    graph = torch.randn(2, 196, 196, device=dev)
    graph = torch.bmm(graph, graph.transpose(-2, -1))
    graph = 0.5 * (graph + graph.transpose(-2, -1))
    head = MomentHead(d_in=768, d_out=1024, use_third_order=True, isqrt_iterations=3,
                      sketch_dim=2048).to(dev)
    with torch.no_grad():
        out = head(tokens, graph)
    print(f"Output moment features shape: {tuple(out.shape)}")


if __name__ == "__main__":
    test_moment_head()
