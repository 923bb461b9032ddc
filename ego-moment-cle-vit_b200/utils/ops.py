"""Matrix helpers - drop-in for the reference's `src/utils/ops.py` (same names, signatures,
return values). The five helpers that share arithmetic with the hot path run on the sm_100a
library (`half_vectorize_symmetric`, `matrix_sqrt_newton_schulz`, `normalize_graph`,
`batch_trace`, `cosine_similarity_matrix`); the eigen-decomposition wrappers and host utilities
delegate to torch exactly as the reference does (ops.py:17-97, 168-235, 274-352).
"""
from __future__ import annotations

import random
from typing import Any, Dict

import numpy as np
import torch
import torch.nn as nn

from .. import functional as EF


def set_seed(seed: int = 42):
    """Seed python, numpy and torch (CPU + all GPUs); deterministic cuDNN (ops.py:17-31)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def count_parameters(model: nn.Module, trainable_only: bool = True) -> int:
    if trainable_only:
        return sum(p.numel() for p in model.parameters() if p.requires_grad)
    return sum(p.numel() for p in model.parameters())


def get_model_info(model: nn.Module) -> Dict[str, Any]:
    total = count_parameters(model, trainable_only=False)
    trainable = count_parameters(model, trainable_only=True)
    param_size = sum(p.nelement() * p.element_size() for p in model.parameters())
    buffer_size = sum(b.nelement() * b.element_size() for b in model.buffers())
    return {
        'total_parameters': total,
        'trainable_parameters': trainable,
        'non_trainable_parameters': total - trainable,
        'model_size_mb': (param_size + buffer_size) / 1024 / 1024,
        'parameter_size_mb': param_size / 1024 / 1024,
        'buffer_size_mb': buffer_size / 1024 / 1024,
    }


def print_model_info(model: nn.Module, input_size=None):
    info = get_model_info(model)
    print("=" * 50)
    print("MODEL INFORMATION")
    print("=" * 50)
    print(f"Total parameters: {info['total_parameters']:,}")
    print(f"Trainable parameters: {info['trainable_parameters']:,}")
    print(f"Non-trainable parameters: {info['non_trainable_parameters']:,}")
    print(f"Model size: {info['model_size_mb']:.2f} MB")
    if input_size is not None:
        print(f"Input size: {input_size}")
    print("=" * 50)


def half_vectorize_symmetric(matrix: torch.Tensor) -> torch.Tensor:
    """[B,D,D] -> [B,D(D+1)/2], row-major upper triangle incl. diagonal (ops.py:100-119)."""
    return EF.half_vectorize(matrix)


def matrix_sqrt_newton_schulz(matrix: torch.Tensor, num_iterations: int = 5,
                              eps: float = 1e-5) -> torch.Tensor:
    """Same coupled iteration as NewtonSchulzSqrtm but post-MULTIPLIED by sqrt(tr+eps)
    (ops.py:122-165) - i.e. sqrt(tr) * A_normalised^(-1/2), reproduced as the reference computes it."""
    return EF.newton_schulz(matrix, num_iterations, eps, post="multiply")


def matrix_power_eigen(matrix: torch.Tensor, power: float) -> torch.Tensor:
    eigenvals, eigenvecs = torch.linalg.eigh(matrix)
    eigenvals = torch.clamp(eigenvals, min=1e-8)
    return torch.bmm(torch.bmm(eigenvecs, torch.diag_embed(torch.pow(eigenvals, power))),
                     eigenvecs.transpose(-2, -1))


def check_psd(matrix: torch.Tensor, tol: float = 1e-6) -> bool:
    try:
        return torch.linalg.eigvals(matrix).real.min().item() >= -tol
    except Exception:
        return False


def ensure_psd(matrix: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    eigenvals, eigenvecs = torch.linalg.eigh(matrix)
    eigenvals = torch.clamp(eigenvals, min=eps)
    return torch.bmm(torch.bmm(eigenvecs, torch.diag_embed(eigenvals)), eigenvecs.transpose(-2, -1))


class _NormalizeGraph(torch.autograd.Function):
    """normalize_graph forward on the library kernel; analytic backward in a few torch ops."""

    @staticmethod
    def forward(ctx, graph, method, eps):
        out, deg = EF.normalize_graph(graph, method, eps)
        raw_deg = graph.sum(dim=-1)
        ctx.save_for_backward(graph, deg, raw_deg)
        ctx.cfg = (method, eps)
        return out

    @staticmethod
    def backward(ctx, dout):
        graph, deg, raw_deg = ctx.saved_tensors
        method, eps = ctx.cfg
        live = (raw_deg >= eps).to(dout.dtype)
        if method == 0:
            s = deg.rsqrt()
            dG = dout * s.unsqueeze(-1) * s.unsqueeze(-2)
            t = dout * graph
            ds = (t * s.unsqueeze(-2)).sum(-1) + (t * s.unsqueeze(-1)).sum(-2)
            ddeg = -0.5 * ds * s / deg * live
        else:
            inv = 1.0 / deg
            dG = dout * inv.unsqueeze(-1)
            ddeg = -(dout * graph).sum(-1) * inv * inv * live
        return dG + ddeg.unsqueeze(-1), None, None


def normalize_graph(graph: torch.Tensor, method: str = 'symmetric') -> torch.Tensor:
    """'symmetric': D^-1/2 A D^-1/2, 'random_walk': D^-1 A, 'none' (ops.py:238-271; eps 1e-8)."""
    if method == 'none':
        return graph
    if method not in ('symmetric', 'random_walk'):
        raise ValueError(f"Unknown normalization method: {method}")
    return _NormalizeGraph.apply(graph, 0 if method == 'symmetric' else 1, 1e-8)


def _normalize_graph_rsqrt(graph: torch.Tensor, eps: float) -> torch.Tensor:
    """MomentHead._normalize_weight_matrix flavour (clamp at `eps`), exposed for the module helper."""
    return _NormalizeGraph.apply(graph, 0, eps)


def compute_graph_statistics(graph: torch.Tensor) -> Dict[str, float]:
    stats: Dict[str, Any] = {}
    with torch.no_grad():
        stats['mean'] = graph.mean().item()
        stats['std'] = graph.std().item()
        stats['min'] = graph.min().item()
        stats['max'] = graph.max().item()
        symmetry_error = (graph - graph.transpose(-2, -1)).abs().max().item()
        stats['symmetry_error'] = symmetry_error
        stats['is_symmetric'] = symmetry_error < 1e-5
        try:
            eigenvals = torch.linalg.eigvals(graph).real
            stats['min_eigenval'] = eigenvals.min().item()
            stats['max_eigenval'] = eigenvals.max().item()
            stats['eigenval_ratio'] = (eigenvals.max() / torch.clamp(eigenvals.min(), min=1e-8)).item()
            stats['is_psd'] = stats['min_eigenval'] >= -1e-6
        except Exception:
            stats['eigenval_error'] = True
        threshold = 0.1 * stats['max']
        stats['sparsity'] = (graph.abs() < threshold).float().mean().item()
    return stats


class _BatchTrace(torch.autograd.Function):
    @staticmethod
    def forward(ctx, matrices):
        ctx.shape = matrices.shape
        return EF.batch_trace(matrices)

    @staticmethod
    def backward(ctx, dtr):
        B, D, _ = ctx.shape
        return dtr.view(B, 1, 1) * torch.eye(D, device=dtr.device, dtype=dtr.dtype).expand(B, D, D)


def batch_trace(matrices: torch.Tensor) -> torch.Tensor:
    """[B,D,D] -> [B] traces (ops.py:316-326)."""
    return _BatchTrace.apply(matrices)


def batch_logdet(matrices: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    eye = torch.eye(matrices.shape[-1], device=matrices.device, dtype=matrices.dtype)
    stabilized = matrices + eps * eye.unsqueeze(0)
    try:
        return torch.logdet(stabilized)
    except Exception:
        eigenvals = torch.clamp(torch.linalg.eigvals(stabilized).real, min=eps)
        return torch.log(eigenvals).sum(-1)


def cosine_similarity_matrix(features: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """Pairwise cosine similarity for [B,N,D] or [N,D] features (ops.py:355-381)."""
    squeeze = features.dim() == 2
    if squeeze:
        features = features.unsqueeze(0)
    sim = EF.similarity_matrix(features, cosine=True, eps=eps)
    return sim.squeeze(0) if squeeze else sim


def test_ops():
    """Smoke test mirroring the reference's module-level test (needs a B200)."""
    dev = torch.device('cuda')
    m = torch.randn(2, 64, 64, device=dev)
    m = torch.bmm(m, m.transpose(-2, -1))
    print("half-vec:", tuple(half_vectorize_symmetric(m).shape))
    print("ns-sqrt:", tuple(matrix_sqrt_newton_schulz(m).shape))
    print("trace:", batch_trace(m))
    print("cos:", tuple(cosine_similarity_matrix(torch.randn(2, 10, 64, device=dev)).shape))


if __name__ == "__main__":
    test_ops()
