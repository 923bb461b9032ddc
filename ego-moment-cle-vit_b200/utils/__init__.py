"""Drop-in for the mathematical half of the reference's `src/utils` package
(src/utils/__init__.py:10-25). The matplotlib/seaborn plotting helpers (`viz.py`) are out of
scope and stay with the reference."""
from .ops import (
    set_seed, count_parameters, get_model_info, print_model_info, half_vectorize_symmetric,
    matrix_sqrt_newton_schulz, matrix_power_eigen, check_psd, ensure_psd, normalize_graph,
    compute_graph_statistics, batch_trace, batch_logdet, cosine_similarity_matrix,
)

__all__ = [
    'set_seed', 'count_parameters', 'get_model_info', 'print_model_info',
    'half_vectorize_symmetric', 'matrix_sqrt_newton_schulz', 'matrix_power_eigen', 'check_psd',
    'ensure_psd', 'normalize_graph', 'compute_graph_statistics', 'batch_trace', 'batch_logdet',
    'cosine_similarity_matrix',
]
