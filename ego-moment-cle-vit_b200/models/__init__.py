"""Drop-in replacements for the hot-path modules of the reference's `src/models` package
(reference exports: src/models/__init__.py:19-28). Backbone / classifier / integration model
are out of scope and stay with the reference."""
from .gpf_kernel import GraphPolynomialFusion, AdaptiveGraphPolynomialFusion, GPFKernel
from .moment_head import MomentHead, NewtonSchulzSqrtm, TensorSketch

__all__ = [
    'GraphPolynomialFusion', 'AdaptiveGraphPolynomialFusion', 'GPFKernel',
    'MomentHead', 'NewtonSchulzSqrtm', 'TensorSketch',
]
