"""B200-native (sm_100a) moment-pooling path of EGO-Moment-CLE-ViT.

Drop-in `nn.Module`s for the reference's hot path (GraphPolynomialFusion, MomentHead,
NewtonSchulzSqrtm, TensorSketch and the utils.ops matrix helpers) over a C-ABI CUDA library
(`include/egm_b200.h`): tcgen05/TMEM GEMM chains fed by TMA for the dense contractions,
bandwidth kernels for the graph / sketch stages. CUDA only - there is no CPU fallback.
"""
from . import _lib, functional
from .functional import get_ns_algorithm, get_precision, precision, set_ns_algorithm, set_precision
from .models import (AdaptiveGraphPolynomialFusion, GPFKernel, GraphPolynomialFusion, MomentHead,
                     NewtonSchulzSqrtm, TensorSketch)
from .dropin import install_into, patch_alignment_loss

__version__ = "0.1.0"
__all__ = [
    'GraphPolynomialFusion', 'AdaptiveGraphPolynomialFusion', 'GPFKernel', 'MomentHead',
    'NewtonSchulzSqrtm', 'TensorSketch', 'functional', 'set_precision', 'get_precision',
    'precision', 'set_ns_algorithm', 'get_ns_algorithm', 'install_into', 'patch_alignment_loss',
]
