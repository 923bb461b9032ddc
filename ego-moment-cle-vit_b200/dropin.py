"""Make the reference's own `ego_moment_clevit.py` run on these modules, unchanged.

The reference imports `.gpf_kernel`, `.moment_head` (ego_moment_clevit.py:21-22) and, via
`src.utils`, `.ops` as sub-modules of its `src` package. `install_into('src')` registers this
package's modules under those names in `sys.modules`, so a later
`from src.models.ego_moment_clevit import EGOMomentCLEViT` picks them up without editing a line
of the reference. See INTEGRATION.md.
"""
from __future__ import annotations

import sys


def install_into(package: str = "src") -> None:
    from .models import gpf_kernel, moment_head
    from .utils import ops
    sys.modules[f"{package}.models.gpf_kernel"] = gpf_kernel
    sys.modules[f"{package}.models.moment_head"] = moment_head
    sys.modules[f"{package}.utils.ops"] = ops


def patch_alignment_loss(model_or_class) -> None:
    """Route `EGOMomentCLEViT._graph_alignment_loss` (ego_moment_clevit.py:278-316, an O(B^2) Python
    loop of autograd in-place writes) to `functional.graph_alignment_loss` - same value and gradient.
    Accepts the reference class or an instance of it; nothing else of the class changes."""
    from . import functional as EF
    cls = model_or_class if isinstance(model_or_class, type) else type(model_or_class)

    def _graph_alignment_loss(self, fused_graph, labels):
        return EF.graph_alignment_loss(fused_graph, labels)

    cls._graph_alignment_loss = _graph_alignment_loss
