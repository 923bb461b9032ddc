"""Parity report of the fused moment head (GPF -> pool -> iSQRT-COV -> half-vector -> Linear) against the
numpy fp64 oracle, for the symmetric-graph fast path and the general chain (run on a B200):

    python tests/gpu_fused_report.py [B N D]

Diagnostic only (no asserts); tests/test_gpu_parity.py holds the gates.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import load_pkg, make_inputs, rel_err  # noqa: E402
from oracle import moment_oracle as O  # noqa: E402


def main():
    B, N, D = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (3, 197, 768)
    P = Q = 3
    K = 5
    n_out = 64
    pkg = load_pkg()
    EF = pkg.functional
    dev = torch.device("cuda")
    anchor, positive = make_inputs(B, N, D)
    torch.manual_seed(0)
    alpha = torch.rand(P + 1, Q + 1) * 0.1
    gd = torch.Generator().manual_seed(4321)
    L = D * (D + 1) // 2
    Wt = torch.randn(n_out, L, generator=gd) / L ** 0.5
    dy = torch.randn(B, n_out, generator=gd)
    a64, p64, al64 = anchor.double().numpy(), positive.double().numpy(), alpha.double().numpy()
    fw = O.gpf_forward(a64, p64, al64)
    st = O.moment_forward(a64, fw["G"], K, 1e-5, None)
    y_o = st["vec"] @ Wt.double().numpy().T
    dvec = dy.double().numpy() @ Wt.double().numpy()
    dZ_o, dG_o = O.moment_backward(a64, fw["G"], K, dvec, 1e-5, None, None)
    da_o, dp_o, dal_o = O.gpf_backward(a64, p64, al64, dG_o)
    da_o = da_o + dZ_o
    dW_o = dy.double().numpy().T @ st["vec"]
    for mode in ("fp32", "bf16"):
        for fast in (True, False):
            EF.set_symmetric_fast_path(fast)
            with EF.precision(mode):
                a = anchor.to(dev).requires_grad_(True)
                p = positive.to(dev).requires_grad_(True)
                al = alpha.to(dev).requires_grad_(True)
                w = Wt.to(dev).requires_grad_(True)
                G = EF.gpf_fused_graph(a, p, torch.nn.functional.softplus(al))
                y = EF.moment_head_linear(a, G, w, None, K, eps=1e-5)
                (y * dy.to(dev)).sum().backward()
                torch.cuda.synchronize()
            row = {"y": rel_err(y.detach().cpu().numpy(), y_o),
                   "d_anchor": rel_err(a.grad.cpu().numpy(), da_o),
                   "d_positive": rel_err(p.grad.cpu().numpy(), dp_o),
                   "d_alpha": rel_err(al.grad.cpu().numpy(), dal_o),
                   "dW": rel_err(w.grad.cpu().numpy(), dW_o)}
            print(f"[{mode:5s} {'symmetric fast path' if fast else 'general chain      '}] " +
                  "  ".join(f"{k}={v:.2e}" for k, v in row.items()))
            sys.stdout.flush()
    EF.set_symmetric_fast_path(True)


if __name__ == "__main__":
    main()
