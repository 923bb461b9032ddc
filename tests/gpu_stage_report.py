"""Stage-by-stage parity report of the CUDA path against the numpy oracle (run on a B200):

    python tests/gpu_stage_report.py [B N D]

Prints rel-Frobenius errors of every stage tap and every gradient for each precision mode.
Not a test (no asserts) - tests/test_gpu_parity.py holds the gates; this is the diagnostic.
"""
import os
import sys
import time

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
    pkg = load_pkg()
    EF = pkg.functional
    dev = torch.device("cuda")
    anchor, positive = make_inputs(B, N, D)
    torch.manual_seed(0)
    alpha = torch.rand(P + 1, Q + 1) * 0.1
    gd = torch.Generator().manual_seed(4321)
    dvec = torch.randn(B, D * (D + 1) // 2, generator=gd)
    S = min(4 * D, 2048)
    hashes = torch.randint(0, S, (3, D), generator=gd)
    signs = torch.randint(0, 2, (3, D), generator=gd) * 2 - 1
    dsk = torch.randn(B, S, generator=gd)

    t0 = time.time()
    a64, p64, al64 = anchor.double().numpy(), positive.double().numpy(), alpha.double().numpy()
    fw = O.gpf_forward(a64, p64, al64)
    third = {"hashes": hashes.numpy(), "signs": signs.numpy(), "sketch_dim": S}
    st = O.moment_forward(a64, fw["G"], K, 1e-5, third)
    dZ_o, dG_o = O.moment_backward(a64, fw["G"], K, dvec.double().numpy(), 1e-5, third, dsk.double().numpy())
    da_o, dp_o, dal_o = O.gpf_backward(a64, p64, al64, dG_o)
    da_o = da_o + dZ_o
    print(f"oracle fp64 done in {time.time() - t0:.1f}s  (B={B} N={N} D={D} P=Q={P} K={K} S={S})")

    for mode in ("fp32_simt", "fp32", "bf16"):
        with EF.precision(mode):
            a = anchor.to(dev).requires_grad_(True)
            p = positive.to(dev).requires_grad_(True)
            al = alpha.to(dev).requires_grad_(True)
            torch.cuda.synchronize()
            t0 = time.time()
            coef = torch.nn.functional.softplus(al)
            G = EF.gpf_fused_graph(a, p, coef)
            M2, u = EF.graph_weighted_pool(a, G, eps=1e-5, third_order=True)
            isq = EF.newton_schulz(M2, K, 1e-5)
            vec = EF.half_vectorize(isq)
            csr = EF.build_sketch_csr(hashes.to(dev), signs.to(dev), S)
            sk = EF.tensor_sketch(u, hashes.to(dev), signs.to(dev), csr, S)
            loss = (vec * dvec.to(dev)).sum() + (sk * dsk.to(dev)).sum()
            G.retain_grad()
            loss.backward()
            torch.cuda.synchronize()
            dt = time.time() - t0
        row = {
            "G": rel_err(G.detach().cpu().numpy(), fw["G"]),
            "M2": rel_err(M2.detach().cpu().numpy(), st["M2"]),
            "u": rel_err(u.detach().cpu().numpy(), st["u"]),
            "isqrt": rel_err(isq.detach().cpu().numpy(), st["isqrt"]),
            "sketch": rel_err(sk.detach().cpu().numpy(), st["sketch"]),
            "dG": rel_err(G.grad.cpu().numpy(), dG_o),
            "d_anchor": rel_err(a.grad.cpu().numpy(), da_o),
            "d_positive": rel_err(p.grad.cpu().numpy(), dp_o),
            "d_alpha": rel_err(al.grad.cpu().numpy(), dal_o),
        }
        off = ~np.eye(D, dtype=bool)
        row["isqrt_offdiag"] = rel_err(isq.detach().cpu().numpy()[:, off], st["isqrt"][:, off])
        print(f"[{mode:9s}] ({dt * 1e3:.0f} ms incl. first-call) " +
              "  ".join(f"{k}={v:.2e}" for k, v in row.items()))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
