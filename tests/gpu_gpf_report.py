"""Timing of GraphPolynomialFusion.forward (fused single pass vs staged path) on a B200 - a report, not a
test. Usage: python tests/gpu_gpf_report.py [B N D P Q]; EGM_GPF_FUSED=0 selects the staged path."""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ego-moment-cle-vit_b200")
EF = pkg.functional


def main():
    B, N, D, P, Q = [int(v) for v in (sys.argv[1:6] if len(sys.argv) >= 6 else (256, 197, 768, 3, 3))]
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(1)
    sets = []
    for _ in range(2):                       # two input sets: 2 x 310 MB > L2
        a = torch.randn(B, N, D, device=dev, generator=g)
        sets.append((a, a + 0.5 * torch.randn(B, N, D, device=dev, generator=g)))
    coef = torch.nn.functional.softplus(torch.rand(P + 1, Q + 1, device=dev, generator=g) * 0.1)
    out = {"shape": [B, N, D, P, Q], "fused": os.environ.get("EGM_GPF_FUSED", "1") != "0"}
    for mode in ("fp32", "bf16"):
        for grad in (False, True):
            def run(i):
                a, p = sets[i % 2]
                if grad:
                    a = a.detach().requires_grad_(True)
                    return EF.gpf_fused_graph(a, p, coef, precision=mode)
                with torch.no_grad():
                    return EF.gpf_fused_graph(a, p, coef, precision=mode)
            for i in range(3):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for i in range(n):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            alg = B * (2 * N * D * 4 + N * N * 4)
            out[f"{mode}_{'grad' if grad else 'nograd'}_us"] = round(us, 1)
            out[f"{mode}_{'grad' if grad else 'nograd'}_alg_GBs"] = round(alg / us / 1e3, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
