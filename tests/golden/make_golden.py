"""Generate the golden fixtures in tests/golden/*.npz FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
It imports the reference's `gpf_kernel.py`, `moment_head.py` and `utils/ops.py` by file path
(the packages themselves need timm / matplotlib, which this image lacks), runs them on seeded
inputs on the CPU and stores inputs, stage taps, outputs and autograd gradients. The fixtures
are what pins `oracle/moment_oracle.py` and, on the GPU box, the CUDA path.
Small cases run the reference in float64 (tight pin for the fp64 oracle); the BASELINE config-1
shape (B=8, N=197, D=768) runs in the reference's native float32.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

REF = os.environ.get("EGM_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_ref(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_inputs(B, N, D, seed=1234, dtype=torch.float32):
    """SURVEY.md section 8d synthetic inputs (iid)."""
    g = torch.Generator().manual_seed(seed)
    anchor = torch.randn(B, N, D, generator=g)
    positive = anchor + 0.5 * torch.randn(B, N, D, generator=g)
    return anchor.to(dtype), positive.to(dtype)


def make_structured_inputs(B, N, D, seed=1234, dtype=torch.float32, rank=16):
    """SURVEY.md section 8d "structured" generator: per-image scale s_b ~ U(0.5, 2), mean m_b ~ N(0, I_D),
    rank-16 factors L_b [N,16], R_b [D,16] and 0.3 noise: anchor = s_b (m_b + L_b R_b^T + 0.3 eps). Images
    differ in scale, mean and covariance, so train-mode BatchNorm sees non-degenerate batch statistics."""
    g = torch.Generator().manual_seed(seed)
    s = 0.5 + 1.5 * torch.rand(B, 1, 1, generator=g)
    m = torch.randn(B, 1, D, generator=g)
    L = torch.randn(B, N, rank, generator=g)
    R = torch.randn(B, D, rank, generator=g)
    anchor = s * (m + torch.bmm(L, R.transpose(1, 2)) + 0.3 * torch.randn(B, N, D, generator=g))
    positive = anchor + 0.5 * s * torch.randn(B, N, D, generator=g)
    return anchor.to(dtype), positive.to(dtype)


def npf(t):
    return t.detach().cpu().numpy()


def run_case(gk, mh, name, B, N, D, P, Q, K, d_out, third, S, similarity="cosine", symmetric=True,
             dtype=torch.float64, full=True, train_bn=False, adaptive=None, patch_sketch=False,
             structured=False):
    torch.manual_seed(0)
    if adaptive:
        gpf = gk.AdaptiveGraphPolynomialFusion(P, Q, similarity=similarity, symmetric_enforce=symmetric,
                                               adaptive_type=adaptive)
    else:
        gpf = gk.GraphPolynomialFusion(P, Q, similarity=similarity, symmetric_enforce=symmetric)
    head = mh.MomentHead(D, d_out, use_third_order=third, isqrt_iterations=K, sketch_dim=S)
    if patch_sketch:
        # the reference caps self.sketch_dim at 4*d_in but still hashes into
        # [0, sketch_dim) and sizes third_net for the uncapped value; this instance-attribute
        # overwrite (no source change) is the documented way to make config 3 runnable.
        head.tensor_sketch.sketch_dim = S
    gpf = gpf.to(dtype)
    head = head.to(dtype)
    for m in head.modules():  # dropout off: parity runs are deterministic
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    head.train(train_bn)
    anchor, positive = (make_structured_inputs if structured else make_inputs)(B, N, D, dtype=dtype)
    anchor.requires_grad_(True)
    positive.requires_grad_(True)
    G = gpf(anchor, positive)
    G.retain_grad()
    # the reference forward, step by step (moment_head.py:279-320), to tap the stages
    W = head._normalize_weight_matrix(G)
    mu = head._graph_weighted_mean(anchor, W)
    Zc = anchor - mu.unsqueeze(1)
    M2 = torch.bmm(Zc.transpose(-2, -1), torch.bmm(W, Zc))
    M2.retain_grad()
    isq = head.isqrt_cov(M2)
    vec = head._half_vectorize(isq)
    vec.retain_grad()
    pre = head.second_net[0](vec)
    out_tap = head(anchor, G)  # the real forward, for the output itself
    gd = torch.Generator().manual_seed(4321)
    dOut = torch.randn(out_tap.shape, generator=gd).to(dtype)
    loss = (out_tap * dOut).sum()
    # gradients of the real forward (taps above share the graph only up to G)
    grads = torch.autograd.grad(loss, [anchor, positive, gpf.alpha_coeffs, G,
                                       head.second_net[0].weight], retain_graph=True)
    # gradient at the half-vector / M2 level through the tapped chain
    second = head.second_net(vec)
    loss_tap = (second * dOut[:, :second.shape[1]]).sum()
    dvec, dM2 = torch.autograd.grad(loss_tap, [vec, M2], retain_graph=True)
    rec = {
        "cfg": np.array([B, N, D, P, Q, K, d_out, int(third), S, int(symmetric), int(train_bn)]),
        "similarity": np.array(similarity),
        "alpha": npf(gpf.alpha_coeffs), "dOut": npf(dOut), "out": npf(out_tap),
        "d_alpha": npf(grads[2]),
        "w_sum": np.array([npf(head.second_net[0].weight).sum(dtype=np.float64),
                           np.abs(npf(head.second_net[0].weight)).sum(dtype=np.float64)]),
        "w_head": npf(head.second_net[0].weight)[:2, :8],
    }
    if third:
        ts = head.tensor_sketch
        rec.update({"hash": np.stack([npf(ts.hash1), npf(ts.hash2), npf(ts.hash3)]),
                    "sign": np.stack([npf(ts.sign1), npf(ts.sign2), npf(ts.sign3)])})
        tw = torch.bmm(W, torch.ones_like(Zc))
        tr_w = torch.diagonal(W, dim1=-2, dim2=-1).sum(-1, keepdim=True)
        u = (Zc * tw).sum(dim=1) / (tr_w + head.eps)
        rec.update({"u": npf(u), "sketch": npf(ts(u))})
    if full:
        rec.update({"anchor": npf(anchor), "positive": npf(positive), "G": npf(G), "W": npf(W),
                    "mu": npf(mu), "M2": npf(M2), "isqrt": npf(isq), "vec": npf(vec), "pre_bn": npf(pre),
                    "d_anchor": npf(grads[0]), "d_positive": npf(grads[1]), "d_G": npf(grads[3]),
                    "d_vec": npf(dvec), "d_M2": npf(dM2),
                    "params": np.array(sorted(head.state_dict().keys()))})
        for k, v in head.state_dict().items():
            if "second_net.0.weight" in k and v.numel() > 200000:
                continue
            rec["p:" + k] = npf(v)
    else:  # config-1 size: compact probes only
        rec.update({
            "in_sum": np.array([npf(anchor).sum(dtype=np.float64), npf(positive).sum(dtype=np.float64)]),
            "G_probe": npf(G)[:, :6, :6], "G_mean": npf(G).mean(axis=(1, 2)),
            "M2_trace": npf(torch.diagonal(M2, dim1=-2, dim2=-1).sum(-1)),
            "M2_probe": npf(M2)[:, :6, :6], "isqrt_probe": npf(isq)[:, :6, :6],
            "isqrt_diag": npf(torch.diagonal(isq, dim1=-2, dim2=-1))[:, ::16],
            "isqrt_fro": npf(isq.pow(2).sum(dim=(1, 2)).sqrt()),
            "vec_probe": npf(vec)[:, :64], "pre_bn": npf(pre),
            "d_anchor_probe": npf(grads[0])[:, :4, :16], "d_positive_probe": npf(grads[1])[:, :4, :16],
            "d_anchor_fro": npf(grads[0].pow(2).sum(dim=(1, 2)).sqrt()),
            "d_positive_fro": npf(grads[1].pow(2).sum(dim=(1, 2)).sqrt()),
            "d_w_probe": npf(grads[4])[:4, :32],
        })
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **rec)
    print(f"{name}: wrote {os.path.getsize(path) / 1024:.1f} KiB  out[0,:3]={npf(out_tap)[0, :3]}")


def run_extgraph(mh):
    """MomentHead on an arbitrary graph: non-symmetric, negative entries, one all-zero row."""
    torch.manual_seed(0)
    B, N, D, K = 3, 11, 16, 4
    head = mh.MomentHead(D, 12, use_third_order=True, isqrt_iterations=K, sketch_dim=48).double()
    for m in head.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    head.eval()
    g = torch.Generator().manual_seed(99)
    tokens = torch.randn(B, N, D, generator=g, dtype=torch.float64).requires_grad_(True)
    graph = torch.randn(B, N, N, generator=g, dtype=torch.float64)
    graph = graph.abs() + 0.1 * torch.randn(B, N, N, generator=g, dtype=torch.float64)
    graph[1, 3, :] = 0.0
    graph.requires_grad_(True)
    out = head(tokens, graph)
    dOut = torch.randn(out.shape, generator=torch.Generator().manual_seed(4321), dtype=torch.float64)
    dt, dg = torch.autograd.grad((out * dOut).sum(), [tokens, graph])
    rec = {"tokens": npf(tokens), "graph": npf(graph), "out": npf(out), "dOut": npf(dOut),
           "d_tokens": npf(dt), "d_graph": npf(dg), "cfg": np.array([B, N, D, K, 12, 48])}
    for k, v in head.state_dict().items():
        rec["p:" + k] = npf(v)
    np.savez_compressed(os.path.join(OUT, "extgraph.npz"), **rec)
    print("extgraph: ok")


def run_ops(ops):
    g = torch.Generator().manual_seed(7)
    X = torch.randn(3, 20, 20, generator=g, dtype=torch.float64)
    M = torch.bmm(X, X.transpose(-2, -1)) + 0.5 * torch.eye(20, dtype=torch.float64)
    graph = torch.rand(3, 9, 9, generator=g, dtype=torch.float64)
    feats = torch.randn(2, 7, 12, generator=g, dtype=torch.float64)
    rec = {
        "M": npf(M), "graph": npf(graph), "feats": npf(feats),
        "sqrt_ns": npf(ops.matrix_sqrt_newton_schulz(M, 5, 1e-5)),
        "halfvec": npf(ops.half_vectorize_symmetric(M)),
        "norm_sym": npf(ops.normalize_graph(graph, "symmetric")),
        "norm_rw": npf(ops.normalize_graph(graph, "random_walk")),
        "trace": npf(ops.batch_trace(M)),
        "cos": npf(ops.cosine_similarity_matrix(feats)),
        "cos2d": npf(ops.cosine_similarity_matrix(feats[0])),
    }
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **rec)
    print("ops: ok")


def run_new(gk, mh):
    # BASELINE config 3: third-order sketch dim 8192 at D=768 (needs the documented instance patch)
    run_case(gk, mh, "cfg3_third_s8192", B=4, N=197, D=768, P=3, Q=3, K=5, d_out=256, third=True,
             S=8192, dtype=torch.float32, full=False, patch_sketch=True)
    # BASELINE config 5: Swin-B 384 px final-stage tokens (N=144, D=1024), adaptive GPF class
    run_case(gk, mh, "cfg5_swin_p3q2", B=4, N=144, D=1024, P=3, Q=2, K=5, d_out=256, third=False,
             S=0, dtype=torch.float32, full=False, adaptive="global")
    run_case(gk, mh, "cfg5_swin_p1q1", B=2, N=144, D=1024, P=1, Q=1, K=5, d_out=256, third=False,
             S=0, dtype=torch.float32, full=False, adaptive="attention")



def run_align():
    """EGOMomentCLEViT._graph_alignment_loss (ego_moment_clevit.py:278-316) of the reference itself:
    the model module is loaded by file path with the reference's own gpf_kernel / moment_head /
    classifier_head next to it and a stub in place of the timm backbone module (timm is absent)."""
    import types
    pkgname = "egm_ref_src"
    root = types.ModuleType(pkgname); root.__path__ = []
    models = types.ModuleType(pkgname + ".models"); models.__path__ = []
    sys.modules.update({pkgname: root, pkgname + ".models": models})
    stub = types.ModuleType(pkgname + ".models.cle_vit_backbone")
    stub.CLEViTDualStream = type("CLEViTDualStream", (torch.nn.Module,), {})
    sys.modules[pkgname + ".models.cle_vit_backbone"] = stub
    for name in ("gpf_kernel", "moment_head", "classifier_head", "ego_moment_clevit"):
        spec = importlib.util.spec_from_file_location(f"{pkgname}.models.{name}",
                                                      os.path.join(REF, "src", "models", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"{pkgname}.models.{name}"] = mod
        spec.loader.exec_module(mod)
    Model = sys.modules[pkgname + ".models.ego_moment_clevit"].EGOMomentCLEViT
    g = torch.Generator().manual_seed(99)
    rec = {}
    for tag, (B, N, dtype) in {"a": (6, 9, torch.float64), "b": (16, 21, torch.float32)}.items():
        G = (torch.rand(B, N, N, generator=g) * 2.0).to(dtype).requires_grad_(True)
        labels = torch.randint(0, 4, (B,), generator=g)
        loss = Model._graph_alignment_loss(None, G, labels)     # the method never touches `self`
        (dG,) = torch.autograd.grad(loss * 3.0, G)
        rec.update({f"{tag}_G": npf(G), f"{tag}_labels": labels.numpy(), f"{tag}_loss": npf(loss),
                    f"{tag}_dG_x3": npf(dG)})
    np.savez_compressed(os.path.join(OUT, "align.npz"), **rec)
    print("align: ok")


def run_dropin():
    """The WHOLE reference model (ego_moment_clevit.py + classifier_head.py + gpf_kernel.py +
    moment_head.py, unchanged, float64, CPU) on a tiny stand-in backbone: forward, the five losses,
    backward. The GPU acceptance test runs the same model file on this repo's drop-in modules."""
    sys.path.insert(0, os.path.dirname(OUT))
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from dropin_stub import StubDualStream
    from baseline import reference_loader as RL
    Model = RL.load_model_class(StubDualStream, root=REF, native=False, pkgname="egm_golden_model")
    torch.manual_seed(0)
    model = Model(num_classes=5, backbone_name="stub", pretrained=False, gpf_degree_p=2, gpf_degree_q=2,
                  moment_d_out=16, use_third_order=True, isqrt_iterations=3, sketch_dim=64,
                  classifier_fusion="concat", lambda_triplet=0.6, lambda_align=0.1, margin=0.3, dropout=0.0)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model = model.double().train()
    state = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(77)
    B = 6
    anchor = torch.randn(B, 3, 16, 16, generator=g, dtype=torch.float64)
    positive = anchor + 0.3 * torch.randn(B, 3, 16, 16, generator=g, dtype=torch.float64)
    labels = torch.randint(0, 5, (B,), generator=g)
    out = model(anchor, positive, labels, return_features=True)
    feats = out["features"]
    watch = {"gpf.alpha_coeffs": model.gpf.alpha_coeffs,
             "backbone.proj.weight": model.backbone.proj.weight,
             "backbone.cls_token": model.backbone.cls_token,
             "moment_head.second_net.0.weight": model.moment_head.second_net[0].weight,
             "moment_head.third_net.0.weight": model.moment_head.third_net[0].weight,
             "classifier.classifier.0.weight": model.classifier.classifier[0].weight}
    grads = torch.autograd.grad(out["loss"], list(watch.values()) + [feats["anchor_tokens"], feats["positive_tokens"]])
    rec = {"anchor": npf(anchor), "positive": npf(positive), "labels": labels.numpy(),
           "logits": npf(out["logits"]), "logits_anchor": npf(out["logits_anchor"]),
           "loss": npf(out["loss"]), "fused_graph": npf(feats["fused_graph"]),
           "moment_features": npf(feats["moment_features"]),
           "d_anchor_tokens": npf(grads[-2]), "d_positive_tokens": npf(grads[-1])}
    for k, v in out["loss_dict"].items():
        rec["loss:" + k] = npf(v)
    for (k, _), gval in zip(watch.items(), grads):
        rec["g:" + k] = npf(gval)
    for k, v in state.items():
        rec["p:" + k] = npf(v)
    np.savez_compressed(os.path.join(OUT, "dropin_model.npz"), **rec)
    print("dropin_model: loss", float(out["loss"]), {k: float(v) for k, v in out["loss_dict"].items()})


def run_structured(gk, mh):
    # BASELINE config-1 shape on the survey's structured inputs, train-mode BatchNorm
    run_case(gk, mh, "cfg1_structured_trainbn", B=8, N=197, D=768, P=3, Q=3, K=5, d_out=256, third=False,
             S=0, dtype=torch.float32, full=False, train_bn=True, structured=True)


def main():
    if "--dropin" in sys.argv:
        run_dropin()
        return
    if "--structured" in sys.argv:
        run_structured(load_ref("ref_gpf_kernel", "src/models/gpf_kernel.py"),
                       load_ref("ref_moment_head", "src/models/moment_head.py"))
        return
    only_new = "--new" in sys.argv
    if not os.path.isdir(REF):
        sys.exit(f"reference tree not found at {REF}")
    gk = load_ref("ref_gpf_kernel", "src/models/gpf_kernel.py")
    mh = load_ref("ref_moment_head", "src/models/moment_head.py")
    ops = load_ref("ref_ops", "src/utils/ops.py")
    if "--align" in sys.argv:
        run_align()
        return
    if only_new:
        run_new(gk, mh)
        return
    run_case(gk, mh, "small_p2q2", B=3, N=12, D=16, P=2, Q=2, K=3, d_out=8, third=False, S=0)
    run_case(gk, mh, "small_p3q3_third", B=3, N=13, D=24, P=3, Q=3, K=5, d_out=10, third=True, S=64)
    run_case(gk, mh, "small_dot_nosym", B=2, N=10, D=16, P=1, Q=2, K=2, d_out=6, third=False, S=0,
             similarity="dot", symmetric=False)
    run_case(gk, mh, "small_trainbn", B=4, N=9, D=16, P=2, Q=1, K=5, d_out=8, third=True, S=32,
             train_bn=True)
    run_extgraph(mh)
    run_ops(ops)
    run_case(gk, mh, "cfg1_b8_n197_d768", B=8, N=197, D=768, P=3, Q=3, K=5, d_out=256, third=False,
             S=0, dtype=torch.float32, full=False)
    run_case(gk, mh, "cfg1_trainbn", B=8, N=197, D=768, P=3, Q=3, K=5, d_out=256, third=False,
             S=0, dtype=torch.float32, full=False, train_bn=True)
    run_new(gk, mh)
    run_structured(gk, mh)
    run_align()
    run_dropin()


if __name__ == "__main__":
    main()
