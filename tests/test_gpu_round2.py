"""Round-2 GPU parity: the drop-in acceptance run of the reference's own model file, parity at the
benchmarked batch, structured inputs through train-mode BatchNorm, and the data-parallel hand-over
hook. Run on a B200: pytest -m gpu. Measured errors are appended to gpurun_out/parity_r02.jsonl.
"""
import importlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, golden, make_inputs, make_structured_inputs, rel_err
from oracle import moment_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda")


def npy(t):
    return t.detach().cpu().double().numpy()


def report(name, **vals):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_r02.jsonl"), "a") as f:
        f.write(json.dumps({"test": name, **{k: (float(v) if not isinstance(v, (str, bool)) else v)
                                             for k, v in vals.items()}}) + "\n")


# ------------------------------------------------------------- drop-in acceptance
@pytest.mark.parametrize("patched_align", [False, True])
def test_reference_model_file_runs_unchanged_on_the_dropin_modules(pkg, dev, patched_align):
    """The reference's ego_moment_clevit.py + classifier_head.py, loaded UNCHANGED by file path, with
    `.gpf_kernel` / `.moment_head` swapped for this package (dropin.install_into) and a stand-in
    backbone: forward + the five losses + backward on the GPU against a golden produced by the
    all-reference model (float64, CPU; tests/golden/make_golden.py --dropin)."""
    from baseline import reference_loader as RL
    from dropin_stub import StubDualStream
    if RL.find_reference_root() is None:
        pytest.skip("reference sources not present (/root/reference, baseline/_ref)")
    rec = golden("dropin_model")
    Model = RL.load_model_class(StubDualStream, native=True, pkgname=f"egm_dropin_test_{int(patched_align)}")
    model = Model(num_classes=5, backbone_name="stub", pretrained=False, gpf_degree_p=2, gpf_degree_q=2,
                  moment_d_out=16, use_third_order=True, isqrt_iterations=3, sketch_dim=64,
                  classifier_fusion="concat", lambda_triplet=0.6, lambda_align=0.1, margin=0.3, dropout=0.0)
    assert type(model.gpf).__module__.startswith("ego-moment-cle-vit_b200")
    assert type(model.moment_head).__module__.startswith("ego-moment-cle-vit_b200")
    state = {k[2:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("p:")}
    missing = model.load_state_dict({k: (v.float() if v.is_floating_point() else v) for k, v in state.items()})
    assert not missing.missing_keys and not missing.unexpected_keys        # same state_dict keys
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    if patched_align:
        pkg.patch_alignment_loss(model)
    model = model.to(dev).train()
    anchor = torch.from_numpy(rec["anchor"]).float().to(dev)
    positive = torch.from_numpy(rec["positive"]).float().to(dev)
    labels = torch.from_numpy(rec["labels"]).to(dev)
    with pkg.functional.precision("fp32"):
        out = model(anchor, positive, labels, return_features=True)
        feats = out["features"]
        feats["anchor_tokens"].retain_grad()
        feats["positive_tokens"].retain_grad()
        out["loss"].backward()
    errs = {
        "fused_graph": rel_err(npy(feats["fused_graph"]), rec["fused_graph"]),
        "moment_features": rel_err(npy(feats["moment_features"]), rec["moment_features"]),
        "logits": rel_err(npy(out["logits"]), rec["logits"]),
        "logits_anchor": rel_err(npy(out["logits_anchor"]), rec["logits_anchor"]),
        "loss": abs(float(out["loss"]) - float(rec["loss"])) / abs(float(rec["loss"])),
        "d_anchor_tokens": rel_err(npy(feats["anchor_tokens"].grad), rec["d_anchor_tokens"]),
        "d_positive_tokens": rel_err(npy(feats["positive_tokens"].grad), rec["d_positive_tokens"]),
    }
    for k, v in out["loss_dict"].items():
        errs["loss:" + k] = abs(float(v) - float(rec["loss:" + k])) / max(abs(float(rec["loss:" + k])), 1e-12)
    named = dict(model.named_parameters())
    for k in [k for k in rec if k.startswith("g:")]:
        errs[k] = rel_err(npy(named[k[2:]].grad), rec[k])
    report("dropin_model", patched_align=patched_align, **errs)
    # fp32 mode vs a float64 reference; train-mode BatchNorm over B=6 sits between the head and the logits
    for k, v in errs.items():
        assert v < 1e-3, (k, v, errs)


# ------------------------------------------------- the benchmarked operator at the benchmarked batch
def _bench_batch(dev):
    """B=256 tokens whose first 8 images are the config-1 fixture's inputs."""
    a8, p8 = make_inputs(8, 197, 768)
    g = torch.Generator().manual_seed(2024)
    ar = torch.randn(248, 197, 768, generator=g)
    pr = ar + 0.5 * torch.randn(248, 197, 768, generator=g)
    return torch.cat([a8, ar]).to(dev), torch.cat([p8, pr]).to(dev)


def test_fused_head_at_B256_matches_the_reference_fixture(pkg, dev):
    """The exact operator bench.py times (fused head, symmetric fast path, split-K Linear) at B=256:
    rows 0..7 of its pre-BatchNorm output against the reference's own fp32 `pre_bn` for those images."""
    rec = golden("cfg1_b8_n197_d768")
    EF = pkg.functional
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(3, 3).to(dev)
    head = pkg.MomentHead(768, 256, use_third_order=False, isqrt_iterations=5).to(dev)
    assert np.array_equal(head.second_net[0].weight.detach().cpu().numpy()[:2, :8], rec["w_head"])
    a, p = _bench_batch(dev)
    lin = head.second_net[0]
    with torch.no_grad(), EF.precision("fp32"):
        G = gpf(a, p)
        assert EF.graph_is_symmetric(G)
        y = EF.moment_head_linear(a, G, lin.weight, lin.bias, 5, eps=head.eps)
    err = rel_err(npy(y[:8]), rec["pre_bn"])
    report("fused_head_B256_vs_cfg1_pre_bn", rel=err)
    assert err < 1e-3
    # eval-mode module output for the same rows (BatchNorm running stats: an affine map)
    head.eval()
    with torch.no_grad(), EF.precision("fp32"):
        out = head(a, G)
    err_out = rel_err(npy(out[:8]), rec["out"])
    report("fused_head_B256_vs_cfg1_out_eval", rel=err_out)
    assert err_out < 1e-3


def test_fused_head_is_shard_invariant_at_B256(pkg, dev):
    """B=256 in one call vs 8 shards of 32 (what data parallelism does): forward and token/graph/weight
    gradients of the FUSED operator agree (the per-image chain is bit-identical; the Linear's split-K
    reduction order may differ by rounding)."""
    EF = pkg.functional
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(3, 3).to(dev)
    head = pkg.MomentHead(768, 256, use_third_order=False, isqrt_iterations=5).to(dev)
    lin = head.second_net[0]
    a, p = _bench_batch(dev)
    dy = torch.randn(256, 256, device=dev, generator=torch.Generator(device="cuda").manual_seed(5))

    def run(sl):
        aa = a[sl].contiguous().requires_grad_(True)
        with EF.precision("fp32"):
            G = gpf(aa, p[sl].contiguous())
            y = EF.moment_head_linear(aa, G, lin.weight, lin.bias, 5, eps=head.eps)
            (y * dy[sl]).sum().backward()
        res = (y.detach(), aa.grad.clone(), lin.weight.grad.clone())
        lin.weight.grad = None
        lin.bias.grad = None
        gpf.alpha_coeffs.grad = None
        return res

    y_all, da_all, dw_all = run(slice(0, 256))
    ys, das, dws = zip(*[run(slice(32 * i, 32 * i + 32)) for i in range(8)])
    y_sh, da_sh, dw_sh = torch.cat(ys), torch.cat(das), sum(dws)
    e_y, e_da, e_dw = rel_err(npy(y_sh), npy(y_all)), rel_err(npy(da_sh), npy(da_all)), rel_err(npy(dw_sh), npy(dw_all))
    report("fused_head_shard_invariance_B256", y=e_y, d_anchor=e_da, d_weight=e_dw,
           y_bit_equal=bool(torch.equal(y_sh, y_all)), d_anchor_bit_equal=bool(torch.equal(da_sh, da_all)))
    assert e_y < 1e-6 and e_da < 1e-6 and e_dw < 1e-5


# -------------------------------------------------- structured inputs through train-mode BatchNorm
@pytest.mark.parametrize("name,maker,gate", [("cfg1_structured_trainbn", make_structured_inputs, 1e-3),
                                             ("cfg1_trainbn", make_inputs, 1e-3)])
def test_train_mode_batchnorm_parity(pkg, dev, name, maker, gate):
    """Config-1 shape, train-mode BatchNorm1d after the head's Linear, against the reference's fp32 run.
    Structured tokens (per-image scale / mean / low-rank covariance) hold the north star's
    1e-3 after BatchNorm; iid tokens have degenerate batch statistics (pre-BN batch std ~2e-5, which BN
    amplifies ~50x) - measured 3.5e-4 there (profiles/r02_parity_report.jsonl), so that
    case is held to 1e-3 as well (round 1 gated it at 5e-3 without recording a value)."""
    rec = golden(name)
    B, N, D, P, Q, K, d_out = [int(v) for v in rec["cfg"][:7]]
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(P, Q).to(dev)
    head = pkg.MomentHead(D, d_out, use_third_order=False, isqrt_iterations=K)
    for m in head.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    head = head.to(dev).train()
    anchor, positive = maker(B, N, D)
    assert abs(float(anchor.double().sum()) - rec["in_sum"][0]) < 1e-6 * max(1.0, abs(rec["in_sum"][0]))   # same inputs
    a = anchor.to(dev).requires_grad_(True)
    p = positive.to(dev).requires_grad_(True)
    with pkg.functional.precision("fp32"):
        out = head(a, gpf(a, p))
        (out * torch.from_numpy(rec["dOut"]).to(dev)).sum().backward()
    errs = {"out_after_train_bn": rel_err(npy(out), rec["out"]),
            "d_alpha": rel_err(npy(gpf.alpha_coeffs.grad), rec["d_alpha"]),
            "d_anchor_probe": rel_err(npy(a.grad)[:, :4, :16], rec["d_anchor_probe"]),
            "d_positive_probe": rel_err(npy(p.grad)[:, :4, :16], rec["d_positive_probe"]),
            "d_anchor_fro": rel_err(np.sqrt((npy(a.grad) ** 2).sum(axis=(1, 2))), rec["d_anchor_fro"]),
            "d_w_probe": rel_err(npy(head.second_net[0].weight.grad)[:4, :32], rec["d_w_probe"])}
    report("train_bn_parity:" + name, **errs)
    assert errs["out_after_train_bn"] < gate, errs


# ------------------------------------------------------------- early gradient hand-over (dist.py)
def test_early_grad_hook_sees_dW_before_the_chain_and_changes_nothing(pkg, dev):
    """functional.set_early_grad_hook: the fused head hands dW of its Linear to the data-parallel
    reducer before the Newton-Schulz backward is enqueued; with a pass-through hook every gradient is
    bit-identical to the run without it."""
    EF = pkg.functional
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(2, 2).to(dev)
    head = pkg.MomentHead(136, 16, use_third_order=False, isqrt_iterations=3).to(dev).eval()
    a, p = make_inputs(4, 50, 136)
    a, p = a.to(dev), p.to(dev)
    dy = torch.randn(4, 16, device=dev)
    lib = pkg._lib.load()
    calls = []

    class Handle:
        def __init__(self):
            self.waited = 0

        def wait(self):
            self.waited += 1

    def hook(ptr, dw):
        h = Handle()
        calls.append((ptr, tuple(dw.shape), lib.egm_launch_count(), h))
        return h

    def run():
        aa = a.clone().requires_grad_(True)
        out = head(aa, gpf(aa, p))
        (out * dy).sum().backward()
        res = [aa.grad.clone(), head.second_net[0].weight.grad.clone(), gpf.alpha_coeffs.grad.clone()]
        for q in list(head.parameters()) + list(gpf.parameters()):
            q.grad = None
        return res

    base = run()
    EF.set_early_grad_hook(hook)
    try:
        l0 = lib.egm_launch_count()
        got = run()
        l1 = lib.egm_launch_count()
    finally:
        EF.set_early_grad_hook(None)
    assert len(calls) == 1
    ptr, shape, at, h = calls[0]
    assert ptr == head.second_net[0].weight.data_ptr() and shape == tuple(head.second_net[0].weight.shape)
    assert h.waited == 1
    assert l0 < at < l1 and (l1 - at) >= 6      # the Newton-Schulz backward is enqueued after the hand-over
    for x, y in zip(base, got):
        assert torch.equal(x, y)


# ------------------------------------------- BatchNorm1d + GELU + Dropout as one kernel
@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("M,N", [(256, 256), (6, 8), (37, 70)])
def test_feature_tail_matches_torch_layers(pkg, dev, train, M, N):
    """functional.feature_tail vs the nn.BatchNorm1d -> nn.GELU -> nn.Dropout(p=0) modules themselves
    (moment_head.py:186-191): outputs, input / gamma / beta gradients, and the in-place running-stat update."""
    EF = pkg.functional
    g = torch.Generator().manual_seed(M * 1000 + N)
    y = (torch.randn(M, N, generator=g) * 0.7 + torch.randn(1, N, generator=g)).to(dev)
    dout = torch.randn(M, N, generator=g).to(dev)

    def make():
        torch.manual_seed(3)
        layers = [torch.nn.BatchNorm1d(N), torch.nn.GELU(), torch.nn.Dropout(0.0)]
        with torch.no_grad():
            layers[0].weight.uniform_(0.5, 1.5)
            layers[0].bias.uniform_(-0.3, 0.3)
            layers[0].running_mean.uniform_(-0.2, 0.2)
            layers[0].running_var.uniform_(0.5, 2.0)
        net = torch.nn.Sequential(*layers).to(dev)
        return net.train(train)

    ref, mine = make(), make()
    y1 = y.clone().requires_grad_(True)
    o1 = ref(y1)
    (o1 * dout).sum().backward()
    y2 = y.clone().requires_grad_(True)
    o2 = EF.feature_tail(y2, list(mine))
    (o2 * dout).sum().backward()
    assert rel_err(npy(o2), npy(o1)) < 2e-6
    assert rel_err(npy(y2.grad), npy(y1.grad)) < 2e-5
    assert rel_err(npy(mine[0].weight.grad), npy(ref[0].weight.grad)) < 2e-5
    assert rel_err(npy(mine[0].bias.grad), npy(ref[0].bias.grad)) < 2e-5
    assert rel_err(npy(mine[0].running_mean), npy(ref[0].running_mean)) < 1e-6
    assert rel_err(npy(mine[0].running_var), npy(ref[0].running_var)) < 1e-6
    assert int(mine[0].num_batches_tracked) == int(ref[0].num_batches_tracked)
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())


def test_feature_tail_dropout_contract(pkg, dev):
    """Inverted dropout inside the fused kernel: keep-rate ~ 1-p, kept entries scaled by 1/(1-p), the
    backward uses the same mask, reproducible under torch.manual_seed, off in eval mode."""
    EF = pkg.functional
    M, N, p = 512, 256, 0.1
    layers = [torch.nn.BatchNorm1d(N).to(dev), torch.nn.GELU(), torch.nn.Dropout(p)]
    for l in layers:
        l.train()
    y = torch.randn(M, N, device=dev) + 2.0
    nodrop = EF.feature_tail(y, [layers[0], layers[1], torch.nn.Dropout(0.0).train()])
    torch.manual_seed(11)
    cpu_state = torch.get_rng_state()
    yy = y.clone().requires_grad_(True)
    out = EF.feature_tail(yy, layers)
    assert torch.equal(torch.get_rng_state(), cpu_state)       # the CPU generator is left alone
    kept = out != 0
    rate = float(kept.float().mean())
    live = float((nodrop != 0).float().mean())
    assert abs(rate - live * (1 - p)) < 0.01
    assert torch.allclose(out[kept], nodrop[kept] / (1 - p), rtol=1e-6, atol=1e-7)
    out.sum().backward()
    assert bool((yy.grad[:, 0].abs().sum() > 0))
    torch.manual_seed(11)
    out2 = EF.feature_tail(y, layers)
    assert torch.equal(out2, out.detach())                     # same seed -> same mask
    out3 = EF.feature_tail(y, layers)
    assert not torch.equal(out3, out.detach())                 # the generator advanced
    for l in layers:
        l.eval()
    ev = EF.feature_tail(y, layers)
    assert float((ev == 0).float().mean()) < 0.01              # no dropout in eval mode
    with pytest.raises(ValueError, match="more than 1 value"):
        for l in layers:
            l.train()
        EF.feature_tail(y[:1], layers)


def test_feature_tail_other_layer_stacks_are_applied_as_they_are(pkg, dev):
    EF = pkg.functional
    y = torch.randn(5, 7, device=dev)
    layers = [torch.nn.LayerNorm(7).to(dev), torch.nn.ReLU()]
    assert torch.equal(EF.feature_tail(y, layers), layers[1](layers[0](y)))


def test_strict_fp32_linear_runs_on_the_library(pkg, dev):
    """fp32_simt mode: F.linear forward / dx / dW on the library's FFMA engine (no torch GEMM)."""
    EF = pkg.functional
    lib = pkg._lib.load()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(7, 100, generator=g).to(dev).requires_grad_(True)
    w = torch.randn(5, 100, generator=g).to(dev).requires_grad_(True)
    b = torch.randn(5, generator=g).to(dev).requires_grad_(True)
    dy = torch.randn(7, 5, generator=g).to(dev)
    l0 = lib.egm_launch_count()
    y = EF.linear(x, w, b, precision="fp32_simt")
    (y * dy).sum().backward()
    assert lib.egm_launch_count() - l0 >= 3
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    y2 = torch.nn.functional.linear(x2, w2, b2)
    (y2 * dy).sum().backward()
    assert rel_err(npy(y), npy(y2)) < 1e-6
    assert rel_err(npy(x.grad), npy(x2.grad)) < 1e-6
    assert rel_err(npy(w.grad), npy(w2.grad)) < 1e-6
    assert rel_err(npy(b.grad), npy(b2.grad)) < 1e-6


# --------------------------------------- F.normalize backward folded into E (no rownorm_bwd pass)
@pytest.mark.parametrize("shape,recompute,raw", [((3, 50, 136), False, False), ((2, 197, 768), False, False),
                                                 ((3, 50, 136), False, True), ((2, 197, 768), False, True),
                                                 ((2, 21, 30), True, False), ((1, 230, 64), True, False)])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_gpf_backward_with_the_normalisation_folded_into_E(pkg, dev, shape, recompute, raw, mode):
    """egm_gpf_bwd without kept operand planes: E'' = (E - diag(s)) / (m_i m_j), dx = E'' X on the raw token
    planes (csrc/egm_kernels.cu, gpf_poly3_bwd_kernel FOLD). Token / coefficient gradients vs the oracle, with
    a zero token row (below the eps clamp: no projection term) in the batch. The last two shapes cannot take
    the fused forward (D % 4 != 0, N > 208) and exercise the fold after the staged forward; `raw` lets the
    fused forward write the raw token planes itself (EGM_GPF_RAW_PLANES) instead of the backward re-deriving them."""
    EF = pkg.functional
    B, N, D = shape
    a0, p0 = make_inputs(B, N, D, seed=7)
    a0[0, 3] = 0.0                                   # ||x|| < eps
    g = torch.Generator().manual_seed(5)
    alpha = torch.rand(4, 4, generator=g) * 0.2
    dG = torch.randn(B, N, N, generator=g)
    fw = O.gpf_forward(npy(a0), npy(p0), npy(alpha))
    da, dp, dal = O.gpf_backward(npy(a0), npy(p0), npy(alpha), npy(dG))
    prev = (EF._gpf_recompute, EF._gpf_raw_planes)
    EF._gpf_recompute, EF._gpf_raw_planes = recompute, raw
    try:
        with EF.precision(mode):
            a = a0.to(dev).requires_grad_(True)
            p = p0.to(dev).requires_grad_(True)
            al = alpha.to(dev).requires_grad_(True)
            G = EF.gpf_fused_graph(a, p, torch.nn.functional.softplus(al))
            (G * dG.to(dev)).sum().backward()
    finally:
        EF._gpf_recompute, EF._gpf_raw_planes = prev
    to, tg = (1e-3, 1e-3) if mode == "fp32" else (2e-2, 6e-2)
    errs = {"G": rel_err(npy(G), fw["G"]), "d_anchor": rel_err(npy(a.grad), da), "d_positive": rel_err(npy(p.grad), dp),
            "d_alpha": rel_err(npy(al.grad), dal),
            "zero_row": rel_err(npy(a.grad)[0, 3], da[0, 3])}
    report(f"gpf_fold:{shape}:{mode}:raw={int(raw)}", **errs)
    assert errs["G"] < to
    assert errs["d_anchor"] < tg and errs["d_positive"] < tg and errs["d_alpha"] < tg
    assert errs["zero_row"] < 10 * tg


def test_large_nograd_batches_are_evaluated_in_slices(pkg, dev):
    """a no-grad call of the fused head must not allocate the whole training state. Above the
    working-state budget the batch is evaluated slice by slice - bit-identical to the single call."""
    EF = pkg.functional
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(2, 2).to(dev)
    head = pkg.MomentHead(136, 16, use_third_order=True, isqrt_iterations=3, sketch_dim=64).to(dev).eval()
    a, p = (t.to(dev) for t in make_inputs(7, 50, 136))
    prev = EF._nograd_state_limit
    try:
        with torch.no_grad():
            G = gpf(a, p)
            ref = head(a, G)
            torch.cuda.reset_peak_memory_stats()
            base = torch.cuda.max_memory_allocated()
            EF._nograd_state_limit = 1.0          # one image per slice
            out = head(a, G)
        # with grad enabled the call is never sliced (the saved state belongs to one autograd node)
        a2 = a.clone().requires_grad_(True)
        EF._nograd_state_limit = 1.0
        out2 = head(a2, gpf(a2, p))
        out2.sum().backward()
    finally:
        EF._nograd_state_limit = prev
    assert torch.equal(out, ref)
    assert rel_err(npy(out2), npy(ref)) < 1e-6 and a2.grad is not None
