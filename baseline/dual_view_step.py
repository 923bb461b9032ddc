"""BASELINE.json configs[3]: the full EGO-Moment-CLE-ViT dual-view training step (bench.py --config 4).

What runs: the reference's own, unmodified `ego_moment_clevit.py` + `classifier_head.py` (loaded by file
path from /root/reference or baseline/_ref - baseline/reference_loader.py) wired through
`dropin.install_into` to THIS repository's GraphPolynomialFusion / MomentHead, i.e. exactly the drop-in
the north star describes (ego_moment_clevit.py:85-99 constructs them, :156/:159 calls them), with
`dropin.patch_alignment_loss` replacing the O(B^2) Python loop of `_graph_alignment_loss` (:278-316,
SURVEY.md 8f row 3). The timm backbone is out of scope (and timm is absent): `TorchvisionDualStream`
stands in for `cle_vit_backbone.CLEViTDualStream` with a random-init torchvision ViT-B/16 and the
reference's token contract (cle_vit_backbone.py:232-236: token 0 is the global feature, the other 196
are patch tokens); it runs both views in ONE batched backbone pass (SURVEY.md 8f row 4 - the reference
runs two sequential passes, cle_vit_backbone.py:313-314).

Step = forward, the model's five losses, backward, gradient all-reduce (GradBuckets), clip_grad_norm_(1.0)
and AdamW (train.py:355-377, configs/ufg_base.yaml: lr 3e-4, wd 0.05, amp false), 80 classes, 64 images
per GPU (global 512 at 8 GPUs). This file is benchmark plumbing: nothing in the product package imports it.
"""
from __future__ import annotations

import contextlib
import importlib
import json
import os
import statistics
import sys

import torch
import torch.nn as nn


class TorchvisionDualStream(nn.Module):
    """Stand-in for the reference's CLEViTDualStream (cle_vit_backbone.py:252-316): same constructor
    signature, `.num_features`, and a forward returning two {'patch_tokens', 'global_features'} dicts."""

    def __init__(self, model_name: str = "vit_b_16", pretrained: bool = False, drop_rate: float = 0.0):
        super().__init__()
        import torchvision
        self.model_name = model_name
        self.vit = torchvision.models.vit_b_16(weights=None, dropout=drop_rate)
        self.vit.heads = nn.Identity()
        self.num_features = self.vit.hidden_dim

    def _features(self, x: torch.Tensor) -> torch.Tensor:
        v = self.vit
        x = v._process_input(x)
        x = torch.cat([v.class_token.expand(x.shape[0], -1, -1), x], dim=1)
        return v.encoder(x)                                    # [B, 197, 768]

    def forward(self, anchor: torch.Tensor, positive: torch.Tensor):
        B = anchor.shape[0]
        f = self._features(torch.cat([anchor, positive], dim=0))     # both views, one pass
        out = []
        for t in (f[:B], f[B:]):
            t = t.float()                                      # the moment path is fp32 in, fp32 out
            out.append({"patch_tokens": t[:, 1:].contiguous(), "global_features": t[:, 0]})
        return out[0], out[1]


def build_model(native: bool, num_classes: int, degree, sketch_dim: int, dev):
    from baseline import reference_loader as RL
    pkg = importlib.import_module("ego-moment-cle-vit_b200")
    Model = RL.load_model_class(TorchvisionDualStream, native=native,
                                pkgname="egm_cfg4_native" if native else "egm_cfg4_reference")
    with contextlib.redirect_stdout(sys.stderr):               # the reference's constructor prints
        torch.manual_seed(0)
        model = Model(num_classes=num_classes, backbone_name="vit_b_16 (torchvision, random init)",
                      pretrained=False, gpf_degree_p=degree[0], gpf_degree_q=degree[1], gpf_similarity="cosine",
                      moment_d_out=1024, use_third_order=True, isqrt_iterations=5, sketch_dim=sketch_dim,
                      classifier_fusion="concat", lambda_triplet=0.6, lambda_align=0.1, margin=0.3, dropout=0.1)
    if native:
        pkg.patch_alignment_loss(model)
    return model.to(dev).train()


def run(args, B):
    """`B` is the bench module (shared helpers and constants)."""
    import torch.distributed as dist
    from baseline import reference_loader as RL
    pkg = importlib.import_module("ego-moment-cle-vit_b200")
    EF = pkg.functional
    egm_dist = importlib.import_module("ego-moment-cle-vit_b200.dist")
    lib = pkg._lib.load()
    world, rank, local, dev = B.dist_setup(not args.nccl_normal_priority)
    if RL.find_reference_root() is None:
        if rank == 0:
            print(json.dumps({"metric": "EGO-Moment-CLE-ViT dual-view training step images/sec",
                              "unavailable": "reference model sources not found (EGM_REFERENCE, /root/reference, "
                                             "baseline/_ref): --config 4 runs the reference's own "
                                             "ego_moment_clevit.py on the drop-in modules"}))
        return
    EF.set_precision(args.precision)
    EF.set_ns_algorithm(args.algorithm)
    bs, classes = args.batch, 80
    sketch = args.sketch_dim or 3072        # the largest the unpatched reference accepts at D=768
    model = build_model(True, classes, args.degree, sketch, dev)
    egm_dist.broadcast_parameters(model)
    params = [p for p in model.parameters() if p.requires_grad]
    buckets = None if args.no_allreduce else egm_dist.GradBuckets(params, bucket_bytes=args.bucket_mb << 20,
                                                                  overlap=not args.no_overlap,
                                                                  chunk_bytes=(args.chunk_mb << 20) or None)
    opt = torch.optim.AdamW(params, lr=3e-4, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, fused=True)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    sets = []
    for _ in range(2):
        a = torch.randn(bs, 3, 224, 224, device=dev, generator=gen)
        p = a + 0.5 * torch.randn(bs, 3, 224, 224, device=dev, generator=gen)
        y = torch.randint(0, classes, (bs,), device=dev, generator=gen)
        sets.append((a, p, y))
    host = [tuple(t.cpu().pin_memory() for t in s) for s in sets]
    stage = [tuple(torch.empty_like(t) for t in sets[0]) for _ in range(2)]
    h2d = sum(t.numel() * t.element_size() for t in sets[0])
    amp = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if args.backbone_amp else contextlib.nullcontext

    def make_step(mdl, optim, bk):
        prm = [p for p in mdl.parameters() if p.requires_grad]

        def step(a, p, y):
            with amp():
                out = mdl(a, p, y)
            loss = out["loss"]
            loss.backward()
            if bk is not None:
                bk.reduce()
            torch.nn.utils.clip_grad_norm_(prm, 1.0)
            optim.step()
            optim.zero_grad(set_to_none=True)
            return loss
        return step

    step = make_step(model, opt, buckets)
    timed, barrier = B.make_timed(dev, world)
    copy_stream = torch.cuda.Stream(device=dev)
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    losses = []

    def prefetch(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for d, h in zip(stage[slot], host[i % 2]):
                d.copy_(h, non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_step(i, last=False):
        slot = i % 2
        if i == 0:
            prefetch(0)
        if not last:
            prefetch(i + 1)
        torch.cuda.current_stream().wait_event(copied[slot])
        loss = step(*stage[slot])
        consumed[slot].record()
        loss_host[slot].copy_(loss.detach(), non_blocking=True)
        loss_ready[slot].record()
        if i > 0:
            loss_ready[1 - slot].synchronize()
            losses.append(float(loss_host[1 - slot]))
        if last:
            loss_ready[slot].synchronize()
            losses.append(float(loss_host[slot]))

    for i in range(max(args.warmup, 3)):
        step(*sets[i % 2])
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    for ev in consumed:
        ev.record()
    torch.cuda.synchronize()
    ms_e2e, _, per_e2e = timed(lambda i: e2e_step(i, last=(i == args.steps - 1)), args.steps)
    lib.egm_prof_reset()
    lib.egm_prof_enable(2)
    l0 = lib.egm_launch_count()
    ms_total, ms_mine, per_step = timed(lambda i: step(*sets[i % 2]), args.steps)
    launches = (lib.egm_launch_count() - l0) / args.steps
    lib.egm_prof_enable(0)
    prof = B.read_prof(lib)
    lib.egm_prof_reset()
    D = model.backbone.num_features
    ns = [r for r in prof if r[2][0] == D and r[2][1] == D and r[2][2] == D]
    ns_ms = sum(r[0] for r in ns) / args.steps
    ns_flops = sum(r[1] for r in ns) / args.steps
    ns_n = sum((-r[2][3] if r[2][3] < 0 else 1) for r in ns) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    ranks = B.gather_ranks(dev, world, [ms_mine / args.steps, statistics.median(per_step), ns_ms])

    # the hot path's share: GPF + MomentHead forward+backward alone on tokens of the same shape
    tok = torch.randn(bs, 196, D, device=dev, generator=gen)
    tok2 = tok + 0.5 * torch.randn(bs, 196, D, device=dev, generator=gen)
    dsel = torch.randn(bs, 1024, device=dev, generator=gen)

    def head_only(i):
        a = tok.detach().requires_grad_(True)
        p = tok2.detach().requires_grad_(True)
        out = model.moment_head(a, model.gpf(a, p))
        (out * dsel).sum().backward()
        for q in params:
            q.grad = None

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        head_only(0)
        n_side = max(5, args.steps // 2)
        head_ms = timed(head_only, n_side)[0] / n_side
        extras["hot_path_alone"] = {"ms_per_step": head_ms, "share_of_step": head_ms / (ms_total / args.steps),
                                    "note": f"GPF + MomentHead fwd+bwd on [B={bs},196,{D}] tokens, same modules"}
        try:
            ref = build_model(False, classes, args.degree, sketch, dev)
            ropt = torch.optim.AdamW([p for p in ref.parameters() if p.requires_grad], lr=3e-4, weight_decay=0.05,
                                     fused=True)
            rstep = make_step(ref, ropt, None)
            rstep(*sets[0])
            ms_r = timed(lambda i: rstep(*sets[i % 2]), 2)[0] / 2
            extras["reference_model_on_b200_fp32"] = {
                "value": bs / (ms_r * 1e-3), "unit": "images/s", "ms_per_step": ms_r,
                "note": "the reference's own gpf_kernel.py / moment_head.py / _graph_alignment_loss under the same "
                        "model file, backbone stand-in, inputs, optimizer and GPU (allow_tf32=False as shipped)"}
            del ref, ropt
        except Exception as exc:
            extras["reference_model_error"] = repr(exc)[:300]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak_tf, peak_src, _ = B.load_peaks()
    passes = 3 if args.precision == "fp32" else 1
    ms_step = ms_total / args.steps
    achieved = ns_flops / (ns_ms * 1e-3) / 1e12 if ns_ms > 0 else None
    col = lambda j: [r[j] for r in ranks]
    line = {
        "metric": "EGO-Moment-CLE-ViT dual-view training step images/sec", "value": bs * world / (ms_step * 1e-3),
        "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "ms_per_step_median": max(col(1)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision != "bf16" else "bf16", "data": "synthetic",
        "config": {
            "workload": f"configs[3]: full EGO-Moment-CLE-ViT dual-view training step, ViT-B/16 224 px random init "
                        f"(torchvision stand-in for the timm DeiT-B backbone), {classes} classes, B={bs}/GPU "
                        f"(global {bs * world}), N=196, D={D}, GPF ({args.degree[0]},{args.degree[1]}), moment_d_out=1024, "
                        f"2nd + 3rd order (S={sketch}), 5 NS iters",
            "model": "reference ego_moment_clevit.py + classifier_head.py unchanged, on this repo's drop-in "
                     "GraphPolynomialFusion / MomentHead (install_into) + patch_alignment_loss",
            "per_gpu_batch": bs, "global_batch": bs * world, "tokens": 196, "d_in": D,
            "parallelism": f"dp{world} (NCCL gradient all-reduce, bucketed, overlapped with backward)",
            "step": "fwd (one batched backbone pass for both views) + 5 losses + bwd + all-reduce + clip 1.0 + AdamW",
            "backbone_precision": "bf16 autocast" if args.backbone_amp else "fp32, allow_tf32=False (the reference ships amp: false)",
            "precision": B_precision_note(args.precision), "ns_algorithm": args.algorithm,
            "l2": "two rotating 77 MB image sets; activations far exceed L2"},
        "e2e": {"value": bs * world / (ms_e2e / args.steps * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": f"gemm_tc2_kernel<{passes}> (Newton-Schulz chain of the moment head)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf if achieved else None, "traffic": None, "peak_source": peak_src,
                     "launches_per_step": ns_n, "kernel_ms_per_step": ns_ms, "share_of_step": ns_ms / ms_step,
                     "note": "the step is dominated by the out-of-scope backbone (cuBLAS/ATen); this is the path's own dominant kernel"},
        "cpu_baseline": None,
        "per_rank": {"step_ms_mean": col(0), "step_ms_median": col(1), "ns_chain_ms": col(2)},
    }
    if extras:
        line["other_modes"] = extras
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def B_precision_note(mode):
    return {"fp32": "moment path: fp32 via bf16 hi/lo split, 3 tcgen05 MMAs per product",
            "bf16": "moment path: single bf16 tcgen05 MMA", "fp32_simt": "moment path: fp32 FFMA"}[mode]
