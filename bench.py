#!/usr/bin/env python
"""Benchmark of the moment-pooling hot path: GPF + MomentHead forward+backward, images/s.

    python bench.py --gpus N --steps K --warmup W [--precision fp32|bf16] [--impl reference]

Workload (BASELINE.json configs[1]): B=256 images per GPU, N=197 tokens, D=768 -> d_out=256,
GPF degree (3,3) cosine, 2nd-order iSQRT-COV with 5 Newton-Schulz iterations, train mode
(BatchNorm batch statistics, Dropout 0.1), synthetic ViT-B/16 token tensors, random-init
weights. One "step" = forward + backward (+ gradient all-reduce for N>1) + SGD update over one
batch. Weak scaling: the per-GPU batch is fixed, `value` is the whole-job images/s.

  value      inputs already resident in HBM (two rotating 310 MB input sets > 126 MB L2)
  e2e        same step through the public nn.Module API with HOST (pinned) token buffers:
             H2D copy of both token tensors (double-buffered on a copy stream) + D2H read of the
             loss of every step, all inside the timed region. Timed FIRST, `value` right after it:
             under the 1 kW power cap the same K steps read ~3 % slower a few seconds later
  roofline   the dominant kernel (tcgen05 GEMM engine, Newton-Schulz chain), timed live with
             CUDA events inside the timed steps, against MEASURED_PEAKS.json
  cpu_baseline  the numpy oracle port of the reference's algorithm on the host cores (bounded
             sample), N=1 rank 0 only

`--impl reference` times that CPU port alone (the reference is pure Python/torch-CPU and is not
present on the GPU box; the oracle restates it - see oracle/moment_oracle.py).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MomentHead+GPF fwd+bwd images/sec"
UNIT = "images/s"
N_TOK, D_IN, D_OUT, DEG, NS_ITERS = 197, 768, 256, 3, 5
DEGS = (DEG, DEG)


def parse():
    global N_TOK, D_IN, DEGS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU")
    ap.add_argument("--algorithm", default="dense", choices=["dense", "lowrank"],
                    help="iSQRT-COV evaluation: D x D Newton-Schulz chain, or the N x N low-rank form")
    ap.add_argument("--tokens", type=int, default=N_TOK, help="tokens per image (197 ViT-B/16; 144 Swin-B 384 px)")
    ap.add_argument("--d-in", type=int, default=D_IN, help="token width (768 ViT-B; 1024 Swin-B)")
    ap.add_argument("--degree", type=int, nargs=2, default=[DEG, DEG], metavar=("P", "Q"),
                    help="GPF polynomial degrees (BASELINE configs[4] sweeps (1,1)...(3,3))")
    ap.add_argument("--no-extras", action="store_true", help="skip bf16-mode / cpu-baseline side runs")
    ap.add_argument("--gemm-breakdown", action="store_true",
                    help="print the per-shape time of the GEMM-engine launches to stderr")
    args = ap.parse_args()
    N_TOK, D_IN, DEGS = args.tokens, args.d_in, tuple(args.degree)   # the named workload unless overridden
    return args


def workload_config(args, extra=None):
    cfg = {
        "workload": ("configs[1]" if (N_TOK, D_IN, tuple(args.degree)) == (197, 768, (DEG, DEG)) else "custom shape") +
                    f": MomentHead+GPFKernel (degree {args.degree[0]},{args.degree[1]}, 2nd-order iSQRT-COV, "
                    f"5 NS iters) fwd+bwd, B={args.batch}/GPU, N={N_TOK}, D={D_IN}->d_out={D_OUT}",
        "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus, "tokens": N_TOK,
        "d_in": D_IN, "d_out": D_OUT, "gpf_degree": list(args.degree), "ns_iterations": NS_ITERS,
        "parallelism": f"dp{args.gpus} (batch sharded per image, NCCL gradient all-reduce)",
        "step": "forward + backward + grad all-reduce + SGD update, train-mode BN, dropout 0.1",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------ clocks during the run
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "nvidia-smi unavailable"}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------- CPU reference leg
def cpu_port_setup(batch, seed=0):
    """Reference-shaped inputs and parameters for the oracle port (fp32, as the reference)."""
    import numpy as np
    import torch
    pkg = importlib.import_module("ego-moment-cle-vit_b200")
    torch.manual_seed(seed)
    gpf = pkg.GraphPolynomialFusion(*DEGS)
    head = pkg.MomentHead(D_IN, D_OUT, use_third_order=False, isqrt_iterations=NS_ITERS)
    params = {k: v.detach().numpy() for k, v in head.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    anchor = torch.randn(batch, N_TOK, D_IN, generator=g)
    positive = anchor + 0.5 * torch.randn(batch, N_TOK, D_IN, generator=g)
    dout = torch.randn(batch, D_OUT, generator=torch.Generator().manual_seed(4321))
    return (anchor.numpy(), positive.numpy(), gpf.alpha_coeffs.detach().numpy(), params, dout.numpy(), np)


def cpu_port_rate(budget_s=20.0, max_steps=8):
    """images/s of the oracle port on the host cores: a bounded sample of the same workload."""
    from oracle import moment_oracle as O
    O.set_matmul_backend("torch")       # the threaded torch.bmm the reference itself runs on
    bs = 8
    a, p, alpha, params, dout, np = cpu_port_setup(bs)
    O.path_step(a[:2], p[:2], alpha, params, NS_ITERS, dout[:2], True, np.float32)   # warm-up
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        O.path_step(a, p, alpha, params, NS_ITERS, dout, True, np.float32)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return bs * n / dt, f"{n} step(s) of B={bs} images (N={N_TOK}, D={D_IN}), fwd+bwd, numpy + torch.bmm fp32, {dt:.1f} s"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import moment_oracle as O
    O.set_matmul_backend("torch")       # the threaded torch.bmm the reference itself runs on
    cores = os.cpu_count() or 1
    bs = 8
    a, p, alpha, params, dout, np = cpu_port_setup(bs)
    t0 = time.perf_counter()
    O.path_step(a[:2], p[:2], alpha, params, NS_ITERS, dout[:2], True, np.float32)
    per_img = (time.perf_counter() - t0) / 2
    # bound the whole run to a few minutes
    total = max(1, args.steps + args.warmup)
    bs = max(1, min(8, int(150.0 / (per_img * total))))
    for _ in range(args.warmup):
        O.path_step(a[:bs], p[:bs], alpha, params, NS_ITERS, dout[:bs], True, np.float32)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.path_step(a[:bs], p[:bs], alpha, params, NS_ITERS, dout[:bs], True, np.float32)
    dt = time.perf_counter() - t0
    value = bs * args.steps / dt
    sample = f"{args.steps} steps of B={bs} images of the same workload, numpy/torch-CPU-bmm fp32 port of the reference algorithm"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, {"sample": sample}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def torch_eager_rates(B, dev, gpf, head, dev_inputs, d_out, timed, steps=3):
    """images/s of the same training step written as plain torch ops (cuBLAS bmm, autograd): fp32 with
    TF32 off (the reference's setting) and with TF32 allowed."""
    import copy
    import torch
    from baseline import torch_eager as TE
    alpha = torch.nn.Parameter(gpf.alpha_coeffs.detach().clone())
    net = copy.deepcopy(head.second_net).to(dev).train()
    params = [alpha] + list(net.parameters())
    opt = torch.optim.SGD(params, lr=1e-6)

    def step(i):
        a, p = dev_inputs[i % 2]
        a = a.detach().requires_grad_(True)
        p = p.detach().requires_grad_(True)
        out = TE.head_forward(a, TE.gpf_forward(a, p, alpha), net, NS_ITERS)
        (out * d_out).sum().backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    out = {}
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32 in (("torch_eager_fp32", False), ("torch_eager_tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            step(0)
            ms = timed(step, steps) / steps
            out[name] = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                         "note": "reference algorithm as stock torch ops + autograd on the same GPU, "
                                 + ("allow_tf32=True" if tf32 else "fp32 (allow_tf32=False, the reference's setting)")}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


# ------------------------------------------------------------------------------ native leg
def run_native(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("ego-moment-cle-vit_b200")
    EF = pkg.functional
    egm_dist = importlib.import_module("ego-moment-cle-vit_b200.dist")
    lib = pkg._lib.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    B = args.batch
    EF.set_precision(args.precision)
    EF.set_ns_algorithm(args.algorithm)
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(args.degree[0], args.degree[1]).to(dev)
    head = pkg.MomentHead(D_IN, D_OUT, use_third_order=False, isqrt_iterations=NS_ITERS).to(dev).train()
    egm_dist.broadcast_parameters(gpf)
    egm_dist.broadcast_parameters(head)
    params = list(gpf.parameters()) + list(head.parameters())
    buckets = egm_dist.GradBuckets(params)
    opt = torch.optim.SGD(params, lr=1e-6)

    # two rotating input sets, each 2 x B x 197 x 768 x 4 B = 310 MB (> 126 MB of L2)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    dev_inputs = []
    for _ in range(2):
        a = torch.randn(B, N_TOK, D_IN, device=dev, generator=gen)
        dev_inputs.append((a, a + 0.5 * torch.randn(B, N_TOK, D_IN, device=dev, generator=gen)))
    host_inputs = [(a.cpu().pin_memory(), p.cpu().pin_memory()) for a, p in dev_inputs]
    stage = [(torch.empty_like(a), torch.empty_like(p)) for a, p in dev_inputs[:1]]
    d_out = torch.randn(B, D_OUT, device=dev, generator=gen)
    h2d_bytes = 2 * B * N_TOK * D_IN * 4

    # live timing of the tcgen05 GEMM engine inside the timed steps: the library brackets every
    # engine launch with two CUDA events on the launching stream (egm_prof_*, include/egm_b200.h)
    def read_prof():
        import ctypes
        recs = []
        ms, fl, dims = ctypes.c_float(), ctypes.c_double(), (ctypes.c_int * 6)()
        for i in range(lib.egm_prof_count()):
            if lib.egm_prof_read(i, ctypes.byref(ms), ctypes.byref(fl), dims) == 0:
                recs.append((ms.value, fl.value, tuple(dims)))
        return recs

    def step(a, p):
        a = a.requires_grad_(True)
        p = p.requires_grad_(True)
        out = head(a, gpf(a, p))
        loss = (out * d_out).sum()
        loss.backward()
        buckets.reduce()
        opt.step()
        opt.zero_grad(set_to_none=True)
        a.requires_grad_(False); p.requires_grad_(False)
        a.grad = None; p.grad = None
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between CUDA events on the current stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def resident_step(i):
        step(*dev_inputs[i % 2])

    # end-to-end: every step's tokens start in pinned HOST memory. The copy of step i+1 is issued
    # on a side stream before step i is computed (a double-buffered loader), and the loss of step
    # i is read back through a pinned buffer one step later, so neither transfer stalls the GPU;
    # all copies and read-backs of the K steps happen inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    stage.append((torch.empty_like(dev_inputs[0][0]), torch.empty_like(dev_inputs[0][1])))
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_ready = [torch.cuda.Event() for _ in range(2)]
    e2e_losses = []

    def e2e_prefetch(i):
        slot = i % 2
        ha, hp = host_inputs[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])       # the step that last used this slot is done
            stage[slot][0].copy_(ha, non_blocking=True)
            stage[slot][1].copy_(hp, non_blocking=True)
            copied[slot].record(copy_stream)

    def e2e_step(i, last=False):
        slot = i % 2
        if i == 0:
            e2e_prefetch(0)
        if not last:
            e2e_prefetch(i + 1)
        torch.cuda.current_stream().wait_event(copied[slot])
        loss = step(*stage[slot])
        consumed[slot].record()
        loss_host[slot].copy_(loss.detach(), non_blocking=True)     # D2H read of the step's result
        loss_ready[slot].record()
        if i > 0:                                                    # consume the previous step's loss
            loss_ready[1 - slot].synchronize()
            e2e_losses.append(float(loss_host[1 - slot]))
        if last:
            loss_ready[slot].synchronize()
            e2e_losses.append(float(loss_host[slot]))

    for i in range(max(args.warmup, 3)):
        resident_step(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- end to end first (the headline): host buffers, copies inside the timed region. The two timed
    # regions run back to back on a GPU that keeps warming up under its power cap (the same K steps
    # read ~3 % slower a few seconds later), so the order is stated in `config`.
    for ev in consumed:
        ev.record()
    torch.cuda.synchronize()
    for i in range(2):
        e2e_step(i, last=(i == 1))
    torch.cuda.synchronize()
    e2e_losses.clear()
    ms_e2e = timed(lambda i: e2e_step(i, last=(i == args.steps - 1)), args.steps)
    assert len(e2e_losses) == args.steps
    # ---- timed region: device-resident inputs, every GEMM-engine launch bracketed by events
    lib.egm_prof_reset()
    # level 2: one pair of events around each Newton-Schulz chain (12 launches), so the launches inside
    # keep their programmatic-dependent-launch overlap; level 1 (--gemm-breakdown) brackets every launch
    per_launch = args.gemm_breakdown or args.algorithm != "dense"
    lib.egm_prof_enable(1 if per_launch else 2)
    l0 = lib.egm_launch_count()
    ms_total = timed(resident_step, args.steps)
    launches = (lib.egm_launch_count() - l0) / args.steps
    lib.egm_prof_enable(0)
    prof = read_prof()
    lib.egm_prof_reset()
    # the Newton-Schulz chain = the D x D x D products (dense algorithm)
    ns = [r for r in prof if r[2][0] == D_IN and r[2][1] == D_IN and r[2][2] == D_IN]
    n_of = lambda r: -r[2][3] if r[2][3] < 0 else 1          # a chain record stands for that many launches
    ns_ms = sum(r[0] for r in ns) / args.steps
    ns_flops = sum(r[1] for r in ns) / args.steps
    ns_launches = sum(n_of(r) for r in ns) / args.steps
    gemm_ms = sum(r[0] for r in prof) / args.steps
    gemm_flops = sum(r[1] for r in prof) / args.steps
    gemm_launches = sum(n_of(r) for r in prof) / args.steps
    if args.gemm_breakdown and rank == 0:
        by = {}
        for ms_i, fl_i, d in prof:
            e = by.setdefault(d, [0, 0.0, 0.0])
            e[0] += 1; e[1] += ms_i; e[2] += fl_i
        for d, (n, ms_i, fl_i) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            print(f"gemm M={d[0]} N={d[1]} K={d[2]}+{d[3]} batch={d[4]} x{d[5]}: {n / args.steps:.1f}/step "
                  f"{ms_i / args.steps:.3f} ms/step {fl_i / ms_i / 1e9:.0f} TFLOP/s algorithmic", file=sys.stderr)
    clocks = sampler.stop() if rank == 0 else None

    ms_step = ms_total / args.steps
    value = B * world / (ms_step * 1e-3)
    e2e_value = B * world / (ms_e2e / args.steps * 1e-3)

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras and args.algorithm == "dense":
        other = "bf16" if args.precision != "bf16" else "fp32"
        EF.set_precision(other)
        for i in range(3):
            resident_step(i)
        ms_o = timed(resident_step, max(5, args.steps // 2)) / max(5, args.steps // 2)
        extras[other] = {"value": B / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o,
                         "tolerance": "rel 2e-2 on outputs" if other == "bf16" else "rel 1e-3"}
        EF.set_precision(args.precision)
        # opt-in algorithmic variant: Newton-Schulz on N x N matrices
        EF.set_ns_algorithm("lowrank")
        for mode in (args.precision, other):
            EF.set_precision(mode)
            for i in range(3):
                resident_step(i)
            ms_o = timed(resident_step, max(5, args.steps // 2)) / max(5, args.steps // 2)
            extras[f"lowrank_{mode}"] = {"value": B / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o,
                                         "note": "same function; iSQRT-COV evaluated in N x N low-rank form"}
        EF.set_ns_algorithm("dense")
        EF.set_precision(args.precision)
        # "PyTorch on B200" comparator: the reference's algorithm as stock torch ops with
        # autograd (baseline/torch_eager.py, pinned to the oracle) on the same GPU, same inputs and step
        try:
            extras.update(torch_eager_rates(B, dev, gpf, head, dev_inputs, d_out, timed))
        except Exception as exc:          # a comparator must never take the benchmark line down
            extras["torch_eager_error"] = repr(exc)[:200]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    if not peak_tf:
        peak_tf, peak_src = 1400.0, "B200_PROFILING.md sustained fallback (of fallback)"
    passes = 3 if args.precision == "fp32" else 1
    dense = args.algorithm == "dense" and ns_ms > 0
    k_ms, k_flops, k_n = (ns_ms, ns_flops, ns_launches) if dense else (gemm_ms, gemm_flops, gemm_launches)
    achieved = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get("gemm_tc2_kernel_dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    roofline = {
        "bound": "tensor",
        "kernel": "gemm_tc2_kernel<%d> (%s, per GPU)" % (passes, "Newton-Schulz chain fwd+bwd: the D x D x D products"
                                                        if dense else "all tcgen05 GEMM launches of the step"),
        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic,
        "peak_source": peak_src,
        "how": ("CUDA events on the launching stream inside the timed steps (egm_prof_*): " +
                ("one pair per launch" if per_launch else
                 "one pair around each chain of 12 launches; avg_launch_ms = chain time / launches")),
        "launches_per_step": k_n, "kernel_ms_per_step": k_ms,
        "avg_launch_ms": (k_ms / k_n) if k_n else None,
        "algorithmic_flops_per_step": k_flops,
        "share_of_step": k_ms / ms_step,
        "timed_gemm_launches": {"per_step": gemm_launches, "ms_per_step": gemm_ms,
                                "algorithmic_tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
                                "note": "every engine launch with --gemm-breakdown, else the chains only"},
        "mma_passes_per_product": passes,
        "executed_tflops": achieved * passes if achieved else None,
        "frac_executed": (achieved * passes / peak_tf) if achieved else None,
    }
    cpu = None
    if world == 1 and not args.no_extras:
        rate, sample = cpu_port_rate()
        cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": sample}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32" if args.precision != "bf16" else "bf16",
        "data": "synthetic",
        "config": workload_config(args, {
            "precision": {"fp32": "fp32 via bf16 hi/lo split, 3 tcgen05 MMAs per product, fp32 accumulate",
                          "bf16": "single bf16 tcgen05 MMA, fp32 accumulate",
                          "fp32_simt": "fp32 FFMA"}[args.precision],
            "ns_algorithm": args.algorithm,
            "l2": "two rotating 310 MB input sets per GPU (> 126 MB L2); no explicit flush",
            "order": "warm-up, e2e region, device-resident region, side modes"}),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    if extras:
        line["other_modes"] = extras
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
