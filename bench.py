#!/usr/bin/env python
"""Benchmark of the moment-pooling hot path: GPF + MomentHead forward+backward, images/s.

    python bench.py --gpus N --steps K --warmup W [--precision fp32|bf16] [--impl reference]
                    [--config 2|3|4|5] [--degree P Q] [--algorithm dense|lowrank]

Default workload (BASELINE.json configs[1], --config 2): B=256 images per GPU, N=197 tokens,
D=768 -> d_out=256, GPF degree (3,3) cosine, 2nd-order iSQRT-COV with 5 Newton-Schulz iterations,
train mode (BatchNorm batch statistics, Dropout 0.1), synthetic ViT-B/16 token tensors, random-init
weights. One "step" = forward + backward (+ gradient all-reduce for N>1) + SGD update over one
batch. Weak scaling: the per-GPU batch is fixed, `value` is the whole-job images/s.
--config 3 adds the third-order Tensor-Sketch branch (S=8192), --config 5 is the Swin-B token shape
(N=144, D=1024; sweep --degree), --config 4 the full dual-view training step (baseline/dual_view_step.py).

  value      inputs already resident in HBM (two rotating input sets, 310 MB each > 126 MB L2)
  e2e        same step through the public nn.Module API with HOST (pinned) token buffers:
             H2D copy of both token tensors (double-buffered on a copy stream) + D2H read of the
             loss of every step, all inside the timed region. Timed FIRST, `value` right after it:
             under the 1 kW power cap the same K steps read ~3 % slower a few seconds later
  roofline   the dominant kernel (tcgen05 GEMM engine, Newton-Schulz chain), timed live with
             CUDA events inside the timed steps, against MEASURED_PEAKS.json; `frac` counts the flops
             this path evaluates, `frac_survey` the (15+30) x 2D^3 of SURVEY.md 8(d)
  per_rank   every rank's own step / chain times (the line's time is the max over ranks)
  cpu_baseline  the reference's CPU path on the host cores (bounded sample), N=1 rank 0 only:
             the reference's own modules when its sources are found (/root/reference, baseline/_ref),
             kind "reference"; else the numpy oracle port, kind "port"

`--impl reference` times that CPU path alone, on every host core (torchrun's OMP_NUM_THREADS=1 is
overridden and the thread count is printed).
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
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json workload (1-based): 2 = configs[1] (default, the metric's config); "
                         "3 = + third-order Tensor-Sketch S=8192; 4 = full dual-view training step; "
                         "5 = Swin-B shape N=144, D=1024 (combine with --degree for the sweep)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (256; 64 for --config 4)")
    ap.add_argument("--algorithm", default="dense", choices=["dense", "lowrank"],
                    help="iSQRT-COV evaluation: D x D Newton-Schulz chain, or the N x N low-rank form")
    ap.add_argument("--tokens", type=int, default=None, help="tokens per image (197 ViT-B/16; 144 Swin-B 384 px)")
    ap.add_argument("--d-in", type=int, default=None, help="token width (768 ViT-B; 1024 Swin-B)")
    ap.add_argument("--degree", type=int, nargs=2, default=None, metavar=("P", "Q"),
                    help="GPF polynomial degrees (BASELINE configs[4] sweeps (1,1)...(3,3))")
    ap.add_argument("--third-order", action="store_true", help="add the 3rd-order Tensor-Sketch branch")
    ap.add_argument("--sketch-dim", type=int, default=0, help="sketch width S (8192 for configs[2])")
    ap.add_argument("--no-extras", action="store_true", help="skip side modes / comparators / cpu baseline")
    ap.add_argument("--no-allreduce", action="store_true", help="diagnostic: skip the gradient all-reduce (N>1)")
    ap.add_argument("--no-overlap", action="store_true", help="diagnostic: all-reduce after backward, no hooks")
    ap.add_argument("--bucket-mb", type=int, default=32, help="gradient bucket size in MB (grouping of small gradients)")
    ap.add_argument("--chunk-mb", type=int, default=0,
                    help="diagnostic: split every gradient all-reduce into pieces of this many MB (0 = one launch)")
    ap.add_argument("--nccl-normal-priority", action="store_true",
                    help="diagnostic: NCCL on a normal-priority stream (default: high priority)")
    ap.add_argument("--backbone-amp", action="store_true",
                    help="--config 4 only: run the ViT backbone under bf16 autocast (the reference ships amp: false)")
    ap.add_argument("--gemm-breakdown", action="store_true",
                    help="print the per-shape time of the GEMM-engine launches to stderr")
    args = ap.parse_args()
    if args.config == 3:
        args.third_order = True
        args.sketch_dim = args.sketch_dim or 8192
    if args.third_order and not args.sketch_dim:
        args.sketch_dim = 2048
    if args.config == 5:
        args.tokens, args.d_in = args.tokens or 144, args.d_in or 1024
    args.tokens = args.tokens or N_TOK
    args.d_in = args.d_in or D_IN
    args.degree = list(args.degree or ((2, 2) if args.config == 4 else (DEG, DEG)))
    args.batch = args.batch or (64 if args.config == 4 else 256)
    N_TOK, D_IN, DEGS = args.tokens, args.d_in, tuple(args.degree)   # the named workload unless overridden
    return args


def workload_name(args):
    shape = (N_TOK, D_IN, tuple(args.degree))
    if args.config == 4:
        return "configs[3]"
    if args.third_order:
        return "configs[2]" if shape == (197, 768, (DEG, DEG)) and args.sketch_dim == 8192 else "custom shape"
    if shape == (197, 768, (DEG, DEG)):
        return "configs[1]"
    if shape[:2] == (144, 1024):
        return "configs[4]"
    return "custom shape"


def workload_config(args, extra=None):
    order = "2nd-order iSQRT-COV" + (f" + 3rd-order Tensor-Sketch S={args.sketch_dim}" if args.third_order else "")
    cfg = {
        "workload": workload_name(args) +
                    f": MomentHead+GPFKernel (degree {args.degree[0]},{args.degree[1]}, {order}, "
                    f"5 NS iters) fwd+bwd, B={args.batch}/GPU, N={N_TOK}, D={D_IN}->d_out={D_OUT}",
        "per_gpu_batch": args.batch, "global_batch": args.batch * args.gpus, "tokens": N_TOK,
        "d_in": D_IN, "d_out": D_OUT, "gpf_degree": list(args.degree), "ns_iterations": NS_ITERS,
        "parallelism": f"dp{args.gpus} (batch sharded per image, NCCL gradient all-reduce)",
        "step": "forward + backward + grad all-reduce + SGD update, train-mode BN, dropout 0.1",
    }
    if args.third_order:
        cfg["third_order"] = {"sketch_dim": args.sketch_dim,
                              "oracle_patch": ("head.tensor_sketch.sketch_dim = S (SURVEY.md 8c: the reference "
                                               "caps the attribute at 4*d_in and then indexes out of bounds)")
                              if args.sketch_dim > 4 * D_IN else None}
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
def host_threads():
    """All host cores for the CPU arm. torchrun exports OMP_NUM_THREADS=1 when N>1, which silently
    made round 1's N>1 reference arm single-threaded; set the pools explicitly and report them."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)          # numpy's BLAS pool (the oracle port's elementwise/BLAS calls)
    except Exception:
        pass
    return torch.get_num_threads()


class CpuReferenceStep:
    """One training step of the path on the host CPU, fp32, same shape/seed recipe as the GPU arm.

    kind "reference": the reference's OWN modules (gpf_kernel.py / moment_head.py loaded unchanged by
    file path from /root/reference or baseline/_ref - baseline/reference_loader.py) with autograd + SGD.
    kind "port"     : the numpy/torch-bmm oracle restatement, only when neither location exists."""

    def __init__(self, third_order=False, sketch_dim=0):
        import torch
        from baseline import reference_loader as RL
        self.third, self.S = third_order, sketch_dim
        self.root = RL.find_reference_root()
        self.kind = "reference" if self.root else "port"
        g = torch.Generator().manual_seed(1234)
        self.anchor = torch.randn(8, N_TOK, D_IN, generator=g)
        self.positive = self.anchor + 0.5 * torch.randn(8, N_TOK, D_IN, generator=g)
        self.dout = torch.randn(8, D_OUT, generator=torch.Generator().manual_seed(4321))
        if self.kind == "reference":
            gk, mh, _ = RL.load_path_modules(self.root)
            torch.manual_seed(0)
            self.gpf = gk.GraphPolynomialFusion(*DEGS)
            self.head = mh.MomentHead(D_IN, D_OUT, use_third_order=third_order, isqrt_iterations=NS_ITERS,
                                      sketch_dim=sketch_dim or 2048).train()
            if third_order and sketch_dim > 4 * D_IN:
                self.head.tensor_sketch.sketch_dim = sketch_dim       # SURVEY.md 8c instance patch
            self.opt = torch.optim.SGD(list(self.gpf.parameters()) + list(self.head.parameters()), lr=1e-6)
            self.where = f"reference modules from {self.root}/src/models (unmodified, autograd, SGD)"
        else:
            import numpy as np
            from oracle import moment_oracle as O
            O.set_matmul_backend("torch")       # the threaded torch.bmm the reference itself runs on
            pkg = importlib.import_module("ego-moment-cle-vit_b200")
            torch.manual_seed(0)
            gpf = pkg.GraphPolynomialFusion(*DEGS)
            head = pkg.MomentHead(D_IN, D_OUT, use_third_order=False, isqrt_iterations=NS_ITERS)
            self.O, self.np = O, np
            self.alpha = gpf.alpha_coeffs.detach().numpy()
            self.params = {k: v.detach().numpy() for k, v in head.state_dict().items()}
            self.where = "numpy/torch-CPU-bmm fp32 port of the reference algorithm (oracle/moment_oracle.py)"

    def __call__(self, bs):
        if self.kind == "reference":
            a = self.anchor[:bs].clone().requires_grad_(True)
            p = self.positive[:bs].clone().requires_grad_(True)
            out = self.head(a, self.gpf(a, p))
            (out * self.dout[:bs, :out.shape[1]]).sum().backward()
            self.opt.step()
            self.opt.zero_grad(set_to_none=True)
        else:
            self.O.path_step(self.anchor[:bs].numpy(), self.positive[:bs].numpy(), self.alpha, self.params,
                             NS_ITERS, self.dout[:bs].numpy(), True, self.np.float32)


def cpu_reference_rate(args, budget_s=20.0, max_steps=8):
    """images/s of the reference CPU path on the host cores: a bounded sample of the same workload."""
    threads = host_threads()
    ref = CpuReferenceStep(getattr(args, "third_order", False), getattr(args, "sketch_dim", 0))
    bs = 8
    ref(2)                                # warm-up
    t0 = time.perf_counter()
    n = 0
    while n < max_steps:
        ref(bs)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    sample = (f"{n} step(s) of B={bs} images (N={N_TOK}, D={D_IN}), fwd+bwd+SGD, fp32, {threads} threads, "
              f"{dt:.1f} s; {ref.where}")
    return bs * n / dt, sample, ref.kind, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    ref = CpuReferenceStep(args.third_order, args.sketch_dim)
    t0 = time.perf_counter()
    ref(2)
    per_img = (time.perf_counter() - t0) / 2
    # bound the whole run to a few minutes
    total = max(1, args.steps + args.warmup)
    bs = max(1, min(8, int(150.0 / (per_img * total))))
    for _ in range(args.warmup):
        ref(bs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref(bs)
    dt = time.perf_counter() - t0
    value = bs * args.steps / dt
    sample = (f"{args.steps} steps of B={bs} images of the same workload, fwd+bwd+SGD fp32 on {threads} host "
              f"threads (OMP_NUM_THREADS env was {os.environ.get('OMP_NUM_THREADS')!r}, overridden); {ref.where}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, {"sample": sample}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": ref.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------- "PyTorch on the same B200" comparator
def gpu_comparator_rates(args, B, dev, gpf, head, dev_inputs, d_out, timed, steps=3):
    """images/s of the same training step through the stock framework on the same GPU and inputs
    (SURVEY.md 8d): the reference's OWN modules moved to CUDA when its sources are present
    (/root/reference or baseline/_ref), else the torch-eager restatement (baseline/torch_eager.py).
    fp32 with TF32 off (the reference's setting) and with TF32 allowed."""
    import copy
    import torch
    from baseline import reference_loader as RL
    root = RL.find_reference_root()
    if root:
        gk, mh, _ = RL.load_path_modules(root)
        torch.manual_seed(0)
        rgpf = gk.GraphPolynomialFusion(*DEGS).to(dev)
        rhead = mh.MomentHead(D_IN, D_OUT, use_third_order=args.third_order, isqrt_iterations=NS_ITERS,
                              sketch_dim=args.sketch_dim or 2048).to(dev).train()
        if args.third_order and args.sketch_dim > 4 * D_IN:
            rhead.tensor_sketch.sketch_dim = args.sketch_dim
        rgpf.load_state_dict(gpf.state_dict())
        rhead.load_state_dict(head.state_dict())
        params = list(rgpf.parameters()) + list(rhead.parameters())
        fwd = lambda a, p: rhead(a, rgpf(a, p))
        what = f"the reference's own modules from {root}/src/models on the same GPU (autograd, cuBLAS)"
        tag = "reference_on_b200"
    else:
        from baseline import torch_eager as TE
        alpha = torch.nn.Parameter(gpf.alpha_coeffs.detach().clone())
        net = copy.deepcopy(head.second_net).to(dev).train()
        params = [alpha] + list(net.parameters())
        fwd = lambda a, p: TE.head_forward(a, TE.gpf_forward(a, p, alpha), net, NS_ITERS)
        what = "reference algorithm restated as stock torch ops + autograd on the same GPU (reference sources absent)"
        tag = "torch_eager"
    opt = torch.optim.SGD(params, lr=1e-6)

    def step(i):
        a, p = dev_inputs[i % 2]
        a = a.detach().requires_grad_(True)
        p = p.detach().requires_grad_(True)
        out = fwd(a, p)
        (out * d_out[:, :out.shape[1]]).sum().backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    out = {}
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32 in ((tag + "_fp32", False), (tag + "_tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            step(0)
            ms = timed(step, steps)[0] / steps
            out[name] = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                         "note": what + ", " + ("allow_tf32=True" if tf32 else
                                                "fp32 (allow_tf32=False, the reference's setting)")}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


def load_peaks():
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
    return peak_tf, peak_src, peaks


def read_prof(lib):
    import ctypes
    recs = []
    ms, fl, dims = ctypes.c_float(), ctypes.c_double(), (ctypes.c_int * 6)()
    for i in range(lib.egm_prof_count()):
        if lib.egm_prof_read(i, ctypes.byref(ms), ctypes.byref(fl), dims) == 0:
            recs.append((ms.value, fl.value, tuple(dims)))
    return recs


def dist_setup(high_priority=True):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # a 302 MB all-reduce needs few CTAs over NVLink/NVSwitch; every CTA it holds is an SM the
        # persistent GEMM CTAs of the Newton-Schulz backward cannot use (8 GPUs: 17.06 -> 16.99 ms/step)
        os.environ.setdefault("NCCL_MAX_CTAS", "8")
        # NCCL's kernels on a high-priority stream: when SMs free up at a kernel boundary the all-reduce's
        # CTAs are placed first instead of queueing behind the persistent GEMM CTAs of the next launch
        opts = None
        if high_priority:
            try:
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            except Exception:
                opts = None
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    return world, rank, local, dev


def make_timed(dev, world):
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between CUDA events on the current stream. Returns (total ms = MAX over ranks,
        this rank's total ms, this rank's per-step ms list): one event after every step, no host sync."""
        barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        for i in range(steps):
            fn(i)
            evs[i + 1].record()
        torch.cuda.synchronize()
        mine = evs[0].elapsed_time(evs[steps])
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        ms = torch.tensor([mine], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item()), mine, per

    return timed, barrier


def gather_ranks(dev, world, values):
    """[world][len(values)] list of every rank's numbers."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, device=dev, dtype=torch.float64)
    if world == 1:
        return [t.tolist()]
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


# ------------------------------------------------------------------------------ native leg
def run_native(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("ego-moment-cle-vit_b200")
    EF = pkg.functional
    egm_dist = importlib.import_module("ego-moment-cle-vit_b200.dist")
    lib = pkg._lib.load()

    world, rank, local, dev = dist_setup(not args.nccl_normal_priority)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    B = args.batch
    EF.set_precision(args.precision)
    EF.set_ns_algorithm(args.algorithm)
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(args.degree[0], args.degree[1]).to(dev)
    head = pkg.MomentHead(D_IN, D_OUT, use_third_order=args.third_order, isqrt_iterations=NS_ITERS,
                          sketch_dim=args.sketch_dim or 2048)
    if args.third_order and args.sketch_dim > 4 * D_IN:
        head.tensor_sketch.sketch_dim = args.sketch_dim       # SURVEY.md 8c instance patch, as for the oracle
    head = head.to(dev).train()
    egm_dist.broadcast_parameters(gpf)
    egm_dist.broadcast_parameters(head)
    params = list(gpf.parameters()) + list(head.parameters())
    buckets = None
    if not args.no_allreduce:
        buckets = egm_dist.GradBuckets(params, bucket_bytes=args.bucket_mb << 20, overlap=not args.no_overlap,
                                       chunk_bytes=(args.chunk_mb << 20) or None)
    opt = torch.optim.SGD(params, lr=1e-6)

    # two rotating input sets, each 2 x B x N x D x 4 B (310 MB at the configs[1] shape, > 126 MB of L2)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    dev_inputs = []
    for _ in range(2):
        a = torch.randn(B, N_TOK, D_IN, device=dev, generator=gen)
        dev_inputs.append((a, a + 0.5 * torch.randn(B, N_TOK, D_IN, device=dev, generator=gen)))
    host_inputs = [(a.cpu().pin_memory(), p.cpu().pin_memory()) for a, p in dev_inputs]
    stage = [(torch.empty_like(a), torch.empty_like(p)) for a, p in dev_inputs[:1]]
    d_out = torch.randn(B, D_OUT, device=dev, generator=gen)
    h2d_bytes = 2 * B * N_TOK * D_IN * 4

    def step(a, p):
        a = a.requires_grad_(True)
        p = p.requires_grad_(True)
        out = head(a, gpf(a, p))
        loss = (out * d_out).sum()
        loss.backward()
        if buckets is not None:
            buckets.reduce()
        opt.step()
        opt.zero_grad(set_to_none=True)
        a.requires_grad_(False); p.requires_grad_(False)
        a.grad = None; p.grad = None
        return loss

    timed, barrier = make_timed(dev, world)

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
    ms_e2e, _, per_e2e = timed(lambda i: e2e_step(i, last=(i == args.steps - 1)), args.steps)
    assert len(e2e_losses) == args.steps
    # ---- timed region: device-resident inputs, every GEMM-engine launch bracketed by events
    lib.egm_prof_reset()
    # level 2: one pair of events around each Newton-Schulz chain (12 launches), so the launches inside
    # keep their programmatic-dependent-launch overlap; level 1 (--gemm-breakdown) brackets every launch
    per_launch = args.gemm_breakdown or args.algorithm != "dense"
    lib.egm_prof_enable(1 if per_launch else 2)
    l0 = lib.egm_launch_count()
    ms_total, ms_mine, per_step = timed(resident_step, args.steps)
    launches = (lib.egm_launch_count() - l0) / args.steps
    lib.egm_prof_enable(0)
    prof = read_prof(lib)
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
    # every rank's own numbers (the line's time is the max over ranks): step mean / median, chain time
    ranks = gather_ranks(dev, world, [ms_mine / args.steps, statistics.median(per_step), ns_ms,
                                      statistics.median(per_e2e)])

    ms_step = ms_total / args.steps
    value = B * world / (ms_step * 1e-3)
    e2e_value = B * world / (ms_e2e / args.steps * 1e-3)
    peak_tf, peak_src, _ = load_peaks()
    passes = 3 if args.precision == "fp32" else 1

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras and args.algorithm == "dense":
        other = "bf16" if args.precision != "bf16" else "fp32"
        EF.set_precision(other)
        for i in range(3):
            resident_step(i)
        n_side = max(5, args.steps // 2)
        ms_o = timed(resident_step, n_side)[0] / n_side
        extras[other] = {"value": B / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o,
                         "tolerance": ("rel 2e-2 on outputs; token gradients measured 3e-2 (tests gate 6e-2)"
                                       if other == "bf16" else "rel 1e-3")}
        EF.set_precision(args.precision)
        # opt-in algorithmic variant: Newton-Schulz on N x N matrices. A first-class
        # record: its own roofline (all tcgen05 launches of the step, one event pair per launch) and its
        # difference from the dense evaluation on the same inputs (the oracle parity is in tests/).
        def probe():
            a, p = dev_inputs[0]
            a = a.detach().clone().requires_grad_(True)
            p = p.detach().clone().requires_grad_(True)
            drops = [m for m in head.modules() if isinstance(m, torch.nn.Dropout)]
            prev = [m.p for m in drops]
            for m in drops:
                m.p = 0.0
            lin = head.second_net[0]
            with torch.no_grad():
                pre = EF.moment_head_linear(a, gpf(a, p), lin.weight, lin.bias, NS_ITERS, eps=head.eps,
                                            third_order=args.third_order)
                pre = pre[0] if isinstance(pre, tuple) else pre
            out = head(a, gpf(a, p))
            (out * d_out).sum().backward()
            for m, q in zip(drops, prev):
                m.p = q
            res = [pre.clone(), out.detach().clone(), a.grad.clone(), p.grad.clone(), gpf.alpha_coeffs.grad.clone()]
            for q in params:
                q.grad = None
            return res

        rel = lambda x, y: float((x - y).norm() / y.norm())
        # both evaluations back to back on identical parameters (no optimizer step in between)
        dense_ref = probe()
        EF.set_ns_algorithm("lowrank")
        lowrank_probe = probe()
        EF.set_ns_algorithm("dense")
        lowrank_parity = {
            "y_before_bn": rel(lowrank_probe[0], dense_ref[0]),
            "y_after_train_bn": rel(lowrank_probe[1], dense_ref[1]),
            "d_anchor": rel(lowrank_probe[2], dense_ref[2]), "d_positive": rel(lowrank_probe[3], dense_ref[3]),
            "d_alpha": rel(lowrank_probe[4], dense_ref[4]),
            "note": "rel. Frobenius difference between the two evaluations of the same function on the benchmark "
                    "inputs and identical parameters (B per GPU, dropout off); train-mode BatchNorm over iid "
                    "tokens has near-degenerate batch statistics and amplifies any upstream difference "
                    "(SURVEY.md 0.8), which is why the pre-BN figure is listed beside it"}
        EF.set_ns_algorithm("lowrank")
        for mode in (args.precision, other):
            EF.set_precision(mode)
            for i in range(3):
                resident_step(i)
            lib.egm_prof_reset(); lib.egm_prof_enable(1)
            ms_o = timed(resident_step, n_side)[0] / n_side
            lib.egm_prof_enable(0)
            pr = read_prof(lib); lib.egm_prof_reset()
            g_ms = sum(r[0] for r in pr) / n_side
            g_fl = sum(r[1] for r in pr) / n_side
            mp = 3 if mode == "fp32" else 1
            rec = {"value": B / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o,
                   "note": "same function; iSQRT-COV evaluated in N x N low-rank form (SURVEY.md 8f row 2)",
                   "roofline": {"bound": "tensor", "kernel": f"gemm_tc2_kernel<{mp}>, all launches of the step",
                                "launches_per_step": len(pr) / n_side, "kernel_ms_per_step": g_ms,
                                "share_of_step": g_ms / ms_o,
                                "achieved": g_fl / (g_ms * 1e-3) / 1e12 if g_ms else None, "peak": peak_tf,
                                "unit": "TFLOP/s", "frac": g_fl / (g_ms * 1e-3) / 1e12 / peak_tf if g_ms else None,
                                "executed_tflops": mp * g_fl / (g_ms * 1e-3) / 1e12 if g_ms else None,
                                "note": "event pair per launch (serialises PDL overlap); short-K products are "
                                        "epilogue/HBM-bound, so the tensor fraction is low by construction"}}
            if mode == args.precision:
                rec["parity_vs_dense"] = lowrank_parity
            extras[f"lowrank_{mode}"] = rec
        EF.set_ns_algorithm("dense")
        EF.set_precision(args.precision)
        # "PyTorch on B200" comparator
        try:
            extras.update(gpu_comparator_rates(args, B, dev, gpf, head, dev_inputs, d_out, timed))
        except Exception as exc:          # a comparator must never take the benchmark line down
            extras["gpu_comparator_error"] = repr(exc)[:300]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    dense = args.algorithm == "dense" and ns_ms > 0
    k_ms, k_flops, k_n = (ns_ms, ns_flops, ns_launches) if dense else (gemm_ms, gemm_flops, gemm_launches)
    achieved = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
    # SURVEY.md 8(d) accounting: (15 + 30) * 2 D^3 per image "actually required" by the minimal
    # bit-identical schedule, whatever this implementation really evaluates (fewer: commuting + symmetric)
    survey_flops = (4 * NS_ITERS - 5 + 8 * (NS_ITERS - 2) + 6) * 2.0 * D_IN ** 3 * B
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
        "frac_survey": (survey_flops / (k_ms * 1e-3) / 1e12 / peak_tf) if dense else None,
        "frac_survey_note": "SURVEY.md 8(d): (15+30) x 2 D^3 x B flops of the minimal bit-identical schedule / "
                            "chain time / peak; `frac` counts only the flops this path evaluates (12+24 "
                            "products on the upper 6/9 tiles)",
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
        rate, sample, kind, threads = cpu_reference_rate(args)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
    col = lambda j: [r[j] for r in ranks]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "ms_per_step_median": max(col(1)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32" if args.precision != "bf16" else "bf16",
        "data": "synthetic",
        "config": workload_config(args, {
            "precision": {"fp32": "fp32 via bf16 hi/lo split, 3 tcgen05 MMAs per product, fp32 accumulate",
                          "bf16": "single bf16 tcgen05 MMA, fp32 accumulate",
                          "fp32_simt": "fp32 FFMA"}[args.precision],
            "ns_algorithm": args.algorithm,
            "l2": "two rotating input sets per GPU (310 MB each at the configs[1] shape, > 126 MB L2); no explicit flush",
            "order": "warm-up, e2e region, device-resident region, side modes",
            "allreduce": ("off (diagnostic)" if args.no_allreduce else
                          (f"{args.chunk_mb} MB chunks" if args.chunk_mb else "one launch per gradient bucket") +
                          f", NCCL AVG (NCCL_MAX_CTAS={os.environ.get('NCCL_MAX_CTAS')}) on a " +
                          ("normal" if args.nccl_normal_priority else "high") + "-priority stream, " +
                          ("after backward" if args.no_overlap else
                           "dW of the Linear handed over before the Newton-Schulz backward"))}),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps,
                "ms_per_step_median": max(col(3))},
        "gpu_launches": launches,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "per_rank": {"step_ms_mean": col(0), "step_ms_median": col(1), "ns_chain_ms": col(2),
                     "note": "each rank's own CUDA-event times; the line's ms_per_step is the max over ranks"},
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
    elif args.config == 4:
        from baseline import dual_view_step
        dual_view_step.run(args, sys.modules[__name__])
    else:
        run_native(args)


if __name__ == "__main__":
    main()
