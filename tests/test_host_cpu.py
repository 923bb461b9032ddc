"""Host-side logic that needs no GPU: C-ABI surface, module construction / state_dict / RNG
parity with the reference, error behaviour, sketch tables, batch sharding under gloo."""
import ctypes
import importlib.util
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden, load_pkg
from oracle import moment_oracle as O

REF = "/root/reference"


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "egm_b200.h")).read()
    declared = set(re.findall(r"\b(egm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 28
    lib = ctypes.CDLL(pkg._lib.lib_path())
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/egm_b200.h but not exported"
    # the ctypes signature table covers the header exactly
    assert set(pkg._lib.SIGNATURES) == declared
    L = pkg._lib.load()
    assert L.egm_version() == 100
    assert L.egm_gpf_ldr(197) == 200
    # size queries are pure host arithmetic: callable without a GPU
    assert L.egm_ns_state_bytes(2, 768, 5, 1) >= 13 * 2 * 768 * 768 * 4
    assert L.egm_ns_state_bytes(2, 768, 1, 1) < L.egm_ns_state_bytes(2, 768, 2, 1)


def test_argument_errors_are_reported_without_a_gpu(pkg):
    L = pkg._lib.load()
    rc = L.egm_gpf_fwd(None, None, None, 1, 1, 1, 1, 1, 1, 1e-6, 1, None, None, None, None, None, None, 1, None, 0, None)
    assert rc == -1 and b"null pointer" in L.egm_last_error()
    rc = L.egm_triu_pack(None, 0, 0, None, None)
    assert rc == -1
    with pytest.raises(ValueError):
        pkg._lib.precision_id("fp8")
    assert pkg._lib.precision_id("fp32") == 1 and pkg._lib.precision_id("bf16") == 2


def test_no_cpu_fallback(pkg):
    gpf = pkg.GraphPolynomialFusion(2, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gpf(torch.randn(2, 5, 8), torch.randn(2, 5, 8))
    head = pkg.MomentHead(8, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        head(torch.randn(2, 5, 8), torch.rand(2, 5, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.NewtonSchulzSqrtm(3)(torch.eye(4).expand(2, 4, 4))


def test_gpf_constructor_contract(pkg):
    g = pkg.GraphPolynomialFusion()
    assert (g.degree_p, g.degree_q, g.similarity, g.eps, g.symmetric_enforce, g.num_terms) == \
        (2, 2, "cosine", 1e-6, True, 9)
    assert list(g.state_dict().keys()) == ["alpha_coeffs"] and g.alpha_coeffs.shape == (3, 3)
    assert pkg.GPFKernel is pkg.GraphPolynomialFusion
    with pytest.raises(ValueError, match="Unknown initialization method"):
        pkg.GraphPolynomialFusion(coeff_init="bogus")
    ident = pkg.GraphPolynomialFusion(2, 2, coeff_init="identity")
    assert ident.alpha_coeffs[0, 0] == 0.5 and ident.alpha_coeffs[1, 1] == 0.5 and ident.alpha_coeffs[0, 1] == 0.01
    # similarity is validated lazily, at forward, before anything touches the device
    bad = pkg.GraphPolynomialFusion(similarity="euclid")
    with pytest.raises(ValueError, match="Unknown similarity function"):
        bad(torch.randn(1, 3, 4), torch.randn(1, 3, 4))
    c = g.get_coefficient_matrix()
    assert torch.allclose(c, torch.nn.functional.softplus(g.alpha_coeffs))
    assert torch.allclose(g.get_sparsity_loss(0.5), 0.5 * c.abs().sum())
    att = pkg.AdaptiveGraphPolynomialFusion(1, 1, adaptive_type="attention")
    assert att.adaptive_type == "attention"
    assert "coeff_attention.in_proj_weight" in att.state_dict()
    assert att.coeff_attention.embed_dim == 4


def test_moment_head_contract_and_state_dict(pkg):
    h = pkg.MomentHead(d_in=64, d_out=33, use_third_order=True, isqrt_iterations=5, sketch_dim=128)
    assert [n for n, _ in h.named_children()] == ["isqrt_cov", "tensor_sketch", "second_net", "third_net"]
    assert (h.d_in, h.d_out, h.d_second, h.d_third, h.use_third_order, h.eps) == (64, 33, 16, 17, True, 1e-5)
    sd = h.state_dict()
    assert sd["tensor_sketch.hash1"].dtype == torch.int64 and sd["tensor_sketch.sign3"].shape == (64,)
    assert sd["second_net.0.weight"].shape == (16, 64 * 65 // 2)
    assert sd["third_net.0.weight"].shape == (17, 128)
    assert h.isqrt_cov.num_iterations == 5 and h.isqrt_cov.eps == 1e-5
    assert isinstance(h.second_net[3], torch.nn.Dropout) and h.second_net[3].p == 0.1
    h2 = pkg.MomentHead(d_in=16)
    assert (h2.d_out, h2.use_third_order, h2.d_third, h2.isqrt_cov.num_iterations) == (512, False, 0, 3)
    assert not hasattr(h2, "tensor_sketch")
    # state_dict round trip
    h3 = pkg.MomentHead(d_in=64, d_out=33, use_third_order=True, isqrt_iterations=5, sketch_dim=128)
    h3.load_state_dict(sd)
    assert all(torch.equal(a, b) for a, b in zip(h3.state_dict().values(), sd.values()))


def test_tensor_sketch_rng_side_effect_matches_reference(pkg):
    # known answers of the reference for seed 42
    ts = pkg.TensorSketch(768, 3072)
    assert ts.hash1[:8].tolist() == [102, 2483, 2908, 1294, 2154, 71, 700, 20]
    assert ts.sign1[:8].tolist() == [-1, -1, 1, 1, -1, -1, 1, 1]
    assert ts.sketch_dim == 3072 and pkg.TensorSketch(8, 2048).sketch_dim == 32
    # the constructor reseeds the global RNG exactly like the reference (moment_head.py:88)
    torch.manual_seed(123)
    pkg.TensorSketch(16, 32)
    after = torch.rand(3)
    torch.manual_seed(42)
    for _ in range(3):
        torch.randint(0, 32, (16,))
    for _ in range(3):
        torch.randint(0, 2, (16,))
    assert torch.equal(after, torch.rand(3))


@pytest.mark.parametrize("name", ["cfg1_b8_n197_d768", "small_p3q3_third"])
def test_initialisation_is_bit_identical_to_the_reference(pkg, name):
    rec = golden(name)
    B, N, D, P, Q, K, d_out, third, S = [int(v) for v in rec["cfg"][:9]]
    torch.manual_seed(0)
    gpf = pkg.GraphPolynomialFusion(P, Q)
    head = pkg.MomentHead(D, d_out, use_third_order=bool(third), isqrt_iterations=K, sketch_dim=S)
    assert np.array_equal(gpf.alpha_coeffs.detach().numpy().astype(rec["alpha"].dtype), rec["alpha"])
    w = head.second_net[0].weight.detach().numpy()
    assert np.array_equal(w[:2, :8].astype(rec["w_head"].dtype), rec["w_head"])
    assert abs(w.sum(dtype=np.float64) - rec["w_sum"][0]) < 1e-6 * max(1.0, abs(rec["w_sum"][0]))
    if third:
        assert np.array_equal(head.tensor_sketch.hash2.numpy(), rec["hash"][1])
        assert np.array_equal(head.tensor_sketch.sign3.numpy(), rec["sign"][2])


def test_sketch_csr_tables_reproduce_count_sketch(pkg):
    EF = pkg.functional
    g = torch.Generator().manual_seed(5)
    D, S, B = 40, 16, 3
    hashes = torch.randint(0, S, (3, D), generator=g)
    signs = torch.randint(0, 2, (3, D), generator=g) * 2 - 1
    off, idx, sgn = EF.build_sketch_csr(hashes, signs, S)
    x = torch.randn(B, D, generator=g).double().numpy()
    for h in range(3):
        ref = O.count_sketch(x, hashes[h].numpy(), signs[h].numpy(), S)
        got = np.zeros_like(ref)
        for s in range(S):
            for e in range(int(off[h, s]), int(off[h, s + 1])):
                got[:, s] += float(sgn[h, e]) * x[:, int(idx[h, e])]
        assert np.allclose(got, ref)
        assert int(off[h, -1]) == D
    with pytest.raises(RuntimeError, match="out of bounds"):
        EF.build_sketch_csr(hashes, signs, S - 8)   # the reference's sketch_dim > 4*d_in failure


def test_precision_context(pkg):
    EF = pkg.functional
    base = EF.get_precision()
    with EF.precision("bf16"):
        assert EF.get_precision() == "bf16"
        with EF.precision("fp32_simt"):
            assert EF.get_precision() == "fp32_simt"
        assert EF.get_precision() == "bf16"
    assert EF.get_precision() == base
    with pytest.raises(ValueError):
        EF.set_precision("int4")


def test_utils_ops_surface(pkg):
    ops = importlib.import_module("ego-moment-cle-vit_b200.utils.ops")
    for name in ["set_seed", "count_parameters", "get_model_info", "print_model_info",
                 "half_vectorize_symmetric", "matrix_sqrt_newton_schulz", "matrix_power_eigen",
                 "check_psd", "ensure_psd", "normalize_graph", "compute_graph_statistics",
                 "batch_trace", "batch_logdet", "cosine_similarity_matrix", "test_ops"]:
        assert callable(getattr(ops, name))
    ops.set_seed(7)
    a = torch.rand(2)
    ops.set_seed(7)
    assert torch.equal(a, torch.rand(2))
    lin = torch.nn.Linear(4, 3)
    assert ops.count_parameters(lin) == 15
    assert ops.get_model_info(lin)["total_parameters"] == 15
    with pytest.raises(ValueError, match="Unknown normalization method"):
        ops.normalize_graph(torch.rand(1, 3, 3), "bogus")
    g = torch.rand(1, 3, 3)
    assert ops.normalize_graph(g, "none") is g
    m = torch.eye(3).unsqueeze(0) * 2.0
    assert ops.check_psd(m) and torch.allclose(ops.matrix_power_eigen(m, 0.5), m.sqrt() * torch.eye(3))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reference_integration_model_constructs_on_these_modules(pkg):
    """The reference's ego_moment_clevit.py, unmodified, wired to this package's modules
    (SURVEY.md 0.12): stub backbone (timm is absent), reference classifier head."""
    import types
    pkg.install_into("egm_src")
    root = types.ModuleType("egm_src"); root.__path__ = []
    models = types.ModuleType("egm_src.models"); models.__path__ = []
    sys.modules.update({"egm_src": root, "egm_src.models": models})

    class CLEViTDualStream(torch.nn.Module):
        def __init__(self, model_name, pretrained, drop_rate):
            super().__init__()
            self.num_features = 64
    stub = types.ModuleType("egm_src.models.cle_vit_backbone"); stub.CLEViTDualStream = CLEViTDualStream
    sys.modules["egm_src.models.cle_vit_backbone"] = stub
    for name in ("classifier_head", "ego_moment_clevit"):
        spec = importlib.util.spec_from_file_location(f"egm_src.models.{name}", f"{REF}/src/models/{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"egm_src.models.{name}"] = mod
        spec.loader.exec_module(mod)
    Model = sys.modules["egm_src.models.ego_moment_clevit"].EGOMomentCLEViT
    model = Model(num_classes=5, backbone_name="stub", pretrained=False, gpf_degree_p=2, gpf_degree_q=1,
                  moment_d_out=32, use_third_order=True, isqrt_iterations=4, sketch_dim=128)
    assert type(model.gpf).__module__.startswith("ego-moment-cle-vit_b200")
    assert type(model.moment_head).__module__.startswith("ego-moment-cle-vit_b200")
    assert (model.gpf.degree_p, model.gpf.degree_q) == (2, 1)
    assert (model.moment_head.d_out, model.moment_head.use_third_order) == (32, True)
    assert model.gpf.get_coefficient_matrix().shape == (3, 2)


# ----------------------------------------------------------------- multi-process (gloo)
def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = load_pkg()
    d = importlib.import_module("ego-moment-cle-vit_b200.dist")
    torch.manual_seed(rank)                       # ranks start different on purpose
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 2))
    d.broadcast_parameters(net, src=0)
    x = torch.arange(7 * 6, dtype=torch.float32).view(7, 6) / 10.0
    xs = d.shard(x, rank, world)
    net(xs).sum().backward()
    buckets = d.GradBuckets(net.parameters(), bucket_bytes=64)   # tiny: several buckets
    buckets.reduce()
    # plain numpy through the queue (torch tensors would be passed by file descriptor)
    q.put((rank, d.shard_bounds(7, rank, world), [p_.grad.numpy().copy() for p_ in net.parameters()],
           [p_.detach().numpy().copy() for p_ in net.parameters()], len(buckets.buckets)))
    dist.destroy_process_group()


def test_sharded_gradients_equal_mean_of_shard_gradients_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, b0, g0, w0, nb), (_, b1, g1, w1, _) = res
    g0, w0, g1, w1 = ([torch.from_numpy(a) for a in lst] for lst in (g0, w0, g1, w1))
    assert b0 == (0, 4) and b1 == (4, 7) and nb > 1
    for a, b in zip(w0, w1):
        assert torch.equal(a, b)                  # broadcast made the replicas identical
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)                  # all-reduce leaves the same gradient everywhere
    # and it is the mean of the two shard gradients
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 2))
    with torch.no_grad():
        for p_, w in zip(net.parameters(), w0):
            p_.copy_(w)
    x = torch.arange(7 * 6, dtype=torch.float32).view(7, 6) / 10.0
    gs = []
    for lo, hi in (b0, b1):
        net.zero_grad()
        net(x[lo:hi]).sum().backward()
        gs.append([p_.grad.clone() for p_ in net.parameters()])
    for got, ga, gb in zip(g0, gs[0], gs[1]):
        assert torch.allclose(got, 0.5 * (ga + gb), atol=1e-6)


def _gloo_accum_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    load_pkg()
    d = importlib.import_module("ego-moment-cle-vit_b200.dist")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 2))
    buckets = d.GradBuckets(net.parameters(), bucket_bytes=64)         # hooks on: overlap mode
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    xs = d.shard(x, rank, world)
    # (1) two backward passes without no_sync(): must raise, not silently mix local and averaged sums
    net(xs[:2]).sum().backward()
    raised = False
    try:
        net(xs[2:]).sum().backward()
    except RuntimeError as exc:
        raised = "no_sync" in str(exc)
    buckets.reduce()
    net.zero_grad(set_to_none=True)
    # (2) accumulation done right: first micro-batch under no_sync(), last one outside, then reduce()
    with buckets.no_sync():
        net(xs[:2]).sum().backward()
    net(xs[2:]).sum().backward()
    buckets.reduce()
    q.put((rank, raised, [p_.grad.numpy().copy() for p_ in net.parameters()]))
    dist.destroy_process_group()


def test_gradient_accumulation_contract_gloo():
    """a second backward before reduce() used to leave (sum_ranks g1 + local g2)/world
    in .grad without any error. Now it raises, and no_sync() gives the mean over ranks of the accumulated sum."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_accum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1], "second backward without no_sync() did not raise"
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 2))
    x = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    net(x).sum().backward()                                  # sum over all 8 rows = sum of both ranks' sums
    for got0, got1, p_ in zip(res[0][2], res[1][2], net.parameters()):
        assert np.array_equal(got0, got1)
        assert np.allclose(got0, 0.5 * p_.grad.numpy(), atol=1e-6)


def _gloo_early_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    load_pkg()
    d = importlib.import_module("ego-moment-cle-vit_b200.dist")
    EF = importlib.import_module("ego-moment-cle-vit_b200.functional")
    torch.manual_seed(0)
    big = torch.nn.Parameter(torch.zeros(4, 50))           # stands in for second_net.0.weight
    small = torch.nn.Parameter(torch.zeros(3))
    buckets = d.GradBuckets([small, big], bucket_bytes=64, chunk_bytes=256)      # hooks on, chunked
    assert EF.get_early_grad_hook() == buckets.early_reduce
    # what the fused head's backward does: dW exists, hand it over, keep computing, wait, return it
    g_big = torch.full((4, 50), float(rank + 1))
    pend = EF.get_early_grad_hook()(big.data_ptr(), g_big)
    assert pend is not None and len(pend.works) > 1         # several chunks in flight
    unknown = EF.get_early_grad_hook()(12345, torch.zeros(2))
    pend.wait()
    # autograd's accumulation (steals the tensor) + the post-accumulate hooks
    big.grad = g_big
    buckets._on_grad(big)
    small.grad = torch.full((3,), float(10 * (rank + 1)))
    buckets._on_grad(small)
    buckets.reduce()
    # under no_sync() the reducer must not take the gradient early (local sums are accumulated instead)
    with buckets.no_sync():
        declined = EF.get_early_grad_hook()(big.data_ptr(), torch.ones(4, 50))
    q.put((rank, unknown is None, declined is None, big.grad.numpy().copy(), small.grad.numpy().copy()))
    buckets.close()
    assert EF.get_early_grad_hook() is None
    dist.destroy_process_group()


def test_early_gradient_hand_over_protocol_gloo():
    """dist.GradBuckets.early_reduce - the hook the fused MomentHead operator calls with dW of its Linear
    before it enqueues the Newton-Schulz backward - averaged in chunks, excluded from its bucket's later
    all-reduce, declined for unknown tensors and under no_sync()."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_early_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, unknown_declined, nosync_declined, gb, gs in res:
        assert unknown_declined and nosync_declined
        assert np.allclose(gb, 1.5) and np.allclose(gs, 15.0)      # mean over the two ranks, once


def test_bench_workload_switches():
    """bench.py --config 3 / 4 / 5 select BASELINE.json's other workloads and name them in the line."""
    import bench
    argv = sys.argv
    try:
        for flags, name, checks in (
                ([], "configs[1]", dict(tokens=197, d_in=768, batch=256, third_order=False)),
                (["--config", "3"], "configs[2]", dict(third_order=True, sketch_dim=8192)),
                (["--config", "5", "--degree", "2", "1"], "configs[4]", dict(tokens=144, d_in=1024, degree=[2, 1])),
                (["--config", "4"], "configs[3]", dict(batch=64, degree=[2, 2]))):
            sys.argv = ["bench.py"] + flags
            args = bench.parse()
            for k, v in checks.items():
                assert getattr(args, k) == v, (flags, k, getattr(args, k))
            assert bench.workload_name(args) == name
            cfg = bench.workload_config(args)
            assert cfg["workload"].startswith(name)
            if args.third_order:
                assert "tensor_sketch.sketch_dim" in cfg["third_order"]["oracle_patch"]
    finally:
        sys.argv = argv
        bench.N_TOK, bench.D_IN, bench.DEGS = 197, 768, (3, 3)


def test_shard_bounds_cover_batch():
    d = importlib.import_module("ego-moment-cle-vit_b200.dist")
    for B in (1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [d.shard_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.shard_bounds(4, 2, 2)


# ------------------------------------------------------------------------- bench.py contract
def _run_reference_arm(env_extra):
    import json
    import subprocess
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the native one): one JSON line
    with the metric / config of the native arm, `impl`, `cpu_baseline` and a zero-copy `e2e`. It times the
    reference's own modules when its sources are found, and it uses every host core even when the
    launcher exported OMP_NUM_THREADS=1 (torchrun does for N>1: round 1's N>1 ratios were void)."""
    from baseline import reference_loader as RL
    d = _run_reference_arm({"OMP_NUM_THREADS": "1"})
    assert d["impl"] == "reference" and d["metric"] == "MomentHead+GPF fwd+bwd images/sec"
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    want = "reference" if RL.find_reference_root() else "port"
    assert d["cpu_baseline"]["kind"] == want
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]") and d["config"]["tokens"] == 197


def test_bench_reference_arm_falls_back_to_the_port(tmp_path, monkeypatch):
    """With no reference sources anywhere the CPU arm is the oracle port and says so."""
    import bench
    from baseline import reference_loader as RL
    monkeypatch.setattr(RL, "find_reference_root", lambda: None)
    monkeypatch.setattr(bench, "N_TOK", 24)
    monkeypatch.setattr(bench, "D_IN", 32)
    ref = bench.CpuReferenceStep()
    assert ref.kind == "port" and "port" in ref.where
    ref(2)


def test_reference_install_is_verbatim():
    """baseline/_ref (git-ignored, travels to the GPU box) holds the reference's files unmodified."""
    from baseline import reference_loader as RL
    if not os.path.isdir(RL.INSTALL_DIR):
        pytest.skip("no reference install on this box")
    assert RL.verify_install()
    if os.path.isdir(REF):
        import filecmp
        for name in ("gpf_kernel.py", "moment_head.py", "ego_moment_clevit.py", "classifier_head.py"):
            assert filecmp.cmp(os.path.join(REF, "src", "models", name),
                               os.path.join(RL.INSTALL_DIR, "src", "models", name), shallow=False)
    gk, mh, ops = RL.load_path_modules(RL.INSTALL_DIR, pkgname="egm_ref_install_probe")
    assert hasattr(gk, "GraphPolynomialFusion") and hasattr(mh, "MomentHead") and hasattr(ops, "batch_trace")


def test_header_is_plain_c():
    """include/egm_b200.h is the C ABI: it must compile as C99 on its own (no C++, no CUDA, no torch)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc on this box")
    hdr = os.path.join(ROOT, "include", "egm_b200.h")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_plain_c_host_links_and_calls_the_abi(pkg, tmp_path):
    """tests/native/abi_smoke.c: a C99 program (no C++/CUDA/Python) linked against libegm_b200.so calls
    the version / size queries and sees argument errors as status codes - the drop-in boundary is a C ABI."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc on this box")
    lib_dir = os.path.dirname(pkg._lib.lib_path())
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([gcc, "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "native", "abi_smoke.c"), "-L", lib_dir, "-legm_b200", "-o", exe],
                   check=True, capture_output=True)
    env = dict(os.environ, LD_LIBRARY_PATH=lib_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("abi ok"), (r.returncode, r.stdout, r.stderr)


def test_feature_tail_has_no_cpu_path_for_the_reference_stack(pkg):
    """BatchNorm1d -> GELU -> Dropout of second_net / third_net runs in the library: CPU tensors raise;
    a Sequential tail that is not the reference's is applied layer by layer (not this library's business)."""
    EF = pkg.functional
    layers = [torch.nn.BatchNorm1d(4), torch.nn.GELU(), torch.nn.Dropout(0.1)]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        EF.feature_tail(torch.randn(3, 4), layers)
    other = [torch.nn.LayerNorm(4), torch.nn.ReLU()]
    y = torch.randn(3, 4)
    assert torch.equal(EF.feature_tail(y, other), other[1](other[0](y)))
    # the data-parallel hand-over hook is a plain module-level slot
    assert EF.get_early_grad_hook() is None
    EF.set_early_grad_hook(print)
    assert EF.get_early_grad_hook() is print
    EF.set_early_grad_hook(None)
