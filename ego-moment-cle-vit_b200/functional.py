"""torch.autograd bindings of the C-ABI operators (forward AND hand-written backward).

Each Function allocates its outputs / saved state / scratch workspace with torch (the library
never allocates), then calls the extern "C" entry point on the caller's current CUDA stream.
The reference leaves all of this to autograd over ~2 200 ATen calls per step.
"""
from __future__ import annotations

import os
import threading

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

_state = threading.local()
_default_precision = os.environ.get("EGM_PRECISION", "fp32")


def set_precision(mode: str) -> None:
    """Select how dense contractions are evaluated: 'fp32' (bf16x3 split on tcgen05, ~1e-5 rel),
    'bf16' (single tcgen05 pass, ~4e-3 rel) or 'fp32_simt' (FFMA on CUDA cores)."""
    global _default_precision
    _lib.precision_id(mode)
    _default_precision = mode


def get_precision() -> str:
    return getattr(_state, "override", None) or _default_precision


class precision:
    """Context manager: `with precision('bf16'): ...` (thread-local)."""

    def __init__(self, mode: str):
        _lib.precision_id(mode)
        self.mode = mode

    def __enter__(self):
        self.prev = getattr(_state, "override", None)
        _state.override = self.mode
        return self

    def __exit__(self, *exc):
        _state.override = self.prev
        return False


def _prec(mode) -> int:
    return _lib.precision_id(mode if mode is not None else get_precision())


def _require_cuda_f32(name: str, t: torch.Tensor, ndim: int) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name}: tensor is on '{t.device}'. This is the B200-native build of the moment-pooling "
            "path: it runs hand-written sm_100a kernels only and has no CPU fallback.")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32, got {t.dtype}")
    if t.dim() != ndim:
        raise RuntimeError(f"{name}: expected a {ndim}-D tensor, got shape {tuple(t.shape)}")
    return t.contiguous()


# EGM_POISON=1 (tests): every buffer this module allocates - outputs, saved state, scratch - starts as
# 0xFF bytes (NaN in fp32 and bf16), so a kernel that reads memory nobody wrote cannot pass by luck.
_poison = os.environ.get("EGM_POISON", "0") == "1"


def set_poison(on: bool) -> None:
    global _poison
    _poison = bool(on)


def _poisoned(t: torch.Tensor) -> torch.Tensor:
    if _poison and t.numel():
        (t if t.dtype == torch.uint8 else t.view(torch.uint8)).fill_(0xFF)
    return t


def _empty(*size, **kw) -> torch.Tensor:
    return _poisoned(torch.empty(*size, **kw))


def _empty_like(x: torch.Tensor) -> torch.Tensor:
    return _poisoned(torch.empty_like(x))


def _ws(nbytes: int, device) -> torch.Tensor:
    return _empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------- GPF
_gpf_recompute = os.environ.get("EGM_GPF_RECOMPUTE", "0") == "1"
# EGM_GPF_RAW_PLANES=1: the fused forward also writes the raw tokens as operand planes for the backward
# (saves the backward's 2 x 54 us re-derivation, costs the forward 137 us of scattered 16-byte stores at
# B=256: measured a net loss of 30 us per step, so it is off by default)
_gpf_raw_planes = os.environ.get("EGM_GPF_RAW_PLANES", "0") == "1"


class _GPFFunction(Function):
    @staticmethod
    def forward(ctx, a, p, coef, cosine, eps, symmetric, prec):
        L = _lib.load()
        B, N, D = a.shape
        P, Q = coef.shape[0] - 1, coef.shape[1] - 1
        dev = a.device
        ldr = L.egm_gpf_ldr(N)
        need_grad = any(ctx.needs_input_grad[:3])
        # one fused pass over the tokens when the library can address them (egm_gpf_fused_ok): it never
        # writes the normalised tokens, and writes R_a / R_p only when a backward will read them
        fused = bool(L.egm_gpf_fused_ok(N, D, P, Q, prec)) and a.data_ptr() % 16 == 0 and p.data_ptr() % 16 == 0
        with torch.cuda.device(dev):
            G = _empty(B, N, N, device=dev, dtype=torch.float32)
            want_R = need_grad or not fused
            Ra = _empty(B, N, ldr, device=dev, dtype=torch.float32) if want_R else None
            Rp = _empty(B, N, ldr, device=dev, dtype=torch.float32) if want_R else None
            nrm = _empty(2, B, N, device=dev, dtype=torch.float32)
            coef_c = coef.detach().contiguous()
            # staged path: keep the normalised tokens (GEMM operand planes, 2 x [B,N,D]) for the backward
            # when one will run; EGM_GPF_RECOMPUTE=1 trades them for a re-normalisation pass in the backward
            # fused path: the RAW token planes, when the backward can fold F.normalize into E (cosine, degrees <= 3)
            raw = (_gpf_raw_planes and fused and bool(cosine)
                   and bool(L.egm_gpf_raw_planes_ok(N, D, P, Q, prec)))
            keep = need_grad and not _gpf_recompute and (raw or not fused)
            xn = _ws(L.egm_gpf_state_bytes(B, N, D, prec), dev) if keep else None
            ws = _ws(L.egm_gpf_fwd_workspace(B, N, D, prec) if not (keep or fused) else 16, dev)
            _lib.check(L.egm_gpf_fwd(a.data_ptr(), p.data_ptr(), coef_c.data_ptr(), B, N, D, P, Q,
                                     int(cosine), float(eps), int(symmetric), G.data_ptr(),
                                     _p(Ra), _p(Rp), nrm[0].data_ptr(), nrm[1].data_ptr(),
                                     _p(xn), prec, ws.data_ptr(), ws.numel(), _stream(dev)), "egm_gpf_fwd")
        if need_grad:
            ctx.save_for_backward(a, p, coef_c, Ra, Rp, nrm, *([xn] if keep else []))
        # bit 1: R_a / R_p are symmetric bit for bit (fused forward) - the backward evaluates each pair once
        # bit 2: the saved planes are the raw tokens (see egm_gpf_raw_planes_ok)
        ctx.cfg = (int(cosine), float(eps), int(symmetric) | (2 if fused else 0) | (4 if (fused and keep) else 0), prec)
        return G

    @staticmethod
    @once_differentiable
    def backward(ctx, dG):
        L = _lib.load()
        a, p, coef, Ra, Rp, nrm = ctx.saved_tensors[:6]
        xn = ctx.saved_tensors[6] if len(ctx.saved_tensors) > 6 else None
        cosine, eps, symmetric, prec = ctx.cfg
        B, N, D = a.shape
        P, Q = coef.shape[0] - 1, coef.shape[1] - 1
        dev = a.device
        dG = dG.contiguous()
        with torch.cuda.device(dev):
            da = _empty_like(a)
            dp = _empty_like(p)
            dcoef = _empty_like(coef)
            ws = _ws(L.egm_gpf_bwd_workspace(B, N, D, P, Q, prec), dev)
            _lib.check(L.egm_gpf_bwd(dG.data_ptr(), a.data_ptr(), p.data_ptr(), coef.data_ptr(),
                                     Ra.data_ptr(), Rp.data_ptr(), nrm[0].data_ptr(), nrm[1].data_ptr(), _p(xn),
                                     B, N, D, P, Q, cosine, eps, symmetric, da.data_ptr(), dp.data_ptr(),
                                     dcoef.data_ptr(), prec, ws.data_ptr(), ws.numel(), _stream(dev)),
                       "egm_gpf_bwd")
        return da, dp, dcoef, None, None, None, None


def gpf_fused_graph(tokens_anchor, tokens_positive, coef, *, cosine=True, eps=1e-6, symmetric=True,
                    precision=None):
    """G = clamp(sym(sum_pq coef[p,q] f_p(R_a) * f_q(R_p)), 0)  (gpf_kernel.py:117-159).

    `coef` is softplus(alpha) [P+1,Q+1]; gradients flow to both token tensors and to coef."""
    a = _require_cuda_f32("tokens_anchor", tokens_anchor, 3)
    p = _require_cuda_f32("tokens_positive", tokens_positive, 3)
    c = _require_cuda_f32("coef", coef, 2)
    if a.shape != p.shape:
        raise RuntimeError(f"token shapes differ: {tuple(a.shape)} vs {tuple(p.shape)}")
    if c.shape[0] > 16 or c.shape[1] > 16:
        raise RuntimeError("polynomial degrees above 15 are not supported")
    G = _GPFFunction.apply(a, p, c, cosine, eps, symmetric, _prec(precision))
    return mark_symmetric(G) if symmetric else G


# -------------------------------------------------------------------------- pool
class _PoolFunction(Function):
    @staticmethod
    def forward(ctx, Z, G, eps, want_u, prec):
        L = _lib.load()
        B, N, D = Z.shape
        dev = Z.device
        with torch.cuda.device(dev):
            M2 = _empty(B, D, D, device=dev, dtype=torch.float32)
            u = _empty(B, D, device=dev, dtype=torch.float32) if want_u else None
            vecs = _empty(B * (4 * N + 2), device=dev, dtype=torch.float32)
            mu = _empty(B, D, device=dev, dtype=torch.float32)
            state = _ws(L.egm_pool_state_bytes(B, N, D, prec), dev)
            ws = _ws(L.egm_pool_fwd_workspace(B, N, D, prec), dev)
            _lib.check(L.egm_pool_fwd(Z.data_ptr(), G.data_ptr(), B, N, D, float(eps), M2.data_ptr(),
                                      _p(u), vecs.data_ptr(), mu.data_ptr(), state.data_ptr(), prec,
                                      ws.data_ptr(), ws.numel(), _stream(dev)), "egm_pool_fwd")
        ctx.save_for_backward(Z, G, vecs, mu, state, *( [u] if want_u else [] ))
        ctx.cfg = (float(eps), bool(want_u), prec)
        if want_u:
            return M2, u
        return M2

    @staticmethod
    @once_differentiable
    def backward(ctx, dM2, du=None):
        L = _lib.load()
        eps, want_u, prec = ctx.cfg
        saved = ctx.saved_tensors
        Z, G, vecs, mu, state = saved[:5]
        u = saved[5] if want_u else None
        B, N, D = Z.shape
        dev = Z.device
        with torch.cuda.device(dev):
            if dM2 is None:
                dM2 = torch.zeros(B, D, D, device=dev, dtype=torch.float32)
            dM2 = dM2.contiguous()
            if du is not None:
                du = du.contiguous()
            dZ = _empty_like(Z)
            dG = _empty_like(G)
            ws = _ws(L.egm_pool_bwd_workspace(B, N, D, prec), dev)
            _lib.check(L.egm_pool_bwd(dM2.data_ptr(), _p(du), Z.data_ptr(), G.data_ptr(), _p(u),
                                      vecs.data_ptr(), mu.data_ptr(), state.data_ptr(), B, N, D, eps,
                                      dZ.data_ptr(), dG.data_ptr(), prec, ws.data_ptr(), ws.numel(),
                                      _stream(dev)), "egm_pool_bwd")
        return dZ, dG, None, None, None


class _MomentLowRankFunction(Function):
    """Pooling + iSQRT-COV fused, Newton-Schulz evaluated on N x N matrices (csrc/egm_lowrank.cu)."""

    @staticmethod
    def forward(ctx, Z, G, iters, eps, want_u, prec):
        L = _lib.load()
        B, N, D = Z.shape
        dev = Z.device
        with torch.cuda.device(dev):
            O = _empty(B, D, D, device=dev, dtype=torch.float32)
            u = _empty(B, D, device=dev, dtype=torch.float32) if want_u else None
            vecs = _empty(B * (4 * N + 2), device=dev, dtype=torch.float32)
            mu = _empty(B, D, device=dev, dtype=torch.float32)
            scal = _empty(5, B, device=dev, dtype=torch.float32)
            state = _ws(L.egm_mlr_state_bytes(B, N, D, iters, prec), dev)
            ws = _ws(L.egm_mlr_fwd_workspace(B, N, D, iters, prec), dev)
            _lib.check(L.egm_mlr_fwd(Z.data_ptr(), G.data_ptr(), B, N, D, int(iters), float(eps),
                                     O.data_ptr(), None, _p(u), vecs.data_ptr(), mu.data_ptr(),
                                     scal.data_ptr(), state.data_ptr(), prec, ws.data_ptr(), ws.numel(),
                                     _stream(dev)), "egm_mlr_fwd")
        ctx.save_for_backward(Z, G, O, vecs, mu, scal, state, *([u] if want_u else []))
        ctx.cfg = (int(iters), float(eps), bool(want_u), prec)
        if want_u:
            return O, u
        return O

    @staticmethod
    @once_differentiable
    def backward(ctx, dO, du=None):
        L = _lib.load()
        iters, eps, want_u, prec = ctx.cfg
        saved = ctx.saved_tensors
        Z, G, O, vecs, mu, scal, state = saved[:7]
        u = saved[7] if want_u else None
        B, N, D = Z.shape
        dev = Z.device
        with torch.cuda.device(dev):
            if dO is None:
                dO = torch.zeros(B, D, D, device=dev, dtype=torch.float32)
            dO = dO.contiguous()
            if du is not None:
                du = du.contiguous()
            dZ = _empty_like(Z)
            dG = _empty_like(G)
            ws = _ws(L.egm_mlr_bwd_workspace(B, N, D, iters, prec), dev)
            _lib.check(L.egm_mlr_bwd(dO.data_ptr(), None, None, _p(du), Z.data_ptr(), G.data_ptr(), O.data_ptr(), _p(u),
                                     vecs.data_ptr(), mu.data_ptr(), scal.data_ptr(), state.data_ptr(),
                                     B, N, D, iters, eps, dZ.data_ptr(), dG.data_ptr(), prec,
                                     ws.data_ptr(), ws.numel(), _stream(dev)), "egm_mlr_bwd")
        return dZ, dG, None, None, None, None


# ---- early hand-over of a parameter gradient produced in the middle of a fused backward ----------
# dist.GradBuckets registers itself here: the fused head computes dW of its Linear before the whole
# Newton-Schulz backward inside ONE autograd node, so a post-accumulate hook on the parameter would
# fire only after the chain. hook(weight_data_ptr, dW) may start an asynchronous all-reduce of dW and
# return a handle; the operator calls handle.wait() (a stream-side wait) before returning dW.
_early_grad_hook = None


def set_early_grad_hook(fn) -> None:
    global _early_grad_hook
    _early_grad_hook = fn


def get_early_grad_hook():
    return _early_grad_hook


class _MomentHeadLinearFunction(Function):
    """pool -> iSQRT-COV -> half-vectorise -> Linear, fused (csrc/egm_api.cu egm_mhd_*, or egm_mlr_*
    for the low-rank evaluation): the packed upper triangle of the normalised covariance is
    written by the last Newton-Schulz product straight into the Linear's operand planes."""

    @staticmethod
    def forward(ctx, Z, G, weight, bias, iters, eps, want_u, lowrank, prec, flags):
        L = _lib.load()
        B, N, D = Z.shape
        n_out, K = weight.shape
        dev = Z.device
        with torch.cuda.device(dev):
            y = _empty(B, n_out, device=dev, dtype=torch.float32)
            u = _empty(B, D, device=dev, dtype=torch.float32) if want_u else None
            vecs = _empty(B * (4 * N + 2), device=dev, dtype=torch.float32)
            mu = _empty(B, D, device=dev, dtype=torch.float32)
            lin_state = _ws(L.egm_linear_state_bytes(B, n_out, K, prec), dev)
            if lowrank:
                scal = _empty(5, B, device=dev, dtype=torch.float32)
                state = _ws(L.egm_mlr_state_bytes(B, N, D, iters, prec), dev)
                ws = _ws(L.egm_mlr_fwd_workspace(B, N, D, iters, prec), dev)
                _lib.check(L.egm_mlr_fwd(Z.data_ptr(), G.data_ptr(), B, N, D, int(iters), float(eps), None,
                                         lin_state.data_ptr(), _p(u), vecs.data_ptr(), mu.data_ptr(),
                                         scal.data_ptr(), state.data_ptr(), prec, ws.data_ptr(), ws.numel(),
                                         _stream(dev)), "egm_mlr_fwd")
            else:
                scal = _empty(3, B, device=dev, dtype=torch.float32)
                state = _ws(L.egm_mhd_state_bytes(B, N, D, iters, prec), dev)
                ws = _ws(L.egm_mhd_fwd_workspace(B, N, D, iters, prec), dev)
                _lib.check(L.egm_mhd_fwd(Z.data_ptr(), G.data_ptr(), B, N, D, int(iters), float(eps), int(flags),
                                         lin_state.data_ptr(), _p(u), vecs.data_ptr(), mu.data_ptr(),
                                         scal.data_ptr(), state.data_ptr(), prec, ws.data_ptr(), ws.numel(),
                                         _stream(dev)), "egm_mhd_fwd")
            ws = _ws(L.egm_linear_fwd_workspace(B, n_out, K, prec), dev)
            _lib.check(L.egm_linear_fwd(None, weight.data_ptr(), _p(bias), B, n_out, K, y.data_ptr(),
                                        lin_state.data_ptr(), prec, ws.data_ptr(), ws.numel(), _stream(dev)),
                       "egm_linear_fwd")
        ctx.save_for_backward(Z, G, vecs, mu, scal, state, lin_state, y, *([bias] if bias is not None else []),
                              *([u] if want_u else []))
        ctx.cfg = (int(iters), float(eps), bool(want_u), bool(lowrank), prec, bias is not None, n_out, K,
                   int(flags))
        ctx.weight_ptr = weight.data_ptr()
        if want_u:
            return y, u
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, du=None):
        L = _lib.load()
        iters, eps, want_u, lowrank, prec, has_bias, n_out, K, flags = ctx.cfg
        saved = list(ctx.saved_tensors)
        Z, G, vecs, mu, scal, state, lin_state, y = saved[:8]
        rest = saved[8:]
        bias = rest.pop(0) if has_bias else None
        u = rest.pop(0) if want_u else None
        B, N, D = Z.shape
        dev = Z.device
        need_w = ctx.needs_input_grad[2]
        need_b = has_bias and ctx.needs_input_grad[3]
        with torch.cuda.device(dev):
            if dy is None:
                dy = torch.zeros(B, n_out, device=dev, dtype=torch.float32)
            dy = dy.contiguous()
            if du is not None:
                du = du.contiguous()
            dv = _empty(B, K, device=dev, dtype=torch.float32)
            dw = _empty(n_out, K, device=dev, dtype=torch.float32) if need_w else None
            db = _empty(n_out, device=dev, dtype=torch.float32) if need_b else None
            ws = _ws(L.egm_linear_bwd_workspace(B, n_out, K, prec), dev)
            _lib.check(L.egm_linear_bwd(dy.data_ptr(), lin_state.data_ptr(), B, n_out, K, dv.data_ptr(),
                                        _p(dw), _p(db), prec, ws.data_ptr(), ws.numel(), _stream(dev)),
                       "egm_linear_bwd")
            # dW is complete as far as the stream is concerned: let the data-parallel reducer start its
            # all-reduce now, underneath the Newton-Schulz backward enqueued below
            pending = _early_grad_hook(ctx.weight_ptr, dw) if (_early_grad_hook is not None and need_w) else None
            dot = _empty(B, device=dev, dtype=torch.float32)   # <dO, O> = <dy, y - bias>
            _lib.check(L.egm_rowdot_bias(dy.data_ptr(), y.data_ptr(), _p(bias), B, n_out, dot.data_ptr(),
                                         _stream(dev)), "egm_rowdot_bias")
            dZ = _empty_like(Z)
            dG = _empty_like(G)
            if lowrank:
                ws = _ws(L.egm_mlr_bwd_workspace(B, N, D, iters, prec), dev)
                _lib.check(L.egm_mlr_bwd(None, dv.data_ptr(), dot.data_ptr(), _p(du), Z.data_ptr(), G.data_ptr(),
                                         None, _p(u), vecs.data_ptr(), mu.data_ptr(), scal.data_ptr(),
                                         state.data_ptr(), B, N, D, iters, eps, dZ.data_ptr(), dG.data_ptr(),
                                         prec, ws.data_ptr(), ws.numel(), _stream(dev)), "egm_mlr_bwd")
            else:
                ws = _ws(L.egm_mhd_bwd_workspace(B, N, D, iters, prec), dev)
                _lib.check(L.egm_mhd_bwd(dv.data_ptr(), dot.data_ptr(), _p(du), Z.data_ptr(), G.data_ptr(), _p(u),
                                         vecs.data_ptr(), mu.data_ptr(), scal.data_ptr(), state.data_ptr(),
                                         B, N, D, iters, eps, flags, dZ.data_ptr(), dG.data_ptr(), prec,
                                         ws.data_ptr(), ws.numel(), _stream(dev)), "egm_mhd_bwd")
            if pending is not None:
                pending.wait()       # stream-side: autograd may read dW from here on
        return dZ, dG, dw, db, None, None, None, None, None, None


def moment_head_linear(tokens, graph, weight, bias, num_iterations, *, eps=1e-5, third_order=False,
                       precision=None, algorithm=None):
    """Linear(half_vectorize(NewtonSchulzSqrtm(Zc^T W Zc)))  (moment_head.py:279-300, first layer of
    second_net) as one fused operator; returns y [B, n_out] or (y, u). Returns None when the fused
    form does not apply (strict fp32 mode, or fewer iterations than the fused chain supports) - the
    caller then composes graph_weighted_pool / newton_schulz / half_vectorize / linear."""
    prec = _prec(precision)
    algo = algorithm or _ns_algorithm
    if prec == _lib.PREC_FP32_SIMT:
        return None
    Z = _require_cuda_f32("tokens", tokens, 3)
    G = _require_cuda_f32("graph", graph, 3)
    if G.shape != (Z.shape[0], Z.shape[1], Z.shape[1]):
        raise RuntimeError(f"graph shape {tuple(G.shape)} does not match tokens {tuple(Z.shape)}")
    lowrank = algo == "lowrank" and num_iterations >= 1 and Z.shape[1] < Z.shape[2]
    if not lowrank and num_iterations < 2:
        return None
    w = _require_cuda_f32("weight", weight, 2)
    D = Z.shape[2]
    if w.shape[1] != D * (D + 1) // 2:
        raise RuntimeError(f"linear: weight {tuple(w.shape)} does not match d_in={D}")
    b = _require_cuda_f32("bias", bias, 1) if bias is not None else None
    flags = _lib.MHD_SYMMETRIC_GRAPH if (not lowrank and graph_is_symmetric(graph)) else 0
    args = (w, b, int(num_iterations), eps, third_order, lowrank, prec, flags)
    # Inference: nothing is kept for a backward, yet the operator's working state is the training state
    # (13 D x D matrices per image at K = 5). The path is independent per image and its results are
    # bit-identical under batch sharding (tests), so a large no-grad batch is evaluated in slices whose
    # state stays under `_nograd_state_limit` instead of allocating all of it at once.
    B = Z.shape[0]
    need_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (Z, G, w, b))
    if not need_grad and B > 1:
        L = _lib.load()
        per = (L.egm_mlr_state_bytes if lowrank else L.egm_mhd_state_bytes)(1, Z.shape[1], D, int(num_iterations), prec)
        step = max(1, int(_nograd_state_limit // max(per, 1)))
        if step < B:
            outs = [_MomentHeadLinearFunction.apply(Z[i:i + step], G[i:i + step], *args) for i in range(0, B, step)]
            if third_order:
                return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
            return torch.cat(outs)
    return _MomentHeadLinearFunction.apply(Z, G, *args)


# working-state budget of a no-grad call of the fused head (bytes); EGM_NOGRAD_STATE_GB overrides
_nograd_state_limit = float(os.environ.get("EGM_NOGRAD_STATE_GB", "8.5")) * 2 ** 30


# ---- exact-symmetry tag of a graph tensor ---------------------------------------------------
# GraphPolynomialFusion with symmetric_enforce emits G with G[i,j] == G[j,i] bit for bit
# (0.5 * (F_ij + F_ji) in one kernel). The tag records the tensor's version counter, so an
# in-place edit of G silently drops it; MomentHead then takes the general (any real matrix) path.
_symmetric_fast_path = os.environ.get("EGM_SYMMETRIC_FAST_PATH", "1") != "0"


def set_symmetric_fast_path(on: bool) -> None:
    """Let MomentHead exploit a graph tagged exactly symmetric (default on): every Newton-Schulz
    product then evaluates its upper tiles only and the backward runs as a symmetric tangent.
    With the fast path the gradient returned for the graph is the symmetric part (dG + dG^T)/2 of the
    reference's (the gradient with respect to a symmetric matrix) - exactly what the symmetrisation in
    GraphPolynomialFusion's backward lets through, so token and parameter gradients are unchanged."""
    global _symmetric_fast_path
    _symmetric_fast_path = bool(on)


def mark_symmetric(graph: torch.Tensor) -> torch.Tensor:
    graph._egm_symmetric_version = graph._version
    return graph


def graph_is_symmetric(graph: torch.Tensor) -> bool:
    """True when MomentHead may take the symmetric fast path for `graph`: the tag written by
    GraphPolynomialFusion matches the tensor's version counter (an in-place edit drops it), and nobody
    observes the graph's own gradient - `graph.retain_grad()` or a tensor hook would see (dG + dG^T)/2
    instead of the reference's dG, so those graphs take the general path. Writes that bypass the
    version counter (`G.data.copy_()`, `set_()`) are not detected; under EGM_POISON=1 (the test
    suite's debug mode) the tag is verified on the device before it is trusted."""
    if not _symmetric_fast_path or getattr(graph, "_egm_symmetric_version", None) != graph._version:
        return False
    if graph.requires_grad and (graph.retains_grad or getattr(graph, "_backward_hooks", None)):
        return False
    if _poison and not torch.equal(graph, graph.transpose(-2, -1)):
        raise RuntimeError("graph is tagged exactly symmetric but is not: it was edited without bumping its "
                           "version counter (e.g. through .data); the symmetric fast path would read only its "
                           "upper blocks")
    return True


_ns_algorithm = os.environ.get("EGM_NS_ALGORITHM", "dense")


def set_ns_algorithm(name: str) -> None:
    """How MomentHead evaluates pooling + iSQRT-COV: 'dense' (D x D Newton-Schulz chain, the
    reference's formulation) or 'lowrank' (same function, all Newton-Schulz products on N x N
    matrices when N < D). NewtonSchulzSqrtm on its own is always dense."""
    global _ns_algorithm
    if name not in ("dense", "lowrank"):
        raise ValueError(f"unknown Newton-Schulz algorithm {name!r}; expected 'dense' or 'lowrank'")
    _ns_algorithm = name


def get_ns_algorithm() -> str:
    return _ns_algorithm


def moment_isqrt(tokens, graph, num_iterations, *, eps=1e-5, third_order=False, precision=None,
                 algorithm=None):
    """iSQRT-COV of the graph-weighted second moment: NewtonSchulzSqrtm(Zc^T W Zc)
    (moment_head.py:279-296) [+ u for the third-order branch]. Returns O [B,D,D] or (O, u)."""
    algo = algorithm or _ns_algorithm
    Z = _require_cuda_f32("tokens", tokens, 3)
    G = _require_cuda_f32("graph", graph, 3)
    if G.shape != (Z.shape[0], Z.shape[1], Z.shape[1]):
        raise RuntimeError(f"graph shape {tuple(G.shape)} does not match tokens {tuple(Z.shape)}")
    if algo == "lowrank" and num_iterations >= 1 and Z.shape[1] < Z.shape[2]:
        return _MomentLowRankFunction.apply(Z, G, int(num_iterations), eps, third_order, _prec(precision))
    if third_order:
        M2, u = _PoolFunction.apply(Z, G, eps, True, _prec(precision))
        return newton_schulz(M2, num_iterations, eps, precision=precision), u
    M2 = _PoolFunction.apply(Z, G, eps, False, _prec(precision))
    return newton_schulz(M2, num_iterations, eps, precision=precision)


def graph_weighted_pool(tokens, graph, *, eps=1e-5, third_order=False, precision=None):
    """W = D^-1/2 G D^-1/2; mu = Z^T W 1/(tr W+eps); Zc = Z - mu; M2 = Zc^T W Zc
    (moment_head.py:246-266, 222-244, 288-293) and, if `third_order`, u = Zc^T W 1/(tr W+eps)
    (moment_head.py:305-311). Returns M2 [B,D,D] or (M2, u [B,D])."""
    Z = _require_cuda_f32("tokens", tokens, 3)
    G = _require_cuda_f32("graph", graph, 3)
    if G.shape != (Z.shape[0], Z.shape[1], Z.shape[1]):
        raise RuntimeError(f"graph shape {tuple(G.shape)} does not match tokens {tuple(Z.shape)}")
    return _PoolFunction.apply(Z, G, eps, third_order, _prec(precision))


# ---------------------------------------------------------------------------- NS
class _NSFunction(Function):
    @staticmethod
    def forward(ctx, M, iters, eps, post_mode, prec):
        L = _lib.load()
        B, D, _ = M.shape
        dev = M.device
        with torch.cuda.device(dev):
            O = _empty_like(M)
            scal = _empty(3, B, device=dev, dtype=torch.float32)
            state = _ws(L.egm_ns_state_bytes(B, D, iters, prec), dev)
            ws = _ws(L.egm_ns_fwd_workspace(B, D, iters, prec), dev)
            _lib.check(L.egm_ns_fwd(M.data_ptr(), B, D, int(iters), float(eps), int(post_mode),
                                    O.data_ptr(), scal.data_ptr(), state.data_ptr(), prec,
                                    ws.data_ptr(), ws.numel(), _stream(dev)), "egm_ns_fwd")
        ctx.save_for_backward(M, O, scal, state)
        ctx.cfg = (int(iters), float(eps), int(post_mode), prec)
        return O

    @staticmethod
    @once_differentiable
    def backward(ctx, dO):
        L = _lib.load()
        M, O, scal, state = ctx.saved_tensors
        iters, eps, post_mode, prec = ctx.cfg
        B, D, _ = M.shape
        dev = M.device
        dO = dO.contiguous()
        with torch.cuda.device(dev):
            dM = _empty_like(M)
            ws = _ws(L.egm_ns_bwd_workspace(B, D, iters, prec), dev)
            _lib.check(L.egm_ns_bwd(dO.data_ptr(), O.data_ptr(), M.data_ptr(), scal.data_ptr(),
                                    state.data_ptr(), B, D, iters, eps, post_mode, dM.data_ptr(), prec,
                                    ws.data_ptr(), ws.numel(), _stream(dev)), "egm_ns_bwd")
        return dM, None, None, None, None


def newton_schulz(matrix, num_iterations, eps=1e-5, *, post="divide", precision=None):
    """Trace-normalised coupled Newton-Schulz, Y0 = I, Z0 = A (moment_head.py:28-70).
    post='divide': Y_K / sqrt(tr+eps) (NewtonSchulzSqrtm); post='multiply': Y_K * sqrt(tr+eps)
    (utils/ops.py:122-165)."""
    M = _require_cuda_f32("matrix", matrix, 3)
    if M.shape[1] != M.shape[2]:
        raise RuntimeError(f"expected square matrices, got {tuple(M.shape)}")
    mode = {"divide": 0, "multiply": 1}[post]
    return _NSFunction.apply(M, int(num_iterations), eps, mode, _prec(precision))


# ------------------------------------------------------------------------ linear
class _LinearFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, prec):
        L = _lib.load()
        M, K = x.shape
        N = weight.shape[0]
        dev = x.device
        with torch.cuda.device(dev):
            y = _empty(M, N, device=dev, dtype=torch.float32)
            state = _ws(L.egm_linear_state_bytes(M, N, K, prec), dev)
            ws = _ws(L.egm_linear_fwd_workspace(M, N, K, prec), dev)
            _lib.check(L.egm_linear_fwd(x.data_ptr(), weight.data_ptr(), _p(bias), M, N, K, y.data_ptr(),
                                        state.data_ptr(), prec, ws.data_ptr(), ws.numel(), _stream(dev)),
                       "egm_linear_fwd")
        ctx.save_for_backward(state)
        ctx.cfg = (M, N, K, prec, bias is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        L = _lib.load()
        (state,) = ctx.saved_tensors
        M, N, K, prec, has_bias = ctx.cfg
        dev = dy.device
        dy = dy.contiguous()
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        with torch.cuda.device(dev):
            dx = _empty(M, K, device=dev, dtype=torch.float32) if need_x else None
            dw = _empty(N, K, device=dev, dtype=torch.float32) if need_w else None
            db = _empty(N, device=dev, dtype=torch.float32) if need_b else None
            ws = _ws(L.egm_linear_bwd_workspace(M, N, K, prec), dev)
            _lib.check(L.egm_linear_bwd(dy.data_ptr(), state.data_ptr(), M, N, K, _p(dx), _p(dw), _p(db), prec,
                                        ws.data_ptr(), ws.numel(), _stream(dev)), "egm_linear_bwd")
        return dx, dw, db, None


class _LinearSimtFunction(Function):
    """F.linear in the strict fp32 mode: three products on the library's FFMA engine (egm_bmm)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        y = bmm(x.unsqueeze(0), weight.unsqueeze(0), trans_b=True, precision="fp32_simt").squeeze(0)
        if bias is not None:
            y += bias
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = bmm(dy.unsqueeze(0), weight.unsqueeze(0), precision="fp32_simt").squeeze(0)
        if ctx.needs_input_grad[1]:
            dw = bmm(dy.unsqueeze(0), x.unsqueeze(0), trans_a=True, precision="fp32_simt").squeeze(0)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(0)
        return dx, dw, db


def linear(x, weight, bias=None, *, precision=None):
    """F.linear(x, weight, bias) for 2-D x on the tcgen05 engine (split-K over all CTA pairs); in the
    strict 'fp32_simt' mode on the library's FFMA engine (egm_bmm). Never a torch / cuBLAS GEMM."""
    prec = _prec(precision)
    if prec == _lib.PREC_FP32_SIMT:
        x = _require_cuda_f32("x", x, 2)
        w = _require_cuda_f32("weight", weight, 2)
        if w.shape[1] != x.shape[1]:
            raise RuntimeError(f"linear: x {tuple(x.shape)} and weight {tuple(w.shape)} do not match")
        b = _require_cuda_f32("bias", bias, 1) if bias is not None else None
        return _LinearSimtFunction.apply(x, w, b)
    x = _require_cuda_f32("x", x, 2)
    w = _require_cuda_f32("weight", weight, 2)
    if w.shape[1] != x.shape[1]:
        raise RuntimeError(f"linear: x {tuple(x.shape)} and weight {tuple(w.shape)} do not match")
    b = _require_cuda_f32("bias", bias, 1) if bias is not None else None
    return _LinearFunction.apply(x, w, b, prec)


# ------------------------------------------------------- BatchNorm1d + GELU + Dropout
class _FeatureTailFunction(Function):
    @staticmethod
    def forward(ctx, y, gamma, beta, run_mean, run_var, training, momentum, bn_eps, drop_p, seed):
        L = _lib.load()
        M, N = y.shape
        dev = y.device
        with torch.cuda.device(dev):
            out = _empty_like(y)
            stats = _empty(2, N, device=dev, dtype=torch.float32)
            _lib.check(L.egm_feature_tail_fwd(y.data_ptr(), M, N, _p(gamma), _p(beta), _p(run_mean), _p(run_var),
                                              int(training), float(momentum), float(bn_eps), float(drop_p),
                                              int(seed), out.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                                              _stream(dev)), "egm_feature_tail_fwd")
        ctx.save_for_backward(y, stats, *([gamma] if gamma is not None else []),
                              *([beta] if beta is not None else []))
        ctx.cfg = (int(training), float(drop_p), int(seed), gamma is not None, beta is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.load()
        training, drop_p, seed, has_g, has_b = ctx.cfg
        saved = list(ctx.saved_tensors)
        y, stats = saved[:2]
        rest = saved[2:]
        gamma = rest.pop(0) if has_g else None
        beta = rest.pop(0) if has_b else None
        M, N = y.shape
        dev = y.device
        dout = dout.contiguous()
        with torch.cuda.device(dev):
            dy = _empty_like(y)
            dgamma = _empty(N, device=dev, dtype=torch.float32) if has_g else None
            dbeta = _empty(N, device=dev, dtype=torch.float32) if has_b else None
            _lib.check(L.egm_feature_tail_bwd(dout.data_ptr(), y.data_ptr(), _p(gamma), _p(beta), stats[0].data_ptr(),
                                              stats[1].data_ptr(), M, N, training, drop_p, seed, dy.data_ptr(),
                                              _p(dgamma), _p(dbeta), _stream(dev)), "egm_feature_tail_bwd")
        return dy, dgamma, dbeta, None, None, None, None, None, None, None


def _draw_seed(device) -> int:
    """A dropout seed drawn the way torch's own CUDA dropout draws its randomness: from the device's
    default Philox generator - (initial seed, current offset) - advancing the offset. Host-side, no
    launch, no sync; `torch.manual_seed` resets it, so runs are reproducible; the CPU generator is not
    touched (the reference's dropout does not touch it either)."""
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("train-mode dropout inside a CUDA-graph capture: the keep-mask seed is a launch constant "
                           "of this library's kernel, every replay would reuse one mask - capture with the "
                           "Dropout modules in eval mode / p = 0, or keep this layer outside the graph")
    gen = torch.cuda.default_generators[device.index if device.index is not None else torch.cuda.current_device()]
    off = gen.get_offset()
    gen.set_offset(off + 4)                      # Philox offsets move in multiples of 4
    return ((gen.initial_seed() * 0x9E3779B97F4A7C15) ^ (off * 0xBF58476D1CE4E5B9 + 0x94D049BB133111EB)) & (2 ** 63 - 1)


def feature_tail(y, layers):
    """The layers after a feature net's Linear - BatchNorm1d -> GELU -> Dropout (moment_head.py:186-191,
    195-200) - as one kernel each way. `layers` are the nn.Modules themselves: their parameters,
    running statistics (updated in place like nn.BatchNorm1d, incl. num_batches_tracked) and
    train/eval flags are used as they are, so state_dicts are unaffected. The dropout keep-mask comes
    from a counter-based hash whose seed is derived from the device's default Philox generator state
    (`_draw_seed`: reproducible under torch.manual_seed, advances the CUDA offset like torch's dropout,
    leaves the CPU generator alone); the mask bits are not torch's: same distribution, different bits. Any other layer combination is applied layer by layer."""
    layers = list(layers)
    nn = torch.nn
    if not (len(layers) == 3 and isinstance(layers[0], nn.BatchNorm1d) and isinstance(layers[1], nn.GELU)
            and getattr(layers[1], "approximate", "none") == "none" and isinstance(layers[2], nn.Dropout)):
        for layer in layers:          # a user-modified Sequential: not this library's business
            y = layer(y)
        return y
    y = _require_cuda_f32("y", y, 2)   # the reference's own stack: CUDA only, like the rest of the path
    bn, _, drop = layers
    use_batch = bn.training or (bn.running_mean is None and bn.running_var is None)
    if use_batch and y.shape[0] <= 1:
        raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(y.shape)}")
    momentum = bn.momentum
    run_mean = run_var = None
    if bn.track_running_stats and bn.running_mean is not None:
        run_mean, run_var = bn.running_mean, bn.running_var
        if bn.training:
            bn.num_batches_tracked.add_(1)
            if momentum is None:          # cumulative moving average
                momentum = 1.0 / float(bn.num_batches_tracked)
    if use_batch and not bn.training:     # eval without running stats: batch statistics, nothing to update
        run_mean = run_var = None
    if not use_batch and run_mean is None:
        for layer in layers:
            y = layer(y)
        return y
    p = float(drop.p) if drop.training else 0.0
    seed = _draw_seed(y.device) if p > 0.0 else 0
    # `training` selects batch statistics; dropout is gated by its own p
    upd_mean, upd_var = (run_mean, run_var) if (bn.training or not use_batch) else (None, None)
    return _FeatureTailFunction.apply(y, bn.weight, bn.bias, upd_mean, upd_var, use_batch, momentum or 0.0,
                                      bn.eps, p, seed)


# -------------------------------------------------------------------------- triu
class _TriuFunction(Function):
    @staticmethod
    def forward(ctx, O):
        L = _lib.load()
        B, D, _ = O.shape
        dev = O.device
        with torch.cuda.device(dev):
            v = _empty(B, D * (D + 1) // 2, device=dev, dtype=torch.float32)
            _lib.check(L.egm_triu_pack(O.data_ptr(), B, D, v.data_ptr(), _stream(dev)), "egm_triu_pack")
        ctx.dims = (B, D)
        return v

    @staticmethod
    @once_differentiable
    def backward(ctx, dv):
        L = _lib.load()
        B, D = ctx.dims
        dev = dv.device
        dv = dv.contiguous()
        with torch.cuda.device(dev):
            dO = _empty(B, D, D, device=dev, dtype=torch.float32)
            _lib.check(L.egm_triu_unpack(dv.data_ptr(), B, D, dO.data_ptr(), _stream(dev)),
                       "egm_triu_unpack")
        return dO


def half_vectorize(matrix):
    """Row-major upper triangle incl. diagonal, [B,D,D] -> [B,D(D+1)/2] (moment_head.py:202-220)."""
    O = _require_cuda_f32("matrix", matrix, 3)
    if O.shape[1] != O.shape[2]:
        raise RuntimeError(f"expected square matrices, got {tuple(O.shape)}")
    return _TriuFunction.apply(O)


# ------------------------------------------------------------------------ sketch
def build_sketch_csr(hashes: torch.Tensor, signs: torch.Tensor, sketch_dim: int):
    """CSR inverse of the three hash maps: bucket s of sketch h sums sgn[h,e]*x[idx[h,e]] for
    e in [off[h,s], off[h,s+1]). Stable sort keeps the reference's CPU accumulation order."""
    H, D = hashes.shape
    if int(hashes.max()) >= sketch_dim or int(hashes.min()) < 0:
        # same failure the reference hits in scatter_add_ (moment_head.py:110)
        raise RuntimeError(
            f"index {int(hashes.max())} is out of bounds for dimension 1 with size {sketch_dim}")
    order = torch.argsort(hashes, dim=1, stable=True)
    sorted_h = torch.gather(hashes, 1, order)
    counts = torch.zeros(H, sketch_dim, dtype=torch.int64, device=hashes.device)
    counts.scatter_add_(1, sorted_h, torch.ones_like(sorted_h))
    off = torch.zeros(H, sketch_dim + 1, dtype=torch.int32, device=hashes.device)
    off[:, 1:] = torch.cumsum(counts, 1).to(torch.int32)
    sgn = torch.gather(signs, 1, order).to(torch.float32)
    return off.contiguous(), order.to(torch.int32).contiguous(), sgn.contiguous()


class _SketchFunction(Function):
    @staticmethod
    def forward(ctx, x, hashes, signs, off, idx, sgn, S):
        L = _lib.load()
        B, D = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            cs = _empty(3, B, S, device=dev, dtype=torch.float32)
            out = _empty(B, S, device=dev, dtype=torch.float32)
            _lib.check(L.egm_sketch_fwd(x.data_ptr(), B, D, S, off.data_ptr(), idx.data_ptr(),
                                        sgn.data_ptr(), cs.data_ptr(), out.data_ptr(), _stream(dev)),
                       "egm_sketch_fwd")
        ctx.save_for_backward(cs, hashes, signs)
        ctx.dims = (B, D, S)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        L = _lib.load()
        cs, hashes, signs = ctx.saved_tensors
        B, D, S = ctx.dims
        dev = dout.device
        dout = dout.contiguous()
        with torch.cuda.device(dev):
            dx = _empty(B, D, device=dev, dtype=torch.float32)
            _lib.check(L.egm_sketch_bwd(dout.data_ptr(), cs.data_ptr(), B, D, S, hashes.data_ptr(),
                                        signs.data_ptr(), dx.data_ptr(), _stream(dev)), "egm_sketch_bwd")
        return dx, None, None, None, None, None, None


def tensor_sketch(x, hashes, signs, csr, sketch_dim):
    """prod_k count_sketch_k(x)  (moment_head.py:100-131). hashes/signs: int64 [3,D]."""
    x = _require_cuda_f32("x", x, 2)
    off, idx, sgn = csr
    return _SketchFunction.apply(x, hashes, signs, off, idx, sgn, int(sketch_dim))


# ---------------------------------------------------------------- alignment loss
class _AlignLossFunction(Function):
    @staticmethod
    def forward(ctx, G, labels):
        L = _lib.load()
        B, N, _ = G.shape
        dev = G.device
        with torch.cuda.device(dev):
            buf = _empty(3, B, device=dev, dtype=torch.float32)       # g, dg, rowloss
            loss = _empty(1, device=dev, dtype=torch.float32)
            _lib.check(L.egm_align_fwd(G.data_ptr(), labels.data_ptr(), B, N, buf[0].data_ptr(),
                                       buf[1].data_ptr(), buf[2].data_ptr(), loss.data_ptr(), _stream(dev)),
                       "egm_align_fwd")
        ctx.save_for_backward(buf)
        ctx.dims = (B, N)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, dloss):
        L = _lib.load()
        (buf,) = ctx.saved_tensors
        B, N = ctx.dims
        dev = buf.device
        dloss = dloss.to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(dev):
            dG = _empty(B, N, N, device=dev, dtype=torch.float32)
            _lib.check(L.egm_align_bwd(buf[1].data_ptr(), dloss.data_ptr(), B, N, dG.data_ptr(), _stream(dev)),
                       "egm_align_bwd")
        return dG, None


def graph_alignment_loss(fused_graph, labels):
    """mse_loss(sigmoid(g g^T), [labels_i == labels_j]) with g = fused_graph.mean((1, 2))
    (EGOMomentCLEViT._graph_alignment_loss, ego_moment_clevit.py:278-316) in three small kernels;
    the reference builds the B x B matrix element by element in Python."""
    G = _require_cuda_f32("fused_graph", fused_graph, 3)
    if G.shape[1] != G.shape[2]:
        raise RuntimeError(f"expected square graphs, got {tuple(G.shape)}")
    if labels.shape != (G.shape[0],):
        raise RuntimeError(f"labels shape {tuple(labels.shape)} does not match batch {G.shape[0]}")
    lab = labels.to(device=G.device, dtype=torch.int64).contiguous()
    return _AlignLossFunction.apply(G, lab)


# ------------------------------------------------------------------ ops helpers
class _GramFunction(Function):
    @staticmethod
    def forward(ctx, x, cosine, eps, prec):
        L = _lib.load()
        B, N, D = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            R = _empty(B, N, N, device=dev, dtype=torch.float32)
            nrm = _empty(B, N, device=dev, dtype=torch.float32)
            ws = _ws(L.egm_gram_workspace(B, N, D, prec), dev)
            _lib.check(L.egm_gram_fwd(x.data_ptr(), B, N, D, int(cosine), float(eps), R.data_ptr(),
                                      nrm.data_ptr(), prec, ws.data_ptr(), ws.numel(), _stream(dev)),
                       "egm_gram_fwd")
        ctx.save_for_backward(x, nrm)
        ctx.cfg = (int(cosine), float(eps), prec)
        return R

    @staticmethod
    @once_differentiable
    def backward(ctx, dR):
        L = _lib.load()
        x, nrm = ctx.saved_tensors
        cosine, eps, prec = ctx.cfg
        B, N, D = x.shape
        dev = x.device
        dR = dR.contiguous()
        with torch.cuda.device(dev):
            dx = _empty_like(x)
            ws = _ws(L.egm_gram_workspace(B, N, D, prec), dev)
            _lib.check(L.egm_gram_bwd(dR.data_ptr(), x.data_ptr(), nrm.data_ptr(), B, N, D, cosine, eps,
                                      dx.data_ptr(), prec, ws.data_ptr(), ws.numel(), _stream(dev)),
                       "egm_gram_bwd")
        return dx, None, None, None


def similarity_matrix(tokens, *, cosine=True, eps=1e-6, precision=None):
    """R = Xn Xn^T with Xn = x/max(||x||,eps) (cosine) or x (dot)  (gpf_kernel.py:75-94)."""
    x = _require_cuda_f32("tokens", tokens, 3)
    return _GramFunction.apply(x, cosine, eps, _prec(precision))


def normalize_graph(graph, method: int, eps: float):
    """Forward-only kernel for utils.ops.normalize_graph; returns (normalised, clamped degrees)."""
    G = _require_cuda_f32("graph", graph, 3)
    L = _lib.load()
    B, N, _ = G.shape
    dev = G.device
    with torch.cuda.device(dev):
        out = _empty_like(G)
        deg = _empty(B, N, device=dev, dtype=torch.float32)
        _lib.check(L.egm_normalize_graph(G.data_ptr(), B, N, int(method), float(eps), out.data_ptr(),
                                         deg.data_ptr(), _stream(dev)), "egm_normalize_graph")
    return out, deg


def batch_trace(matrices):
    M = _require_cuda_f32("matrices", matrices, 3)
    L = _lib.load()
    B, D, _ = M.shape
    dev = M.device
    with torch.cuda.device(dev):
        tr = _empty(B, device=dev, dtype=torch.float32)
        _lib.check(L.egm_batch_trace(M.data_ptr(), B, D, tr.data_ptr(), _stream(dev)), "egm_batch_trace")
    return tr


def bmm(A, B, *, trans_a=False, trans_b=False, alpha=1.0, precision=None):
    """alpha * op(A) @ op(B) through the active GEMM engine (diagnostics / benchmark probe)."""
    A = _require_cuda_f32("A", A, 3)
    B = _require_cuda_f32("B", B, 3)
    L = _lib.load()
    prec = _prec(precision)
    nb = A.shape[0]
    M, K = (A.shape[2], A.shape[1]) if trans_a else (A.shape[1], A.shape[2])
    N = B.shape[1] if trans_b else B.shape[2]
    dev = A.device
    with torch.cuda.device(dev):
        C = _empty(nb, M, N, device=dev, dtype=torch.float32)
        ws = _ws(L.egm_bmm_workspace(nb, M, N, K, prec), dev)
        _lib.check(L.egm_bmm(A.data_ptr(), int(trans_a), B.data_ptr(), int(trans_b), nb, M, N, K,
                             float(alpha), C.data_ptr(), prec, ws.data_ptr(), ws.numel(), _stream(dev)),
                   "egm_bmm")
    return C
