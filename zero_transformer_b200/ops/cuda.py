"""CUDA op layer: ``torch.autograd.Function`` wrappers around the hand-written
sm_100a kernels in ``csrc/`` (loaded from ``zero_transformer_b200/_ext``).

Importing this module on a box where the extension is not built raises -- there
is no silent eager fallback on the product path.  Library GEMMs (cuBLAS through
``torch.matmul``) are used only for shapes the tcgen05 kernel does not cover
(tiny/odd problem sizes such as the 64-wide test model or 1-token decode).

Fusions implemented here:
  K6/K8+K9  projection GEMM with dropout+residual epilogue (Philox mask regenerated in backward)
  K7        fc_in GEMM with tanh-GELU epilogue that also emits the pre-activation
  K7 bwd    fc_out dgrad GEMM with GELU-backward epilogue (dZ = (dY W2) * gelu'(Z))
  K12       wgrad GEMMs accumulate straight into the flat gradient buffer (beta = 1, no
            separate .grad add) when the parameter already has a ``.grad`` buffer
  K10+K11   LM head GEMM -> single-pass CE; dlogits overwrite the logits in place
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import reference
from ._build import load



class _LaunchCounter:
    """Counts calls into the extension (each call launches one of our kernels; a few launch two)."""

    def __init__(self):
        self.count = 0

    def reset(self):
        self.count = 0


LAUNCHES = _LaunchCounter()


class _Counting:
    def __init__(self, mod):
        self._mod = mod

    def __getattr__(self, name):
        fn = getattr(self._mod, name)
        if not callable(fn):
            return fn
        if name == "gemm":
            def wrapped(*a, **k):
                LAUNCHES.count += 1
                if "two_cta" not in k and CONFIG.gemm_2cta:
                    # round 1 measured the CTA-pair kernel ahead everywhere except the vocabulary-sized GEMMs
                    # (N or K = 50304) with its row-block-group raster; `lmhead_2cta` sends those to the CTA-pair kernel
                    # too (column-block-group raster for the forward GEMM, see ztb_gemm_2cta)
                    big = max(a[0].shape[0], a[0].shape[1], a[1].shape[0], a[1].shape[1]) > 20000
                    k["two_cta"] = (not big) or CONFIG.lmhead_2cta
                return fn(*a, **k)
        else:
            def wrapped(*a, **k):
                LAUNCHES.count += 1
                return fn(*a, **k)

        setattr(self, name, wrapped)
        return wrapped


_C = _Counting(load())

EPI_NONE, EPI_GELU, EPI_DROPOUT_ADD, EPI_GELU_BWD, EPI_ROWDOT = 0, 1, 2, 3, 4


class _Config:
    # escape hatches for bring-up / A-B measurements (the defaults are the product path)
    tc_gemm = os.environ.get("ZTB_TC_GEMM", "1") == "1"          # tcgen05 GEMM (0 -> cuBLAS everywhere)
    tc_gemm_bwd = os.environ.get("ZTB_TC_GEMM_BWD", "1") == "1"  # MN-major dgrad/wgrad on tcgen05
    fuse_wgrad = os.environ.get("ZTB_FUSE_WGRAD", "1") == "1"    # accumulate dW into param.grad in the epilogue
    gemm_2cta = os.environ.get("ZTB_GEMM_2CTA", "1") == "1"      # cta_group::2 256x256 kernel for the big GEMMs
    ce_chunk_rows = int(os.environ.get("ZTB_CE_CHUNK_ROWS", "0"))   # > 0: chunked LM head + CE (logits of one chunk at a time)
    lmhead_2cta = os.environ.get("ZTB_LMHEAD_2CTA", "0") == "1"   # vocabulary-sized GEMMs on the CTA-pair kernel as well
    attn_impl = os.environ.get("ZTB_ATTN", "auto")                # auto | tc (own kernel) | flash (flash_attn library) | ref
    # head_dim 128: accumulate dQ over key blocks directly in the bf16 dqkv buffer (no fp32 scratch, zero-fill, convert pass).
    # "auto" = on up to 2048 tokens (<= 16 bf16 roundings per element), fp32 scratch beyond that and in deterministic mode;
    # measured -3.7 ms per 1.3B step (profiles/bench_r2_ab.jsonl), same loss trajectory.
    attn_dq_bf16 = os.environ.get("ZTB_ATTN_DQ_BF16", "auto")
    # order-fixed reductions where the fast path uses atomics (LayerNorm weight gradient): bitwise run-to-run
    # reproducibility for T <= 256 (longer sequences: attention dQ has > 2 reduce-add contributions per element)
    deterministic = os.environ.get("ZTB_DETERMINISTIC", "0") == "1"
    fuse_attn_delta = os.environ.get("ZTB_FUSE_ATTN_DELTA", "1") == "1"   # delta from the out-projection dgrad GEMM's epilogue
    embedding_bwd_sorted = os.environ.get("ZTB_EMB_BWD_SORTED", "1") == "1"  # fp32 per-token sums, one add per gradient row
    # LayerNorm-backward emits the dropout-masked gradient as a second output (one extra store instead of a read + write +
    # launch).  The keep decisions come as 1 bit per element from the forward GEMM epilogue, which computes them anyway.
    # (First version regenerated them with Philox inside LayerNorm backward -- a kernel that runs 8 warps per SM because of its
    # 64 accumulator registers: 56 -> 159 us per call, measured in r2 call 1; that variant is `fold_dropout_philox`, off.)
    fold_dropout_bwd = os.environ.get("ZTB_FOLD_DROPOUT_BWD", "1") == "1"
    fold_dropout_philox = os.environ.get("ZTB_FOLD_DROPOUT_PHILOX", "0") == "1"


CONFIG = _Config()

# Gradient write mode for parameters that already own a ``.grad`` buffer (the ZeRO-1 flat gradient
# buffer): accumulate (beta = 1) or overwrite (beta = 0).  The engine switches to overwrite for the
# first micro-batch of a step, which removes the gradient memset and the read half of the first
# accumulation.
GRAD_MODE = {"accumulate": True, "lazy_zero": False}   # lazy_zero: the engine skipped the memset; the next writer of each buffer overwrites

# ---------------------------------------------------------------------------
# Dropout RNG: one device-resident (seed, offset) pair per device so the whole
# step is CUDA-graph capturable.  Each dropout site of a forward pass gets a
# distinct op_id; the offset advances once per forward (= micro-step).
# ---------------------------------------------------------------------------
_rng = {}
_op_counter = [0]


def rng_state(device) -> torch.Tensor:
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _rng:
        _rng[key] = torch.tensor([0x5EED5EED, 0], dtype=torch.int64, device=device)
    return _rng[key]


def manual_seed(seed: int, device=None) -> None:
    st = rng_state(device or torch.device("cuda", torch.cuda.current_device()))
    st.copy_(torch.tensor([seed, 0], dtype=torch.int64))


def begin_forward(device) -> None:
    """Call once per training forward: new dropout stream, op ids restart at 0."""
    rng_state(device)[1:2].add_(1)
    _op_counter[0] = 0


def _next_op_id() -> int:
    _op_counter[0] += 1
    return _op_counter[0]


# ---------------------------------------------------------------------------
# GEMM front-end
# ---------------------------------------------------------------------------
def _tc_ok(M: int, N: int, K: int) -> bool:
    return CONFIG.tc_gemm and M >= 128 and N >= 256 and K >= 64 and N % 8 == 0 and K % 8 == 0


def gemm_nt(a: torch.Tensor, b: torch.Tensor, **kw) -> torch.Tensor:
    """``a[M,K] @ b[N,K].T`` on the tcgen05 kernel."""
    return _C.gemm(a, b, **kw)


def _mm_fwd(x2: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    M, K = x2.shape
    ag = getattr(w, "_ztb_ag", None)
    if ag is not None and ag.pending():
        # this weight's ZeRO-1 bucket has not been all-gathered yet: gather it inside this GEMM
        if _tc_ok(M, w.shape[0], K) and x2.is_contiguous():
            LAUNCHES.count += 1
            return ag.gemm(x2, w)
        ag.fallback()
    if _tc_ok(M, w.shape[0], K):
        return _C.gemm(x2, w)
    return x2 @ w.t()


def _mm_dgrad(dy: torch.Tensor, w: torch.Tensor, **kw) -> torch.Tensor:
    """dX[M,K] = dY[M,N] @ W[N,K]  (W consumed MN-major, no transpose copy)."""
    M, N = dy.shape
    K = w.shape[1]
    if CONFIG.tc_gemm_bwd and _tc_ok(M, K, N) and M % 8 == 0:
        return _C.gemm(dy, w, b_mn=True, **kw)
    assert not kw, "fused dgrad epilogue needs the tcgen05 path"
    return dy @ w


def _mm_wgrad(dy: torch.Tensor, x2: torch.Tensor, w: torch.Tensor) -> Optional[torch.Tensor]:
    """dW[N,K] = dY[M,N]^T @ X[M,K].  When the parameter has a gradient buffer (the ZeRO-1 flat gradient buffer, bf16, or
    the engine's fp32 main-gradient buffer) the product is written / accumulated into it inside the GEMM epilogue
    (TMA store or reduce-add) and ``None`` is returned."""
    M, N = dy.shape
    K = x2.shape[1]
    tc = CONFIG.tc_gemm_bwd and _tc_ok(N, K, M) and N % 8 == 0
    tgt = _grad_target(w)
    if tgt is not None:
        acc = GRAD_MODE["accumulate"]
        if tc:
            _C.gemm(dy, x2, C=tgt, accumulate=acc, a_mn=True, b_mn=True)
        elif acc:
            tgt.addmm_(dy.t(), x2) if tgt.dtype == dy.dtype else tgt.add_((dy.t() @ x2).to(tgt.dtype))
        else:
            tgt.copy_(dy.t() @ x2)
        return None
    if tc:
        return _C.gemm(dy, x2, a_mn=True, b_mn=True)
    return dy.t() @ x2


# ---------------------------------------------------------------------------
# Dropout backward without its own kernel.  ``residual + dropout(y)`` epilogues hand their output to a LayerNorm;
# in backward that LayerNorm's kernel produces dx, which is exactly the ``dout`` whose masked copy the epilogue op's
# dgrad / wgrad GEMMs need.  The LayerNorm-backward kernel therefore emits mask(dx) / keep as a second output
# (one extra store instead of a read + write + launch): the forward op leaves a note (``_LAST_DROP``), the LayerNorm
# forward that consumes the tensor picks it up, its backward stashes the masked gradient under the op id.
_LAST_DROP = {"ptr": 0, "numel": 0, "p": 0.0, "op_id": 0, "mask": None}
_DROP_STASH = {}


def _new_keep_mask(M: int, N: int, device, p: float):
    """(M, N/32) int32 buffer for the keep bits the dropout GEMM epilogue records (1 bit per output element), or None
    when the backward mask will not be folded into LayerNorm backward anyway."""
    if p > 0.0 and CONFIG.fold_dropout_bwd and CONFIG.gemm_2cta and N % 32 == 0 and max(M, N) <= 20000:
        return torch.empty(M, N // 32, dtype=torch.int32, device=device)
    return None


def _note_dropout_output(out: torch.Tensor, p: float, op_id: int, mask: Optional[torch.Tensor] = None) -> None:
    if p > 0.0 and CONFIG.fold_dropout_bwd and (mask is not None or CONFIG.fold_dropout_philox):
        _LAST_DROP.update(ptr=out.data_ptr(), numel=out.numel(), p=float(p), op_id=int(op_id), mask=mask)


def _claim_dropout_note(x: torch.Tensor):
    note = None
    if _LAST_DROP["ptr"] and _LAST_DROP["ptr"] == x.data_ptr() and _LAST_DROP["numel"] == x.numel():
        note = (_LAST_DROP["p"], _LAST_DROP["op_id"], _LAST_DROP["mask"])
    _LAST_DROP["ptr"] = 0
    _LAST_DROP["mask"] = None
    return note


def _dropout_bwd(dout2: torch.Tensor, p: float, training: bool, op_id: int, rng: torch.Tensor) -> torch.Tensor:
    if not training or p <= 0.0:
        return dout2
    hit = _DROP_STASH.pop(op_id, None)
    if hit is not None and hit[0] == dout2.data_ptr() and hit[1].numel() == dout2.numel():
        return hit[1].view(dout2.shape)
    return _C.dropout_add(dout2, None, p, rng, op_id)


# ---------------------------------------------------------------------------
# Attention-backward `delta = rowsum(dO o O)` without its own kernel.  dO is produced by the dgrad GEMM of the output
# projection, whose saved input IS O: that GEMM's ROWDOT epilogue emits delta[b, h, t] from the rows it has in registers
# (the standalone kernel re-read 2 x 67 MB per layer).  The attention forward leaves a note, the projection that consumes
# its output picks it up, its backward stashes delta under the address of the dO tensor it returns.
_LAST_ATTN = {"ptr": 0, "numel": 0, "H": 0, "T": 0}
_DELTA_STASH = {}


def _note_attention_output(out: torch.Tensor, n_head: int, T: int) -> None:
    if CONFIG.fuse_attn_delta:
        _LAST_ATTN.update(ptr=out.data_ptr(), numel=out.numel(), H=int(n_head), T=int(T))


def _claim_attention_note(x2: torch.Tensor):
    note = None
    if _LAST_ATTN["ptr"] and _LAST_ATTN["ptr"] == x2.data_ptr() and _LAST_ATTN["numel"] == x2.numel():
        note = (_LAST_ATTN["H"], _LAST_ATTN["T"])
    _LAST_ATTN["ptr"] = 0
    return note


_LN_SCRATCH = {}


def _ln_scratch(device: torch.device, D: int):
    """Persistent (fp32 [D], int32 [1]) scratch of the LayerNorm-backward weight-gradient reduction; the kernel leaves
    both zeroed.  One pair per (device, stream, D): launches on one stream are serialised."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, D)
    if key not in _LN_SCRATCH:
        _LN_SCRATCH[key] = (torch.zeros(D, dtype=torch.float32, device=device), torch.zeros(1, dtype=torch.int32, device=device))
    return _LN_SCRATCH[key]


def _grad_target(w: torch.Tensor):
    """Where a parameter's gradient is written by the fused paths: the engine's fp32 main-gradient view if it set one,
    else the (flat-buffer) ``.grad``."""
    if not CONFIG.fuse_wgrad:
        return None
    mg = getattr(w, "_ztb_main_grad", None)
    if mg is not None:
        return mg
    g = w.grad
    return g if (g is not None and g.is_contiguous()) else None


def _ln_backward(ctx, dy, dres):
    x, w, mean, rstd = ctx.saved_tensors
    dr = dres.contiguous() if dres is not None else None
    tgt = _grad_target(w)
    kw = {}
    if tgt is not None:
        scratch, ticket = _ln_scratch(x.device, x.shape[-1])
        kw.update(grad_out=tgt, accumulate=GRAD_MODE["accumulate"], scratch=scratch, ticket=ticket, deterministic=CONFIG.deterministic)
    if ctx.drop is not None:
        kw.update(drop_p=ctx.drop[0], op_id=ctx.drop[1])
        if ctx.drop[2] is not None:
            kw.update(drop_mask=ctx.drop[2])          # keep bits recorded by the forward GEMM epilogue
        else:
            kw.update(rng=rng_state(x.device))        # regenerate the mask with Philox (slow inside this kernel, see CONFIG)
    dx, dw, dmask = _C.layernorm_bwd(dy.contiguous(), x, w.contiguous(), mean, rstd, dr, **kw)
    if ctx.drop is not None:
        _DROP_STASH[ctx.drop[1]] = (dx.data_ptr(), dmask)
    return dx, (None if tgt is not None else dw.to(w.dtype))


# ---------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, eps):
        xc = x.contiguous()
        ctx.drop = _claim_dropout_note(xc)
        y, mean, rstd = _C.layernorm_fwd(xc, w.contiguous(), eps)
        ctx.save_for_backward(xc, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        dx, dw = _ln_backward(ctx, dy, None)
        return dx, dw, None


def layernorm(x, weight, eps):
    return _LayerNorm.apply(x, weight, eps)


class _LayerNormRes(torch.autograd.Function):
    """LayerNorm that also hands back its input for the residual connection.  Because both uses
    of ``x`` leave through this node, backward receives the residual-branch gradient together with
    the LayerNorm gradient and adds them inside the LayerNorm-backward kernel (no separate add)."""

    @staticmethod
    def forward(ctx, x, w, eps):
        xc = x.contiguous()
        ctx.drop = _claim_dropout_note(xc)
        y, mean, rstd = _C.layernorm_fwd(xc, w.contiguous(), eps)
        ctx.save_for_backward(xc, w, mean, rstd)
        return y, xc.view_as(xc)

    @staticmethod
    def backward(ctx, dy, dres):
        dx, dw = _ln_backward(ctx, dy, dres)
        return dx, dw, None


def layernorm_res(x, weight, eps):
    return _LayerNormRes.apply(x, weight, eps)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        x2 = x.reshape(-1, x.shape[-1])
        ctx.save_for_backward(x2, w)
        ctx.xshape = x.shape
        return _mm_fwd(x2, w).view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = _mm_dgrad(dy2, w).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = _mm_wgrad(dy2, x2, w) if ctx.needs_input_grad[1] else None
        return dx, dw


def linear(x, w):
    return _Linear.apply(x, w)


class _LinearDropoutAdd(torch.autograd.Function):
    """``residual + dropout(x @ w.T)`` in one GEMM (epilogue does mask + scale + add)."""

    @staticmethod
    def forward(ctx, x, w, residual, p, training):
        x2 = x.reshape(-1, x.shape[-1])
        r2 = residual.reshape(-1, residual.shape[-1])
        M, K = x2.shape
        N = w.shape[0]
        pe = float(p) if training else 0.0
        rng = rng_state(x.device)
        op_id = _next_op_id() if pe > 0 else 0
        mask = None
        if _tc_ok(M, N, K):
            mask = _new_keep_mask(M, N, x2.device, pe)
            out = _C.gemm(x2, w, epilogue=EPI_DROPOUT_ADD, aux_in=r2, p=pe, rng=rng, op_id=op_id,
                          **({"mask_out": mask, "two_cta": True} if mask is not None else {}))
        else:
            y = x2 @ w.t()
            out = _C.dropout_add(y, r2.contiguous(), pe, rng, op_id) if y.numel() % 8 == 0 else r2 + y
        ctx.save_for_backward(x2, w)
        ctx.meta = (x.shape, pe, op_id)
        ctx.attn = _claim_attention_note(x2)      # (heads, T) when x is the output of our attention kernel
        _note_dropout_output(out, pe, op_id, mask)
        return out.view(residual.shape)

    @staticmethod
    def backward(ctx, dout):
        x2, w = ctx.saved_tensors
        xshape, pe, op_id = ctx.meta
        d2 = dout.reshape(-1, dout.shape[-1])
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        dy = _dropout_bwd(d2, pe, pe > 0, op_id, rng_state(dout.device))
        dx = None
        if ctx.needs_input_grad[0]:
            M, N = dy.shape
            Kd = w.shape[1]
            hd = Kd // ctx.attn[0] if ctx.attn else 0
            if (ctx.attn and hd in (64, 128) and CONFIG.gemm_2cta and CONFIG.tc_gemm_bwd and _tc_ok(M, Kd, N) and M % 8 == 0
                    and M % ctx.attn[1] == 0 and max(M, N, Kd) <= 20000):
                H, T = ctx.attn
                delta = torch.empty(M // T, H, T, dtype=torch.float32, device=dy.device)
                dx = _C.gemm(dy, w, b_mn=True, two_cta=True, epilogue=EPI_ROWDOT, aux_in=x2, rowdot_out=delta, rowdot_T=T, rowdot_hd=hd)
                _DELTA_STASH[dx.data_ptr()] = delta
                dx = dx.view(xshape)
            else:
                dx = _mm_dgrad(dy, w).view(xshape)
        dw = _mm_wgrad(dy, x2, w) if ctx.needs_input_grad[1] else None
        return dx, dw, dout, None, None


def linear_dropout_add(x, w, residual, p, training):
    return _LinearDropoutAdd.apply(x, w, residual, p, training)


class _MLP(torch.autograd.Function):
    """``residual + dropout(gelu(x W1^T) W2^T)`` -- 2 GEMMs forward, 4 backward, every
    elementwise op folded into a GEMM epilogue except the backward dropout mask."""

    @staticmethod
    def forward(ctx, x, w1, w2, residual, p, training):
        x2 = x.reshape(-1, x.shape[-1])
        r2 = residual.reshape(-1, residual.shape[-1])
        M, D = x2.shape
        H = w1.shape[0]
        pe = float(p) if training else 0.0
        rng = rng_state(x.device)
        op_id = _next_op_id() if pe > 0 else 0
        if _tc_ok(M, H, D):
            z = torch.empty(M, H, dtype=x2.dtype, device=x2.device)
            g = _C.gemm(x2, w1, epilogue=EPI_GELU, aux_out=z)
        else:
            z = x2 @ w1.t()
            g = _C.gelu_fwd(z) if z.numel() % 8 == 0 else torch.nn.functional.gelu(z, approximate="tanh")
        mask = None
        if _tc_ok(M, D, H):
            mask = _new_keep_mask(M, D, x2.device, pe)
            out = _C.gemm(g, w2, epilogue=EPI_DROPOUT_ADD, aux_in=r2, p=pe, rng=rng, op_id=op_id,
                          **({"mask_out": mask, "two_cta": True} if mask is not None else {}))
        else:
            y = g @ w2.t()
            out = _C.dropout_add(y, r2.contiguous(), pe, rng, op_id) if y.numel() % 8 == 0 else r2 + y
        ctx.save_for_backward(x2, w1, w2, z, g)
        ctx.meta = (x.shape, pe, op_id)
        _note_dropout_output(out, pe, op_id, mask)
        return out.view(residual.shape)

    @staticmethod
    def backward(ctx, dout):
        x2, w1, w2, z, g = ctx.saved_tensors
        xshape, pe, op_id = ctx.meta
        d2 = dout.reshape(-1, dout.shape[-1])
        if not d2.is_contiguous():
            d2 = d2.contiguous()
        dy = _dropout_bwd(d2, pe, pe > 0, op_id, rng_state(dout.device))
        M, D = dy.shape
        H = w2.shape[1]
        dw2 = _mm_wgrad(dy, g, w2)
        if CONFIG.tc_gemm_bwd and _tc_ok(M, H, D):
            dz = _C.gemm(dy, w2, epilogue=EPI_GELU_BWD, aux_in=z, b_mn=True)  # (dY W2) * gelu'(Z)
        else:
            dg = dy @ w2
            dz = _C.gelu_bwd(dg.contiguous(), z) if z.numel() % 8 == 0 else dg * _gelu_grad_ref(z)
        dw1 = _mm_wgrad(dz, x2, w1)
        dx = _mm_dgrad(dz, w1).view(xshape)
        return dx, dw1, dw2, dout, None, None


def _gelu_grad_ref(z):
    zf = z.float().requires_grad_(True)
    with torch.enable_grad():
        (g,) = torch.autograd.grad(torch.nn.functional.gelu(zf, approximate="tanh").sum(), zf)
    return g.to(z.dtype)


def mlp(x, w1, w2, residual, p, training):
    return _MLP.apply(x, w1, w2, residual, p, training)


class _LinearGelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        x2 = x.reshape(-1, x.shape[-1])
        M, K = x2.shape
        if _tc_ok(M, w.shape[0], K):
            z = torch.empty(M, w.shape[0], dtype=x2.dtype, device=x2.device)
            g = _C.gemm(x2, w, epilogue=EPI_GELU, aux_out=z)
        else:
            z = x2 @ w.t()
            g = torch.nn.functional.gelu(z, approximate="tanh")
        ctx.save_for_backward(x2, w, z)
        ctx.xshape = x.shape
        return g.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dg):
        x2, w, z = ctx.saved_tensors
        dg2 = dg.reshape(-1, dg.shape[-1]).contiguous()
        dz = _C.gelu_bwd(dg2, z) if z.numel() % 8 == 0 else dg2 * _gelu_grad_ref(z)
        return _mm_dgrad(dz, w).view(ctx.xshape), _mm_wgrad(dz, x2, w)


def linear_gelu(x, w):
    return _LinearGelu.apply(x, w)


class _DropoutAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, residual, p, training):
        pe = float(p) if training else 0.0
        op_id = _next_op_id() if pe > 0 else 0
        ctx.meta = (pe, op_id)
        return _C.dropout_add(y.contiguous(), residual.contiguous(), pe, rng_state(y.device), op_id)

    @staticmethod
    def backward(ctx, dout):
        pe, op_id = ctx.meta
        return _dropout_bwd(dout.contiguous(), pe, pe > 0, op_id, rng_state(dout.device)), dout, None, None


def dropout_add(y, residual, p, training):
    if y.numel() % 8:
        return reference.dropout_add(y, residual, p, training)
    return _DropoutAdd.apply(y, residual, p, training)


# ---------------------------------------------------------------------------
class _Embedding(torch.autograd.Function):
    @staticmethod
    def forward(ctx, idx, wte):
        idx32 = idx.to(torch.int32).contiguous()
        ctx.save_for_backward(idx32)
        ctx.wte = wte
        return _C.embedding_fwd(idx32, wte)

    @staticmethod
    def backward(ctx, dout):
        (idx32,) = ctx.saved_tensors
        wte = ctx.wte
        d = dout.contiguous()
        tgt = _grad_target(wte)
        dw = tgt if tgt is not None else torch.zeros_like(wte)
        # Scatter-add into the shared (tied) gradient buffer.  The LM-head wgrad of the same backward pass has already
        # written (overwrite mode) or accumulated into it, so this is always an accumulation.  Default: sort the ids, sum
        # each token's occurrences in fp32 and add once per row (deterministic, no per-occurrence bf16 rounding).
        if CONFIG.embedding_bwd_sorted:
            stok, order = torch.sort(idx32.view(-1).to(torch.int64), stable=True)
            _C.embedding_bwd_sorted_(stok.to(torch.int32), order.to(torch.int32), d.view(-1, d.shape[-1]), dw)
        else:
            _C.embedding_bwd_(idx32, d, dw)
        return None, (None if tgt is not None else dw)


def embedding(idx, wte):
    if wte.shape[1] % 8:
        return reference.embedding(idx, wte)
    return _Embedding.apply(idx, wte)


# ---------------------------------------------------------------------------
class _LMHeadLoss(torch.autograd.Function):
    """Tied LM head + shifted cross entropy.  The bf16 logits are produced by the tcgen05 GEMM,
    reduced to per-row log-sum-exp in one pass, and overwritten in place by dlogits in backward
    (no fp32 logits, no one-hot, no softmax tensor; reference GPT.py:100-111 materialises all three)."""

    @staticmethod
    def forward(ctx, h, wte, tokens):
        B, T, D = h.shape
        h2 = h.reshape(B * T, D)
        logits = _mm_fwd(h2, wte)  # (B*T, V) bf16
        labels = torch.full((B, T), -1, dtype=torch.int32, device=h.device)
        labels[:, :-1] = tokens[:, 1:].to(torch.int32)
        labels = labels.view(-1)
        lse, nll = _C.ce_fwd(logits, labels)
        count = B * (T - 1)
        ctx.save_for_backward(h2, wte, labels, lse)
        ctx.logits = logits
        ctx.meta = (h.shape, count)
        return nll.sum() / count

    @staticmethod
    def backward(ctx, gout):
        h2, wte, labels, lse = ctx.saved_tensors
        hshape, count = ctx.meta
        logits = ctx.logits
        ctx.logits = None
        _C.ce_bwd_(logits, labels, lse, gout.reshape(1).float().contiguous(), 1.0 / count)
        dh = _mm_dgrad(logits, wte).view(hshape)
        dw = _mm_wgrad(logits, h2, wte)
        return dh, dw, None


class _LMHeadLossChunked(torch.autograd.Function):
    """Chunked LM head + cross entropy (``training.ce_chunk_rows`` / ``ZTB_CE_CHUNK_ROWS``): the token
    rows are processed ``chunk`` at a time -- logits GEMM, single-pass CE forward, dlogits in place, dgrad GEMM into the
    rows' slice of dh, wgrad accumulated into the tied embedding's gradient buffer -- so only ``chunk x V`` logits ever
    exist (1.65 GB -> 0.1-0.2 GB at 16k tokens per micro-batch) and nothing is recomputed in backward.

    Trade: the gradients are produced during *forward* under the assumption that the loss is differentiated with a unit
    upstream gradient (what ``Zero1Engine.backward`` / ``loss.backward()`` do); backward only scales the stored dh by the
    actual upstream gradient, so a caller that backpropagates ``k * loss`` gets dh right but the embedding's head gradient
    unscaled -- use the unchunked path for that.  The per-chunk wgrad (K = chunk rows) re-reads and re-writes the whole
    (V, D) gradient each time, which is why this is an option for memory-bound shapes and not the default on a 180 GB part."""

    @staticmethod
    def forward(ctx, h, wte, tokens, chunk):
        B, T, D = h.shape
        M = B * T
        h2 = h.reshape(M, D)
        labels = torch.full((B, T), -1, dtype=torch.int32, device=h.device)
        labels[:, :-1] = tokens[:, 1:].to(torch.int32)
        labels = labels.view(-1)
        count = B * (T - 1)
        need_grad = any(ctx.needs_input_grad[:2])   # (grad mode is off inside Function.forward; this is the caller's view)
        dh = torch.empty_like(h2) if need_grad else None
        one = torch.ones(1, dtype=torch.float32, device=h.device)
        total = torch.zeros((), dtype=torch.float32, device=h.device)
        tgt = _grad_target(wte) if need_grad else None
        dw = None
        first_acc = GRAD_MODE["accumulate"] and not GRAD_MODE.get("lazy_zero", False)
        for i, s0 in enumerate(range(0, M, chunk)):
            e0 = min(s0 + chunk, M)
            logits = _mm_fwd(h2[s0:e0], wte)
            lse, nll = _C.ce_fwd(logits, labels[s0:e0])
            total = total + nll.sum()
            if not need_grad:
                continue
            _C.ce_bwd_(logits, labels[s0:e0], lse, one, 1.0 / count)
            if CONFIG.tc_gemm_bwd and _tc_ok(e0 - s0, D, wte.shape[0]) and (e0 - s0) % 8 == 0:
                _C.gemm(logits, wte, C=dh[s0:e0], b_mn=True)      # dgrad straight into the rows' slice of dh
            else:
                dh[s0:e0].copy_(logits @ wte)
            if tgt is not None:
                acc = first_acc or i > 0
                if CONFIG.tc_gemm_bwd and _tc_ok(wte.shape[0], D, e0 - s0):
                    _C.gemm(logits, h2[s0:e0], C=tgt, accumulate=acc, a_mn=True, b_mn=True)
                elif acc:
                    tgt.add_((logits.t() @ h2[s0:e0]).to(tgt.dtype))
                else:
                    tgt.copy_(logits.t() @ h2[s0:e0])
            else:
                part = logits.t() @ h2[s0:e0]
                dw = part if dw is None else dw + part
        ctx.save_for_backward(dh if dh is not None else h2.new_empty(0), dw if dw is not None else h2.new_empty(0))
        ctx.meta = (h.shape, dw is not None)
        return total / count

    @staticmethod
    def backward(ctx, gout):
        dh, dw = ctx.saved_tensors
        hshape, has_dw = ctx.meta
        g = gout.reshape(1).to(dh.dtype)
        return (dh * g).view(hshape), ((dw * g).to(dw.dtype) if has_dw else None), None, None


def lm_head_loss(h, wte, tokens):
    if wte.shape[0] % 8 or h.dtype != torch.bfloat16:
        return reference.lm_head_loss(h, wte, tokens)
    if CONFIG.ce_chunk_rows > 0 and h.shape[0] * h.shape[1] > CONFIG.ce_chunk_rows:
        return _LMHeadLossChunked.apply(h, wte, tokens, int(CONFIG.ce_chunk_rows))
    return _LMHeadLoss.apply(h, wte, tokens)


def cross_entropy(logits, labels):
    # logits-returning API path (not on the training hot path): reference math on the GPU
    return reference.cross_entropy(logits, labels)


# ---------------------------------------------------------------------------
def attention(qkv, slopes, n_head, scale, p, training):
    impl = CONFIG.attn_impl
    if impl in ("auto", "tc"):
        from .attention_cuda import alibi_attention, supported

        if supported(qkv, n_head):
            return alibi_attention(qkv, slopes, n_head, scale, p if training else 0.0)
        if impl == "tc":
            raise RuntimeError("tcgen05 attention needs bf16, T % 128 == 0 and head_dim in {64, 96, 128}")
    if impl in ("auto", "flash"):
        from flash_attn import flash_attn_qkvpacked_func

        B, T, D3 = qkv.shape
        hd = D3 // 3 // n_head
        out = flash_attn_qkvpacked_func(qkv.view(B, T, 3, n_head, hd), dropout_p=p if training else 0.0, softmax_scale=scale,
                                        causal=True, alibi_slopes=slopes.float() if slopes is not None else None)
        return out.reshape(B, T, n_head * hd)
    return reference.attention(qkv, slopes, n_head, scale, p, training)
