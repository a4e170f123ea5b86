"""Drop-in for the batch similarity-matrix contrastive losses of /root/reference/lib/losses.py,
computed on sm_100a:

  NTXentLoss(temperature=0.1).forward(z_label, z_idx, z, extra=None) -> (loss, logdict)   lib/losses.py:10-73
  CLEWSLoss(gamma, b, eps, epsilon, uniformity_weight, warmup_steps)
      .forward(z_label, z_idx, z, extra=None, numerically_friendly=True) -> (loss, logdict) lib/losses.py:176-285
  TripletLoss(margin=0.2, p=2, eps=1e-6, swap=False, reduction='mean')
      .forward(z_label, z_idx, z, extra=None) -> (loss, logdict)                          lib/losses.py:76-171
      (the Python mining loop becomes one kernel, the margin loss a fused fwd / bwd pair)

Same constructor arguments, forward signature, logdict keys and side effects (a single-label batch
mutates the caller's z_label in place, lib/losses.py:34-35 / 221-222).  Forward and backward run in
the fused tcgen05 kernels behind wealy_loss_forward / wealy_loss_backward (include/wealy_b200.h):
the B x B similarity matrix is never stored; the backward recomputes it on the tensor cores.
`loss` is differentiable w.r.t. z through a torch.autograd.Function; the logdict entries are
detached diagnostics.  Output dtype: like the reference (lib/losses.py:65-66 computes the loss from z's
own dtype), `loss` and the logdict entries carry z's dtype -- the kernels accumulate in fp32 / fp64 and the
result is rounded once at the end; pass `loss_dtype=torch.float32` to the constructor to keep the unrounded
value for fp16 / bf16 batches.  float64 batches are rounded to float32 on entry (the tensor-core path is
fp32-grade: loss within 1e-6 relative of the float64 value) and loss / gradient are returned as float64.
Inputs must be CUDA tensors.
"""
import ctypes

import torch
import torch.nn as nn

from . import _native as N
from .tensor_ops import passes_of


def _label_noise_(z_label):
    """lib/losses.py:34-35: if the batch holds a single label, overwrite the first max(2, 1%) labels
    with -1 IN PLACE.  Done without a host sync (the reference's .unique() forces one)."""
    n = len(z_label)
    k = max(2, int(n * 0.01))
    single = (z_label == z_label[0]).all()
    head = z_label[:k]
    head.copy_(torch.where(single, torch.full_like(head, -1), head))


_HALF = (torch.float16, torch.bfloat16)


class _FusedLoss(torch.autograd.Function):
    """forward: ONE C call (ids packed + single-label noise in place + prep + statistics sweep + merge + finish) that also
    writes the loss / logdict numbers in z's dtype; backward: ONE C call that reads the upstream gradient in whatever
    dtype autograd hands it over.  No ATen kernel runs between the module call and z.grad."""

    @staticmethod
    def forward(ctx, z, z_label, z_idx, cfg_items, loss_dtype, noise):
        N.require_cuda(z, z_label, z_idx)
        ctx.in_dtype = z.dtype
        if z.dtype == torch.float64:
            z = z.float()
        code = N.dtype_code(z.dtype)
        zz = z if z.stride(1) == 1 else z.contiguous()
        b, d = zz.shape
        # the single-label noise of lib/losses.py:34-35 mutates the CALLER's labels: fused into the id-packing kernel when
        # they are an int64 tensor the kernel can write in place, applied up front otherwise
        fuse = noise and z_label.dtype == torch.long and z_label.is_contiguous() and b <= 65536
        if noise and not fuse:
            _label_noise_(z_label)
        lab = z_label if fuse else z_label.to(torch.long).contiguous()
        idx = z_idx.to(torch.long).contiguous()
        cfg = N.LossCfg(**dict(cfg_items), label_noise=1 if fuse else 0)
        want = z.dtype if loss_dtype is None else loss_dtype
        with torch.cuda.device(zz.device):
            ws_bytes = N.lib.wealy_loss_workspace_bytes(b, d, cfg.passes)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=zz.device)
            out = torch.empty(N.OUT_COUNT, dtype=torch.float64, device=zz.device)
            cast = torch.empty(N.OUT_COUNT, dtype=z.dtype, device=zz.device)
            N.check(N.lib.wealy_loss_forward(ctypes.byref(cfg), zz.data_ptr(), b, zz.stride(0), d, code, lab.data_ptr(),
                                             idx.data_ptr(), out.data_ptr(), cast.data_ptr(), ws.data_ptr(), ws_bytes,
                                             N.stream_ptr(zz.device)))
        ctx.save_for_backward(zz, ws)
        ctx.cfg_items = cfg_items
        if ctx.in_dtype == torch.float64 and loss_dtype is None:
            want = torch.float64
        stats = cast if want == z.dtype else out.to(want)
        ctx.mark_non_differentiable(stats)
        return stats[0], stats

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        zz, ws = ctx.saved_tensors
        b, d = zz.shape
        dz = torch.empty_like(zz)
        g = grad_loss.detach()
        if g.dtype not in (torch.float32,) + _HALF:
            g = g.to(torch.float32)
        g = g.reshape(1)
        cfg = N.LossCfg(**dict(ctx.cfg_items), grad_dtype=N.dtype_code(g.dtype))
        with torch.cuda.device(zz.device):
            N.check(N.lib.wealy_loss_backward(ctypes.byref(cfg), zz.data_ptr(), b, zz.stride(0), d,
                                              N.dtype_code(zz.dtype), g.data_ptr(), dz.data_ptr(), dz.stride(0),
                                              ws.data_ptr(), ws.numel(), N.stream_ptr(zz.device)))
        return (dz if ctx.in_dtype == zz.dtype else dz.to(ctx.in_dtype)), None, None, None, None, None


def _passes(precision, z):
    """precision=None: fp16x3 (fp32-grade) for fp32 inputs; one fp16 pass for fp16 / bf16 inputs, whose own
    8-11 significant bits are fewer than the pass keeps (loss error ~1e-5, gradient rel-L2 ~3e-4)."""
    if precision is None and z.dtype in (torch.float16, torch.bfloat16):
        return 1
    return passes_of(precision)


def _run(z, z_label, z_idx, loss_dtype=None, noise=True, **cfg):
    """-> (loss, stats[WEALY_OUT_COUNT]) both in the output dtype (z's unless loss_dtype says otherwise)."""
    base = dict(kind=0, passes=3, temperature=1.0, gamma=0.0, b=0.0, eps=1e-8, epsilon=1e-6, uw=0.0,
                numerically_friendly=1)
    base.update(cfg)
    return _FusedLoss.apply(z, z_label, z_idx, tuple(sorted(base.items())), loss_dtype, noise)


class NTXentLoss(nn.Module):
    """lib/losses.py:10-73."""

    def __init__(self, temperature=0.1, precision=None, loss_dtype=None):
        super().__init__()
        self.tau = temperature
        self.precision = precision
        self.loss_dtype = loss_dtype

    def forward(self, z_label, z_idx, z, extra=None):
        assert len(z_label) == len(z_idx) and len(z_label) == len(z)
        N.require_cuda(z, z_label, z_idx)
        loss, stats = _run(z, z_label, z_idx, self.loss_dtype, kind=N.LOSS_NTXENT, passes=_passes(self.precision, z),
                           temperature=float(self.tau))            # (label noise of :34-35 happens inside)
        logdict = {"l_main": loss, "v_zmax": stats[1], "v_zmean": stats[2], "v_zstd": stats[3]}
        return loss, logdict


class CLEWSLoss(nn.Module):
    """lib/losses.py:176-285 (CLEWS-style alignment + uniformity on cosine distances)."""

    def __init__(self, gamma: float = 8.0, b: float = 1.0, eps: float = 1e-8, epsilon: float = 1e-6,
                 uniformity_weight: float = 0.5, warmup_steps: int = 1000, precision=None, loss_dtype=None):
        super().__init__()
        self.gamma = float(gamma)
        self.b = float(b)
        self.eps = float(eps)
        self.epsilon = float(epsilon)
        self.uniformity_weight = float(uniformity_weight)
        self.warmup_steps = int(warmup_steps)
        self.precision = precision
        self.loss_dtype = loss_dtype

    def forward(self, z_label, z_idx, z, extra=None, numerically_friendly=True):
        if z.dim() == 3:
            assert z.size(1) == 1, f"CLEWS (vector) expects S=1, got S={z.size(1)}"
            z = z.squeeze(1)
        assert z.dim() == 2
        B = z.size(0)
        assert len(z_label) == len(z_idx) == B and B >= 4
        N.require_cuda(z, z_label, z_idx)
        # warm-up of the uniformity weight (lib/losses.py:248-258); the label noise of :221-222 happens inside _run
        uw = self.uniformity_weight
        if self.warmup_steps > 0:
            step = None
            if isinstance(extra, dict) and "global_step" in extra:
                step = int(extra["global_step"])
            elif hasattr(self, "global_step"):
                step = int(self.global_step)
            if step is not None:
                uw = float(min(self.uniformity_weight, self.uniformity_weight * (step + 1) / self.warmup_steps))
        loss, stats = _run(z, z_label, z_idx, self.loss_dtype, kind=N.LOSS_CLEWS, passes=_passes(self.precision, z),
                           gamma=self.gamma, b=self.b, eps=self.eps, epsilon=self.epsilon, uw=uw,
                           numerically_friendly=1 if numerically_friendly else 0)
        logdict = {
            "l_main": loss,
            "l_cent": stats[4],
            "l_cont": stats[5],
            "cnt_pos_pairs": stats[6],
            "cnt_neg_pairs": stats[7],
            "anchors_with_pos": stats[8],
            "v_dpos": stats[9],
            "v_dneg": stats[10],
            "uniformity_weight": torch.full((), uw, device=z.device),
            "z_max": stats[1],
            "z_mean": stats[2],
            "z_std": stats[3],
        }
        return loss, logdict


class _TripletFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, pos, neg, margin, p, eps, swap, reduction):
        b, d = z.shape
        zz = z if z.stride(1) == 1 else z.contiguous()
        rows = torch.empty(b, 4, dtype=torch.float32, device=z.device)
        acc = torch.empty(2, dtype=torch.float64, device=z.device)
        with torch.cuda.device(z.device):
            N.check(N.lib.wealy_triplet_forward(zz.data_ptr(), zz.stride(0), b, d, N.dtype_code(zz.dtype), pos.data_ptr(),
                                                neg.data_ptr(), margin, p, eps, 1 if swap else 0, rows.data_ptr(),
                                                acc.data_ptr(), N.stream_ptr(z.device)))
        ctx.save_for_backward(zz, pos, neg, rows, acc)
        ctx.cfg = (p, eps, swap, reduction)
        out_dtype = torch.float32 if z.dtype in (torch.float16, torch.bfloat16) else z.dtype
        if reduction == "none":
            out = rows[:, 3].to(out_dtype)          # per anchor; -1 marks anchors without a triplet
        elif reduction == "sum":
            out = acc[0].to(out_dtype)
        else:
            out = (acc[0] / acc[1].clamp(min=1.0)).to(out_dtype)
        ctx.mark_non_differentiable(acc)
        return out, acc

    @staticmethod
    def backward(ctx, g, _gacc):
        zz, pos, neg, rows, acc = ctx.saved_tensors
        p, eps, swap, reduction = ctx.cfg
        b, d = zz.shape
        dz = torch.empty(b, d, dtype=torch.float32, device=zz.device)
        up = g.detach().to(torch.float32).reshape(-1).contiguous()
        with torch.cuda.device(zz.device):
            N.check(N.lib.wealy_triplet_backward(zz.data_ptr(), zz.stride(0), b, d, N.dtype_code(zz.dtype), pos.data_ptr(),
                                                 neg.data_ptr(), p, eps, 1 if swap else 0, rows.data_ptr(), up.data_ptr(),
                                                 1 if reduction == "none" else 0, 1 if reduction == "mean" else 0,
                                                 acc.data_ptr(), dz.data_ptr(), N.stream_ptr(zz.device)))
        return dz.to(zz.dtype), None, None, None, None, None, None, None


def mine_triplets(z_label, z_idx):
    """lib/losses.py:140-171 without the Python loop: per anchor the FIRST index with the same label and a different
    idx, and the FIRST index with a different label (-1 where there is none).  -> (positives[B], negatives[B]) int64."""
    N.require_cuda(z_label, z_idx)
    lab = z_label.to(torch.long).contiguous()
    idx = z_idx.to(torch.long).contiguous()
    b = lab.numel()
    pos = torch.empty(b, dtype=torch.long, device=lab.device)
    neg = torch.empty(b, dtype=torch.long, device=lab.device)
    with torch.cuda.device(lab.device):
        N.check(N.lib.wealy_triplet_mine(lab.data_ptr(), idx.data_ptr(), b, pos.data_ptr(), neg.data_ptr(),
                                         N.stream_ptr(lab.device)))
    return pos, neg


class TripletLoss(nn.Module):
    """lib/losses.py:76-171.  Same constructor / forward / logdict; one host read (the triplet count, needed for
    the reference's empty-batch branch) instead of the reference's 2*B `.item()` calls."""

    def __init__(self, margin=0.2, p=2, eps=1e-6, swap=False, reduction="mean"):
        super().__init__()
        assert reduction in ("mean", "sum", "none")
        self.margin, self.p, self.eps, self.swap, self.reduction = float(margin), float(p), float(eps), bool(swap), reduction

    def _create_triplets(self, z_label, z_idx):
        """-> (anchors, positives, negatives) index tensors, as lib/losses.py:140-171."""
        pos, neg = mine_triplets(z_label, z_idx)
        anchors = torch.nonzero((pos >= 0) & (neg >= 0)).flatten()
        return anchors, pos[anchors], neg[anchors]

    def forward(self, z_label, z_idx, z, extra=None):
        assert len(z_label) == len(z_idx) and len(z_label) == len(z)
        N.require_cuda(z, z_label, z_idx)
        _label_noise_(z_label)                                  # lib/losses.py:107-108
        pos, neg = mine_triplets(z_label, z_idx)
        out, acc = _TripletFn.apply(z, pos, neg, self.margin, self.p, self.eps, self.swap, self.reduction)
        n_triplets = int(acc[1].item())
        stats = {"v_zmax": z.detach().abs().max(), "v_zmean": z.detach().mean(), "v_zstd": z.detach().std()}
        if n_triplets == 0:                                     # lib/losses.py:113-123
            loss = torch.tensor(0.0, device=z.device, requires_grad=True)
            return loss, {"l_main": loss, **stats, "n_triplets": 0}
        loss = out[out >= 0] if self.reduction == "none" else out
        return loss, {"l_main": loss, **stats}
