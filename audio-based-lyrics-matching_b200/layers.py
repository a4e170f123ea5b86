"""The pooled-embedding producer immediately upstream of the hot path, on sm_100a:

  MeanPool().forward(x, mask=None) -> (B, C)              /root/reference/lib/layers.py:6-30
  avg_pool_tracks(frames, offsets) -> (K, E) fp32         the `use_avg_pooling` branch of the collate,
      /root/reference/lib/embedding_dataset/collate_functions.py:131-172 (`emb.mean(dim=0)` per track), fed by the
      fp16-stored / fp32-used embeddings of lib/embedding_dataset/base_dataset.py:229-233

Both are one HBM pass behind the C ABI (wealy_mean_pool / wealy_segment_mean, include/wealy_b200.h); the output of
`avg_pool_tracks` is exactly the [N, E] fp32 matrix `wealy_b200.evaluation.evaluate` consumes, so pooled
embeddings never make a host round trip.  CUDA tensors only (no CPU fallback).
"""
import torch
import torch.nn as nn

from . import _native as N


def _mean_pool_call(x, mask_u8, b, c, t, out, backward):
    with torch.cuda.device(x.device):
        N.check(N.lib.wealy_mean_pool(x.data_ptr(), mask_u8.data_ptr() if mask_u8 is not None else None, b, c, t,
                                      N.dtype_code(x.dtype), out.data_ptr(), 1 if backward else 0,
                                      N.stream_ptr(x.device)))


class _MeanPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mask_u8):
        b, c, t = x.shape
        xx = x.contiguous()
        out = torch.empty(b, c, dtype=x.dtype, device=x.device)
        _mean_pool_call(xx, mask_u8, b, c, t, out, False)
        ctx.shape = (b, c, t)
        ctx.mask = mask_u8
        return out

    @staticmethod
    def backward(ctx, g):
        b, c, t = ctx.shape
        gg = g.contiguous()
        dx = torch.empty(b, c, t, dtype=g.dtype, device=g.device)
        _mean_pool_call(gg, ctx.mask, b, c, t, dx, True)
        return dx, None


class MeanPool(nn.Module):
    """lib/layers.py:6-30: masked temporal mean, `sum_t x*mask / (sum_t mask + 1e-8)`; plain mean without a mask."""

    def forward(self, x, mask=None):
        assert x.dim() == 3, "MeanPool expects (B, C, T)"
        N.require_cuda(x)
        m = None
        if mask is not None:
            assert mask.shape == (x.shape[0], x.shape[2]), "mask must be (B, T)"
            N.require_cuda(mask)
            m = (mask != 0).to(torch.uint8).contiguous()
        return _MeanPoolFn.apply(x, m)


def avg_pool_tracks(frames, offsets=None):
    """Per-track temporal mean of ragged frame embeddings -> [K, E] fp32 (what the evaluator consumes).

    `frames`: a list of K tensors [T_k, E] (the per-track `.pt` payloads, fp16 or fp32), or one concatenated
    [sum_T, E] CUDA tensor together with `offsets` [K + 1] (int64, track k owns rows offsets[k]:offsets[k+1]).
    A track with a single frame is copied (the SBERT branch, collate_functions.py:163-166), a track without
    frames gives zeros (the missing-embedding branch, :158-161)."""
    if offsets is None:
        lens = torch.tensor([0] + [int(f.shape[0]) for f in frames], dtype=torch.long)
        offsets = torch.cumsum(lens, 0)
        frames = torch.cat([f.reshape(f.shape[0], f.shape[-1]) for f in frames], 0) if len(frames) else torch.empty(0, 1)
        frames = frames.cuda() if not frames.is_cuda else frames
    N.require_cuda(frames)
    assert frames.dim() == 2
    x = frames.contiguous()
    off = offsets.to(device=x.device, dtype=torch.long).contiguous()
    k = off.numel() - 1
    assert k >= 0
    out = torch.empty(k, x.shape[1], dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib.wealy_segment_mean(x.data_ptr(), off.data_ptr(), k, x.shape[1], N.dtype_code(x.dtype),
                                         out.data_ptr(), N.stream_ptr(x.device)))
    return out
