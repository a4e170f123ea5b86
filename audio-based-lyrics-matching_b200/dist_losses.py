"""Data-parallel (global-batch) NT-Xent / CLEWS over the GPUs of one box -- SURVEY.md 8(f) row f2.

The reference trains single-process (lib/losses.py sees one batch).  With one process per GPU every rank holds
B/G anchors; the contrastive losses want ALL B columns (negatives from every rank).  Scheme, one process per GPU:

  1. all-gather z, labels, ids over NCCL (NVLink / NVSwitch)               -> the global batch on every rank
     -- the all-gather of z is asynchronous: while it is in flight the rank already normalises its OWN rows and sweeps
     its anchors against its OWN column block (wealy_loss_dp_forward_phase 1: needs nothing from the other ranks)
  2. this rank's anchors [row0, row0 + nb) are swept against the other B - nb columns (phase 2, behind the all-gather;
     the two launches write disjoint partial records that one merge kernel folds)
  3. all-reduce the batch sums (16 doubles + 2 maxima), all-gather the per-anchor records (4 floats per anchor)
  4. loss + logdict of the GLOBAL batch, identical on every rank             (wealy_loss_dp_forward_finish)
  5. backward: dz of this rank's rows only, and complete -- the symmetrised dL/dS contains the terms in which these
     rows are columns of other ranks' anchors, so there is NO reduce-scatter    (wealy_loss_dp_backward)

`loss` is the global-batch loss of lib/losses.py:19-73 / 210-285 evaluated on the concatenated batch (rank order);
`z.grad` is d(global loss)/dz for the local rows.  Same logdict keys as the single-GPU modules.
"""
import ctypes
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _native as N
from .losses import _label_noise_, _passes


def _dev_view(ptr, shape, typestr, device):
    class _Dev:
        __cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}
    return torch.as_tensor(_Dev(), device=device)


class ShardState:
    """Workspace + geometry of one rank's share of a global-batch loss (thin handle over the C ABI)."""

    def __init__(self, cfg_items, zg, labg, idxg, row0, nb):
        self.cfg = N.LossCfg(**dict(cfg_items))
        N.require_cuda(zg, labg, idxg)
        self.zg = zg if zg.stride(1) == 1 else zg.contiguous()
        self.labg = labg.to(torch.long).contiguous()
        self.idxg = idxg.to(torch.long).contiguous()
        self.bg, self.d = self.zg.shape
        self.row0, self.nb = int(row0), int(nb)
        self.code = N.dtype_code(self.zg.dtype)
        dev = self.zg.device
        with torch.cuda.device(dev):
            self.ws_bytes = N.lib.wealy_loss_dp_workspace_bytes(self.bg, self.d, self.cfg.passes, self.nb)
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)

    def forward_local(self):
        with torch.cuda.device(self.zg.device):
            N.check(N.lib.wealy_loss_dp_forward_local(
                ctypes.byref(self.cfg), self.zg.data_ptr(), self.bg, self.zg.stride(0), self.d, self.code,
                self.labg.data_ptr(), self.idxg.data_ptr(), self.row0, self.nb, self.ws.data_ptr(), self.ws_bytes,
                N.stream_ptr(self.zg.device)))

    def forward_phase(self, phase):
        """phase 1: own rows / own column block only (run it while the all-gather of the other rows is in flight);
        phase 2: everything else + merge (zg must be complete)."""
        with torch.cuda.device(self.zg.device):
            N.check(N.lib.wealy_loss_dp_forward_phase(
                ctypes.byref(self.cfg), self.zg.data_ptr(), self.bg, self.zg.stride(0), self.d, self.code,
                self.labg.data_ptr(), self.idxg.data_ptr(), self.row0, self.nb, int(phase), self.ws.data_ptr(), self.ws_bytes,
                N.stream_ptr(self.zg.device)))

    def buffers(self):
        """-> (acc float64[16] to SUM, acc_max int32[2] to MAX (bit patterns of non-negative floats),
        rowstat float32[B, 4] whose rows [row0, row0 + nb) this rank has written) -- views, no copies."""
        acc, accm, rs, cnt = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        N.check(N.lib.wealy_loss_dp_buffers(ctypes.byref(self.cfg), self.ws.data_ptr(), self.ws_bytes, self.bg, self.d,
                                            self.nb, ctypes.byref(acc), ctypes.byref(cnt), ctypes.byref(accm),
                                            ctypes.byref(rs)))
        dev = self.zg.device
        return (_dev_view(acc.value, (cnt.value,), "<f8", dev), _dev_view(accm.value, (2,), "<i4", dev),
                _dev_view(rs.value, (self.bg, 4), "<f4", dev))

    def exchange(self, group, world):
        """Exchange 2 in one collective: pack this rank's {batch sums, maxima, anchor records}, all-gather the records,
        fold all of them back into the workspace (equal shards)."""
        dev = self.zg.device
        rec_bytes = N.lib.wealy_loss_dp_record_bytes(self.nb)
        recs = torch.empty((world, rec_bytes), dtype=torch.uint8, device=dev)
        mine = recs[self.row0 // self.nb]
        with torch.cuda.device(dev):
            N.check(N.lib.wealy_loss_dp_pack(ctypes.byref(self.cfg), self.ws.data_ptr(), self.ws_bytes, self.bg, self.d,
                                             self.row0, self.nb, mine.data_ptr(), N.stream_ptr(dev)))
        dist.all_gather_into_tensor(recs, mine, group=group)             # in place: every rank's record in rank order
        with torch.cuda.device(dev):
            N.check(N.lib.wealy_loss_dp_unpack(ctypes.byref(self.cfg), self.ws.data_ptr(), self.ws_bytes, self.bg, self.d,
                                               self.nb, recs.data_ptr(), world, N.stream_ptr(dev)))
        recs.record_stream(torch.cuda.current_stream(dev))

    def forward_finish(self):
        out = torch.empty(N.OUT_COUNT, dtype=torch.float64, device=self.zg.device)
        with torch.cuda.device(self.zg.device):
            N.check(N.lib.wealy_loss_dp_forward_finish(ctypes.byref(self.cfg), self.bg, self.d, self.nb, out.data_ptr(), None,
                                                       0, self.ws.data_ptr(), self.ws_bytes, N.stream_ptr(self.zg.device)))
        return out

    def backward(self, grad_loss):
        dz = torch.empty((self.nb, self.d), dtype=self.zg.dtype, device=self.zg.device)
        g = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        with torch.cuda.device(self.zg.device):
            N.check(N.lib.wealy_loss_dp_backward(
                ctypes.byref(self.cfg), self.zg.data_ptr(), self.bg, self.zg.stride(0), self.d, self.code, self.row0,
                self.nb, g.data_ptr(), dz.data_ptr(), dz.stride(0), self.ws.data_ptr(), self.ws_bytes,
                N.stream_ptr(self.zg.device)))
        return dz


def _world(group):
    on = dist.is_available() and dist.is_initialized()
    return (dist.get_world_size(group), dist.get_rank(group)) if on else (1, 0)


def _all_gather_rows(t, group):
    world, _ = _world(group)
    if world == 1:
        return t
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    try:
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):      # backends without the flat variant (CPU tests of the host logic)
        dist.all_gather(list(out.view((world, t.shape[0]) + tuple(t.shape[1:])).unbind(0)), t.contiguous(), group=group)
    return out


class _DistLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, labg, idxg, cfg_items, group):
        world, rank = _world(group)
        nb = z.shape[0]
        # Overlapping the all-gather of z with the rank's own column block pays once the gathered batch is large; at
        # 8 x 4096 x 1024 bf16 (64 MB) the exchange is latency-bound (~0.09 ms of transfer) and splitting the sweep in two
        # launches costs as much as it hides (measured on 8 x B200: 1.53 vs 1.49 ms NT-Xent, 1.58 vs 1.46 ms CLEWS), so
        # the default switches on at 128 MB; WEALY_DP_OVERLAP=1 / 0 forces it.
        mode = os.environ.get("WEALY_DP_OVERLAP", "auto")
        big = world * nb * z.shape[1] * z.element_size() >= (128 << 20)
        overlap = world > 1 and dist.get_backend(group) == "nccl" and (mode == "1" or (mode == "auto" and big))
        if overlap:
            # exchange 1, asynchronous and in place: the rank's rows sit in their slot of the global batch, NCCL fills the
            # other slots on its own stream while this stream already works on the local block
            zd = z.detach()
            zg = torch.empty((world * nb,) + tuple(zd.shape[1:]), dtype=zd.dtype, device=zd.device)
            mine = zg[rank * nb:(rank + 1) * nb]
            mine.copy_(zd)
            work = dist.all_gather_into_tensor(zg, mine, group=group, async_op=True)
            st = ShardState(cfg_items, zg, labg, idxg, rank * nb, nb)
            st.forward_phase(1)
            work.wait()                                              # the compute stream waits for the gathered rows
            st.forward_phase(2)
        else:
            zg = _all_gather_rows(z.detach(), group)                 # exchange 1: the global batch
            st = ShardState(cfg_items, zg, labg, idxg, rank * nb, nb)
            st.forward_local()
        if world > 1 and dist.get_backend(group) == "nccl":
            st.exchange(group, world)                                # exchange 2: batch sums + anchor records, one all-gather
        elif world > 1:
            acc, accm, rs = st.buffers()
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(accm, op=dist.ReduceOp.MAX, group=group)
            dist.all_gather_into_tensor(rs, rs[rank * nb:(rank + 1) * nb].clone(), group=group)
        out = st.forward_finish()
        ctx.st = st
        ctx.mark_non_differentiable(out)
        return out[0].to(z.dtype), out              # like the single-GPU modules: the loss carries z's dtype

    @staticmethod
    def backward(ctx, grad_loss, _g):
        return ctx.st.backward(grad_loss), None, None, None, None


def _gather_ids(z_label, z_idx, group):
    """Global labels / ids (rank order) with the reference's single-label noise applied to the GLOBAL batch
    (lib/losses.py:34-35); the local slice of the caller's z_label is updated in place like upstream."""
    world, rank = _world(group)
    if world > 1:
        both = _all_gather_rows(torch.stack((z_label.to(torch.long), z_idx.to(torch.long))), group)   # ONE collective
        both = both.view(world, 2, -1)
        labg, idxg = both[:, 0].reshape(-1), both[:, 1].reshape(-1)
    else:
        labg, idxg = z_label.to(torch.long), z_idx.to(torch.long)
    if labg.data_ptr() == z_label.data_ptr():
        _label_noise_(z_label)
        return z_label, idxg
    _label_noise_(labg)
    nb = z_label.shape[0]
    z_label.copy_(labg[rank * nb:(rank + 1) * nb])
    return labg, idxg


def _cfg(**kw):
    base = dict(kind=0, passes=3, temperature=1.0, gamma=0.0, b=0.0, eps=1e-8, epsilon=1e-6, uw=0.0,
                numerically_friendly=1)
    base.update(kw)
    return tuple(sorted(base.items()))


class DistributedNTXentLoss(nn.Module):
    """NTXentLoss (lib/losses.py:10-73) over the global batch of a process group; every rank passes its own
    equally sized (z_label, z_idx, z) shard."""

    def __init__(self, temperature=0.1, precision=None, group=None):
        super().__init__()
        self.tau, self.precision, self.group = temperature, precision, group

    def forward(self, z_label, z_idx, z, extra=None):
        assert len(z_label) == len(z_idx) and len(z_label) == len(z)
        N.require_cuda(z, z_label, z_idx)
        labg, idxg = _gather_ids(z_label, z_idx, self.group)
        loss, st = _DistLoss.apply(z, labg, idxg, _cfg(kind=N.LOSS_NTXENT, passes=_passes(self.precision, z),
                                                       temperature=float(self.tau)), self.group)
        stats = st.to(loss.dtype)
        return loss, {"l_main": loss, "v_zmax": stats[1], "v_zmean": stats[2], "v_zstd": stats[3]}


class DistributedCLEWSLoss(nn.Module):
    """CLEWSLoss (lib/losses.py:176-285) over the global batch of a process group."""

    def __init__(self, gamma=8.0, b=1.0, eps=1e-8, epsilon=1e-6, uniformity_weight=0.5, warmup_steps=1000,
                 precision=None, group=None):
        super().__init__()
        self.gamma, self.b, self.eps, self.epsilon = float(gamma), float(b), float(eps), float(epsilon)
        self.uniformity_weight, self.warmup_steps = float(uniformity_weight), int(warmup_steps)
        self.precision, self.group = precision, group

    def forward(self, z_label, z_idx, z, extra=None, numerically_friendly=True):
        if z.dim() == 3:
            assert z.size(1) == 1, f"CLEWS (vector) expects S=1, got S={z.size(1)}"
            z = z.squeeze(1)
        assert z.dim() == 2
        world, _ = _world(self.group)
        assert len(z_label) == len(z_idx) == z.size(0) and z.size(0) * world >= 4
        N.require_cuda(z, z_label, z_idx)
        labg, idxg = _gather_ids(z_label, z_idx, self.group)
        uw = self.uniformity_weight
        if self.warmup_steps > 0:                                    # lib/losses.py:248-258
            step = None
            if isinstance(extra, dict) and "global_step" in extra:
                step = int(extra["global_step"])
            elif hasattr(self, "global_step"):
                step = int(self.global_step)
            if step is not None:
                uw = float(min(self.uniformity_weight, self.uniformity_weight * (step + 1) / self.warmup_steps))
        loss, st = _DistLoss.apply(z, labg, idxg, _cfg(
            kind=N.LOSS_CLEWS, passes=_passes(self.precision, z), gamma=self.gamma, b=self.b, eps=self.eps,
            epsilon=self.epsilon, uw=uw, numerically_friendly=1 if numerically_friendly else 0), self.group)
        stats = st.to(loss.dtype)
        return loss, {
            "l_main": loss, "l_cent": stats[4], "l_cont": stats[5], "cnt_pos_pairs": stats[6],
            "cnt_neg_pairs": stats[7], "anchors_with_pos": stats[8], "v_dpos": stats[9], "v_dneg": stats[10],
            "uniformity_weight": torch.full((), uw, device=z.device), "z_max": stats[1], "z_mean": stats[2],
            "z_std": stats[3],
        }
