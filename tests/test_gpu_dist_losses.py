"""-m gpu: data-parallel (global-batch) NT-Xent / CLEWS -- SURVEY.md 8(f) row f2.

* the ranks of a sharded loss emulated one after the other on ONE GPU (the collectives done by hand on the exposed
  buffers): loss, logdict and the concatenated per-rank gradients must equal the single-GPU module on the global batch
  and the reference's own outputs (tests/golden/losses.npz);
* the real thing over NCCL on two GPUs when the box has them (skipped otherwise)."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as ol

pytestmark = pytest.mark.gpu


def _cfgs():
    from wealy_b200 import _native as N
    from wealy_b200.dist_losses import _cfg
    return {
        "ntx": _cfg(kind=N.LOSS_NTXENT, passes=3, temperature=0.1),
        "clews": _cfg(kind=N.LOSS_CLEWS, passes=3, gamma=8.0, b=1.0, eps=1e-8, epsilon=1e-6, uw=0.5, numerically_friendly=1),
    }


def _emulated(cfg_items, z, lab, idx, bounds, phased=False, packed=False):
    """Run the shards [bounds[r], bounds[r+1]) as if they were ranks; collectives by hand.  phased: the forward in the
    two phases of the overlapped path -- phase 1 sees a batch in which ONLY the shard's own rows are valid."""
    from wealy_b200.dist_losses import ShardState
    states = []
    for r in range(len(bounds) - 1):
        if phased:
            zz = torch.full_like(z, float("nan"))                   # rows of the other ranks: not arrived yet
            zz[bounds[r]:bounds[r + 1]] = z[bounds[r]:bounds[r + 1]]
            st = ShardState(cfg_items, zz, lab, idx, bounds[r], bounds[r + 1] - bounds[r])
            st.forward_phase(1)
            torch.cuda.synchronize()
            zz.copy_(z)                                             # "all-gather complete"
            st.forward_phase(2)
        else:
            st = ShardState(cfg_items, z, lab, idx, bounds[r], bounds[r + 1] - bounds[r])
            st.forward_local()
        states.append(st)
    sizes = {bounds[r + 1] - bounds[r] for r in range(len(bounds) - 1)}
    if packed and len(sizes) == 1:
        # exchange 2 as the product does it over NCCL: one record per rank, "all-gathered" by hand, unpacked everywhere
        import ctypes
        from wealy_b200 import _native as N
        nb, world = sizes.pop(), len(states)
        rec = N.lib.wealy_loss_dp_record_bytes(nb)
        recs = torch.zeros((world, rec), dtype=torch.uint8, device=z.device)
        for r, st in enumerate(states):
            N.check(N.lib.wealy_loss_dp_pack(ctypes.byref(st.cfg), st.ws.data_ptr(), st.ws_bytes, st.bg, st.d, st.row0, st.nb,
                                             recs[r].data_ptr(), N.stream_ptr(z.device)))
        outs, grads = [], []
        for st in states:
            N.check(N.lib.wealy_loss_dp_unpack(ctypes.byref(st.cfg), st.ws.data_ptr(), st.ws_bytes, st.bg, st.d, st.nb,
                                               recs.data_ptr(), world, N.stream_ptr(z.device)))
            outs.append(st.forward_finish())
            grads.append(st.backward(torch.ones((), device=z.device)))
        torch.cuda.synchronize()
        return outs, torch.cat(grads)
    bufs = [st.buffers() for st in states]
    acc = sum(b[0].clone() for b in bufs)                                   # all-reduce SUM
    accm = torch.stack([b[1] for b in bufs]).max(dim=0).values              # all-reduce MAX
    rows = torch.cat([b[2][bounds[r]:bounds[r + 1]] for r, b in enumerate(bufs)])   # all-gather
    outs, grads = [], []
    for st, b in zip(states, bufs):
        b[0].copy_(acc); b[1].copy_(accm); b[2].copy_(rows)
        outs.append(st.forward_finish())
        grads.append(st.backward(torch.ones((), device=z.device)))
    torch.cuda.synchronize()
    return outs, torch.cat(grads)


@pytest.mark.parametrize("bounds", [[0, 256, 512], [0, 128, 256, 384, 512], [0, 100, 333, 512]])
@pytest.mark.parametrize("kind", ["ntx", "clews"])
def test_emulated_ranks_equal_single_gpu(kind, bounds):
    from wealy_b200 import losses as wl
    from wealy_b200.data import synth
    s = synth.make_loss_batch(512, 192, seed=3, device="cuda")
    z, lab, idx = s["z"], s["label"], s["idx"]
    mod = wl.NTXentLoss(0.1) if kind == "ntx" else wl.CLEWSLoss()
    zz = z.clone().requires_grad_(True)
    loss, logd = mod(lab.clone(), idx, zz)
    loss.backward()
    outs, grad = _emulated(_cfgs()[kind], z, lab, idx, bounds)
    for o in outs:
        assert abs(float(o[0]) - float(loss)) <= 1e-6 * max(1.0, abs(float(loss)))
        assert torch.allclose(o, outs[0])                                   # identical on every "rank"
    assert float((grad - zz.grad).norm()) <= 2e-6 * float(zz.grad.norm())
    # and against the oracle (autograd of the restated reference loss) on the global batch
    zr = z.cpu().clone().requires_grad_(True)
    lo, _ = (ol.ntxent(lab.cpu().clone(), idx.cpu(), zr) if kind == "ntx" else ol.clews(lab.cpu().clone(), idx.cpu(), zr))
    lo.backward()
    assert abs(float(outs[0][0]) - float(lo)) <= 1e-3 * abs(float(lo))
    assert float((grad.cpu() - zr.grad).norm()) <= 1e-5 * float(zr.grad.norm())


@pytest.mark.parametrize("bounds", [[0, 256, 512], [0, 256, 512, 768, 1000], [0, 100, 333, 512]])
@pytest.mark.parametrize("kind", ["ntx", "clews"])
def test_two_phase_forward_equals_one_shot(kind, bounds):
    """The overlapped data-parallel forward (own column block while the all-gather is in flight, the rest behind it)
    gives the loss / logdict / gradients of the one-shot forward; unaligned shards fall back to doing everything in
    phase 2.  Phase 1 is run on a batch whose foreign rows are NaN: it must not read them."""
    from wealy_b200.data import synth
    b = bounds[-1]
    s = synth.make_loss_batch(b, 160, seed=9, device="cuda")
    z, lab, idx = s["z"], s["label"], s["idx"]
    o1, g1 = _emulated(_cfgs()[kind], z, lab, idx, bounds)
    o2, g2 = _emulated(_cfgs()[kind], z, lab, idx, bounds, phased=True)
    o3, g3 = _emulated(_cfgs()[kind], z, lab, idx, bounds, phased=True, packed=True)   # + exchange 2 as one packed record per rank
    for a, c, e in zip(o1, o2, o3):
        assert torch.isfinite(c[:11]).all() and torch.allclose(a, c, rtol=1e-6, atol=1e-9)
        assert torch.allclose(a, e, rtol=1e-6, atol=1e-9)
    assert torch.isfinite(g2).all() and float((g1 - g2).norm()) <= 1e-6 * float(g1.norm())
    assert float((g1 - g3).norm()) <= 1e-6 * float(g1.norm())


def test_emulated_ranks_against_reference_outputs(golden):
    G = golden("losses.npz")
    name = "f32_big"
    z = torch.from_numpy(G[f"{name}_z"]).float().cuda()
    lab, idx = torch.from_numpy(G[f"{name}_label"]).cuda(), torch.from_numpy(G[f"{name}_idx"]).cuda()
    rows = torch.from_numpy(G[f"{name}_gradrows"])
    b = z.shape[0]
    bounds = [0, b // 4, b // 2, b]
    for tag, kind in (("ntx", "ntx"), ("clews", "clews")):
        outs, grad = _emulated(_cfgs()[kind], z, lab.clone(), idx, bounds)
        ref_l = float(G[f"{name}_{tag}_loss"])
        assert abs(float(outs[0][0]) - ref_l) <= 1e-3 * abs(ref_l)
        ref_g = torch.from_numpy(G[f"{name}_{tag}_grad"]).double()
        assert float((grad.cpu().double()[rows] - ref_g).norm()) <= 1e-5 * float(ref_g.norm())


def _nccl_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from wealy_b200 import losses as wl
        from wealy_b200.dist_losses import DistributedNTXentLoss, DistributedCLEWSLoss
        from wealy_b200.data import synth
        s = synth.make_loss_batch(1024, 256, seed=5, device="cuda")
        nb = 1024 // world
        sl = slice(rank * nb, (rank + 1) * nb)
        ok = True
        for single, multi in ((wl.NTXentLoss(0.1), DistributedNTXentLoss(0.1)), (wl.CLEWSLoss(), DistributedCLEWSLoss())):
            zz = s["z"].clone().requires_grad_(True)
            l1, d1 = single(s["label"].clone(), s["idx"], zz)
            l1.backward()
            zl = s["z"][sl].clone().requires_grad_(True)
            l2, d2 = multi(s["label"][sl].clone(), s["idx"][sl].clone(), zl)
            l2.backward()
            torch.cuda.synchronize()
            ok &= abs(float(l1) - float(l2)) <= 1e-6 * max(1.0, abs(float(l1)))
            ok &= float((zl.grad - zz.grad[sl]).norm()) <= 2e-6 * float(zz.grad[sl].norm())
            ok &= set(d1) == set(d2) and all(abs(float(d1[k]) - float(d2[k])) <= 1e-5 * max(1.0, abs(float(d1[k]))) for k in d1)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_nccl_two_ranks():
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 200)
    ret = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
