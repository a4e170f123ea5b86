"""-m gpu: evaluation of chunked tracks: the s x s chunk distances of every track pair are
reduced like distance_tensor_redux (lib/tensor_ops.py:288-373) inside the sweep's epilogue.  Checked against the
oracle evaluator fed with the reduced distance matrix of the restated (reference-pinned) distance_tensor_redux."""
import pytest
import torch

from oracle import evaluator as oev

pytestmark = pytest.mark.gpu


def _chunked_set(n, s, d, seed):
    from wealy_b200.data import synth
    base = synth.make_eval_set(n, d, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    # chunk embeddings = the track embedding + chunk-level noise (some chunks of a cover match better than others)
    z = base["z"][:, None, :] + 0.8 * base["z"].norm(dim=1).mean() / d ** 0.5 * torch.randn(n, s, d, generator=g)
    return base["c"], base["i"], z.contiguous()


@pytest.mark.parametrize("redux", ["min", "max", "mean", "meanmin", "minmean"])
@pytest.mark.parametrize("n,s,d", [(900, 4, 96), (333, 8, 64), (200, 16, 48), (1500, 2, 128)])
def test_chunked_parity_with_oracle(n, s, d, redux):
    from wealy_b200 import evaluation as we
    c, i, z = _chunked_set(n, s, d, seed=n + s)
    aps_o, r1_o = oev.evaluate_argsort(c, i, z, c, i, z, redux=redux)
    cq, iq, zq = c.cuda(), i.cuda(), z.cuda()
    aps, r1s = we.evaluate(cq, iq, zq, cq, iq, zq, redux=redux)
    torch.cuda.synchronize()
    aps, r1s = aps.cpu().double(), r1s.cpu().double()
    assert abs(float(aps.mean()) - float(aps_o.mean())) <= 1e-4            # MAP within 1e-4
    assert abs(float(r1s.mean()) - float(r1_o.mean())) <= 1e-4 * max(1.0, float(r1_o.mean()))
    lo, hi = oev.rank_tolerance(c, i, z, c, i, z, gap=1e-5, redux=redux)
    assert bool(((r1s >= lo) & (r1s <= hi)).all())
    exact = lo == hi
    assert torch.equal(r1s[exact], r1_o[exact])


def test_chunked_topk_and_flat_layout():
    from wealy_b200 import evaluation as we
    n, s, d, k = 700, 4, 64, 10
    c, i, z = _chunked_set(n, s, d, seed=7)
    _, _, idx_o, sim_o = oev.evaluate_argsort(c, i, z, c, i, z, topk=k, redux="meanmin")
    cq, iq = c.cuda(), i.cuda()
    zf = z.reshape(n * s, d).cuda()                                         # flat [N * s, D] + chunks=
    aps, r1s, idx, sim = we.evaluate(cq, iq, zf, cq, iq, zf, topk=k, chunks=s, redux="meanmin")
    torch.cuda.synchronize()
    idx, sim = idx.cpu(), sim.cpu()
    assert (sim - sim_o).abs().max() <= 4e-6
    ok = torch.ones_like(idx_o, dtype=torch.bool)
    ok[:, 1:] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    ok[:, :-1] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    ok[:, -1] = False   # the last position also depends on the (k+1)-th best, which the lists do not show
    assert torch.equal(idx[ok], idx_o[ok])
    assert bool((idx != torch.arange(n)[:, None]).all())                    # a track never retrieves itself


def test_chunked_queries_disjoint_from_corpus_and_errors():
    from wealy_b200 import evaluation as we
    c, i, z = _chunked_set(600, 4, 48, seed=9)
    q, cand = slice(0, 100), slice(100, 600)
    keep = torch.tensor([bool((c[cand] == c[k]).any()) for k in range(100)])
    qc, qi, qz = c[q][keep], i[q][keep], z[q][keep]
    aps_o, r1_o = oev.evaluate_argsort(qc, qi, qz, c[cand], i[cand], z[cand], redux="min")
    aps, r1s = we.evaluate(qc.cuda(), qi.cuda(), qz.cuda(), c[cand].cuda(), i[cand].cuda(), z[cand].cuda(), redux="min")
    assert abs(float(aps.double().mean().cpu()) - float(aps_o.mean())) <= 1e-4
    assert int((r1s.cpu().double() != r1_o).sum()) <= 1
    with pytest.raises(NotImplementedError):
        we.evaluate(c.cuda(), i.cuda(), z.cuda(), c.cuda(), i.cuda(), z.cuda(), redux="bpwr")
    with pytest.raises(NotImplementedError):
        z3 = z[:, :3].contiguous().cuda()
        we.evaluate(c.cuda(), i.cuda(), z3, c.cuda(), i.cuda(), z3)


def _ragged(n, s, d, seed):
    c, i, z = _chunked_set(n, s, d, seed)
    g = torch.Generator().manual_seed(seed + 500)
    lens = torch.randint(1, s + 1, (n,), generator=g)
    for t in range(n):                                                      # padding chunks hold junk, not zeros
        z[t, lens[t]:] = 23.0 * torch.randn(s - int(lens[t]), d, generator=g)
    return c, i, z.contiguous(), lens


@pytest.mark.parametrize("redux", ["min", "max", "mean", "meanmin", "minmean"])
@pytest.mark.parametrize("n,s,d", [(700, 4, 96), (333, 8, 64), (200, 16, 48), (900, 2, 128)])
def test_ragged_tracks_parity_with_oracle(n, s, d, redux):
    """Tracks with 1 .. s valid chunks: padding is excluded like distance_tensor_redux's mask (lib/tensor_ops.py:288)."""
    from wealy_b200 import evaluation as we
    c, i, z, lens = _ragged(n, s, d, seed=n + s)
    aps_o, r1_o = oev.evaluate_argsort(c, i, z, c, i, z, redux=redux, q_len=lens, c_len=lens)
    cq, iq, zq = c.cuda(), i.cuda(), z.cuda()
    aps, r1s = we.evaluate(cq, iq, zq, cq, iq, zq, redux=redux, q_chunks=lens, c_chunks=lens)
    torch.cuda.synchronize()
    aps, r1s = aps.cpu().double(), r1s.cpu().double()
    assert abs(float(aps.mean()) - float(aps_o.mean())) <= 1e-4
    assert abs(float(r1s.mean()) - float(r1_o.mean())) <= 1e-4 * max(1.0, float(r1_o.mean()))
    lo, hi = oev.rank_tolerance(c, i, z, c, i, z, gap=1e-5, redux=redux, q_len=lens, c_len=lens)
    assert bool(((r1s >= lo) & (r1s <= hi)).all())
    exact = lo == hi
    assert torch.equal(r1s[exact], r1_o[exact])


def test_ragged_tracks_topk_disjoint_queries_and_full_counts():
    from wealy_b200 import evaluation as we
    n, s, d, k = 600, 4, 64, 8
    c, i, z, lens = _ragged(n, s, d, seed=11)
    q, cand = slice(0, 120), slice(120, 600)
    keep = torch.tensor([bool((c[cand] == c[t]).any()) for t in range(120)])
    qc, qi, qz, ql = c[q][keep], i[q][keep], z[q][keep].contiguous(), lens[q][keep]
    _, _, idx_o, sim_o = oev.evaluate_argsort(qc, qi, qz, c[cand], i[cand], z[cand], topk=k, redux="meanmin",
                                              q_len=ql, c_len=lens[cand])
    _, _, idx, sim = we.evaluate(qc.cuda(), qi.cuda(), qz.cuda(), c[cand].cuda(), i[cand].cuda(), z[cand].cuda(), topk=k,
                                 redux="meanmin", q_chunks=ql, c_chunks=lens[cand])
    idx, sim = idx.cpu(), sim.cpu()
    assert (sim - sim_o).abs().max() <= 4e-6
    ok = torch.ones_like(idx_o, dtype=torch.bool)
    ok[:, 1:] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    ok[:, :-1] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    ok[:, -1] = False   # the last position also depends on the (k+1)-th best, which the lists do not show
    assert torch.equal(idx[ok], idx_o[ok])
    # all chunks valid == the dense chunked path (same ranks up to rounding of the means)
    full = torch.full((n,), s)
    cq, iq, zq = c.cuda(), i.cuda(), z.cuda()
    a1, r1 = we.evaluate(cq, iq, zq, cq, iq, zq, redux="min", q_chunks=full, c_chunks=full)
    a0, r0 = we.evaluate(cq, iq, zq, cq, iq, zq, redux="min")
    assert torch.equal(a1, a0) and torch.equal(r1, r0)
    with pytest.raises(ValueError):
        we.evaluate(cq, iq, zq[:, 0].contiguous(), cq, iq, zq[:, 0].contiguous(), q_chunks=full, c_chunks=full)


@pytest.mark.parametrize("redux", ["min", "max", "mean", "minmean", "meanmin"])
def test_ragged_tracks_topk_similarities_against_reference_outputs(redux):
    """Golden vectors produced by the unmodified reference (distance_tensor_redux with the ragged mask): the fused
    sweep's top-k similarities over ALL candidates are 1 - those distances."""
    import os
    import numpy as np
    from wealy_b200 import evaluation as we
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "redux_ragged.npz"))
    zq, zc = torch.from_numpy(G["zq"]).cuda(), torch.from_numpy(G["zc"]).cuda()
    lq, lc = torch.from_numpy(G["lq"]), torch.from_numpy(G["lc"])
    n, m = zq.shape[0], zc.shape[0]
    qc, qi = torch.zeros(n, dtype=torch.long).cuda(), (torch.arange(n) + 1000).cuda()     # one clique, no self matches
    cc, ci = torch.zeros(m, dtype=torch.long).cuda(), torch.arange(m).cuda()
    _, _, idx, sim = we.evaluate(qc, qi, zq, cc, ci, zc, topk=m, redux=redux, q_chunks=lq, c_chunks=lc)
    want = 1 - torch.from_numpy(G[f"r_{redux}"])                                          # [n, m]
    got = torch.empty_like(want)
    got.scatter_(1, idx.cpu(), sim.cpu())
    assert (got - want).abs().max() <= 4e-6


def _all_item_ranks_chunked(plan, c, i, z, redux, lens=None, gap=1e-5, sim_tol=4e-6):
    """Every relevant item's rank (plan.ranks() of the run that just finished) inside the band its `gap` neighbours allow,
    exact where the band is a single rank -- the contract of tests/test_gpu_eval.py::_check_all_item_ranks on tracks."""
    off_g, ranks_g, sims_g = (t.cpu() for t in plan.ranks())
    off_o, sims_o, exact, lo, hi = oev.rank_bands(c, i, z, c, i, z, gap=gap, redux=redux, q_len=lens, c_len=lens)
    assert torch.equal(off_g, off_o)
    r, sg = ranks_g.long(), sims_g.double()
    assert float((sg - sims_o).abs().max()) <= sim_tol
    assert bool(((r >= lo) & (r <= hi)).all()), "a relevant item's rank left the band its 1e-5 neighbours allow"
    single = lo == hi
    assert torch.equal(r[single], exact[single])
    return int(single.sum()), int(single.numel())


@pytest.mark.parametrize("redux", ["min", "mean", "max"])
@pytest.mark.parametrize("n,s,d", [(2500, 4, 64), (1100, 8, 96), (3001, 2, 64), (130, 16, 32)])
def test_chunked_all_vs_all_symmetric_sweep_all_item_ranks(n, s, d, redux):
    """Chunked all-vs-all with a reduction that is the same in both directions runs the half sweep (tiles above the
    diagonal on the CTA-pair core, every track pair scored for its row AND its column query): the rank of EVERY relevant
    item against the oracle, at sizes that span several super row blocks / column chunks and end in ragged tiles."""
    from wealy_b200 import evaluation as we
    c, i, z = _chunked_set(n, s, d, seed=3 * n + s)
    cq, iq, zq = c.cuda(), i.cuda(), z.cuda()
    plan = we.EvalPlan(cq, iq, cq, iq)
    res = plan.run(zq, zq, redux=redux)
    torch.cuda.synchronize()
    n_exact, n_all = _all_item_ranks_chunked(plan, c, i, z, redux)
    assert n_exact > 0.5 * n_all
    aps_o, r1_o = oev.evaluate_argsort(c, i, z, c, i, z, redux=redux)
    assert abs(float(res["aps"].double().mean().cpu()) - float(aps_o.mean())) <= 1e-4
    plan.close()


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("redux", ["min", "max", "mean"])
def test_chunked_symmetric_sweep_equals_rectangle_sweep(redux, ragged, monkeypatch):
    """The half sweep against the full rectangle (WEALY_SYM_TRACKS=0) on the same data: min / max reduce the same 64
    numbers either way (identical results); the mean's summation order differs between the directions (ranks may move
    inside their 1e-5 band: MAP within 1e-6)."""
    from wealy_b200 import evaluation as we
    n, s, d = 1800, 8, 64
    c, i, z, lens = _ragged(n, s, d, seed=21) if ragged else (*_chunked_set(n, s, d, seed=21), None)
    cq, iq, zq = c.cuda(), i.cuda(), z.cuda()
    kw = dict(q_chunks=lens, c_chunks=lens) if ragged else {}
    a1, r1 = we.evaluate(cq, iq, zq, cq, iq, zq, redux=redux, **kw)
    monkeypatch.setenv("WEALY_SYM_TRACKS", "0")
    a0, r0 = we.evaluate(cq, iq, zq, cq, iq, zq, redux=redux, **kw)
    if redux != "mean":
        assert torch.equal(a1, a0) and torch.equal(r1, r0)
    else:
        assert abs(float(a1.double().mean()) - float(a0.double().mean())) <= 1e-6
        assert float((r1 != r0).float().mean()) <= 0.01


def test_chunked_symmetric_sweep_id_collisions_and_singletons():
    """Half sweep, both directions of a pair: a version-id collision between two tracks of DIFFERENT cliques removes the
    pair for both of them (lib/losses.py:40-42: self is decided by the version id), and a track without any relevant
    candidate (singleton clique, allow_empty) is still a candidate of every other query."""
    from wealy_b200 import evaluation as we
    n, s, d = 700, 4, 48
    c, i, z = _chunked_set(n, s, d, seed=31)
    c, i = c.clone(), i.clone()
    a, b = 5, 400
    assert c[a] != c[b]
    i[b] = i[a]                                                    # collision across cliques, far apart in the sweep
    c[650] = int(c.max()) + 1                                      # singleton clique
    cq, iq, zq = c.cuda(), i.cuda(), z.cuda()
    plan = we.EvalPlan(cq, iq, cq, iq)
    with pytest.raises(ValueError):
        plan.run(zq, zq, redux="min")
    res = plan.run(zq, zq, redux="min", allow_empty=True)
    torch.cuda.synchronize()
    keep = ((c[:, None] == c[None, :]) & (i[:, None] != i[None, :])).any(dim=1)     # queries with a relevant candidate
    assert not bool(keep[650])
    off_g, ranks_g, sims_g = (t.cpu() for t in plan.ranks())
    off_o, sims_o, exact, lo, hi = oev.rank_bands(c[keep], i[keep], z[keep], c, i, z, gap=1e-5, redux="min")
    sel = torch.cat([torch.arange(int(off_g[q]), int(off_g[q + 1])) for q in torch.nonzero(keep).view(-1).tolist()])
    r = ranks_g[sel].long()
    assert r.numel() == exact.numel() == ranks_g.numel() and int(off_g[651] - off_g[650]) == 0
    assert bool(((r >= lo) & (r <= hi)).all())
    assert torch.equal(r[lo == hi], exact[lo == hi])
    plan.close()
