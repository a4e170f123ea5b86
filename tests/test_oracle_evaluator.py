"""Known-answer and metamorphic tests of the evaluator oracle.  The reference has no evaluator
(parity unpinned): these hand-computed cases ARE the specification the CUDA
path is held to."""
import numpy as np
import pytest
import torch

from oracle import evaluator as oev


def _unit(vs):
    z = torch.tensor(vs, dtype=torch.float32)
    return z


def test_kat_all_relevant_first():
    # clique 0 = {0,1,2} tightly clustered, clique 1 = {3,4} far away -> every query ranks its clique first
    z = _unit([[1, 0.01, 0], [1, 0.02, 0], [1, 0.03, 0], [0, 1, 0.01], [0, 1, 0.02]])
    c = torch.tensor([0, 0, 0, 1, 1]); i = torch.arange(5)
    aps, r1s = oev.evaluate_argsort(c, i, z, c, i, z)
    assert torch.allclose(aps, torch.ones(5, dtype=torch.float64)) and torch.all(r1s == 1)


def test_kat_hand_computed_ap():
    # query 0 (clique 0): similarities to 1..5 are strictly decreasing; relevant = items 2 and 4
    # ranks: item1 ->1, item2 ->2, item3 ->3, item4 ->4 ; AP = (1/2 + 2/4)/2 = 0.5 ; R1 = 2
    ang = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5]
    z = _unit([[np.cos(a), np.sin(a)] for a in ang])
    c = torch.tensor([0, 1, 0, 2, 0, 3]); i = torch.arange(6)
    aps, r1s = oev.evaluate_argsort(c[:1], i[:1], z[:1], c, i, z)
    assert abs(float(aps[0]) - 0.5) < 1e-12 and float(r1s[0]) == 2
    a2, r2 = oev.evaluate_rankcount(c[:1], i[:1], z[:1], c, i, z)
    assert abs(float(a2[0]) - 0.5) < 1e-12 and float(r2[0]) == 2


def test_kat_worst_case():
    # the single relevant item is the farthest of 4 candidates (self excluded): AP = 1/4, R1 = 4
    ang = [0.0, 0.1, 0.2, 0.3, 1.5]
    z = _unit([[np.cos(a), np.sin(a)] for a in ang])
    c = torch.tensor([0, 1, 2, 3, 0]); i = torch.arange(5)
    aps, r1s = oev.evaluate_argsort(c[:1], i[:1], z[:1], c, i, z)
    assert abs(float(aps[0]) - 0.25) < 1e-12 and float(r1s[0]) == 4


def test_self_is_by_id_not_position():
    # a DIFFERENT track that collides on the version id is masked out like self (lib/losses.py:40-42)
    ang = [0.0, 0.05, 0.2, 0.3]
    z = _unit([[np.cos(a), np.sin(a)] for a in ang])
    c = torch.tensor([0, 1, 0, 2]); i = torch.tensor([7, 7, 8, 9])   # item 1 collides with query 0
    aps, r1s = oev.evaluate_argsort(c[:1], i[:1], z[:1], c, i, z)
    assert float(r1s[0]) == 1 and float(aps[0]) == 1.0               # item 1 would otherwise rank first
    # queries that are not part of the corpus: nothing is masked
    aps2, r1s2 = oev.evaluate_argsort(torch.tensor([0]), torch.tensor([99]), z[:1], c[1:], i[1:], z[1:])
    assert float(r1s2[0]) == 2


def test_no_relevant_is_flagged():
    z = torch.eye(3)
    c = torch.tensor([0, 1, 2]); i = torch.arange(3)
    with pytest.raises(ValueError):
        oev.evaluate_argsort(c, i, z, c, i, z)


def test_topk_matches_sort():
    g = torch.Generator().manual_seed(5)
    z = torch.randn(60, 16, generator=g)
    c = torch.arange(60) // 3; i = torch.arange(60)
    aps, r1s, idx, sim = oev.evaluate_argsort(c, i, z, c, i, z, topk=7)
    assert idx.shape == (60, 7)
    for q in (0, 17, 59):
        s = torch.nn.functional.cosine_similarity(z[q:q + 1], z)
        s[q] = -2
        assert torch.equal(idx[q], torch.argsort(-s, stable=True)[:7])
        assert q not in idx[q].tolist()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_argsort_equals_rankcount(seed):
    g = torch.Generator().manual_seed(seed)
    n, d = 300, 24
    c = torch.randint(0, 60, (n,), generator=g)
    # make sure every clique has >= 2 members
    c = torch.cat([c, c])[:n] if False else c
    counts = torch.bincount(c, minlength=60)
    for k in torch.nonzero(counts == 1).flatten().tolist():
        c[(c == k).nonzero()[0, 0]] = int(torch.argmax(counts))
    z = torch.randn(n, d, generator=g, dtype=torch.float64)
    i = torch.arange(n)
    a1, r1 = oev.evaluate_argsort(c, i, z, c, i, z)
    a2, r2 = oev.evaluate_rankcount(c, i, z, c, i, z)
    assert torch.allclose(a1, a2, atol=1e-12) and torch.equal(r1, r2)


def test_metamorphic_invariances():
    g = torch.Generator().manual_seed(9)
    n, d = 120, 16
    c = torch.arange(n) // 4; i = torch.arange(n) + 1000
    z = torch.randn(n, d, generator=g, dtype=torch.float64)
    a0, r0 = oev.evaluate_argsort(c, i, z, c, i, z)
    # permutation of the corpus does not change per-query results
    perm = torch.randperm(n, generator=g)
    a1, r1 = oev.evaluate_argsort(c, i, z, c[perm], i[perm], z[perm])
    assert torch.allclose(a0, a1, atol=1e-12) and torch.equal(r0, r1)
    # cosine ranking is invariant to per-row positive scaling
    scale = torch.rand(n, 1, generator=g, dtype=torch.float64) * 5 + 0.1
    a2, r2 = oev.evaluate_argsort(c, i, z * scale, c, i, z * scale)
    assert torch.allclose(a0, a2, atol=1e-9) and torch.equal(r0, r2)
    m, r = oev.mean_metrics(a0, r0)
    assert 0 < m <= 1 and r >= 1


@pytest.mark.parametrize("redux", ["min", "max", "mean", "meanmin", "minmean"])
def test_chunked_oracle_degenerates_to_the_plain_evaluator(redux):
    """Chunked tracks: with one chunk per track, or with every chunk of a track equal, the
    track-level distance of any redux is the plain cosine distance -- same AP / R1 as the unchunked evaluator."""
    from wealy_b200.data import synth
    s = synth.make_eval_set(180, 24, seed=5)
    c, i, z = s["c"], s["i"], s["z"]
    aps, r1s = oev.evaluate_argsort(c, i, z, c, i, z)
    a1, r1 = oev.evaluate_argsort(c, i, z[:, None, :], c, i, z[:, None, :], redux=redux)
    assert torch.allclose(a1, aps) and torch.equal(r1, r1s)
    z4 = z[:, None, :].repeat(1, 4, 1).contiguous()
    a4, r4 = oev.evaluate_argsort(c, i, z4, c, i, z4, redux=redux)
    assert torch.allclose(a4, aps, atol=1e-6)
    assert (r4 != r1s).sum() <= 2          # fp32 rounding of a mean of equal numbers may swap a near-tie


def test_chunked_oracle_redux_ordering():
    """min <= meanmin <= mean <= max and min <= minmean <= mean hold for the reduced DISTANCES of every track pair."""
    from oracle.masked import distance_tensor_redux
    g = torch.Generator().manual_seed(3)
    d = torch.rand(7, 9, 4, 4, generator=g)
    r = {k: distance_tensor_redux(d, k) for k in ("min", "max", "mean", "meanmin", "minmean")}
    eps = 1e-6
    assert bool((r["min"] <= r["meanmin"] + eps).all()) and bool((r["meanmin"] <= r["mean"] + eps).all())
    assert bool((r["min"] <= r["minmean"] + eps).all()) and bool((r["minmean"] <= r["mean"] + eps).all())
    assert bool((r["mean"] <= r["max"] + eps).all())


def _ragged_set(n, s, d, seed):
    from wealy_b200.data import synth
    base = synth.make_eval_set(n, d, seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    z = base["z"][:, None, :] + 0.5 * base["z"].norm(dim=1).mean() / d ** 0.5 * torch.randn(n, s, d, generator=g)
    lens = torch.randint(1, s + 1, (n,), generator=g)
    return base["c"], base["i"], z.contiguous(), lens


@pytest.mark.parametrize("redux", ["min", "max", "mean", "meanmin", "minmean"])
def test_ragged_oracle_against_explicit_loops(redux):
    """Ragged tracks: the mask handed to the restated distance_tensor_redux (True = excluded, lib/tensor_ops.py:186)
    reproduces reductions written out over the valid chunks of every track pair; padding content is irrelevant."""
    c, i, z, lens = _ragged_set(40, 4, 16, seed=2)
    sim = oev._sim_block(z, z, "cos", redux, lens, lens)
    zn = z / (z.norm(dim=-1, keepdim=True) + 1e-6)
    for a in range(0, 40, 7):
        for b in range(0, 40, 5):
            dd = 1 - zn[a, :lens[a]] @ zn[b, :lens[b]].T               # valid chunk pairs only
            want = {"min": dd.min(), "max": dd.max(), "mean": dd.mean(), "meanmin": dd.min(dim=1).values.mean(),
                    "minmean": dd.mean(dim=1).min()}[redux]
            assert abs(float(sim[a, b]) - float(1 - want)) < 1e-5
    junk = z.clone()
    for t in range(40):
        junk[t, lens[t]:] = 37.0 * torch.randn(4 - int(lens[t]), 16)
    assert torch.allclose(oev._sim_block(junk, junk, "cos", redux, lens, lens), sim, atol=1e-6)
    full = torch.full((40,), 4)
    assert torch.allclose(oev._sim_block(z, z, "cos", redux, full, full), oev._sim_block(z, z, "cos", redux), atol=1e-6)


@pytest.mark.parametrize("case", ["a", "b"])
def test_oracle_ranks_by_the_reference_distance_matrix(case):
    """The reference holds no evaluator, but the distances the ranking is made of are its own: AP / R1
    of the oracle equal the ones read off an argsort of the UNMODIFIED reference's pairwise_distance_matrix(mode="cos")
    output (tests/golden/sim_modes.npz)."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sim_modes.npz"))
    x, y = torch.from_numpy(G[f"{case}_x"]), torch.from_numpy(G[f"{case}_y"])
    dist = torch.from_numpy(G[f"{case}_cos"]).double()
    n, m = dist.shape
    g = torch.Generator().manual_seed(5)
    qc, cc = torch.randint(0, 6, (n,), generator=g), torch.randint(0, 6, (m,), generator=g)
    qi, ci = torch.arange(n) + 10_000, torch.arange(m)
    aps, r1s = oev.evaluate_argsort(qc, qi, x, cc, ci, y)
    order = torch.argsort(dist, dim=1, stable=True)
    sd = torch.gather(dist, 1, order)
    for q in range(n):
        rel = (cc[order[q]] == qc[q]).double()
        if rel.sum() == 0 or bool(((sd[q, 1:] - sd[q, :-1]) <= 1e-6).any()):
            continue
        hits = torch.cumsum(rel, 0)
        ap = float((hits / torch.arange(1, m + 1) * rel).sum() / rel.sum())
        assert abs(float(aps[q]) - ap) <= 1e-6 and float(r1s[q]) == float(torch.nonzero(rel)[0, 0] + 1)


def test_rank_bands_cover_every_relevant_item():
    """rank_bands: the per-item form of the parity rule.  Its exact ranks reproduce AP / R1 of the rank-count
    evaluator, bands contain the exact rank, and a perturbation below gap / 2 keeps every k-th best rank inside."""
    from wealy_b200.data import synth  # noqa: F401  (needs the built library only for the package import)
    s = synth.make_eval_set(400, 32, seed=77)
    c, i, z = s["c"], s["i"], s["z"]
    off, sims, exact, lo, hi = oev.rank_bands(c, i, z, c, i, z, gap=1e-5)
    aps, r1s = oev.evaluate_rankcount(c, i, z, c, i, z)
    assert off.numel() == 401 and int(off[-1]) == exact.numel()
    assert bool(((lo <= exact) & (exact <= hi)).all())
    for q in range(400):
        r = exact[off[q]:off[q + 1]].double()
        k = torch.arange(1, r.numel() + 1, dtype=torch.float64)
        assert abs(float((k / r).mean()) - float(aps[q])) < 1e-12 and float(r[0]) == float(r1s[q])
        assert bool((sims[off[q]:off[q + 1]][:-1] >= sims[off[q]:off[q + 1]][1:]).all())     # best first
    # perturbed similarities (|err| < gap / 2): recompute the ranks of the k-th best relevant items by brute force
    g = torch.Generator().manual_seed(1)
    zn = z / (z.norm(dim=1, keepdim=True) + 1e-6)
    S = (zn @ zn.T).double() + (torch.rand(400, 400, generator=g).double() - 0.5) * 0.9e-5
    for q in range(0, 400, 7):
        is_self = i == i[q]
        rel = (c == c[q]) & ~is_self
        thr = torch.sort(S[q][rel], descending=True).values
        others = S[q][~is_self]
        r = 1 + (others[None, :] > thr[:, None]).sum(1)
        assert bool(((r >= lo[off[q]:off[q + 1]]) & (r <= hi[off[q]:off[q + 1]])).all())
