"""-m gpu: fused evaluation (wealy_eval_run through the C ABI) vs the evaluator oracle."""
import numpy as np
import pytest
import torch

from oracle import evaluator as oev

pytestmark = pytest.mark.gpu


def _we():
    from wealy_b200 import evaluation as we
    return we


def _synth():
    from wealy_b200.data import synth
    return synth


def _gpu_eval(c, i, z, cc=None, ci=None, cz=None, **kw):
    cc = c if cc is None else cc
    ci = i if ci is None else ci
    cz = z if cz is None else cz
    same = cc is c and cz is z
    cq, iq, zq = c.cuda(), i.cuda(), z.cuda()
    if same:
        res = _we().evaluate(cq, iq, zq, cq, iq, zq, **kw)
    else:
        res = _we().evaluate(cq, iq, zq, cc.cuda(), ci.cuda(), cz.cuda(), **kw)
    torch.cuda.synchronize()
    return [r.cpu() for r in res]


def _unit(angles):
    return torch.tensor([[np.cos(a), np.sin(a)] for a in angles], dtype=torch.float32)


def test_kat_hand_computed():
    # same cases as tests/test_oracle_evaluator.py, now through the CUDA path
    z = torch.tensor([[1, 0.01, 0], [1, 0.02, 0], [1, 0.03, 0], [0, 1, 0.01], [0, 1, 0.02]], dtype=torch.float32)
    c = torch.tensor([0, 0, 0, 1, 1]); i = torch.arange(5)
    aps, r1s = _gpu_eval(c, i, z)
    assert torch.allclose(aps, torch.ones(5)) and torch.all(r1s == 1)

    z = _unit([0.0, 0.1, 0.2, 0.3, 0.4, 0.5])
    c = torch.tensor([0, 1, 0, 2, 0, 3]); i = torch.arange(6)
    aps, r1s = _gpu_eval(c[:1], i[:1], z[:1], c, i, z)
    assert abs(float(aps[0]) - 0.5) < 1e-6 and float(r1s[0]) == 2

    z = _unit([0.0, 0.1, 0.2, 0.3, 1.5])
    c = torch.tensor([0, 1, 2, 3, 0]); i = torch.arange(5)
    aps, r1s = _gpu_eval(c[:1], i[:1], z[:1], c, i, z)
    assert abs(float(aps[0]) - 0.25) < 1e-6 and float(r1s[0]) == 4


def test_self_is_by_id_not_position():
    z = _unit([0.0, 0.05, 0.2, 0.3])
    c = torch.tensor([0, 1, 0, 2]); i = torch.tensor([7, 7, 8, 9])
    aps, r1s = _gpu_eval(c[:1], i[:1], z[:1], c, i, z)
    assert float(r1s[0]) == 1 and float(aps[0]) == 1.0
    aps, r1s = _gpu_eval(torch.tensor([0]), torch.tensor([99]), z[:1], c[1:], i[1:], z[1:])
    assert float(r1s[0]) == 2


def test_no_relevant_is_flagged():
    z = torch.eye(3)
    c = torch.tensor([0, 1, 2]); i = torch.arange(3)
    with pytest.raises(ValueError):
        _gpu_eval(c, i, z)
    c = torch.tensor([0, 0, 2])
    aps, r1s = _gpu_eval(c, i, z, allow_empty=True)
    assert torch.isnan(aps[2]) and torch.isnan(r1s[2]) and float(aps[0]) == 1.0


def _check_all_item_ranks(plan, c, i, z, cc=None, ci=None, cz=None, queries=None, gap=1e-5, sim_tol=4e-6, truth64=False):
    """The parity contract on the quantities AP is made of: the rank of EVERY relevant item (not only the best one)
    must lie in the band the candidates within `gap` of it allow, and be exact where that band is a single rank.
    `plan` has just been run; `queries` (index tensor) restricts the oracle to a sample of the queries.
    truth64: bands and similarities from the float64 evaluation of the same formula (the arbiter between two
    fp32-accumulating implementations; then |GPU - truth| <= gap / 2 is exactly the condition under which the band
    property is guaranteed)."""
    cc, ci, cz = (c if cc is None else cc), (i if ci is None else ci), (z if cz is None else cz)
    if truth64:
        z, cz = z.double(), cz.double()
    off_g, ranks_g, sims_g = (t.cpu() for t in plan.ranks())
    qsel = torch.arange(len(c)) if queries is None else torch.as_tensor(queries).long().cpu()
    off_o, sims_o, exact, lo, hi = oev.rank_bands(c[qsel], i[qsel], z[qsel], cc, ci, cz, gap=gap)
    # gather the sampled queries' CSR slices of the device result
    lens = off_g[qsel + 1] - off_g[qsel]
    assert torch.equal(lens, off_o[1:] - off_o[:-1])                 # same relevant sets
    pos = torch.cat([torch.arange(int(off_g[q]), int(off_g[q + 1])) for q in qsel.tolist()]) if len(qsel) else torch.empty(0, dtype=torch.long)
    r, sg = ranks_g[pos].long(), sims_g[pos].double()
    assert float((sg - sims_o).abs().max()) <= sim_tol               # k-th best relevant similarity
    assert bool(((r >= lo) & (r <= hi)).all()), "a relevant item's rank left the band its 1e-5 neighbours allow"
    single = lo == hi
    assert torch.equal(r[single], exact[single])                     # bit-exact ranks wherever the gap exceeds 1e-5
    return int(single.sum()), int(single.numel())


def _run_plan(c, i, z, cc=None, ci=None, cz=None, **kw):
    """Like _gpu_eval, but keeps the plan (for plan.ranks())."""
    we = _we()
    same = cc is None
    cq, iq, zq = c.cuda(), i.cuda(), z.cuda()
    if same:
        plan = we.EvalPlan(cq, iq, cq, iq)
        res = plan.run(zq, zq, **kw)
    else:
        plan = we.EvalPlan(cq, iq, cc.cuda(), ci.cuda())
        res = plan.run(zq, cz.cuda(), **kw)
    torch.cuda.synchronize()
    return plan, res


def _parity(s, precision="fp16x3", topk=None, d_map=1e-4, d_mr1=1e-4, gap=1e-5):
    aps_o, r1_o = oev.evaluate_argsort(s["c"], s["i"], s["z"], s["c"], s["i"], s["z"])
    res = _gpu_eval(s["c"], s["i"], s["z"], precision=precision, topk=topk)
    aps, r1s = res[0].double(), res[1].double()
    assert abs(float(aps.mean()) - float(aps_o.mean())) <= d_map            # MAP within 1e-4
    assert abs(float(r1s.mean()) - float(r1_o.mean())) <= d_mr1 * max(1.0, float(r1_o.mean()))
    if precision == "fp16x3":
        # ranks bit-exact wherever the similarity gap exceeds 1e-5: R1 must lie in the range the
        # candidates within 1e-5 of the best relevant item allow
        lo, hi = oev.rank_tolerance(s["c"], s["i"], s["z"], s["c"], s["i"], s["z"], gap=gap)
        assert bool(((r1s >= lo) & (r1s <= hi)).all())
        exact = (lo == hi)
        assert torch.equal(r1s[exact], r1_o[exact])
        # ... and so must the rank of every other relevant item (what AP is made of)
        plan, _ = _run_plan(s["c"], s["i"], s["z"], precision=precision, topk=topk)
        _check_all_item_ranks(plan, s["c"], s["i"], s["z"], gap=gap)
        plan.close()
    return res, aps_o, r1_o


@pytest.mark.parametrize("n,d,seed", [(700, 64, 0), (2000, 128, 1), (3000, 1024, 2), (1300, 200, 3)])
def test_parity_with_oracle(n, d, seed):
    s = _synth().make_eval_set(n, d, seed=seed)
    _parity(s)


def test_single_pass_mode_is_map_accurate():
    s = _synth().make_eval_set(3000, 256, seed=4)
    _parity(s, precision="fp16", d_map=1e-4, d_mr1=2e-3)


def test_topk_matches_oracle():
    s = _synth().make_eval_set(2500, 256, seed=5)
    k = 20
    (aps, r1s, idx, sim), _, _ = _parity(s, topk=k)
    _, _, idx_o, sim_o = oev.evaluate_argsort(s["c"], s["i"], s["z"], s["c"], s["i"], s["z"], topk=k)
    assert idx.shape == (2500, k) and idx.dtype == torch.long
    assert (sim - sim_o).abs().max() <= 4e-6                                  # the k best similarities agree
    # indices exact wherever neighbouring similarities are more than 1e-5 apart
    gaps_ok = torch.ones_like(idx_o, dtype=torch.bool)
    gaps_ok[:, 1:] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    gaps_ok[:, :-1] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    gaps_ok[:, -1] = False   # the last position also depends on the (k+1)-th best, which the lists do not show
    assert torch.equal(idx[gaps_ok], idx_o[gaps_ok])
    assert bool((idx != torch.arange(2500)[:, None]).all())                   # self never returned


@pytest.mark.parametrize("k", [150, 300, 800])
def test_topk_large_k(k):
    # long candidate lists: the out-of-line selection variants and the generic finalize kernel
    s = _synth().make_eval_set(3000, 128, seed=11)
    aps, r1s, idx, sim = _gpu_eval(s["c"], s["i"], s["z"], topk=k)
    _, _, idx_o, sim_o = oev.evaluate_argsort(s["c"][:200], s["i"][:200], s["z"][:200], s["c"], s["i"], s["z"], topk=k)
    assert (sim[:200] - sim_o).abs().max() <= 4e-6
    ok = torch.ones_like(idx_o, dtype=torch.bool)
    ok[:, 1:] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    ok[:, :-1] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    ok[:, -1] = False   # the last position also depends on the (k+1)-th best, which the lists do not show
    assert torch.equal(idx[:200][ok], idx_o[ok])
    with pytest.raises(NotImplementedError):
        _gpu_eval(s["c"], s["i"], s["z"], topk=801)


def test_topk_larger_than_corpus_and_k1():
    s = _synth().make_eval_set(40, 16, seed=6)
    aps, r1s, idx, sim = _gpu_eval(s["c"], s["i"], s["z"], topk=100)
    assert idx.shape == (40, 40)
    assert bool((idx[:, -1] == -1).all()) and bool(torch.isinf(sim[:, -1]).all())   # only 39 candidates exist
    aps, r1s, idx1, sim1 = _gpu_eval(s["c"], s["i"], s["z"], topk=1)
    assert torch.equal(idx1[:, 0], idx[:, 0])


def test_queries_disjoint_from_corpus_and_collisions():
    syn = _synth()
    s = syn.make_eval_set(1500, 96, seed=7)
    q = slice(0, 300)
    cand = slice(300, 1500)
    # give some candidates the version id of a query from ANOTHER clique (md5 id collisions at scale)
    i2 = s["i"].clone()
    i2[400] = s["i"][5]
    i2[401] = s["i"][6]
    # keep only queries that still have a relevant candidate in the reduced corpus
    keep = torch.tensor([bool(((s["c"][cand] == s["c"][k]) & (i2[cand] != s["i"][k])).any()) for k in range(300)])
    qc, qi, qz = s["c"][q][keep], s["i"][q][keep], s["z"][q][keep]
    aps_o, r1_o = oev.evaluate_argsort(qc, qi, qz, s["c"][cand], i2[cand], s["z"][cand])
    aps, r1s = _gpu_eval(qc, qi, qz, s["c"][cand], i2[cand], s["z"][cand])
    assert abs(float(aps.double().mean()) - float(aps_o.mean())) <= 1e-4
    assert (r1s.double() != r1_o).sum() <= 2


def test_exact_duplicates_and_giant_clique():
    g = torch.Generator().manual_seed(8)
    n_big = 359                                         # largest clique of SHS100K-TRAIN
    n = n_big + 540
    z = torch.randn(n, 64, generator=g)
    c = torch.cat([torch.zeros(n_big, dtype=torch.long), 1 + torch.arange(n - n_big) // 3])
    z[:n_big] += 0.8 * torch.randn(1, 64, generator=g)
    z[10] = z[11]                                       # exact duplicates inside the clique (ties)
    z[500] = z[20]                                      # a negative that ties with a relevant item
    i = torch.arange(n)
    aps_o, r1_o = oev.evaluate_argsort(c, i, z, c, i, z)
    aps, r1s = _gpu_eval(c, i, z)
    # exact ties (duplicates) may be ordered either way: AP of the affected queries moves by O(1/P)
    assert abs(float(aps.double().mean()) - float(aps_o.mean())) <= 1e-3
    lo, hi = oev.rank_tolerance(c, i, z, c, i, z, gap=1e-5)
    assert bool(((r1s.double() >= lo) & (r1s.double() <= hi)).all())


def test_counts_beyond_16_bits_and_topk_consistency():
    """A query whose relevant item is ranked near the bottom of a 40k corpus: the per-unit rank counters
    (16-bit fields in shared memory, spilled to the global histogram) must not wrap, with and without the
    top-k path (which sweeps the whole corpus in a single unit)."""
    g = torch.Generator().manual_seed(12)
    n, d = 40000, 64
    z = torch.randn(n, d, generator=g)
    c = torch.arange(n) // 2
    i = torch.arange(n)
    z[1] = -z[0] + 0.01 * torch.randn(d, generator=g)      # query 0's only relevant item is its antipode
    z[3] = z[2]
    aps, r1s = _gpu_eval(c, i, z)
    aps2, r1s2, idx, sim = _gpu_eval(c, i, z, topk=5)
    # the symmetric sweep (clique-sorted, thresholds from the clique-block MMA kernel) and the top-k sweep
    # (general rectangle kernel) accumulate the relevant similarities in different orders: a rank may move by
    # one where a candidate lies within ~1e-7 of the relevant item (allowed below the 1e-5 gap of the contract;
    # this corpus has a candidate every 8e-6)
    dr = (r1s - r1s2).abs()
    assert float(dr.max()) <= 2 and float((dr > 0).float().mean()) < 0.05
    assert torch.allclose(aps, aps2, rtol=2e-2, atol=1e-4)
    assert float(r1s[0]) == n - 1 and float(r1s[1]) == n - 1      # ranked last of the n - 1 candidates
    assert float(r1s[2]) == 1 and float(r1s[3]) == 1
    nq = 64
    aps_o, r1_o = oev.evaluate_argsort(c[:nq], i[:nq], z[:nq], c, i, z)
    assert torch.equal(r1s[:nq].double(), r1_o)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_symmetric_sweep_emulated_on_one_gpu(world):
    """The multi-GPU all-vs-all path: every rank sweeps the row blocks rank (mod world) of the symmetric
    problem, the rank counters are summed (here by hand instead of the NCCL all-reduce), finish() yields the
    complete result.  The ranks are emulated one after the other on a single GPU."""
    s = _synth().make_eval_set(1700, 96, seed=13)
    we = _we()
    c, i, z = s["c"].cuda(), s["i"].cuda(), s["z"].cuda()
    ref = we.EvalPlan(c, i, c, i).run(z, z)
    total = None
    plans = []
    for r in range(world):
        pl = we.EvalPlan(c, i, c, i)
        pl.sweep_shard(z, r, world)
        cnt = pl.counts_tensor()
        assert cnt.dtype == torch.int32 and cnt.numel() == max(pl.total_pairs, 1)
        total = cnt.clone() if total is None else total + cnt
        plans.append(pl)
    plans[0].counts_tensor().copy_(total)
    out = plans[0].finish()
    torch.cuda.synchronize()
    assert torch.equal(out["aps"], ref["aps"]) and torch.equal(out["r1s"], ref["r1s"])
    assert torch.allclose(out["sums"], ref["sums"], rtol=1e-9)
    with pytest.raises(AssertionError):
        we.EvalPlan(c[:100], i[:100], c, i).sweep_shard(z, 0, 2)      # needs queries == candidates


def test_plan_reuse_and_half_precision_inputs():
    s = _synth().make_eval_set(1200, 128, seed=9)
    we = _we()
    c, i = s["c"].cuda(), s["i"].cuda()
    plan = we.EvalPlan(c, i, c, i)
    assert plan.total_pairs > 0 and plan.queries_without_relevant == 0 and plan.max_relevant >= 1
    z1 = s["z"].cuda()
    r1 = plan.run(z1, z1)
    z2 = (s["z"] * 0.5 + 0.1).cuda()
    r2 = plan.run(z2, z2)
    r1b = plan.run(z1, z1)
    assert torch.equal(r1["aps"], r1b["aps"]) and not torch.equal(r1["aps"], r2["aps"])
    zh = s["z"].half().cuda()
    rh = plan.run(zh, zh)
    aps_o, _ = oev.evaluate_argsort(s["c"], s["i"], s["z"].half().float(), s["c"], s["i"], s["z"].half().float())
    assert abs(float(rh["aps"].double().mean()) - float(aps_o.mean())) <= 1e-4
    m, r = we.mean_metrics(r1["sums"])
    assert abs(m - float(r1["aps"].double().mean())) < 1e-6
    plan.close()


def test_full_size_properties():
    """BASELINE configs[1] size (100k x 1024): per-query results are invariant under a permutation of
    the corpus, and a query sample agrees with the CPU oracle."""
    syn, we = _synth(), _we()
    n = 100_000
    s = syn.make_eval_set(n, 1024, seed=0, device="cuda", md5_ids=False)
    z, c, i = s["z"], s["c"], s["i"]
    aps, r1s = we.evaluate(c, i, z, c, i, z)
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    aps_p, r1_p = we.evaluate(c, i, z, c[perm], i[perm], z[perm])
    # the two calls take different kernels (symmetric clique-sorted sweep vs general rectangle) whose relevant
    # similarities are accumulated in different orders: a rank may move by one where a candidate lies within
    # ~1e-7 of a relevant item (far inside the 1e-5 gap of the contract), nothing else may change
    dr = (r1s - r1_p).abs()
    assert float(dr.max()) <= 2 and float((dr > 0).float().mean()) < 1e-2
    da = (aps - aps_p).abs()
    assert abs(float(aps.double().mean() - aps_p.double().mean())) <= 1e-5 and float((da > 1e-6).float().mean()) < 5e-2
    nq = 128
    zc, cc, ic = z.cpu(), c.cpu(), i.cpu()
    aps_o, r1_o = oev.evaluate_argsort(cc[:nq], ic[:nq], zc[:nq], cc, ic, zc)
    assert abs(float(aps[:nq].double().mean().cpu()) - float(aps_o.mean())) <= 1e-4
    lo, hi = oev.rank_tolerance(cc[:nq], ic[:nq], zc[:nq], cc, ic, zc, gap=1e-5)
    r = r1s[:nq].double().cpu()
    assert bool(((r >= lo) & (r <= hi)).all())


def _oracle_match(c, i, z, r1_slack=0):
    aps_o, r1_o = oev.evaluate_argsort(c, i, z, c, i, z)
    plan, res = _run_plan(c, i, z)
    aps, r1s = res["aps"].cpu(), res["r1s"].cpu()
    lo, hi = oev.rank_tolerance(c, i, z, c, i, z, gap=1e-5)
    r = r1s.double()
    assert bool(((r >= lo) & (r <= hi)).all())
    assert abs(float(aps.double().mean()) - float(aps_o.mean())) <= 1e-4
    _check_all_item_ranks(plan, c, i, z)                              # every relevant item, not only the best
    plan.close()
    return aps, r1s, aps_o, r1_o


def test_symmetric_sweep_version_id_collisions_across_cliques():
    """The clique-sorted symmetric sweep skips id tests in 'clean' tiles: pairs of DIFFERENT cliques that share a
    version id (md5-derived 31-bit ids collide at scale, lib/embedding_dataset/utils.py:7-13) are not candidates of
    each other and must be found by the plan wherever they fall (far from the diagonal included)."""
    s = _synth().make_eval_set(6000, 64, seed=21, md5_ids=False)
    c, i, z = s["c"], s["i"].clone(), s["z"].clone()
    order = torch.argsort(c, stable=True)
    far = [(int(order[10]), int(order[5500])), (int(order[700]), int(order[3100])), (int(order[4000]), int(order[4001 + 300]))]
    for a, b in far:
        assert c[a] != c[b]
        i[b] = i[a]                       # collision between cliques
        z[b] = z[a] * 1.5                 # ... of near-identical tracks: scoring the pair would put b at rank 1 of a
    _, r1s, _, r1_o = _oracle_match(c, i, z)
    for a, b in far:
        assert float(r1s[a]) == float(r1_o[a]) and float(r1s[b]) == float(r1_o[b])


def test_symmetric_sweep_many_equal_version_ids():
    """More than 16 tracks with one version id (pathological input): the plan gives up on locating the colliding
    pairs and runs every tile with id tests -- still the oracle's result."""
    s = _synth().make_eval_set(2500, 48, seed=22, md5_ids=False)
    c, i, z = s["c"], s["i"].clone(), s["z"]
    i[torch.arange(0, 2500, 100)] = 123456789          # 25 tracks share a version id
    keep = torch.ones(2500, dtype=torch.bool)
    for q in range(2500):                              # drop queries left without a relevant candidate
        keep[q] = bool(((c == c[q]) & (i != i[q])).any())
    c, i, z = c[keep], i[keep], z[keep]
    _oracle_match(c, i, z)


@pytest.mark.parametrize("n", [127, 128, 129, 255, 257, 1000])
def test_symmetric_sweep_ragged_sizes(n):
    # partial row blocks / column tiles: padded rows and columns must never be scored
    s = _synth().make_eval_set(n, 40, seed=30 + n)
    _oracle_match(s["c"], s["i"], s["z"])


@pytest.mark.parametrize("n,d,per", [(3000, 32, 6), (2500, 16, 12), (1100, 24, 40)])
def test_symmetric_sweep_dense_deep_path(n, d, per):
    """Uncorrelated clique members: a query's relevant items are scattered over the whole similarity range, so
    about half of ALL pairs lie above its third-lowest threshold -- every chunk overflows the warp queue (the
    batched 4-columns-at-a-time path), row caches overflow for the larger cliques, counters run high."""
    g = torch.Generator().manual_seed(n + per)
    z = torch.randn(n, d, generator=g)
    c = torch.arange(n) // per
    perm = torch.randperm(n, generator=g)
    c, i = c[perm], torch.arange(n)
    aps, r1s, aps_o, r1_o = _oracle_match(c, i, z)
    assert float((aps.double() - aps_o).abs().max()) <= 5e-3      # per-query AP: only near-tie rank swaps may differ


def test_symmetric_sweep_randomised_shapes():
    """Many small random problems (md5-derived version ids, shuffled rows, ragged sizes) against the oracle."""
    syn = _synth()
    g = torch.Generator().manual_seed(99)
    for trial in range(12):
        n = int(torch.randint(60, 2600, (1,), generator=g))
        d = int(torch.randint(8, 300, (1,), generator=g))
        s = syn.make_eval_set(n, d, seed=1000 + trial, sigma=float(1.0 + 3.0 * torch.rand(1, generator=g)))
        _oracle_match(s["c"], s["i"], s["z"])


def test_config1_shs100k_test_shape_full_parity():
    """BASELINE.json configs[0]: the SHS100K-TEST-shaped evaluation the reference can run on CPU -- 10 547 versions in
    1 692 cliques (the exact clique-size multiset of the shipped split), 1024-d.  Every query against the oracle."""
    s = _synth().make_eval_set(10547, 1024, seed=0)
    assert int(torch.bincount(s["c"]).gt(0).sum()) == 1692
    aps, r1s, aps_o, r1_o = _oracle_match(s["c"], s["i"], s["z"])
    assert abs(float(r1s.double().mean()) - float(r1_o.mean())) <= 1e-4 * float(r1_o.mean())      # MR1 within 1e-4


def test_equal_ids_in_distinct_tensors_take_the_all_vs_all_path():
    we = _we()
    s = _synth().make_eval_set(1500, 64, seed=41)
    c, i, z = s["c"].cuda(), s["i"].cuda(), s["z"].cuda()
    a1, r1 = we.evaluate(c, i, z, c, i, z)
    a2, r2 = we.evaluate(c, i, z, c.clone(), i.clone(), z)          # same values, other buffers
    assert torch.equal(a1, a2) and torch.equal(r1, r2)
    p = we.EvalPlan(c, i, c.clone(), i.clone())
    p.sweep_shard(z, 0, 1)                                          # the symmetric sweep needs an all-vs-all plan
    i2 = i.clone(); i2[3] += 1
    with pytest.raises(AssertionError):
        we.EvalPlan(c, i, c.clone(), i2).sweep_shard(z, 0, 2)


@pytest.mark.parametrize("d", [1, 7, 65, 4096])
def test_tiny_sets_and_odd_embedding_sizes(d):
    """Two tracks of one clique up to a few dozen tracks; D = 1, not a multiple of the k-block, and 64 k-blocks."""
    g = torch.Generator().manual_seed(d)
    z = torch.randn(2, d, generator=g) + 0.1
    aps, r1s = _gpu_eval(torch.tensor([5, 5]), torch.tensor([10, 11]), z)
    assert torch.equal(aps, torch.ones(2)) and torch.equal(r1s, torch.ones(2))
    n = 37
    c = torch.arange(n) // 3
    c[-1] = c[-2]                                                         # 37 = 12 * 3 + 1: no singleton clique
    i = torch.arange(n) + 100
    z = torch.randn(n, d, generator=g)
    aps_o, r1_o = oev.evaluate_argsort(c, i, z, c, i, z)
    aps, r1s = _gpu_eval(c, i, z)
    if d == 1:                                                            # every similarity is +-1: all ties
        assert abs(float(aps.mean()) - float(aps_o.mean())) < 0.5
        return
    lo, hi = oev.rank_tolerance(c, i, z, c, i, z, gap=1e-5)
    assert bool(((r1s.double() >= lo) & (r1s.double() <= hi)).all())
    assert abs(float(aps.double().mean()) - float(aps_o.mean())) <= 1e-4
    # general (queries != corpus tensors) path on the same data
    aps2, r1s2 = _gpu_eval(c[:9], i[:9], z[:9].clone(), c, i, z)
    assert bool(((r1s2.double() >= lo[:9]) & (r1s2.double() <= hi[:9])).all())


def test_zero_vectors_and_float64_inputs():
    """Adversarial rows: all-zero embeddings have similarity 0 to everything (x / (|x| + eps), lib/tensor_ops.py:152-176);
    float64 embeddings are rounded to float32 on entry (documented in wealy_b200.evaluation): same ranks, never computed on the CPU."""
    s = _synth().make_eval_set(600, 48, seed=21)
    z = s["z"].clone()
    z[::50] = 0
    aps_o, r1_o = oev.evaluate_argsort(s["c"], s["i"], z, s["c"], s["i"], z)
    aps, r1s = _gpu_eval(s["c"], s["i"], z)
    assert bool(torch.isfinite(aps).all())
    lo, hi = oev.rank_tolerance(s["c"], s["i"], z, s["c"], s["i"], z, gap=1e-5)
    nz = torch.ones(600, dtype=torch.bool)
    nz[::50] = False                       # a zero query ties with every candidate: its rank is a tie-break, not compared
    assert abs(float(aps.double()[nz].mean()) - float(aps_o[nz].mean())) <= 1e-4
    assert bool(((r1s.double() >= lo) & (r1s.double() <= hi))[nz].all())
    a64, r64 = _gpu_eval(s["c"], s["i"], s["z"].double())
    a32, r32 = _gpu_eval(s["c"], s["i"], s["z"])
    assert torch.equal(a64, a32) and torch.equal(r64, r32)


@pytest.mark.parametrize("case", ["a", "b"])
def test_ranking_follows_the_reference_distance_matrix(case):
    """The reference has no evaluator, but it does define the distances the ranking is made of: the fused sweep's
    full candidate order (top-k with k = Nc) and AP / R1 must be the ones an argsort of the UNMODIFIED reference's
    pairwise_distance_matrix(x, y, mode="cos") output gives (tests/golden/sim_modes.npz; case a holds a zero vector
    and an exact duplicate)."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "sim_modes.npz"))
    x, y = torch.from_numpy(G[f"{case}_x"]), torch.from_numpy(G[f"{case}_y"])
    dist = torch.from_numpy(G[f"{case}_cos"]).double()                       # [n, m], reference output
    n, m = dist.shape
    g = torch.Generator().manual_seed(5)
    qc, cc = torch.randint(0, 6, (n,), generator=g), torch.randint(0, 6, (m,), generator=g)
    qi, ci = torch.arange(n) + 10_000, torch.arange(m)                       # queries are not in the corpus
    aps, r1s, idx, sim = _gpu_eval(qc, qi, x, cc, ci, y, topk=m)
    got = torch.empty(n, m).scatter_(1, idx, sim)
    assert (got.double() - (1 - dist)).abs().max() <= 4e-6
    order = torch.argsort(dist, dim=1, stable=True)
    sd = torch.gather(dist, 1, order)
    ok = torch.ones(n, m, dtype=torch.bool)
    ok[:, 1:] &= (sd[:, 1:] - sd[:, :-1]) > 1e-5
    ok[:, :-1] &= (sd[:, 1:] - sd[:, :-1]) > 1e-5
    assert torch.equal(idx[ok], order[ok])                                   # same order wherever the gap exceeds 1e-5
    for q in range(n):
        rel = (cc[order[q]] == qc[q]).double()
        if rel.sum() == 0 or not bool(ok[q].all()):
            continue
        hits = torch.cumsum(rel, 0)
        ap = float((hits / torch.arange(1, m + 1) * rel).sum() / rel.sum())
        assert abs(float(aps[q]) - ap) <= 1e-6 and float(r1s[q]) == float(torch.nonzero(rel)[0, 0] + 1)


def test_all_item_ranks_general_and_topk_paths():
    """plan.ranks() through the other two kernels: queries disjoint from the corpus (rectangle sweep) and the
    all-vs-all top-k sweep; bands for every relevant item."""
    s = _synth().make_eval_set(2600, 160, seed=51)
    q, cand = slice(0, 500), slice(500, 2600)
    keep = torch.tensor([bool((s["c"][cand] == s["c"][k]).any()) for k in range(500)])
    qc, qi, qz = s["c"][q][keep], s["i"][q][keep], s["z"][q][keep]
    plan, res = _run_plan(qc, qi, qz, s["c"][cand], s["i"][cand], s["z"][cand])
    single, total = _check_all_item_ranks(plan, qc, qi, qz, s["c"][cand], s["i"][cand], s["z"][cand])
    assert single > 0.9 * total                                      # the exactness clause is not vacuous
    plan.close()
    plan, res = _run_plan(s["c"], s["i"], s["z"], topk=10)
    _check_all_item_ranks(plan, s["c"], s["i"], s["z"])
    plan.close()


def _topk_sample_parity(s, k, n_sample, seed, sim_tol=4e-6, expect_path=None):
    """Top-k of the whole set on the GPU vs the oracle on a query sample: similarities within 4e-6, indices exact
    wherever neighbouring similarities differ by more than 1e-5, self never returned."""
    we = _we()
    n = s["c"].numel()
    c, i, z = s["c"].cuda(), s["i"].cuda(), s["z"].cuda()
    plan = we.EvalPlan(c, i, c, i)
    res = plan.run(z, z, topk=k)
    torch.cuda.synchronize()
    if expect_path is not None:
        assert plan.last_topk_path() == expect_path
    qs = torch.randperm(n, generator=torch.Generator().manual_seed(seed))[:n_sample]
    cc, ic, zc = s["c"].cpu(), s["i"].cpu(), s["z"].cpu()
    aps_o, r1_o, idx_o, sim_o = oev.evaluate_argsort(cc[qs], ic[qs], zc[qs], cc, ic, zc, topk=k)
    idx, sim = res["topk_idx"][qs.cuda()].cpu(), res["topk_sim"][qs.cuda()].cpu()
    assert (sim - sim_o).abs().max() <= sim_tol
    ok = torch.ones_like(idx_o, dtype=torch.bool)
    ok[:, 1:] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    ok[:, :-1] &= (sim_o[:, :-1] - sim_o[:, 1:]) > 1e-5
    ok[:, -1] = False   # the last position also depends on the (k+1)-th best, which the lists do not show
    assert float(ok.float().mean()) > 0.5
    assert torch.equal(idx[ok], idx_o[ok])
    assert bool((idx != qs[:, None]).all())
    aps, r1s = res["aps"][qs.cuda()].double().cpu(), res["r1s"][qs.cuda()].double().cpu()
    assert abs(float(aps.mean()) - float(aps_o.mean())) <= 1e-4
    assert abs(float(r1s.mean()) - float(r1_o.mean())) <= 1e-4 * max(1.0, float(r1_o.mean()))
    _check_all_item_ranks(plan, cc, ic, zc, queries=qs, sim_tol=sim_tol, truth64=sim_tol > 4e-6)
    plan.close()


def test_config5_lyric_covers_shape_top100():
    """BASELINE.json configs[4]: lyric-covers multimodal retrieval -- ~50k tracks, 2048-d (text + audio concatenated),
    top-100 output; clique sizes bootstrapped from the shipped lyric-covers test split.  256 sampled queries against
    the oracle: top-100 indices / similarities, AP / R1 and the rank band of every relevant item."""
    s = _synth().make_eval_set(50_000, 2048, seed=5, dist="lyric_covers_test", device="cuda", md5_ids=False)
    # 2048-term fp32 accumulations on both sides: similarities within gap / 2 = 5e-6 (of the float64 value for the
    # relevant items, of the oracle's float32 value for the top-k lists); this size takes the symmetric top-k sweep
    _topk_sample_parity(s, 100, 256, seed=17, sim_tol=5e-6, expect_path=1)


def test_config3_discogs_full_scale_shape():
    """BASELINE.json configs[2]: Discogs-VI full scale, 500k x 1024 all-vs-all on one GPU (2.5e11 pairs): 64 sampled
    queries against the oracle -- MAP / MR1 within 1e-4, rank bands of every relevant item."""
    we = _we()
    n = 500_000
    s = _synth().make_eval_set(n, 1024, seed=3, device="cuda", md5_ids=False)
    c, i, z = s["c"], s["i"], s["z"]
    plan = we.EvalPlan(c, i, c, i)
    res = plan.run(z, z)
    torch.cuda.synchronize()
    qs = torch.randperm(n, generator=torch.Generator().manual_seed(23))[:64]
    cc, ic, zc = c.cpu(), i.cpu(), z.cpu()
    aps_o, r1_o = oev.evaluate_argsort(cc[qs], ic[qs], zc[qs], cc, ic, zc)
    aps, r1s = res["aps"][qs.cuda()].double().cpu(), res["r1s"][qs.cuda()].double().cpu()
    assert abs(float(aps.mean()) - float(aps_o.mean())) <= 1e-4
    assert abs(float(r1s.mean()) - float(r1_o.mean())) <= 1e-4 * max(1.0, float(r1_o.mean()))
    lo, hi = oev.rank_tolerance(cc[qs], ic[qs], zc[qs], cc, ic, zc, gap=1e-5)
    assert bool(((r1s >= lo) & (r1s <= hi)).all())
    _check_all_item_ranks(plan, cc, ic, zc, queries=qs)
    # size-independent property at full scale: the sums behind MAP / MR1 equal the per-query outputs
    sm = res["sums"].cpu()
    assert int(sm[2]) == n and abs(float(sm[0]) - float(res["aps"].double().sum())) <= 1e-6 * n
    plan.close()


@pytest.mark.parametrize("n,d,k", [(20000, 96, 10), (33000, 64, 128), (33000, 128, 50)])
def test_symmetric_topk_sweep_matches_oracle_and_rectangle(n, d, k):
    """All-vs-all top-k through the SYMMETRIC sweep (sampled per-query bounds, candidates collected in both directions;
    csrc/topk_sym_kernels.cuh) vs the oracle on a query sample and vs the rectangle sweep's streaming top-k on ALL
    queries (md5-derived version ids: collisions between cliques are candidates of nobody)."""
    import os
    we = _we()
    s = _synth().make_eval_set(n, d, seed=60 + k, device="cuda")
    _topk_sample_parity(s, k, 96, seed=k, expect_path=1)
    c, i, z = s["c"], s["i"], s["z"]
    plan = we.EvalPlan(c, i, c, i)
    a = plan.run(z, z, topk=k)
    assert plan.last_topk_path() == 1
    os.environ["WEALY_SYM_TOPK"] = "0"
    try:
        b = plan.run(z, z, topk=k)
        assert plan.last_topk_path() == 2
    finally:
        del os.environ["WEALY_SYM_TOPK"]
    torch.cuda.synchronize()
    assert float((a["topk_sim"] - b["topk_sim"]).abs().max()) <= 2e-6         # same planes, other accumulation order
    gap_ok = torch.ones_like(a["topk_idx"], dtype=torch.bool)
    sb = b["topk_sim"]
    gap_ok[:, 1:] &= (sb[:, :-1] - sb[:, 1:]) > 1e-5
    gap_ok[:, :-1] &= (sb[:, :-1] - sb[:, 1:]) > 1e-5
    gap_ok[:, -1] = False   # the last position also depends on the (k+1)-th best, which the lists do not show
    assert torch.equal(a["topk_idx"][gap_ok], b["topk_idx"][gap_ok])
    assert abs(float(a["aps"].double().mean() - b["aps"].double().mean())) <= 1e-5
    plan.close()


def test_symmetric_topk_falls_back_when_a_list_overflows():
    """Thousands of exact duplicates of one track: every one of them has thousands of candidates tied at similarity 1,
    far more than a candidate list holds -- the overflow is detected and the call recomputed by the rectangle sweep."""
    we = _we()
    s = _synth().make_eval_set(18000, 64, seed=71, device="cuda", md5_ids=False)
    c, i, z = s["c"], s["i"], s["z"].clone()
    z[:3000] = z[0]
    plan = we.EvalPlan(c, i, c, i)
    res = plan.run(z, z, topk=20)
    torch.cuda.synchronize()
    assert plan.last_topk_path() == 3
    sim = res["topk_sim"].cpu()
    assert bool((sim[:3000] > 1 - 1e-5).all())                               # the duplicates fill each other's lists
    qs = torch.arange(3000, 3064)
    _, _, idx_o, sim_o = oev.evaluate_argsort(c.cpu()[qs], i.cpu()[qs], z.cpu()[qs], c.cpu(), i.cpu(), z.cpu(), topk=20)
    assert (sim[qs] - sim_o).abs().max() <= 4e-6
    plan.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_staged_sharded_sweep_emulated_on_one_gpu(world):
    """The multi-GPU path as the product runs it: the relevant similarities are computed ONCE across the ranks (each rank
    its share of the queries, the float buffer summed -- every element has one non-zero contribution), then every rank
    sweeps its row blocks and the rank counters are summed.  Ranks emulated one after the other on a single GPU."""
    s = _synth().make_eval_set(2300, 96, seed=14)
    we = _we()
    c, i, z = s["c"].cuda(), s["i"].cuda(), s["z"].cuda()
    ref = we.EvalPlan(c, i, c, i).run(z, z)
    plans = [we.EvalPlan(c, i, c, i) for _ in range(world)]
    for r, pl in enumerate(plans):
        pl.shard_prepare(z, r, world)
    thr = [pl.thresholds_tensor() for pl in plans]
    nz = torch.stack([(t != 0) & torch.isfinite(t) for t in thr]).sum(0)
    assert int(nz.max()) <= 1                                       # disjoint shares: the sum is exact
    total = torch.stack(thr).sum(0)
    for t in thr:
        t.copy_(total)
    counts = None
    for r, pl in enumerate(plans):
        pl.shard_sweep(r, world)
        cnt = pl.counts_tensor()
        counts = cnt.clone() if counts is None else counts + cnt
    plans[0].counts_tensor().copy_(counts)
    out = plans[0].finish()
    torch.cuda.synchronize()
    assert torch.equal(out["aps"], ref["aps"]) and torch.equal(out["r1s"], ref["r1s"])
    _check_all_item_ranks(plans[0], s["c"], s["i"], s["z"])         # plan.ranks() after a staged run
    for pl in plans:
        pl.close()


def test_eval_pipeline_overlaps_requests_and_matches_evaluate():
    """EvalPipeline: back-to-back requests from pinned host memory (upload of request k + 1 on the copy stream while
    request k sweeps); every result equals the synchronous evaluate() of the same request."""
    we = _we()
    sets = [_synth().make_eval_set(n, 64, seed=80 + k) for k, n in enumerate((1500, 1500, 900, 1500, 2100))]
    want = []
    for s in sets:
        c, i, z = s["c"].cuda(), s["i"].cuda(), s["z"].cuda()    # the SAME tensors on both sides: the symmetric sweep
        a, r = we.evaluate(c, i, z, c, i, z)
        want.append((a.cpu(), r.cpu()))
    pipe = we.EvalPipeline()
    tickets = []
    for k, s in enumerate(sets):
        tickets.append(pipe.submit(s["c"].pin_memory(), s["i"].pin_memory(), s["z"].pin_memory()))
        if k >= 1:
            a, r = pipe.result(tickets[k - 1])
            assert torch.equal(a, want[k - 1][0]) and torch.equal(r, want[k - 1][1])
    a, r = pipe.result(tickets[-1])
    assert torch.equal(a, want[-1][0]) and torch.equal(r, want[-1][1])
    with pytest.raises(ValueError):
        pipe.result(tickets[0])                                   # overwritten: more than `depth` requests ago
    # a set with a singleton clique is refused like evaluate() refuses it, and the pipeline stays usable
    bad = sets[0]
    c_bad = bad["c"].clone()
    c_bad[0] = int(c_bad.max()) + 1
    with pytest.raises(ValueError):
        pipe.submit(c_bad.pin_memory(), bad["i"].pin_memory(), bad["z"].pin_memory())
    a, r = pipe.result(pipe.submit(sets[1]["c"].pin_memory(), sets[1]["i"].pin_memory(), sets[1]["z"].pin_memory()))
    assert torch.equal(a, want[1][0]) and torch.equal(r, want[1][1])
    pipe.close()
    lax = we.EvalPipeline(allow_empty=True)                      # ... or scored without that query when asked to
    a, r = lax.result(lax.submit(c_bad.pin_memory(), bad["i"].pin_memory(), bad["z"].pin_memory()))
    cb, ib, zb = c_bad.cuda(), bad["i"].cuda(), bad["z"].cuda()
    ae, re_ = we.evaluate(cb, ib, zb, cb, ib, zb, allow_empty=True)
    assert torch.allclose(a, ae.cpu(), rtol=0, atol=0, equal_nan=True) and torch.allclose(r, re_.cpu(), rtol=0, atol=0, equal_nan=True)
    lax.close()
    # chunked tracks go through the same pipeline ([N, s, D] embeddings, the redux of the constructor)
    g = torch.Generator().manual_seed(5)
    zt = (sets[0]["z"][:, None, :] + 0.05 * torch.randn(1500, 4, 64, generator=g)).contiguous()
    ct, it = sets[0]["c"].cuda(), sets[0]["i"].cuda()
    ztc = zt.cuda()
    at, rt = we.evaluate(ct, it, ztc, ct, it, ztc, redux="mean")
    chunked = we.EvalPipeline(redux="mean")
    a, r = chunked.result(chunked.submit(sets[0]["c"].pin_memory(), sets[0]["i"].pin_memory(), zt.pin_memory()))
    assert torch.equal(a, at.cpu()) and torch.equal(r, rt.cpu())
    chunked.close()


def test_library_scratch_is_reused_and_returned(monkeypatch):  # noqa: C901
    """Plans come and go (evaluate() builds one per call): the library's device scratch must reach a steady state
    (no growth from call to call: the operand planes of one plan are handed to the next) and be given back to the driver
    when the caller asks for it (WEALY_POOL_KEEP_MB=0)."""
    we = _we()
    s = _synth().make_eval_set(20000, 512, seed=90)                 # planes: 20k x 512 x 2 B x 2 = 41 MB
    c, i, z = s["c"].cuda(), s["i"].cuda(), s["z"].cuda()
    import gc
    gc.collect()                                                    # plans other tests left to the garbage collector
    torch.cuda.synchronize()
    base = we.pool_stats()["used"]
    slack = 32 << 20                                                # the driver reserves in whole 2 MB pages
    stats = []
    for _ in range(6):
        we.evaluate(c, i, z, c, i, z)
        torch.cuda.synchronize()
        stats.append(we.pool_stats())
    assert all(st["used"] == base for st in stats)                  # nothing is held once the plan is gone
    # steady state after the second call (a few MB of slack: the driver's pool reserves in whole pages)
    assert all(st["reserved"] <= stats[1]["reserved"] + (16 << 20) for st in stats[2:])
    assert stats[-1]["reserved_high"] <= stats[1]["reserved_high"] + (16 << 20)
    a, r = we.evaluate(c, i, z, c, i, z)
    we.release_scratch()
    after = we.pool_stats()
    assert after["used"] == base and after["reserved"] <= base + slack   # idle scratch released
    plan = we.EvalPlan(c, i, c, i)
    plan.run(z, z)
    we.release_scratch()                                            # a live plan keeps what it holds
    assert we.pool_stats()["used"] >= base + (40 << 20)
    res = plan.run(z, z)
    assert torch.equal(res["aps"], a) and torch.equal(res["r1s"], r)
    plan.close()
    monkeypatch.setenv("WEALY_POOL_KEEP_MB", "0")                   # no caching at all: every destroy returns its memory
    a2, r2 = we.evaluate(c, i, z, c, i, z)
    assert torch.equal(a, a2) and torch.equal(r, r2)
    we.release_scratch()
    assert we.pool_stats()["reserved"] <= base + slack
