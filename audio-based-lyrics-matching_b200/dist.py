"""Multi-GPU evaluation: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on
the GPU box; gloo in the CPU tests of the merge logic).

Two schemes:

* all-vs-all (queries ARE the corpus, the headline case) -- `evaluate_all_vs_all`: every rank holds the whole
  corpus and sweeps the row blocks rb = rank (mod world) of the SAME symmetric problem (only tiles above the
  diagonal, each element scoring its row and its column query).  A rank therefore holds partial rank counts for
  ALL queries; the one real exchange step of the path is an all-reduce (SUM) of those int32 counters -- the
  "rank counts merged over NCCL" of the north star -- after which every rank owns the complete AP / R1.
* general queries vs corpus -- `evaluate_sharded`: queries are independent units, so rank r scores the contiguous
  query slice shard_range(Nq, r, world) against the whole (replicated) corpus with the single-GPU fused kernel and
  the data path needs no collective.  The only exchange is the result merge: one all-reduce of {sum AP, sum R1,
  count} (24 bytes) for MAP / MR1 and, when per-query values are wanted, one all-gather of [Nq/world] x {ap, r1}
  (and of the top-k lists).

Errors are raised on EVERY rank or on none: whatever one rank finds wrong with its share (a query without relevant
candidates) is agreed on with a tiny all-reduce before any rank enters a data collective.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced partition of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_sums(local_sums, group=None):
    """All-reduce {sum AP, sum R1, count} -> (MAP, MR1, count) identical on every rank."""
    s = local_sums.detach().clone().double()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    s = s.cpu()
    n = max(float(s[2]), 1.0)
    return float(s[0]) / n, float(s[1]) / n, int(s[2])


def gather_rows(local, n_total, group=None):
    """All-gather row-sharded results (shard_range layout) into the full [n_total, ...] tensor."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    # (gloo has no all_gather of CUDA tensors: results are staged through the host there -- CPU tests, and two
    #  processes sharing one GPU)
    via_host = local.is_cuda and dist.get_backend(group) == "gloo"
    src = local.cpu() if via_host else local
    pad = torch.zeros((longest,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    pad[: src.shape[0]] = src
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    full = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
    return full.to(local.device) if via_host else full


def upload_sharded(z_host, device, group=None):
    """Host embeddings -> the full [n, d] matrix on every rank's device without every rank pulling the whole corpus
    over PCIe: rank r uploads only its contiguous 1/world slice of the rows (pinned host memory -> HBM), then ONE
    all-gather over NVLink / NVSwitch replicates the corpus.  Host traffic per rank drops from n*d to n*d/world
    bytes (8 ranks x 1.16 GB through one host bridge at 282k x 1024 was 2/3 of the end-to-end step)."""
    on = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if on else 1
    rank = dist.get_rank(group) if on else 0
    n, d = z_host.shape
    if world == 1:
        return z_host.to(device, non_blocking=True)
    chunk = -(-n // world)
    lo, hi = min(rank * chunk, n), min((rank + 1) * chunk, n)
    mine = torch.zeros((chunk, d), dtype=z_host.dtype, device=device) if hi - lo < chunk else \
        torch.empty((chunk, d), dtype=z_host.dtype, device=device)
    if hi > lo:
        mine[: hi - lo].copy_(z_host[lo:hi], non_blocking=True)
    full = torch.empty((world * chunk, d), dtype=z_host.dtype, device=device)
    try:
        dist.all_gather_into_tensor(full, mine, group=group)
    except (RuntimeError, NotImplementedError):      # backends without the flat variant (CPU tests)
        dist.all_gather(list(full.view(world, chunk, d).unbind(0)), mine, group=group)
    return full[:n]


def _agree(flag, device, group=None):
    """MAX of a small non-negative integer over the ranks: every rank learns whether ANY rank wants to raise."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(flag)
    t = torch.tensor([int(flag)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


STAGED_MIN_WORLD = 4


def sharded_step(plan, z, rank, world, *, eps=1e-6, precision=None, group=None):
    """The multi-GPU all-vs-all sweep up to (not including) finish(): the relevant similarities are computed once across
    the ranks (each rank its share of the queries; one all-reduce of floats assembles them -- every element has a single
    non-zero contribution, so the sum is exact), every rank sweeps the row blocks rank (mod world) of the symmetric
    problem, and the per-(query, relevant item) rank counters are summed (the second all-reduce)."""
    if world >= STAGED_MIN_WORLD:
        plan.shard_prepare(z, rank, world, eps=eps, precision=precision)
        dist.all_reduce(plan.thresholds_tensor(), op=dist.ReduceOp.SUM, group=group)
        plan.shard_sweep(rank, world)
    else:
        # two or three ranks: recomputing the relevant similarities on every rank (0.5 ms at 141k tracks) costs less
        # than the extra collective (measured at N = 2: 23.1 vs 23.5 ms per step)
        plan.sweep_shard(z, rank, world, eps=eps, precision=precision)
    dist.all_reduce(plan.counts_tensor(), op=dist.ReduceOp.SUM, group=group)


def evaluate_all_vs_all(c, i, z, *, precision=None, eps=1e-6, group=None, plan=None, allow_empty=False, topk=None):
    """All-vs-all evaluation of one set (clique ids c, version ids i, embeddings z) over all ranks of `group`.
    Every rank passes the full tensors (host embeddings are uploaded 1/world per rank and all-gathered over
    NVLink).  -> dict(map, mr1, count, aps, r1s, plan); identical on every rank.

    topk: per-query top-k lists are per-query state, not additive counters, so with more than one rank they take the
    query-partitioned scheme (`evaluate_sharded`: every rank scores its slice of the queries against the replicated
    corpus with the streaming top-k, the lists are all-gathered) -> also topk_idx / topk_sim.

    Like EvalPlan.run, a query without any relevant candidate raises ValueError unless allow_empty -- on every
    rank (all ranks build the same id plan), before the first collective."""
    from .evaluation import EvalPlan
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    if topk and world > 1:
        if not torch.as_tensor(z).is_cuda:
            z = upload_sharded(torch.as_tensor(z), torch.device("cuda", torch.cuda.current_device()), group)
        return evaluate_sharded(c, i, z, c, i, z, topk=topk, precision=precision, eps=eps, group=group,
                                allow_empty=allow_empty)
    side = None
    if world > 1 and not torch.as_tensor(z).is_cuda:
        # 1/world of the rows per rank + NVLink all-gather, issued on a side stream BEFORE the id plan is built: the
        # copy engine and NCCL move the embeddings while the plan's small kernels and host read-backs run
        device = plan.device if plan is not None else torch.device("cuda", torch.cuda.current_device())
        from .evaluation import side_stream
        side = side_stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            z = upload_sharded(torch.as_tensor(z), device, group)
    if plan is None:
        plan = EvalPlan(c, i, c, i)
    if side is not None:
        torch.cuda.current_stream(plan.device).wait_stream(side)
        z.record_stream(torch.cuda.current_stream(plan.device))
    if plan.queries_without_relevant and not allow_empty:
        raise ValueError(f"{plan.queries_without_relevant} queries have no relevant candidate "
                         "(every clique needs >= 2 versions; pass allow_empty=True to score the rest)")
    if world == 1:
        res = plan.run(z, z, eps=eps, precision=precision, allow_empty=allow_empty, topk=topk)
    else:
        sharded_step(plan, z, rank, world, eps=eps, precision=precision, group=group)
        res = plan.finish()
    s = res["sums"].cpu()                                            # one 24-byte device -> host read
    n = max(float(s[2]), 1.0)
    out = {"map": float(s[0]) / n, "mr1": float(s[1]) / n, "count": int(s[2]), "aps": res["aps"], "r1s": res["r1s"],
           "plan": plan}
    if topk and "topk_idx" in res:
        out["topk_idx"], out["topk_sim"] = res["topk_idx"], res["topk_sim"]
    return out


def evaluate_sharded(queries_c, queries_i, queries_z, candidates_c, candidates_i, candidates_z, *, topk=None,
                     precision=None, eps=1e-6, gather=True, group=None, plan=None, allow_empty=False):
    """Every rank passes the FULL query / candidate tensors (or its own copy of them); rank r
    scores its slice.  Returns dict(map, mr1, count[, aps, r1s, topk_idx, topk_sim]) on every rank.

    A rank whose slice is empty (world > Nq) contributes nothing; a query without relevant candidates raises
    ValueError on EVERY rank (agreed on before the merge collectives) unless allow_empty."""
    from .evaluation import EvalPlan
    on = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if on else 0
    world = dist.get_world_size(group) if on else 1
    nq = len(queries_c)
    lo, hi = shard_range(nq, rank, world)
    device = plan.device if plan is not None else torch.device("cuda", torch.cuda.current_device())
    own = plan is None
    if own and hi > lo:
        plan = EvalPlan(queries_c[lo:hi], queries_i[lo:hi], candidates_c, candidates_i, device=device)
    bad = 0 if plan is None else int(plan.queries_without_relevant)
    bad = _agree(bad if not allow_empty else 0, device, group)
    if bad:
        raise ValueError("some queries have no relevant candidate (every clique needs >= 2 versions; "
                         "pass allow_empty=True to score the rest)")
    k = 0 if topk is None else min(int(topk), len(candidates_c))
    if hi > lo:
        res = plan.run(queries_z[lo:hi], candidates_z, topk=topk, eps=eps, precision=precision, allow_empty=allow_empty)
    else:                                                            # empty shard: neutral contribution
        res = {"aps": torch.empty(0, dtype=torch.float32, device=device),
               "r1s": torch.empty(0, dtype=torch.float32, device=device),
               "sums": torch.zeros(3, dtype=torch.float64, device=device)}
        if k:
            res["topk_idx"] = torch.empty((0, k), dtype=torch.long, device=device)
            res["topk_sim"] = torch.empty((0, k), dtype=torch.float32, device=device)
    m, r1, cnt = merge_sums(res["sums"], group)
    out = {"map": m, "mr1": r1, "count": cnt, "plan": plan}
    if gather:
        out["aps"] = gather_rows(res["aps"], nq, group)
        out["r1s"] = gather_rows(res["r1s"], nq, group)
        if topk:
            out["topk_idx"] = gather_rows(res["topk_idx"], nq, group)
            out["topk_sim"] = gather_rows(res["topk_sim"], nq, group)
    return out
