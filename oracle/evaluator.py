"""Oracle for the retrieval evaluator: self / same-clique masking, per-query ranking,
AP / MAP / MR1 and top-k (TEST INFRASTRUCTURE).

**Parity unpinned**: the reference has not released its evaluator (no argsort / AP / MR1
code under /root/reference/lib, SURVEY.md section 8(a7),(c)).  This restatement follows
what the reference does pin down:
  * positives / self by *id equality*, never by position  -- lib/losses.py:40-42, 225-228
  * distance = pairwise_distance_matrix(q, cands, mode="cos") -- lib/tensor_ops.py:167-173
  * argument names and order (queries_c, queries_i, candidates_c, candidates_i)
                                                   -- lib/audio_dataset/dataset.py:82-86, 448-449
and the standard definitions stated in SURVEY.md section 8(a7):
  self stays in the candidate list at +inf distance; candidates are ranked by ascending
  distance; AP is normalised by the number of relevant items; R1 is 1-based.

Two forms are provided and tested equal on tie-free data:
  evaluate_argsort    -- per-query sort (the textbook definition)
  evaluate_rankcount  -- rank(p) = 1 + #{j != self : s_qj > s_qp}   (what the CUDA path uses)
"""
import numpy as np
import torch

from .similarity import distance_matrix


def _as_long(t):
    return torch.as_tensor(t).long().reshape(-1)


def _sim_block(qz, cz, mode, redux=None, q_len=None, c_len=None, dist_fn=None):
    """similarity (higher = closer) of a block of queries against all candidates.

    Chunked tracks (qz [b, s1, D], cz [nc, s2, D]): the chunk-level cosine DISTANCES (b, nc, s1, s2) are reduced with
    the restated distance_tensor_redux (lib/tensor_ops.py:288-373) and returned as 1 - distance.  Ragged tracks:
    q_len [b] / c_len [nc] valid chunks per track -> the redux's mask (True = excluded, lib/tensor_ops.py:186) is
    "query chunk >= q_len or candidate chunk >= c_len"."""
    if mode not in ("cos", "cossim", "dot", "dotsim"):
        raise NotImplementedError(mode)
    base = "cossim" if mode in ("cos", "cossim") else "dotsim"
    if qz.ndim == 3:
        from .masked import distance_tensor_redux
        b, s1, d = qz.shape
        nc, s2, _ = cz.shape
        dist = 1 - distance_matrix(qz.reshape(b * s1, d), cz.reshape(nc * s2, d), mode=base)
        dist = dist.reshape(b, s1, nc, s2).permute(0, 2, 1, 3)            # (b1, b2, s1, s2)
        mask = None
        if q_len is not None:
            qm = torch.arange(s1)[None, :] >= torch.as_tensor(q_len).long()[:, None]      # (b, s1)
            cm = torch.arange(s2)[None, :] >= torch.as_tensor(c_len).long()[:, None]      # (nc, s2)
            mask = qm[:, None, :, None] | cm[None, :, None, :]
        return 1 - distance_tensor_redux(dist, redux or "min", mask)
    if dist_fn is not None:   # the unmodified reference's pairwise_distance_matrix (baseline/_ref), when installed
        return dist_fn(qz, cz, mode=base)
    return distance_matrix(qz, cz, mode=base)


def evaluate_argsort(queries_c, queries_i, queries_z, candidates_c, candidates_i, candidates_z,
                     *, topk=None, mode="cos", block=256, redux=None, q_len=None, c_len=None, dist_fn=None):
    """Returns (aps[Nq], r1s[Nq]) and, when topk is given, (topk_idx[Nq,k], topk_sim[Nq,k]).
    `dist_fn`: similarity backend with the signature of lib/tensor_ops.py:152 (default: the restatement in
    oracle/similarity.py; bench.py passes the unmodified reference function when baseline/_ref is installed).

    A query without any relevant candidate raises ValueError (the reference data pipeline
    guarantees >= 2 versions per clique: lib/embedding_dataset/filters.py:87-109)."""
    qc, qi, cc, ci = map(_as_long, (queries_c, queries_i, candidates_c, candidates_i))
    qz = torch.as_tensor(queries_z)
    cz = torch.as_tensor(candidates_z)
    nq, nc = qz.shape[0], cz.shape[0]
    aps = torch.empty(nq, dtype=torch.float64)
    r1s = torch.empty(nq, dtype=torch.float64)
    if topk is not None:
        k = min(int(topk), nc)
        tk_idx = torch.full((nq, k), -1, dtype=torch.long)
        tk_sim = torch.full((nq, k), float("-inf"), dtype=qz.dtype)
    for b0 in range(0, nq, block):
        sim = _sim_block(qz[b0:b0 + block], cz, mode, redux,
                         None if q_len is None else torch.as_tensor(q_len)[b0:b0 + block], c_len, dist_fn)   # (b, nc)
        dist = 1 - sim                                          # "cos"/"dot" distance
        for r in range(sim.shape[0]):
            q = b0 + r
            is_self = ci == qi[q]
            rel = (cc == qc[q]) & ~is_self
            if not bool(rel.any()):
                raise ValueError(f"query {q} has no relevant candidate")
            d = torch.where(is_self, torch.full_like(dist[r], float("inf")), dist[r])
            order = torch.argsort(d, stable=True)
            hit = rel[order]
            pos = torch.nonzero(hit).reshape(-1).double() + 1.0   # 1-based ranks of the relevant items
            kth = torch.arange(1, pos.numel() + 1, dtype=torch.float64)
            aps[q] = (kth / pos).mean()
            r1s[q] = pos[0]
            if topk is not None:
                n_valid = int(nc - is_self.sum())
                kk = min(k, n_valid)
                tk_idx[q, :kk] = order[:kk]
                tk_sim[q, :kk] = sim[r][order[:kk]]
    if topk is None:
        return aps, r1s
    return aps, r1s, tk_idx, tk_sim


def evaluate_rankcount(queries_c, queries_i, queries_z, candidates_c, candidates_i, candidates_z,
                       *, mode="cos", block=256):
    """Sort-free form: for every relevant p of query q
         rank_all(p) = 1 + #{j != self : s_qj > s_qp},  rank_rel(p) = 1 + #{p' in rel : s_qp' > s_qp}
       AP_q = mean_p rank_rel(p) / rank_all(p),  R1_q = min_p rank_all(p).
    Identical to evaluate_argsort whenever no two candidates of a query tie exactly."""
    qc, qi, cc, ci = map(_as_long, (queries_c, queries_i, candidates_c, candidates_i))
    qz = torch.as_tensor(queries_z)
    cz = torch.as_tensor(candidates_z)
    nq = qz.shape[0]
    aps = torch.empty(nq, dtype=torch.float64)
    r1s = torch.empty(nq, dtype=torch.float64)
    for b0 in range(0, nq, block):
        sim = _sim_block(qz[b0:b0 + block], cz, mode)
        for r in range(sim.shape[0]):
            q = b0 + r
            is_self = ci == qi[q]
            rel = (cc == qc[q]) & ~is_self
            if not bool(rel.any()):
                raise ValueError(f"query {q} has no relevant candidate")
            s = sim[r]
            thr = s[rel]                                        # (P,)
            others = s[~is_self]
            rank_all = 1 + (others[None, :] > thr[:, None]).sum(dim=1).double()
            rank_rel = 1 + (thr[None, :] > thr[:, None]).sum(dim=1).double()
            aps[q] = (rank_rel / rank_all).mean()
            r1s[q] = rank_all.min()
    return aps, r1s


def rank_tolerance(queries_c, queries_i, queries_z, candidates_c, candidates_i, candidates_z,
                   *, gap=1e-5, mode="cos", block=256, redux=None, q_len=None, c_len=None):
    """For the parity rule "ranks exact wherever the similarity gap exceeds `gap`": per query,
    returns (r1_lo, r1_hi): the range the rank of the best relevant item may take when every
    candidate whose similarity is within `gap` of it may fall on either side."""
    qc, qi, cc, ci = map(_as_long, (queries_c, queries_i, candidates_c, candidates_i))
    qz = torch.as_tensor(queries_z)
    cz = torch.as_tensor(candidates_z)
    nq = qz.shape[0]
    lo = torch.empty(nq, dtype=torch.float64)
    hi = torch.empty(nq, dtype=torch.float64)
    for b0 in range(0, nq, block):
        sim = _sim_block(qz[b0:b0 + block], cz, mode, redux,
                         None if q_len is None else torch.as_tensor(q_len)[b0:b0 + block], c_len).double()
        for r in range(sim.shape[0]):
            q = b0 + r
            is_self = ci == qi[q]
            rel = (cc == qc[q]) & ~is_self
            s = sim[r]
            best = s[rel].max()
            others = s[~is_self & ~rel]
            lo[q] = 1 + (others > best + gap).sum()
            hi[q] = 1 + (others > best - gap).sum()
    return lo, hi


def rank_bands(queries_c, queries_i, queries_z, candidates_c, candidates_i, candidates_z,
               *, gap=1e-5, mode="cos", block=256, redux=None, q_len=None, c_len=None):
    """The parity rule "ranks bit-exact wherever the similarity gap exceeds `gap`" for EVERY relevant item (the
    quantities AP is made of, not only R1).  Per query the relevant candidates are taken best first (descending
    similarity; lib/losses.py:40-42 defines relevant / self, lib/tensor_ops.py:167-173 the similarity) and for the
    k-th best, with similarity s_k:
        exact[k] = 1 + #{j != self : s_j > s_k}                      (rank among all non-self candidates)
        lo[k]    = 1 + #{j != self : s_j > s_k + gap}                every candidate within `gap` of s_k may fall
        hi[k]    =     #{j != self : s_j > s_k - gap}                on either side (the item itself is in that set)
    An implementation whose similarities are within gap/2 of these has its k-th best relevant rank inside [lo, hi]
    (order statistics are 1-Lipschitz under sup-norm perturbations), and equal to exact[k] where lo == hi.
    -> (offsets[nq + 1], sims, exact, lo, hi): CSR over the queries, float64 / int64 tensors."""
    qc, qi, cc, ci = map(_as_long, (queries_c, queries_i, candidates_c, candidates_i))
    qz = torch.as_tensor(queries_z)
    cz = torch.as_tensor(candidates_z)
    nq = qz.shape[0]
    offsets = [0]
    sims, exact, lo, hi = [], [], [], []
    for b0 in range(0, nq, block):
        sim = _sim_block(qz[b0:b0 + block], cz, mode, redux,
                         None if q_len is None else torch.as_tensor(q_len)[b0:b0 + block], c_len).double()
        for r in range(sim.shape[0]):
            q = b0 + r
            is_self = ci == qi[q]
            rel = (cc == qc[q]) & ~is_self
            s = sim[r]
            thr = torch.sort(s[rel], descending=True).values
            others = torch.sort(s[~is_self]).values
            n = others.numel()
            above = lambda x: n - torch.searchsorted(others, x.contiguous(), right=True)   # #{s_j > x}
            sims.append(thr)
            exact.append(1 + above(thr))
            lo.append(1 + above(thr + gap))
            hi.append(above(thr - gap))
            offsets.append(offsets[-1] + thr.numel())
    cat = lambda xs, dt: torch.cat(xs).to(dt) if xs else torch.empty(0, dtype=dt)
    return (torch.tensor(offsets, dtype=torch.long), cat(sims, torch.float64), cat(exact, torch.long),
            cat(lo, torch.long), cat(hi, torch.long))


def mean_metrics(aps, r1s):
    return float(torch.as_tensor(aps).double().mean()), float(torch.as_tensor(r1s).double().mean())
