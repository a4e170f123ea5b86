"""Oracle restatement of the batch similarity-matrix contrastive losses (TEST INFRASTRUCTURE).

Follows /root/reference/lib/losses.py:
  * NTXentLoss.forward   lines 19-73   (tau at line 15)
  * CLEWSLoss.forward    lines 210-285 (_per_anchor_mean 202-208, ctor 185-200)

Written as plain differentiable torch functions (autograd gives the reference gradient);
`ntxent_grad` / `clews_grad` are the closed-form gradients the CUDA backward implements
(SURVEY.md section 8(a5),(a6)) and are tested equal to autograd in fp64.
Pinned by tests/golden/loss_*.npz (outputs of the reference's own modules).

Reference behaviours reproduced on purpose:
  * a single-label batch MUTATES the caller's z_label in place (lines 34-35, 221-222);
  * NT-Xent masks the diagonal by POSITION (torch.eye, line 52) while positives are by id;
  * NT-Xent normalises with x/(|x|+1e-6) (tensor_ops.py:169), CLEWS with F.normalize (eps 1e-12);
  * CLEWS v_dpos / v_dneg are means over the COMPLEMENT of the pos / neg masks (lines 267-268).
"""
import torch

from .similarity import l2_scale


def _label_noise(z_label):
    if torch.unique(z_label).numel() == 1:  # losses.py:34-35 / 221-222 (in place!)
        z_label[: max(2, int(len(z_label) * 0.01))] = -1


def _masks(z_label, z_idx):
    same_label = z_label[:, None] == z_label[None, :]
    same_idx = z_idx[:, None] == z_idx[None, :]
    return same_label & ~same_idx, ~same_label  # positives, negatives


def _zstats(z):
    return z.abs().max(), z.mean(), z.std()


def ntxent(z_label, z_idx, z, temperature=0.1):
    """-> (loss, logdict) as lib/losses.py:19-73."""
    assert len(z_label) == len(z_idx) == len(z)
    _label_noise(z_label)
    pos, _ = _masks(z_label, z_idx)
    u = l2_scale(z, 1e-6)
    logits = (u @ u.t()) / temperature
    eye = torch.eye(len(z), dtype=torch.bool, device=z.device)
    logits = logits.masked_fill(eye, -1e9)
    logits = logits - logits.max(dim=1, keepdim=True)[0].detach()
    e = torch.exp(logits)
    p_sum = (e * pos.float()).sum(dim=1)
    a_sum = e.sum(dim=1)
    loss = -torch.log(p_sum / (a_sum + 1e-8) + 1e-8).mean()
    zmax, zmean, zstd = _zstats(z)
    return loss, {"l_main": loss, "v_zmax": zmax, "v_zmean": zmean, "v_zstd": zstd}


def ntxent_grad(z_label, z_idx, z, temperature=0.1):
    """Closed-form d loss / d z (what the CUDA backward computes)."""
    z = z.detach()
    B = len(z)
    pos, _ = _masks(z_label, z_idx)
    r = torch.sqrt((z * z).sum(dim=1, keepdim=True))
    u = z / (r + 1e-6)
    logits = (u @ u.t()) / temperature
    eye = torch.eye(B, dtype=torch.bool)
    logits = logits.masked_fill(eye, -1e9)
    e = torch.exp(logits - logits.max(dim=1, keepdim=True)[0])
    posf = pos.to(z.dtype)
    P = (e * posf).sum(dim=1, keepdim=True)
    A = e.sum(dim=1, keepdim=True) + 1e-8
    rho = P / A + 1e-8
    G = -e * (posf / A - P / (A * A)) / (B * rho)
    G = G.masked_fill(eye, 0.0)
    dU = (G + G.t()) @ u / temperature
    proj = (u * dU).sum(dim=1, keepdim=True)
    return (dU - u * proj * (r + 1e-6) / r) / (r + 1e-6)


def clews(z_label, z_idx, z, *, gamma=8.0, b=1.0, eps=1e-8, epsilon=1e-6, uniformity_weight=0.5,
          warmup_steps=1000, step=None, numerically_friendly=True):
    """-> (loss, logdict) as lib/losses.py:210-285.  `step` is the resolved global step
    (extra["global_step"] or module attribute) or None."""
    if z.dim() == 3:
        assert z.size(1) == 1
        z = z.squeeze(1)
    assert z.dim() == 2
    B = z.size(0)
    assert len(z_label) == len(z_idx) == B and B >= 4
    _label_noise(z_label)
    pos, neg = _masks(z_label, z_idx)
    zn = torch.nn.functional.normalize(z, p=2, dim=-1)
    d = 1.0 - zn @ zn.t()

    def anchor_mean(x, m):
        w = m.float()
        return (x * w).sum(dim=1) / w.sum(dim=1).clamp_min(eps)

    align = anchor_mean(d, pos)
    has_pos = pos.any(dim=1)
    l_align = align[has_pos].mean() if bool(has_pos.any()) else zn.sum() * 0.0
    uni = anchor_mean(torch.exp(b - gamma * d), neg)
    l_uni = torch.log1p(uni).mean() if numerically_friendly else torch.log(uni + epsilon).mean()
    uw = uniformity_weight
    if warmup_steps > 0 and step is not None:
        uw = float(min(uniformity_weight, uniformity_weight * (int(step) + 1) / warmup_steps))
    loss = l_align + uw * l_uni
    with torch.no_grad():
        n_pos, n_neg = pos.float().sum(), neg.float().sum()
        keep_notpos = (~pos).to(d.dtype)
        keep_notneg = (~neg).to(d.dtype)
        v_dpos = (d * keep_notpos).sum() / keep_notpos.sum().clamp(min=1e-7) if n_pos > 0 else torch.tensor(0.0)
        v_dneg = (d * keep_notneg).sum() / keep_notneg.sum().clamp(min=1e-7) if n_neg > 0 else torch.tensor(0.0)
    zmax, zmean, zstd = _zstats(zn)
    logdict = {
        "l_main": loss, "l_cent": l_align, "l_cont": l_uni,
        "cnt_pos_pairs": n_pos, "cnt_neg_pairs": n_neg,
        "anchors_with_pos": has_pos.float().mean(),
        "v_dpos": v_dpos, "v_dneg": v_dneg,
        "uniformity_weight": torch.tensor(uw),
        "z_max": zmax, "z_mean": zmean, "z_std": zstd,
    }
    return loss, logdict


def clews_grad(z_label, z_idx, z, *, gamma=8.0, b=1.0, eps=1e-8, epsilon=1e-6, uw=0.5,
               numerically_friendly=True):
    """Closed-form d loss / d z for CLEWS (uw = the resolved uniformity weight)."""
    z = z.detach()
    B = len(z)
    pos, neg = _masks(z_label, z_idx)
    r = torch.sqrt((z * z).sum(dim=1, keepdim=True)).clamp_min(1e-12)
    u = z / r
    d = 1.0 - u @ u.t()
    posf, negf = pos.to(z.dtype), neg.to(z.dtype)
    npos = posf.sum(dim=1, keepdim=True)
    nneg = negf.sum(dim=1, keepdim=True)
    H = (npos > 0).sum().clamp_min(1)
    X = torch.exp(b - gamma * d)
    uni = (X * negf).sum(dim=1, keepdim=True) / nneg.clamp_min(eps)
    outer = 1.0 / (1.0 + uni) if numerically_friendly else 1.0 / (uni + epsilon)
    dS = -posf / (npos.clamp_min(eps) * H) + uw * gamma * X * negf * outer / (B * nneg.clamp_min(eps))
    dU = (dS + dS.t()) @ u
    proj = (u * dU).sum(dim=1, keepdim=True)
    return (dU - u * proj) / r


def triplet_mine(z_label, z_idx):
    """lib/losses.py:140-171 restated without the loop: per anchor the first positive (same label, different idx)
    and the first negative (different label); anchors lacking either are dropped.  -> (anchors, positives, negatives)."""
    pos, neg = _masks(z_label, z_idx)
    has = pos.any(dim=1) & neg.any(dim=1)
    first_pos = pos.float().argmax(dim=1)           # argmax returns the FIRST maximal index
    first_neg = neg.float().argmax(dim=1)
    anchors = torch.nonzero(has).flatten()
    return anchors, first_pos[anchors], first_neg[anchors]


def triplet(z_label, z_idx, z, margin=0.2, p=2, eps=1e-6, swap=False, reduction="mean"):
    """-> (loss, logdict) as lib/losses.py:92-137 (torch.nn.TripletMarginLoss on the mined triplets)."""
    assert len(z_label) == len(z_idx) == len(z)
    _label_noise(z_label)                           # :107-108
    a, pp, nn_ = triplet_mine(z_label, z_idx)
    zmax, zmean, zstd = _zstats(z)
    if len(a) == 0:                                 # :113-123
        loss = torch.tensor(0.0, requires_grad=True)
        return loss, {"l_main": loss, "v_zmax": zmax, "v_zmean": zmean, "v_zstd": zstd, "n_triplets": 0}
    d_ap = torch.nn.functional.pairwise_distance(z[a], z[pp], p=p, eps=eps)
    d_an = torch.nn.functional.pairwise_distance(z[a], z[nn_], p=p, eps=eps)
    if swap:
        d_an = torch.minimum(d_an, torch.nn.functional.pairwise_distance(z[pp], z[nn_], p=p, eps=eps))
    li = torch.clamp_min(margin + d_ap - d_an, 0)
    loss = li.mean() if reduction == "mean" else (li.sum() if reduction == "sum" else li)
    return loss, {"l_main": loss, "v_zmax": zmax, "v_zmean": zmean, "v_zstd": zstd}
