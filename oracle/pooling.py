"""Oracle restatement of the pooled-embedding producer (TEST INFRASTRUCTURE; never imported by the product).

Follows /root/reference:
  * lib/layers.py:6-30                                   MeanPool.forward
  * lib/embedding_dataset/collate_functions.py:131-172   the `use_avg_pooling` branch (per-track emb.mean(dim=0))
Pinned by tests/golden/pooling.npz (outputs of the reference's own MeanPool module; the collate branch is
restated from the cited lines -- the collate module itself needs omegaconf and is not importable here).
"""
import torch


def mean_pool(x, mask=None):
    """x (B, C, T), mask (B, T) bool (True = valid) -> (B, C)."""
    if mask is not None:
        xt = x.transpose(1, 2)                      # layers.py:21
        m = mask.unsqueeze(-1).float()              # :22
        return (xt * m).sum(dim=1) / (m.sum(dim=1) + 1e-8)   # :23-24
    return x.mean(dim=2)                            # :27


def avg_pool_tracks(frames, embed_dim):
    """frames: list of [T_k, E] tensors or None -> ([K, E] fp32, valid[K] bool)  (collate_functions.py:141-168)."""
    out = torch.zeros(len(frames), embed_dim)
    valid = torch.ones(len(frames), dtype=torch.bool)
    for i, emb in enumerate(frames):
        if emb is None or emb.shape[0] == 0:
            valid[i] = False                        # :158-161 (zeros, marked invalid)
        elif emb.shape[0] == 1:
            out[i] = emb[0]                         # :163-166
        else:
            out[i] = emb.float().mean(dim=0)        # :169
    return out, valid
