"""Synthetic inputs shaped like the reference's evaluation / training sets.

No dataset or checkpoint can be downloaded, so every benchmark and parity test runs on
synthetic embeddings whose clique structure is bootstrapped from the clique-size
multisets of the reference's shipped split files (clique_sizes.json, extracted by
make_clique_sizes.py):

  * clique ids: dense ints (lib/embedding_dataset/base_dataset.py:178-189 convention);
  * version ids: the reference's 31-bit md5-derived id of "<clique>-<version>"
    (lib/embedding_dataset/utils.py:7-13), so real id collisions occur at scale;
  * embeddings: z = (centroid[clique] + sigma_item * noise) * lognormal row norm, fp32,
    with per-item sigma spread so that MAP lands mid-range (non-saturated);
  * rows shuffled: no clique-contiguity assumption.
"""
import hashlib
import json
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def clique_size_multiset(dist="shs100k_test"):
    with open(os.path.join(_HERE, "clique_sizes.json")) as f:
        hist = json.load(f)[dist]
    sizes = []
    for s, c in hist.items():
        sizes.extend([int(s)] * int(c))
    return np.asarray(sorted(sizes), dtype=np.int64)


def sample_clique_sizes(n, dist="shs100k_test", seed=0):
    """Clique sizes summing to exactly n: the exact multiset when n matches the split's size,
    otherwise a bootstrap resample (every clique keeps >= 2 versions,
    lib/embedding_dataset/filters.py:87-109)."""
    base = clique_size_multiset(dist)
    if int(base.sum()) == n:
        return base.copy()
    rng = np.random.default_rng(seed)
    out, total = [], 0
    while total < n:
        draw = rng.choice(base, size=max(16, (n - total) // max(1, int(base.mean()))))
        for s in draw:
            if total >= n:
                break
            s = int(min(s, n - total))
            if s == 1:                       # never leave a singleton: grow the previous clique
                if out:
                    out[-1] += 1
                else:
                    out.append(1)
                total += 1
                break
            out.append(s)
            total += s
    sizes = np.asarray(out, dtype=np.int64)
    if sizes.size and sizes[-1] < 2 and sizes.size > 1:
        sizes[-2] += sizes[-1]
        sizes = sizes[:-1]
    assert int(sizes.sum()) == n
    return sizes


def deterministic_song_id(clique, version):
    """Same arithmetic as lib/embedding_dataset/utils.py:7-13 (md5 -> first 4 bytes -> 31 bits)."""
    digest = hashlib.md5(f"{clique}-{version}".encode("utf-8")).digest()
    return int.from_bytes(digest[:4], byteorder="big") & 0x7FFFFFFF


def make_ids(n, dist="shs100k_test", seed=0, md5_ids=True):
    """-> (clique_ids[n], version_ids[n]) int64 numpy arrays, rows shuffled."""
    sizes = sample_clique_sizes(n, dist, seed)
    cliques = np.repeat(np.arange(sizes.size, dtype=np.int64), sizes)
    within = np.concatenate([np.arange(s, dtype=np.int64) for s in sizes])
    if md5_ids:
        vers = np.fromiter((deterministic_song_id(int(c), int(v)) for c, v in zip(cliques, within)),
                           dtype=np.int64, count=n)
    else:
        vers = np.arange(n, dtype=np.int64)
    perm = np.random.default_rng(seed + 1).permutation(n)
    return cliques[perm], vers[perm]


def make_eval_set(n, d, seed=0, dist="shs100k_test", sigma=2.4, sigma_spread=0.35,
                  device="cpu", md5_ids=True):
    """-> dict(z=[n,d] fp32, c=[n] int64, i=[n] int64) on `device`."""
    c_np, i_np = make_ids(n, dist, seed, md5_ids)
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    c = torch.from_numpy(c_np).to(dev)
    n_cliques = int(c_np.max()) + 1
    centroids = torch.randn(n_cliques, d, generator=g, device=dev)
    item_sigma = sigma * torch.exp(sigma_spread * torch.randn(n, 1, generator=g, device=dev))
    z = centroids[c] + item_sigma * torch.randn(n, d, generator=g, device=dev)
    z = z / z.norm(dim=1, keepdim=True)
    z = z * torch.exp(0.25 * torch.randn(n, 1, generator=g, device=dev)) * 7.0
    return {"z": z.float().contiguous(), "c": c, "i": torch.from_numpy(i_np).to(dev)}


def make_loss_batch(b, d, seed=0, per_clique=4, dtype=torch.float32, device="cpu", dup_idx=2):
    """Training-batch shaped input: z ~ N(0,1) [b,d], labels = `per_clique` items per clique
    (n_per_class batches, lib/embedding_dataset/base_dataset.py:281-289), z_idx = arange with a
    few duplicates (two augmentations of the same sample)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    z = torch.randn(b, d, generator=g, device=dev).to(dtype)
    label = (torch.arange(b, device=dev) // per_clique).long()
    idx = torch.arange(b, device=dev).long()
    for k in range(dup_idx):
        j = 2 + k * per_clique
        if j + 1 < b:
            idx[j + 1] = idx[j]
    perm = torch.randperm(b, generator=g, device=dev)
    return {"z": z[perm].contiguous(), "label": label[perm].contiguous(), "idx": idx[perm].contiguous()}
