"""Extract the clique-size multisets of the reference's shipped split files into
clique_sizes.json (run once in the build container; the JSON is committed because
/root/reference does not exist on the GPU box).

Sources (data, not code): /root/reference/datasets/shs/SHS100K-{TRAIN,VAL,TEST}
("set_id<TAB>ver_id" per line) and /root/reference/datasets/lyric-covers/*_no_dup.csv
(clique = `original_id` column).  Only the histogram {clique size: number of cliques}
is kept -- it drives the synthetic generator in synth.py.
"""
import collections
import csv
import json
import os

ROOT = os.environ.get("WEALY_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def shs(name):
    cnt = collections.Counter()
    with open(os.path.join(ROOT, "datasets", "shs", name)) as f:
        for line in f:
            parts = line.split()
            if len(parts) >= 2:
                cnt[parts[0]] += 1
    return cnt


def lyric(name):
    cnt = collections.Counter()
    with open(os.path.join(ROOT, "datasets", "lyric-covers", name)) as f:
        for row in csv.DictReader(f):
            cnt[row["original_id"]] += 1
    return cnt


def hist(cnt):
    h = collections.Counter(cnt.values())
    return {str(k): h[k] for k in sorted(h)}


if __name__ == "__main__":
    out = {
        "shs100k_test": hist(shs("SHS100K-TEST")),
        "shs100k_train": hist(shs("SHS100K-TRAIN")),
        "shs100k_val": hist(shs("SHS100K-VAL")),
        "lyric_covers_test": hist(lyric("test_no_dup.csv")),
    }
    with open(os.path.join(HERE, "clique_sizes.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    for k, h in out.items():
        n = sum(int(s) * c for s, c in h.items())
        print(k, "versions", n, "cliques", sum(h.values()), "max", max(map(int, h)))
