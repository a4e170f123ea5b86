"""In-tree build of lib/libwealy_b200.so for sm_100a (nvcc cross-compiles without a GPU).

Staleness is decided by a content hash of the sources (stored next to the library), not by mtimes:
the library travels to the GPU box inside a snapshot whose timestamps are not preserved."""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(PKG, "csrc")
HDR = os.path.join(PKG, "..", "include", "wealy_b200.h")
OUT = os.path.join(PKG, "lib", "libwealy_b200.so")
STAMP = OUT + ".srchash"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"] + os.environ.get("WEALY_NVCC_EXTRA", "").split()


def source_hash():
    h = hashlib.sha256()
    for path in sorted(os.path.join(SRC, f) for f in os.listdir(SRC)) + [HDR]:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _stale():
    if not (os.path.isfile(OUT) and os.path.isfile(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, os.path.join(SRC, "api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libwealy_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
