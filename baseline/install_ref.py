"""Stand-in for `pip install --target baseline/_ref /root/reference`.

The reference has no setup.py / pyproject.toml (it is a bare `lib/` directory of Python modules), so pip has nothing to
install.  This recipe does what an install would: it places the UNMODIFIED modules of the path
(`lib/tensor_ops.py`, `lib/losses.py`) under `baseline/_ref/lib/`, which is git-ignored (never part of the history) but
travels to the GPU box with the snapshot, so that `bench.py --impl reference` and the `cpu_baseline` legs can time the
reference's own `pairwise_distance_matrix` / `NTXentLoss` / `CLEWSLoss` on the box's host cores.  Runs only where
/root/reference exists (the build container); elsewhere it is a no-op and whatever was installed earlier is used.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("WEALY_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("lib/tensor_ops.py", "lib/losses.py")


def install():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return os.path.isdir(os.path.join(DST, "lib"))
    os.makedirs(os.path.join(DST, "lib"), exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    open(os.path.join(DST, "lib", "__init__.py"), "a").close()
    return True


def load():
    """-> (tensor_ops, losses) of the installed reference, or None when it is not installed.
    `pytorch_metric_learning` (imported but never used by lib/losses.py:4-5, not installed here) is stubbed."""
    import importlib.util
    import sys
    import types
    if not os.path.isfile(os.path.join(DST, FILES[0])):
        return None
    for name in ("pytorch_metric_learning", "pytorch_metric_learning.losses", "pytorch_metric_learning.miners"):
        sys.modules.setdefault(name, types.ModuleType(name))
    mods = []
    pkg = types.ModuleType("_wealy_ref_lib")
    pkg.__path__ = [os.path.join(DST, "lib")]
    sys.modules.setdefault("_wealy_ref_lib", pkg)
    for f in FILES:
        name = "_wealy_ref_lib." + os.path.basename(f)[:-3]
        if name in sys.modules:
            mods.append(sys.modules[name])
            continue
        spec = importlib.util.spec_from_file_location(name, os.path.join(DST, f))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


if __name__ == "__main__":
    print("installed" if install() else "reference not available")
