"""Stage the UNMODIFIED reference model next to the oracle so that it can be timed on the GPU box's host cores.

    python oracle/make_ref.py            # copies /root/reference/src/{__init__,model,config,audio,mel_griffinlim}.py
                                         # to oracle/_ref/src/ (git-ignored, not gpurun-ignored: it travels to the box)

Test / measurement infrastructure only (see oracle/__init__.py): `bench.py --impl reference` and `bench.py`'s
`cpu_baseline` leg import `oracle/_ref/src/model.py` through `load_reference()` below; nothing under `b200taco/` does.
The files are byte-for-byte copies (no edits); `librosa`, which `src/audio.py` imports at module level but only uses
inside function bodies, is replaced by an empty stub module at import time.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref", "src")
FILES = ("__init__.py", "model.py", "config.py", "audio.py", "mel_griffinlim.py")


def stage(force=False):
    """Copy the reference files if /root/reference is present (build container); returns True if _ref is usable."""
    if os.path.isdir(REF_SRC):
        os.makedirs(DST, exist_ok=True)
        for f in FILES:
            dst = os.path.join(DST, f)
            if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(os.path.join(REF_SRC, f)):
                shutil.copyfile(os.path.join(REF_SRC, f), dst)
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def load_reference():
    """Import the staged reference package; returns its `src.model` module or None when it was never staged."""
    if not all(os.path.exists(os.path.join(DST, f)) for f in FILES):
        return None
    for name in ("librosa", "librosa.filters", "librosa.feature", "librosa.effects"):
        sys.modules.setdefault(name, types.ModuleType(name))
    root = os.path.dirname(DST)
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    return importlib.import_module("src.model")


if __name__ == "__main__":
    print("oracle/_ref staged:", stage(force="--force" in sys.argv))
