"""The reference's own CPU MAS (`b_mas`, numba) for the timed baseline -- TEST / BASELINE INFRASTRUCTURE ONLY.

`__graft_entry__.build()` stages the reference's two files, unmodified, from /root/reference into the git-ignored
oracle/_ref/ (/root/reference/tts/modules/aligner/{__init__,mas}.py); like the built libraries they travel to the GPU
box with the working tree, where /root/reference itself does not exist.  Nothing here is imported by the product.

    b_mas, why = load()        # the numba function, or (None, reason) -- bench.py then falls back to the C port and says so
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = os.environ.get("ISP_REFERENCE_ROOT", "/root/reference")
_FILES = ["tts/__init__.py", "tts/modules/__init__.py", "tts/modules/aligner/__init__.py", "tts/modules/aligner/mas.py"]


def stage() -> bool:
    """Copy the reference's files into oracle/_ref/ (build container only).  Returns False when there is no reference here."""
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, _FILES[-1])):
        return False
    for rel in _FILES:
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFERENCE_ROOT, rel), dst)
    return True


def load():
    """-> (b_mas, None) or (None, reason)."""
    if not os.path.isfile(os.path.join(REF_DIR, _FILES[-1])):
        return None, "oracle/_ref is empty (run __graft_entry__.build() where /root/reference exists)"
    try:
        import numba  # noqa: F401
    except Exception as exc:                               # pragma: no cover - depends on the box
        return None, f"numba is not importable here: {exc!r}"
    saved = {k: v for k, v in sys.modules.items() if k == "tts" or k.startswith("tts.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_DIR)
    try:
        mod = importlib.import_module("tts.modules.aligner")
        fn = mod.b_mas
    except Exception as exc:                               # pragma: no cover
        return None, f"importing the staged reference failed: {exc!r}"
    finally:
        sys.path.remove(REF_DIR)
        for k in [k for k in sys.modules if k == "tts" or k.startswith("tts.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return fn, None
