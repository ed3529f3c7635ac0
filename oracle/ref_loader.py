"""Loads the UNMODIFIED reference from /root/reference -- build container only.

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so
nothing that runs there (gpu tests, smoke(), bench.py) may import this module;
it is used by oracle/gen_golden.py and by the optional `-m "not gpu"` tests that
re-validate the oracle against the live reference when it is present.

  b_mas            <- tts/modules/aligner/mas.py:30 (numba; imported as is)
  alignment module <- tts/models/acoustic/modules/alignment.py, loaded by file
                      path because `import tts.models` drags in phonemizer,
                      matplotlib, ... which this image lacks.
                      A 15-line stand-in for the absent `omegaconf` package is
                      written to a temp dir; it is never part of the product.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile

REFERENCE_ROOT = os.environ.get("ISP_REFERENCE_ROOT", "/root/reference")

_OMEGACONF_STUB = '''
MISSING = "???"
class DictConfig(dict):
    def _get_flag(self, name): return False
class ListConfig(list): pass
class OmegaConf:
    merge = staticmethod(lambda *a: {k: v for d in a for k, v in d.items()})
    set_readonly = staticmethod(lambda *a, **k: None)
    to_container = staticmethod(lambda x, **k: x)
    register_new_resolver = staticmethod(lambda *a, **k: None)
    create = staticmethod(lambda x: DictConfig(x))
'''


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "tts", "modules", "aligner", "mas.py"))


def _ensure_path():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_b_mas():
    """The reference's numba `b_mas` (first call JIT-compiles, ~20 s)."""
    _ensure_path()
    from tts.modules.aligner import b_mas  # type: ignore
    return b_mas


def load_mas_width1():
    _ensure_path()
    from tts.modules.aligner.mas import mas_width1  # type: ignore
    return mas_width1


_alignment = None


def load_alignment():
    """The reference's alignment.py as a module (Aligner, ConvAttention, ...)."""
    global _alignment
    if _alignment is not None:
        return _alignment
    _ensure_path()
    try:
        import omegaconf  # noqa: F401
    except ImportError:
        d = tempfile.mkdtemp(prefix="omegaconf_stub_")
        os.makedirs(os.path.join(d, "omegaconf"))
        with open(os.path.join(d, "omegaconf", "__init__.py"), "w") as f:
            f.write(_OMEGACONF_STUB)
        sys.path.insert(0, d)
    path = os.path.join(REFERENCE_ROOT, "tts", "models", "acoustic", "modules", "alignment.py")
    spec = importlib.util.spec_from_file_location("ref_alignment", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_alignment"] = mod
    spec.loader.exec_module(mod)
    _alignment = mod
    return mod


_temporal = None


def load_temporal_adaptor():
    """The reference's temporal_adaptor.py as a module (LengthRegulator, TemporalAverager, generate_soft_path), loaded by
    file path like alignment.py (its own imports -- tts.modules.transformer, tts.utils -- resolve in this image)."""
    global _temporal
    if _temporal is not None:
        return _temporal
    load_alignment()                                    # installs the omegaconf stand-in and sys.path
    path = os.path.join(REFERENCE_ROOT, "tts", "models", "acoustic", "modules", "temporal_adaptor.py")
    spec = importlib.util.spec_from_file_location("ref_temporal_adaptor", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_temporal_adaptor"] = mod
    spec.loader.exec_module(mod)
    _temporal = mod
    return mod
