"""No-GPU checks: the C-ABI library loads and exports every symbol the headers declare; without a
device the entry points fail loudly instead of falling back to CPU code."""
import ctypes as C
import os
import re

import pytest

from atomorph_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(amx_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run python -m atomorph_b200.build"
    L = C.CDLL(_lib.LIB_PATH)
    for header in ("amx.h", "amx_morph.h"):
        names = _declared(header)
        assert len(names) > 20
        for n in names:
            assert hasattr(L, n), "%s declared in %s but not exported" % (n, header)


def test_python_binding_lists_match_headers():
    from atomorph_b200 import morph
    assert sorted(_lib.AMX_SYMBOLS) == _declared("amx.h")
    assert sorted(morph.MORPH_SYMBOLS) == _declared("amx_morph.h")


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from atomorph_b200.engine import Engine
    with pytest.raises(_lib.AmxError):
        Engine(0)
    from atomorph_b200.morph import Morph
    m = Morph()
    m.add_pixel(0, 1, 1, 0xff102030)
    m.set_resolution(4, 4)
    assert not m.synchronize()
    assert "no CPU fallback" in m.last_error()
    assert m.get_state() == 4          # STATE_DONE so that polling callers terminate
    assert (m.get_pixels(0.0) == 0).all()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "atomorph_b200")
    for d, _, fs in os.walk(pkg):
        for f in fs:
            if f == "build.py":
                continue        # builds the checker (allowed); never imports or runs it
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(d, f), errors="ignore").read()
                for needle in ("import oracle", "from oracle", "amref", "libamoracle", "amoracle_"):
                    assert needle not in src, "%s uses the oracle (%s)" % (f, needle)
