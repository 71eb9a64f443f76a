"""Import shim: make the reference's own `model/face_tts.py` pick up this
implementation unmodified.

The reference does `from model import monotonic_align` (face_tts.py:12) and
calls `monotonic_align.maximum_path(log_prior, attn_mask.squeeze(1))`
(face_tts.py:173).  `install()` registers this package's drop-in under the
names the reference imports, so no reference file has to change:

    import face_gan_tts_b200; face_gan_tts_b200.install()
    from model.face_tts import FaceTTS            # reference code, untouched
"""
from __future__ import annotations

import sys
import types

_NAMES = ("model.monotonic_align", "model.monotonic_align.model",
          "model.monotonic_align.model.monotonic_align", "model.monotonic_align.model.monotonic_align.core")
_saved = {}


def install():
    from . import monotonic_align as ma

    for n in _NAMES:
        _saved.setdefault(n, sys.modules.get(n))
    sys.modules["model.monotonic_align"] = ma
    # the reference's odd nested import path (monotonic_align/__init__.py:5)
    pkg1 = types.ModuleType("model.monotonic_align.model")
    pkg2 = types.ModuleType("model.monotonic_align.model.monotonic_align")
    pkg1.monotonic_align = pkg2
    pkg2.core = ma.core
    sys.modules["model.monotonic_align.model"] = pkg1
    sys.modules["model.monotonic_align.model.monotonic_align"] = pkg2
    sys.modules["model.monotonic_align.model.monotonic_align.core"] = ma.core
    model_pkg = sys.modules.get("model")
    if model_pkg is not None:
        setattr(model_pkg, "monotonic_align", ma)
    return ma


def uninstall():
    for n in _NAMES:
        old = _saved.pop(n, None)
        if old is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = old
