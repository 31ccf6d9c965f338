"""Load the UNMODIFIED reference loss heads from /root/reference (build container only).

Test infrastructure.  The reference file
``/root/reference/cvap/module/decoder/loss_head.py`` imports two packages that
are not installed here (``fvcore`` for its ``Registry``, ``ftfy`` through the
vendored CLIP tokenizer).  Neither touches arithmetic, so two stub modules are
registered in ``sys.modules`` before the file is executed by path; the
reference source itself is not copied, patched or re-typed.

``/root/reference`` does not exist on the GPU box: everything that runs there
uses the committed vectors in ``tests/golden/`` instead (see make_golden.py).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VIPANT_REFERENCE_ROOT", "/root/reference")
_LOSS_HEAD = os.path.join(REFERENCE_ROOT, "cvap", "module", "decoder", "loss_head.py")
_CACHE = {}


def available() -> bool:
    return os.path.isfile(_LOSS_HEAD)


class _Registry:
    """Dict-backed stand-in for fvcore.common.registry.Registry (loss_head.py:3,17-23)."""

    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(cls):
                self._obj_map[cls.__name__] = cls
                return cls
            return deco
        self._obj_map[obj.__name__] = obj
        return obj

    def get(self, name):
        return self._obj_map[name]


def _install_stubs():
    if "fvcore.common.registry" not in sys.modules:
        fv = types.ModuleType("fvcore")
        fvc = types.ModuleType("fvcore.common")
        fvr = types.ModuleType("fvcore.common.registry")
        fvr.Registry = _Registry
        fv.common = fvc
        fvc.registry = fvr
        sys.modules.update({"fvcore": fv, "fvcore.common": fvc, "fvcore.common.registry": fvr})
    if "ftfy" not in sys.modules:
        ft = types.ModuleType("ftfy")
        ft.fix_text = lambda s: s
        sys.modules["ftfy"] = ft


def load_reference_loss_head():
    """Return the executed reference module (CELossHead, LossHead, ClassificationHead, ...)."""
    if "mod" in _CACHE:
        return _CACHE["mod"]
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)  # `from clip import ...` (loss_head.py:15)
    spec = importlib.util.spec_from_file_location("_vipant_reference_loss_head", _LOSS_HEAD)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _CACHE["mod"] = mod
    return mod


class Cfg:
    """Plain attribute bag standing in for the OmegaConf node the reference reads."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
