"""Registers this package under the reference's import names so that user code written against
the reference (``from df.enhance import enhance, init_df``, ``from libdf import DF``,
``from df import enhance, init_df`` -- DeepFilterNet/df/__init__.py:1-6) runs unchanged."""
from __future__ import annotations

import sys
import types


def install_dropin(force: bool = False) -> None:
    from . import enhance as _enh
    from . import libdf as _libdf

    if "libdf" in sys.modules and not force and sys.modules["libdf"] is not _libdf:
        raise RuntimeError("a different `libdf` module is already imported; pass force=True to replace it")
    sys.modules["libdf"] = _libdf
    df = types.ModuleType("df")
    df.__doc__ = "deepfilternet_b200 drop-in for the reference package `df` (inference path only)"
    df.enhance = _enh.enhance
    df.init_df = _enh.init_df
    df.__version__ = "0.5.7-pre+b200"
    df.version = df.__version__
    enh_mod = types.ModuleType("df.enhance")
    for k in ("init_df", "enhance", "df_features", "get_model_basedir", "PRETRAINED_MODELS", "DEFAULT_MODEL"):
        setattr(enh_mod, k, getattr(_enh, k))
    sys.modules["df"] = df
    sys.modules["df.enhance"] = enh_mod
