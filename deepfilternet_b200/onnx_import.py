"""Weights of DeepFilterNet3_ll ship only as ONNX (enc.onnx / erb_dec.onnx / df_dec.onnx inside
models/DeepFilterNet3_ll_onnx.tar.gz).  This module transplants the ONNX initialisers into a
reference-style ``state_dict`` with a minimal protobuf wire reader (the
`onnx` package is not installed).  Implemented in a later step of round 1."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from .config import ModelConfig


def state_dict_from_onnx_dir(model_dir: str, cfg: ModelConfig) -> Optional[Dict[str, torch.Tensor]]:
    if not os.path.isfile(os.path.join(model_dir, "enc.onnx")):
        return None
    raise NotImplementedError("ONNX -> state_dict transplant is not built yet")
