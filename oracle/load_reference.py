"""TEST INFRASTRUCTURE ONLY -- import the reference's own hot-path source files, unmodified,
from /root/reference (build container only; the path does not exist on the GPU box).

The package `__init__` files of the reference pull in `torch_geometric.data/datasets`, sklearn
pipelines and networkx loaders that the hot path does not need, so instead of importing the
package we register *empty* package modules whose `__path__` points at the real directories
and let the import system execute only the leaf files that are asked for.  Relative imports
inside those files (`from ...utils.directed.get_magnetic_Laplacian import ...`) then resolve
to the reference's real files.

Used by `tests/golden/make_golden.py` (fixture generation) and by the `not gpu` tests that
cross-check `oracle/port.py` against the live reference when /root/reference is present.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

from . import pyg_shim

REFERENCE_ROOT = os.environ.get("PGSD_REFERENCE_ROOT", "/root/reference")
_PKG = "torch_geometric_signed_directed"

_SUBPACKAGES = [
    "", "nn", "nn.directed", "nn.signed", "nn.general",
    "utils", "utils.directed", "utils.signed", "utils.general",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, _PKG, "nn"))


def _ensure_tree() -> bool:
    """Returns True if PyG is shimmed (always the case in this image)."""
    shimmed = pyg_shim.install()
    base = os.path.join(REFERENCE_ROOT, _PKG)
    for sub in _SUBPACKAGES:
        name = _PKG + ("." + sub if sub else "")
        if name in sys.modules:
            continue
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(base, *sub.split("."))] if sub else [base]
        m.__package__ = name
        sys.modules[name] = m
        if sub:
            parent, _, leaf = name.rpartition(".")
            setattr(sys.modules[parent], leaf, m)
    return shimmed


def load(dotted: str):
    """load('nn.directed.MagNetConv') -> module object of the reference file."""
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _ensure_tree()
    return importlib.import_module(f"{_PKG}.{dotted}")


def ref_classes():
    """The reference symbols the parity suite exercises, keyed by their public names."""
    out = {}
    out["MagNetConv"] = load("nn.directed.MagNetConv").MagNetConv
    out["get_magnetic_Laplacian"] = load("utils.directed.get_magnetic_Laplacian").get_magnetic_Laplacian
    out["MSConv"] = load("nn.general.MSConv").MSConv
    out["get_magnetic_signed_Laplacian"] = load(
        "utils.general.get_magnetic_signed_Laplacian").get_magnetic_signed_Laplacian
    out["DiGCNConv"] = load("nn.directed.DiGCNConv").DiGCNConv
    out["DiGCN_InceptionBlock"] = load("nn.directed.DiGCN_Inception_Block").DiGCN_InceptionBlock
    out["SGCNConv"] = load("nn.signed.SGCNConv").SGCNConv
    cb = load("nn.general.conv_base")
    out["Conv_Base"], out["conv_norm_rw"] = cb.Conv_Base, cb.conv_norm_rw
    out["DIMPA"] = load("nn.directed.DIMPA").DIMPA
    out["SIMPA"] = load("nn.signed.SIMPA").SIMPA
    out["DGCNConv"] = load("nn.directed.DGCNConv").DGCNConv
    out["complex_relu_layer"] = load("nn.directed.complex_relu").complex_relu_layer
    out["SNEAConv"] = load("nn.signed.SNEAConv").SNEAConv
    # SDGNN.py imports losses / feature builders from utils.signed (sklearn, scipy pipelines that the
    # SDRLayer itself never touches): give the empty package module placeholder attributes
    us = sys.modules[_PKG + ".utils.signed"]
    import torch as _torch
    for name in ("create_spectral_features", "Sign_Product_Entropy_Loss", "Sign_Direction_Loss",
                 "Sign_Triangle_Loss", "Link_Sign_Product_Loss", "Link_Sign_Entropy_Loss", "Sign_Structure_Loss"):
        if getattr(us, name, None) is None:      # the model constructors instantiate their loss modules
            setattr(us, name, lambda *a, **k: _torch.nn.Identity())
    out["SDRLayer"] = load("nn.signed.SDGNN").SDRLayer
    out["MagNet_node_classification"] = load("nn.directed.MagNet_node_classification").MagNet_node_classification
    out["DiGCN_Inception_Block_node_classification"] = load(
        "nn.directed.DiGCN_Inception_Block_node_classification").DiGCN_Inception_Block_node_classification
    # SGCN.py imports the TSVD initialiser and two loss modules from utils.signed (scikit-learn pipelines
    # outside the hot path); its constructor instantiates the losses, forward() never touches them
    out["SGCN"] = load("nn.signed.SGCN").SGCN
    # SDGNN / SiGAT: the constructors instantiate their loss modules; forward() and build_adj_lists() (the parts
    # restated here) never touch them
    out["SDGNN"] = load("nn.signed.SDGNN").SDGNN
    out["SiGAT"] = load("nn.signed.SiGAT").SiGAT
    return out
