"""Parameter containers that reproduce the reference's ``state_dict`` key layout without its arithmetic.

The modules here hold ``nn.Parameter``s under exactly the names the reference registers (see
gst_visdial_b200/weights.py) so that ``load_state_dict`` / ``.to(device)`` / ``nn.DataParallel`` / attribute swaps such as
``decoder.decoder.bert.embeddings = encoder.bert_pretrained.bert.embeddings`` (generate.py:65) keep working.  All
computation is delegated to the CUDA engine; these modules have no ``forward``.
"""
from __future__ import annotations

import re
from typing import Mapping

import torch
from torch import nn


class ParamNode(nn.Module):
    """A node of the parameter tree: children and parameters are registered under the reference's names."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("ParamNode holds weights only; call the owning VisualDialog* module")


def build_tree(spec: Mapping[str, tuple], prefix: str) -> ParamNode:
    """Builds the sub-tree of ``spec`` (flat dotted names) under ``prefix`` (e.g. 'encoder.bert_pretrained.')."""
    root = ParamNode()
    for name, shape in spec.items():
        if not name.startswith(prefix):
            continue
        parts = name[len(prefix):].split(".")
        node = root
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, ParamNode())
            node = node._modules[p]
        leaf = parts[-1]
        if re.search(r"LayerNorm[12]?\.weight$", name):
            t = torch.ones(shape)
        else:
            t = torch.zeros(shape)
        node.register_parameter(leaf, nn.Parameter(t, requires_grad=False))
    return root


class WeightVersion:
    """Shared, mutable change counter: bumped whenever parameters may have changed (load_state_dict, .to(), .float())."""

    def __init__(self):
        self.value = 1

    def bump(self):
        self.value += 1
