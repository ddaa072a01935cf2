"""Reference checkpoint compatibility.

The reference saves either whole modules -- `torch.save(model, path)` (nlp_classifier_train.py:159,
cv_classifier_train_daodian.py:298-306), unpickled later by `torch.load` with `arcface.ArcMarginProduct` on the
import path (daodian_infer.py:354-355, goodssku_emb.py:174-175) -- or `state_dict()`s whose head entry is
`classifier.weight` (`module.classifier.weight` under nn.DataParallel, `classifier{1,2,3}.weight` for the
multi-label model).  Host logic only: nothing here launches a kernel.
"""
from __future__ import annotations

import sys
import types
from typing import Dict, Optional

import torch


def install_reference_shim() -> types.ModuleType:
    """Make `arcface.ArcMarginProduct` resolve to the B200 head, so `torch.load(path, weights_only=False)` of a
    module pickled by the reference rebuilds it around this package's class.  (Pickle restores `__dict__`
    without calling `__init__`; attributes the reference never had -- `validate_labels`, `use_cuda_graph` --
    fall back to the class defaults.)  Returns the shim module; a real `arcface` module already imported is
    left alone and returned unchanged."""
    from .head import ArcMarginProduct

    mod = sys.modules.get("arcface")
    if mod is not None and getattr(mod, "ArcMarginProduct", None) is not None and \
            not getattr(mod, "__b200_shim__", False):
        return mod
    mod = types.ModuleType("arcface")
    mod.__doc__ = "shim: arcface.ArcMarginProduct -> multimodalsimilar_b200.ArcMarginProduct"
    mod.ArcMarginProduct = ArcMarginProduct
    mod.__b200_shim__ = True
    sys.modules["arcface"] = mod
    return mod


def head_weights(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """{head name: [C, D] weight} of every ArcFace head in a reference state_dict: keys `<head>.weight` whose
    head is `classifier`, `classifier1` ... ; `module.` prefixes (nn.DataParallel) are dropped."""
    out = {}
    for key, value in state_dict.items():
        k = key
        while k.startswith("module."):
            k = k[len("module."):]
        parts = k.split(".")
        if len(parts) == 2 and parts[1] == "weight" and parts[0].startswith("classifier") and value.dim() == 2:
            out[parts[0]] = value
    return out


def load_reference_head(head, source, name: Optional[str] = None) -> None:
    """Load a reference head weight into `head` (ArcMarginProduct or ShardedArcMarginProduct).

    `source`: a [C, D] tensor, a reference `state_dict`, or a reference module / model (anything with
    `state_dict()`).  `name` picks the head when several are present (default: `classifier`, else the only one)."""
    if hasattr(source, "state_dict") and not isinstance(source, dict):
        sd = source.state_dict()
        if "weight" in sd and sd["weight"].dim() == 2 and len(sd) == 1:
            source = sd["weight"]        # a bare ArcMarginProduct
        else:
            source = sd
    if isinstance(source, dict):
        if "weight" in source and len(source) == 1:
            weight = source["weight"]
        else:
            heads = head_weights(source)
            if not heads:
                raise KeyError("no `classifier*.weight` entry in the state_dict")
            if name is None:
                name = "classifier" if "classifier" in heads else (next(iter(heads)) if len(heads) == 1 else None)
            if name is None or name not in heads:
                raise KeyError("several heads in the state_dict (%s): pass name=" % ", ".join(sorted(heads)))
            weight = heads[name]
    else:
        weight = source
    weight = weight.detach()
    if tuple(weight.shape) != (head.out_feature, head.in_feature):
        raise ValueError("checkpoint head is %s, this head is %s" % (tuple(weight.shape),
                                                                     (head.out_feature, head.in_feature)))
    if hasattr(head, "load_full_weight"):      # class-sharded: keep this rank's rows
        head.load_full_weight(weight)
    else:
        with torch.no_grad():
            head.weight.copy_(weight.to(head.weight.device, head.weight.dtype))
        if hasattr(head, "invalidate_weight_cache"):
            head.invalidate_weight_cache()


def reference_state_dict(head, prefix: str = "classifier.") -> Dict[str, torch.Tensor]:
    """The head's entry of a reference-layout state_dict: {prefix + 'weight': full [C, D] fp32 weight on the CPU}.
    On a class-sharded head every rank must call this (it all-gathers the shards)."""
    w = head.gather_weight() if hasattr(head, "gather_weight") else head.weight.detach()
    return {prefix + "weight": w.detach().to("cpu", torch.float32).clone()}
