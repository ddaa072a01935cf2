"""Mint reference CHECKPOINT fixtures from the unmodified reference (build container only):

    python tests/golden/make_ref_checkpoints.py

  ref_head_module.pt   torch.save(ArcMarginProduct(16, 32, s=30, m=0.5))   -- whole-module pickle of the reference
                       class `arcface.ArcMarginProduct` (what nlp_classifier_train.py:159 does for the full model)
  ref_model_state.pt   a DataParallel-style state_dict: {'module.backbone.weight', 'module.classifier.weight'}
  ref_multilabel_state.pt  {'classifier1.weight', 'classifier2.weight', 'classifier3.weight'} (nlp_classifier_multilabel.py:15-17)
"""
import os
import sys

import torch

sys.path.insert(0, "/root/reference")
from arcface import ArcMarginProduct  # noqa: E402  (the reference, imported verbatim)

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(0)
head = ArcMarginProduct(16, 32, s=30.0, m=0.5)
head.update_m(0.04)
torch.save(head, os.path.join(HERE, "ref_head_module.pt"))
torch.save({"module.backbone.weight": torch.randn(4, 4), "module.classifier.weight": head.weight.detach().clone()},
           os.path.join(HERE, "ref_model_state.pt"))
torch.save({"classifier%d.weight" % i: ArcMarginProduct(16, 8 * i).weight.detach().clone() for i in (1, 2, 3)},
           os.path.join(HERE, "ref_multilabel_state.pt"))
print("wrote fixtures; head.m =", head.m)
