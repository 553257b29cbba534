"""Shared helpers for the parity tests."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fusion_golden.json")
METHOD_NAMES = ["semantic", "sparse", "domain"]


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


def golden_lists(case):
    lists = [case["semantic"], case["sparse"]] + ([case["domain"]] if case["domain"] else [])
    weights = [case["dense_weight"], case["sparse_weight"], 0.2][: len(lists)]
    return lists, weights


def id_map(lists):
    """string ids -> dense ints in first-seen order, and back."""
    fwd = {}
    for lst in lists:
        for x in lst:
            fwd.setdefault(x, len(fwd))
    back = {v: k for k, v in fwd.items()}
    return fwd, back


def to_bits_view(t):
    """torch 16-bit tensor -> numpy uint16 bit patterns."""
    import torch
    return t.view(torch.int16).cpu().numpy().view(np.uint16)
