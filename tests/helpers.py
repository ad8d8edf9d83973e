"""Shared helpers for the parity tests (golden loading, tolerances, bin-mismatch classifier)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Parity tolerances (BASELINE.json north_star; SURVEY.md §8(d) "parity gates")
RTOL_FP32 = 1e-5      # |a-b| / max(1, |b|), per layer, fp32, identical inputs/params
RTOL_BF16 = 1e-2      # same measure, bf16 tensor-core conditioner


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(dtype) if dtype is not None else t


def sub_sd(npz, prefix, dtype=None):
    """state-dict slice of an npz under ``prefix`` -> {key: tensor}."""
    return {k[len(prefix):]: T(npz[k], dtype) for k in npz.files if k.startswith(prefix)}


def rel_err(a, b):
    """max |a-b| / max(1, |b|) — the measure the parity gate is defined on."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    if a.numel() == 0:
        return 0.0
    return float(((a - b).abs() / b.abs().clamp_min(1.0)).max())


def parse_masks(npz):
    return [[int(c) for c in s.split(",")] for s in npz["masks"].tolist()]
