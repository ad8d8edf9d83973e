#!/usr/bin/env python
"""GPU probe: which cumsum association does torch.cumsum use on this box, and does the EXACT
knot chain reproduce the ATen-on-CUDA chain bit for bit?  Prints one line per scan order.
Run on the GPU box: python tools/probe_exact.py
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from normalizingflow_b200 import _lib, _ops  # noqa: E402


def aten_knots(logits, B, layer_norm):
    v = logits
    if layer_norm:
        v = 2 * B * torch.softmax(v, dim=-1)
    K = v.shape[-1]
    p = F.softmax(v, dim=-1)
    p = 1e-3 + (1 - 1e-3 * K) * p
    c = torch.cumsum(p, dim=-1)
    c = F.pad(c, pad=(1, 0), mode="constant", value=0.0)
    c = (B - (-B)) * c + (-B)
    c[..., 0] = -B
    c[..., -1] = B
    return c


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    for K in (8, 5, 32):
        logits = (torch.randn(1 << 18, K, generator=g) * 2).to(dev)
        for B in (3.0, 4.0):
            for ln in (True, False):
                ref = aten_knots(logits, B, ln)
                res = []
                for order in (0, 1, 2):
                    _lib.lib.nfk_set_scan_order(order)
                    mine = _ops.debug_knots(logits, B, ln, True)
                    res.append(int((mine.view(torch.int32) != ref.view(torch.int32)).sum()))
                fast = _ops.debug_knots(logits, B, ln, False)
                print(f"K={K} B={B} layer_norm={ln}: bit mismatches by scan order 0/1/2 = {res} of {ref.numel()};"
                      f" fast max abs err {float((fast - ref).abs().max()):.3e}")
    # stage-wise: softmax alone, cumsum alone
    x = (torch.randn(1 << 18, 8, generator=g) * 2).to(dev)
    sm = torch.softmax(x, -1)
    e = torch.exp(x - x.max(-1, keepdim=True).values)
    s_butter = ((e[:, 0] + e[:, 4]) + (e[:, 2] + e[:, 6])) + ((e[:, 1] + e[:, 5]) + (e[:, 3] + e[:, 7]))
    print("softmax == exp/butterfly-sum bitwise:", bool(((e / s_butter[:, None]) == sm).all()))
    c = torch.cumsum(sm, -1)
    seq = sm.clone()
    for j in range(1, 8):
        seq[:, j] = seq[:, j] + seq[:, j - 1]
    print("cumsum == sequential bitwise:", bool((c == seq).all()), " mismatches:", int((c != seq).sum()))


if __name__ == "__main__":
    main()
