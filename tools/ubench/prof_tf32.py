import os, sys, torch
sys.path.insert(0, '.')
from normalizingflow_b200 import _ops
dev = torch.device("cuda:0")
x = torch.randn(1 << 20, 128, device=dev); w = torch.randn(736, 128, device=dev) / 11.3; b = torch.randn(736, device=dev)
for _ in range(3): _ops.linear_f32(x, w, b, 0)
torch.cuda.synchronize()
