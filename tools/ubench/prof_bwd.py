"""Three launches of the one-launch layer backward (csrc/nsf_fused_bwd.cu) at 65,536 rows, for ncu."""
import sys
import torch
sys.path.insert(0, ".")
from normalizingflow_b200 import _fused, flows
torch.manual_seed(0)
lay = flows.NSF_CL(32, dim=2, K=8, B=3.0, hidden_dim=128, mask=[0]).cuda()
lay.psi.precision = "bf16"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = torch.randn(N, 64, device="cuda")
g = torch.randn(N, 64, device="cuda")
with torch.no_grad():
    for _ in range(3):
        _fused.layer_backward(lay, x, g, None, 1.0, False)
    _fused.run(lay, x, False)
torch.cuda.synchronize()
print("ok")
