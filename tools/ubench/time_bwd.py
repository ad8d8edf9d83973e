"""Average time of the one-launch layer backward and of the fused forward at 65,536 rows (CUDA events, 50 launches)."""
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for name, fn in (("backward", lambda: _fused.layer_backward(lay, x, g, None, 1.0, False)), ("forward", lambda: _fused.run(lay, x, False))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us at {N} rows", flush=True)
