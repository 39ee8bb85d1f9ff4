"""Times the fused residual stack alone: 409600 rows x 100, 12 layers (one person net of config B/C)."""
import sys
import torch
sys.path.insert(0, ".")
from fastace_b200 import fused_mlp
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 409600
torch.manual_seed(0)
layers = [torch.nn.Linear(100, 100).cuda() for _ in range(12)]
x = torch.randn(rows, 100, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = fused_mlp.residual_tanh_stack(x, layers)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = fused_mlp.residual_tanh_stack(x, layers)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{ms:.3f} ms  {rows * 12 * 2 * 100 * 100 / ms / 1e9:.1f} TFLOP/s (useful)  {rows * 100 * 4 * 2 / ms / 1e6:.0f} GB/s activations")
