"""Profiling driver: a few launches of the fused bilinear-pool forward/backward at BASELINE size (B=32, 512x14x14)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hawkeye_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.rand(B, 512, 14, 14, device='cuda', requires_grad=True)
dy = torch.randn(B, 512 * 512, device='cuda')
for _ in range(3):
    y = ops.bilinear_pool(x)
    (dx,) = torch.autograd.grad(y, x, dy)
torch.cuda.synchronize()
