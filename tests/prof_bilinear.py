"""Profiling driver: a few launches of the fused bilinear-pool forward (and backward) at BASELINE size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hawkeye_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bwd = len(sys.argv) > 2 and sys.argv[2] == 'bwd'
x = torch.rand(B, 512, 14, 14, device='cuda', requires_grad=True)
dy = torch.randn(B, 512 * 512, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3):
    flush.zero_()
    y = ops.bilinear_pool(x)
    if bwd:
        (dx,) = torch.autograd.grad(y, x, dy)
torch.cuda.synchronize()
