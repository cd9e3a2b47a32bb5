"""Profiling driver: the Fast MPN-COV head (covariance + 5-iteration Newton-Schulz + triu-vec) forward + backward, B=32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hawkeye_b200 import ops
x = torch.rand(32, 256, 14, 14, device='cuda', requires_grad=True)
for _ in range(2):
    v = ops.TriuvecLayer(ops.SqrtmLayer(ops.CovpoolLayer(x), 5))
    v.backward(torch.ones_like(v))
    x.grad = None
torch.cuda.synchronize()
