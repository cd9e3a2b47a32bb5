"""Profiling driver: 2 MPN (ResNet-50 + MPN-COV) train steps at 448x448, batch 32 (eager launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('HAWKEYE_ALLOW_RANDOM_INIT', '1')
import torch
import hawkeye_b200 as hb
from hawkeye_b200 import engine, ops
class Cfg(dict):
    __getattr__ = dict.__getitem__
net = hb.MODEL.get('MPN')(Cfg(name='MPN', iter_num=5, is_sqrt=True, is_vec=True, input_dim=2048, dimension_reduction=256, num_classes=200)).cuda().train()
flat = engine.FlatParams(net); opt = engine.FusedSGD(flat, lr=1e-3, momentum=0.9); crit = ops.CrossEntropyLS(0.1)
x = torch.randn(32, 3, 448, 448, device='cuda'); y = torch.randint(0, 200, (32,), device='cuda')
for i in range(2):
    loss = crit(net(x), y); opt.zero_grad(); loss.backward(); opt.step()
torch.cuda.synchronize()
