import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ.setdefault('HAWKEYE_ALLOW_RANDOM_INIT', '1')
import torch, detgen
import hawkeye_b200 as hb
from hawkeye_b200 import ops
from oracle.hop_oracle import VGG16_D
from conftest import rel_l2
class Cfg(dict):
    __getattr__ = dict.__getitem__
net = hb.MODEL.get('BCNN')(Cfg(name='BCNN', stage=2, num_classes=200))
net.load_state_dict(detgen.vgg_bcnn_state(VGG16_D, 200, seed=100))
net = net.cuda().train()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = detgen.det((2, 3, S, S), 41).cuda()
labels = detgen.det_labels(2, 200, 42).cuda()
def backward():
    ops.CrossEntropyLS(0.1)(net(x), labels).backward()
net.zero_grad(set_to_none=True)
backward()
ref = {k: p.grad.clone() for k, p in net.named_parameters()}
net.zero_grad(set_to_none=True)
backward()
for k, p in net.named_parameters():
    e = rel_l2(p.grad.cpu(), ref[k].cpu())
    if e > 1e-6: print('fresh vs fresh', k, e)
for p in net.parameters():
    p.grad = torch.zeros_like(p)
backward()
for k, p in net.named_parameters():
    e = rel_l2(p.grad.cpu(), ref[k].cpu())
    pass
acc1 = {k: p.grad.clone() for k, p in net.named_parameters()}
for p in net.parameters():
    p.grad = torch.zeros_like(p)
backward()
worst = max(rel_l2(p.grad.cpu(), acc1[k].cpu()) for k, p in net.named_parameters())
print('acc vs acc worst', worst)
worst = max(rel_l2(p.grad.cpu(), ref[k].cpu()) for k, p in net.named_parameters())
print('acc vs fresh worst', worst)
print('done')
