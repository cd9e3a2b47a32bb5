"""Diagnostic (not a test): where does the BCNN backward lose accuracy at 64x64 (2x2 feature map)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
import detgen, matched
from conftest import rel_l2
from oracle import hop_oracle as O
import hawkeye_b200 as hb
from hawkeye_b200 import ops


class Cfg(dict):
    __getattr__ = dict.__getitem__


torch.set_num_threads(16)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
net = hb.MODEL.get('BCNN')(Cfg(name='BCNN', stage=2, num_classes=200))
state = detgen.vgg_bcnn_state(O.VGG16_D, 200, seed=100)
net.load_state_dict(state)
net = net.cuda().train()
x, labels = detgen.det((2, 3, size, size), 41), detgen.det_labels(2, 200, 42)
ops.CAPTURE = []
feat = net.backbone(x.cuda())
cap = ops.CAPTURE
ops.CAPTURE = None
feat.retain_grad()
y = net.bilinear_pooling(feat)
y.retain_grad()
logits = ops.linear(y, net.classifier.weight, net.classifier.bias)
loss = ops.CrossEntropyLS(0.1)(logits, labels.cuda())
loss.backward()
tape = O.MaskTape(matched.tape_items(cap))
st = {k: v.double().requires_grad_(True) for k, v in state.items()}
f64 = O.vgg_features_fwd(x.double(), st, nl=tape)
f64.retain_grad()
y64 = O.bilinear_pool_fwd(f64)
y64.retain_grad()
lg64 = F.linear(y64, st['classifier.weight'], st['classifier.bias'])
l64 = O.cross_entropy_ls(lg64, labels)
l64.backward()
print('size', size, 'feat', tuple(feat.shape))
print('feat fwd rel', rel_l2(feat.detach().cpu(), f64.detach()), 'y fwd rel', rel_l2(y.detach().cpu(), y64.detach()))
print('dlogits->dy rel', rel_l2(y.grad.cpu(), y64.grad), ' dfeat rel', rel_l2(feat.grad.cpu(), f64.grad))
# the bilinear backward alone, fed the ORACLE's x and dy
xo = f64.detach().float().cuda().requires_grad_(True)
yo = ops.bilinear_pool(xo)
(dxo,) = torch.autograd.grad(yo, xo, y64.grad.float().cuda())
print('bilinear bwd alone (oracle inputs) rel', rel_l2(dxo.cpu(), f64.grad))
xo2 = detgen.tf32_rna(f64.detach().float()).cuda().requires_grad_(True)
yo2 = ops.bilinear_pool(xo2)
(dxo2,) = torch.autograd.grad(yo2, xo2, y64.grad.float().cuda())
print('bilinear bwd alone (tf32-rounded oracle x) rel', rel_l2(dxo2.cpu(), f64.grad))
for k in ('backbone.28.bias', 'backbone.28.weight', 'backbone.0.weight'):
    print(k, rel_l2(dict(net.named_parameters())[k].grad.cpu(), st[k].grad))
