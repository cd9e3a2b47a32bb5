"""Debug aid: run the VGG feature stack twice (fused conv+pool) and once unfused; report where outputs / codes differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import detgen
from hawkeye_b200 import ops
from oracle.hop_oracle import VGG16_D

st = detgen.vgg_bcnn_state(VGG16_D, 200, seed=100)
params = []
i = 0
for v in VGG16_D:
    if v == 'M':
        i += 1
        continue
    params += [st[f'backbone.{i}.weight'].cuda().requires_grad_(True), st[f'backbone.{i}.bias'].cuda().requires_grad_(True)]
    i += 2
x = detgen.det((2, 3, 64, 64), 41).cuda()

def run(fuse):
    ops.FUSE_CONV_POOL = fuse
    with torch.enable_grad():
        f = ops.vgg_features(x, VGG16_D, params)
        node = f.grad_fn
        codes = [r['code'].clone() for r in node.records if r['kind'] == 'pool']
        g = torch.autograd.grad(f, params, torch.ones_like(f))
    return f.detach().clone(), codes, [t.clone() for t in g]

a = run(True); b = run(True); c = run(False)
for name, (u, v) in (('fused vs fused', (a, b)), ('fused vs unfused', (a, c))):
    print(name, 'feat equal', torch.equal(u[0], v[0]), 'max diff', (u[0] - v[0]).abs().max().item())
    for k, (cu, cv) in enumerate(zip(u[1], v[1])):
        print('  pool', k, 'code equal', torch.equal(cu, cv), 'mismatches', (cu != cv).sum().item(), 'of', cu.numel())
    for k, (gu, gv) in enumerate(zip(u[2], v[2])):
        d = (gu - gv).norm().item() / (gv.norm().item() + 1e-30)
        if d > 1e-6:
            print('  grad', k, 'rel diff', d)
