"""Debug aid: repeat one BCNN step and report which intermediate first differs between runs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ.setdefault('HAWKEYE_ALLOW_RANDOM_INIT', '1')
import torch, detgen
import hawkeye_b200 as hb
from hawkeye_b200 import ops
from oracle.hop_oracle import VGG16_D
class Cfg(dict):
    __getattr__ = dict.__getitem__
net = hb.MODEL.get('BCNN')(Cfg(name='BCNN', stage=2, num_classes=200))
net.load_state_dict(detgen.vgg_bcnn_state(VGG16_D, 200, seed=100))
net = net.cuda().train()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = detgen.det((2, 3, S, S), 41).cuda()
labels = detgen.det_labels(2, 200, 42).cuda()
def step(acc):
    if acc:
        for p in net.parameters():
            p.grad = torch.zeros_like(p)
    else:
        net.zero_grad(set_to_none=True)
    feat = net.features(x); feat.retain_grad()
    y = net.bilinear_pooling(feat); y.retain_grad()
    logits = ops.linear(y, net.classifier.weight, net.classifier.bias); logits.retain_grad()
    loss = ops.CrossEntropyLS(0.1)(logits, labels)
    loss.backward()
    return dict(feat=feat.detach().clone(), y=y.detach().clone(), logits=logits.detach().clone(), dlogits=logits.grad.clone(),
                dy=y.grad.clone(), dfeat=feat.grad.clone(), g28=net.backbone.features[28].weight.grad.clone() if hasattr(net.backbone, 'features') else None,
                g0=next(net.backbone.parameters()).grad.clone())
runs = [step(a) for a in (0, 0, 1, 1, 1, 0)]
def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
for i in range(1, len(runs)):
    print('run', i, 'vs run 0:', ' '.join(f'{k}={rel(runs[i][k], runs[0][k]):.1e}' for k in runs[0] if runs[0][k] is not None))
