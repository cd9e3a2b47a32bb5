"""Golden fixture for the full MPN model (ResNet-50 trunk + Fast MPN-COV head), from the UNMODIFIED reference.
Run here only:  python tests/golden/make_golden_mpn.py   -> tests/golden/reference_mpn.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
from oracle import ref_harness as rh  # noqa: E402
import detgen  # noqa: E402

rh.load_reference()
from model.registry import MODEL  # noqa: E402

torch.set_num_threads(8)
net = MODEL.get('MPN')(rh.cfg(name='MPN', iter_num=5, is_sqrt=True, is_vec=True, input_dim=2048,
                              dimension_reduction=256, num_classes=200))
net.load_state_dict(detgen.state_like(net))
net.train()
x = detgen.det((4, 3, 128, 128), 51)
labels = detgen.det_labels(4, 200, 52)
feat = net.backbone(x)
logits = net(x)      # second forward: running stats move twice; irrelevant for train-mode outputs
loss = torch.nn.CrossEntropyLoss(label_smoothing=0.1)(logits, labels)
net.zero_grad()
loss.backward()
out = {'feat_slice': feat.detach().numpy()[:, ::16], 'logits': logits.detach().numpy(), 'loss': np.float32(loss.item()),
       'g_classifier_bias': net.classifier.bias.grad.numpy(),
       'g_dr_conv': net.pool.conv_dr_block[0].weight.grad.numpy()[:, ::8, 0, 0],
       'g_layer4_bn3_w': net.backbone[7][2].bn3.weight.grad.numpy(),
       'g_stem_w': net.backbone[0].weight.grad.numpy()}
np.savez_compressed(os.path.join(HERE, 'reference_mpn.npz'), **out)
print({k: np.asarray(v).shape for k, v in out.items()}, float(loss))
