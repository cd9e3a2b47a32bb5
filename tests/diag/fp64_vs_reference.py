import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import torch, detgen, numpy as np
from conftest import rel_l2
from oracle import hop_oracle as O
torch.set_num_threads(8)
ref=np.load('/root/repo/tests/golden/reference_448.npz')
state={k:v.double() for k,v in detgen.vgg_bcnn_state(O.VGG16_D,200,seed=100).items()}
x,labels=detgen.det((2,3,448,448),41).double(),detgen.det_labels(2,200,42)
lg,loss,g=O.loss_and_grads(lambda xx,st:O.bcnn_forward(xx,st,2),x,labels,state)
print('fp64 plain oracle vs fp32 reference: logits',rel_l2(lg,ref['bcnn_s2_logits']))
for k in ('backbone.0.weight','backbone.0.bias','backbone.2.bias','backbone.14.bias','backbone.28.bias','classifier.bias'):
    print(k, rel_l2(g[k],ref['bcnn_s2_g_'+k]))
