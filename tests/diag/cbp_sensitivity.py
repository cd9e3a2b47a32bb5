import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import torch, detgen, numpy as np
import torch.nn.functional as F
from conftest import rel_l2
from oracle import hop_oracle as O
torch.set_num_threads(8)
d=8192
state={k:v.double() for k,v in detgen.vgg_bcnn_state(O.VGG16_D,200,seed=100,head_in=d).items()}
x,labels=detgen.det((2,3,448,448),41).double(),detgen.det_labels(2,200,42)
with torch.no_grad(): f=O.vgg_features_fwd(x,state)
hashes=O.cbp_hashes(512,d)
def dfeat(f, gram_noise, seed):
    f=f.clone().requires_grad_(True)
    B,C,H,W=f.shape
    xf=f.reshape(B,C,H*W); g=torch.bmm(xf,xf.transpose(1,2))
    if gram_noise:
        gen=torch.Generator().manual_seed(seed)
        g=g*(1+gram_noise*torch.randn(g.shape,generator=gen,dtype=torch.float64))
    h1,s1,h2,s2=hashes
    idx=torch.from_numpy((h1[:,None]+h2[None,:])%d).reshape(-1); sgn=torch.from_numpy(s1[:,None]*s2[None,:]).double().reshape(-1)
    pre=torch.zeros(B,d,dtype=torch.float64).index_add(1,idx,g.reshape(B,-1)*sgn)
    y=F.normalize(torch.sign(pre)*torch.sqrt(pre.abs()+1e-10))
    lg=F.linear(y,state['classifier.weight'],state['classifier.bias'])
    loss=O.cross_entropy_ls(lg,labels); (gf,)=torch.autograd.grad(loss,f)
    return gf, pre.detach()
g0,pre0=dfeat(f,0,0)
print('bin |v| quantiles', [float(q) for q in torch.quantile(pre0.abs().flatten(), torch.tensor([0.0,1e-4,1e-3,0.01,0.5],dtype=torch.float64))])
for noise in (1e-7,3e-7,1e-6):
    print(noise, [f'{rel_l2(dfeat(f,noise,s)[0],g0):.1e}' for s in range(6)])
