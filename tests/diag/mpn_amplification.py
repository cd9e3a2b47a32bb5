import sys; sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
import torch, detgen, time
import torch.nn.functional as F
from conftest import rel_l2
from oracle import hop_oracle as O
import hawkeye_b200 as hb
class Cfg(dict): __getattr__=dict.__getitem__
torch.set_num_threads(8)
net = hb.MODEL.get('MPN')(Cfg(name='MPN', iter_num=5, is_sqrt=True, is_vec=True, input_dim=2048, dimension_reduction=256, num_classes=200))
st = {k:(v.double() if v.is_floating_point() else v) for k,v in detgen.state_like(net).items()}
size,B=448,2
x = detgen.det((B,3,size,size),51).double()
with torch.no_grad():
    f = O.resnet50_trunk_fwd(x, st)
    def head(f):
        d = F.relu(O._bn_train(F.conv2d(f, st['pool.conv_dr_block.0.weight']), st, 'pool.conv_dr_block.1'))
        c = O.covpool_fwd(d)
        s,_ = O.sqrtm_fwd(c,5)
        v = O.triuvec_fwd(s)
        lg = F.linear(v.reshape(B,-1), st['classifier.weight'], st['classifier.bias'])
        return d,c,s,lg
    d,c,s,lg = head(f)
    eps=1e-6
    g=torch.Generator().manual_seed(0)
    fp = f*(1+eps*torch.randn(f.shape,generator=g,dtype=torch.float64))
    d2,c2,s2,lg2 = head(fp)
    print('feature pert 1e-6 ->  dr', rel_l2(d2,d), 'cov', rel_l2(c2,c), 'sqrtm', rel_l2(s2,s), 'logits', rel_l2(lg2,lg))
    # perturb cov directly
    cp = c*(1+eps*torch.randn(c.shape,generator=g,dtype=torch.float64)); cp=(cp+cp.transpose(1,2))/2
    s3,_=O.sqrtm_fwd(cp,5)
    print('cov pert 1e-6 -> sqrtm', rel_l2(s3,s))
    ev=torch.linalg.eigvalsh(c[0]/c[0].trace())
    print('eig of A: min',ev.min().item(),'max',ev.max().item(),'n<1e-6',(ev<1e-6).sum().item())
    # image perturbation -> features
    xp = x*(1+eps*torch.randn(x.shape,generator=g,dtype=torch.float64))
    f2 = O.resnet50_trunk_fwd(xp, st)
    print('image pert 1e-6 -> feat', rel_l2(f2,f))
    # logits magnitude relative to bias
    print('logits norm', lg.norm().item(), 'bias norm', st['classifier.bias'].norm().item(), 'lg - bias norm', (lg-st['classifier.bias']).norm().item())
