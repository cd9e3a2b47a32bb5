import os, subprocess, sys
code = r'''
import sys, torch, torch.nn.functional as F
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import detgen
from hawkeye_b200 import _lib
s=_lib.stream_ptr()
for (N,H,W,Cin,Cout) in ((2,16,16,64,64),(1,4,4,512,512)):
    x = detgen.det((N,Cin,H,W),1,positive=True); dy = detgen.det((N,Cout,H,W),4)
    xg = x.permute(0,2,3,1).contiguous().cuda(); dg = dy.permute(0,2,3,1).contiguous().cuda()
    dw = torch.empty(Cout,Cin,3,3,device='cuda'); db = torch.empty(Cout,device='cuda')
    nb=_lib.query('hk_conv3x3_wgrad_workspace_bytes',Cin,Cout); ws=torch.empty(nb,dtype=torch.uint8,device='cuda')
    _lib.call('hk_conv3x3_wgrad', xg, dg, dw, db, N,H,W,Cin,Cout, ws, nb, s); torch.cuda.synchronize()
    ref = dy.double().sum((0,2,3)); got = db.cpu().double()
    print('   ', (N,H,W,Cin,Cout), 'db rel', ((got-ref).norm()/ref.norm()).item(), 'ratio', (got[:4]/ref[:4]).tolist())
'''
for v in (0,1,2,4):
    print('variant', v, flush=True)
    r = subprocess.run([sys.executable,'-c',code], env=dict(os.environ, HK_DBG_WG=str(v)), capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr[-400:] if r.returncode else '', flush=True)
