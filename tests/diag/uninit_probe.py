"""Debug aid: run workspace-taking ops with NaN-poisoned vs zeroed workspaces / outputs; any difference = uninitialised read."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from hawkeye_b200 import _lib
s = _lib.stream_ptr()
B, C, HW = 2, 512, 4
x = torch.rand(B, C, 2, 2, device='cuda'); dy = torch.randn(B, C * C, device='cuda')
nb = _lib.query('hk_bilinear_pool_bwd_workspace_bytes', B, C, HW)
outs = []
for fill in (0.0, float('nan')):
    ws = torch.full((nb // 4 + 4,), fill, device='cuda'); dx = torch.full_like(x, fill)
    _lib.call('hk_bilinear_pool_bwd', x, dy, dx, B, C, HW, ws, nb, s); outs.append(dx.clone())
print('bilinear bwd (2x2 map): equal', torch.equal(outs[0], outs[1]), 'nan', torch.isnan(outs[1]).any().item())
# pool bwd
N, H, W, Cc = 2, 8, 8, 64
code = torch.randint(0, 8, (N, H // 2, W // 2, Cc), device='cuda', dtype=torch.uint8); g = torch.randn(N, H // 2, W // 2, Cc, device='cuda')
outs = []
for fill in (0.0, float('nan')):
    dx = torch.full((N, H, W, Cc), fill, device='cuda')
    _lib.call('hk_maxpool2x2_bwd_idx', code, g, dx, N, H, W, Cc, 0, s); outs.append(dx.clone())
print('pool bwd: equal', torch.equal(outs[0], outs[1]))
# wgrad acc / dgrad
for (N, H, W, Cin, Cout) in ((2, 64, 64, 64, 64), (2, 32, 32, 64, 128), (2, 4, 4, 512, 512), (2, 8, 8, 256, 512)):
    xx = torch.randn(N, H, W, Cin, device='cuda'); gg = torch.randn(N, H, W, Cout, device='cuda')
    nbw = _lib.query('hk_conv3x3_wgrad_workspace_bytes', Cin, Cout)
    res = []
    for fill in (0.0, float('nan')):
        ws = torch.full((nbw // 4 + 4,), fill, device='cuda')
        dw = torch.zeros(Cout, Cin, 3, 3, device='cuda'); db = torch.zeros(Cout, device='cuda')
        _lib.call('hk_conv3x3_wgrad_acc', xx, gg, dw, db, N, H, W, Cin, Cout, ws, nbw, 1, s); res.append((dw.clone(), db.clone()))
    print('wgrad_acc', (N, H, W, Cin, Cout), 'equal', torch.equal(res[0][0], res[1][0]), torch.equal(res[0][1], res[1][1]),
          'nan', torch.isnan(res[1][0]).any().item(), 'maxdiff', (res[0][0] - res[1][0]).abs().max().item())
    w = torch.randn(Cout, Cin, 3, 3, device='cuda') * 0.05
    wf = torch.empty(9 * Cout * Cin, device='cuda'); wd = torch.empty(9 * Cout * Cin, device='cuda')
    _lib.call('hk_conv3x3_pack_weights', w, wf, wd, Cout, Cin, s)
    res = []
    for fill in (0.0, float('nan')):
        dx = torch.full((N, H, W, Cin), fill, device='cuda')
        _lib.call('hk_conv3x3_dgrad', gg, wd, None, dx, N, H, W, Cin, Cout, s); res.append(dx.clone())
    print('dgrad', 'equal', torch.equal(res[0], res[1]))
