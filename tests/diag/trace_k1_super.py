"""Profiling aid: per-CTA phase timestamps of the K1 super-tile kernel (not a test).  usage: trace_k1_super.py B"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from hawkeye_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
xs = [torch.rand(B, 512, 14, 14, device='cuda') for _ in range(max(2, 700 // (B * 2) + 1))]
for x in xs[:3]:
    ops.bilinear_pool(x)
torch.cuda.synchronize()
tr = torch.zeros(148 * 16, dtype=torch.int64, device='cuda')
lib = ctypes.CDLL(_lib.LIB_PATH)
lib.hk_debug_k1_trace.argtypes = [ctypes.c_void_p]
for x in xs:                      # cold inputs, steady state: trace the last launch of a back-to-back series
    ops.bilinear_pool(x)
lib.hk_debug_k1_trace(tr.data_ptr())
ops.bilinear_pool(xs[0])
torch.cuda.synchronize()
lib.hk_debug_k1_trace(None)
t = tr.view(148, 16).cpu().double()
live = t[:, 0] > 0
names = {0: 'start', 1: 'it0 first stage full', 2: 'it0 MMAs issued', 3: 'it0 acc_full', 4: 'it0 sums out', 5: 'it0 norm ready',
         6: 'it0 stores done', 7: 'it1 first stage full', 8: 'it1 MMAs issued', 9: 'it1 acc_full', 10: 'it1 sums out',
         11: 'it1 norm ready', 12: 'it1 stores done', 13: 'end'}
t0 = t[live, 0].min()
for kind, sel in (('D', [c for c in range(148) if c % 4 < 2]), ('O', [c for c in range(148) if c % 4 >= 2])):
    tt = t[sel]
    tt = tt[tt[:, 0] > 0]
    print(f'B={B} {kind}-CTAs: {tt.shape[0]} traced; us since the earliest CTA start (min / median / max)')
    for i, n in names.items():
        v = tt[:, i]
        v = v[v > 0]
        if len(v):
            v = (v - t0) / 1e3
            print(f'  {n:24s} {v.min():8.2f} {v.median():8.2f} {v.max():8.2f}')
if B <= 64 and len(sys.argv) > 2:
    r = (t - t0) / 1e3
    print('per image: sums-out of its 4 CTAs | norm-ready | stores-done')
    for b in range(B):
        c = [4 * b + i for i in range(4)]
        print(f'  img {b:2d}: ' + ' '.join(f'{r[i, 4]:6.2f}' for i in c) + ' | ' + ' '.join(f'{r[i, 5]:6.2f}' for i in c) + ' | ' +
              ' '.join(f'{r[i, 6]:6.2f}' for i in c))
