"""Profiling aid: per-CTA phase timestamps of the K1 cluster kernel (not a test).  usage: trace_k1.py B"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from hawkeye_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.rand(B, 512, 14, 14, device='cuda')
for _ in range(3):
    ops.bilinear_pool(x)
torch.cuda.synchronize()
tr = torch.zeros(148 * 16, dtype=torch.int64, device='cuda')
lib = ctypes.CDLL(_lib.LIB_PATH)
lib.hk_debug_k1_trace.argtypes = [ctypes.c_void_p]
lib.hk_debug_k1_trace(tr.data_ptr())
ops.bilinear_pool(x)
torch.cuda.synchronize()
lib.hk_debug_k1_trace(None)
t = tr.view(148, 16).cpu().double()
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
names = {0: 'start', 1: 'img0 first stage full', 2: 'img0 MMAs issued', 3: 'img0 acc_full (epi start)', 4: 'img0 epi done',
         5: 'img1 first stage full', 6: 'img1 MMAs issued', 7: 'img1 acc_full', 8: 'img1 epi done', 9: 'img0 norm ready',
         10: 'img1 norm ready', 11: 'end'}
print(f'B={B}: {t.shape[0]} CTAs traced; times in us relative to the earliest CTA start (min / median / max over CTAs)')
for i, n in names.items():
    v = t[:, i]
    v = v[v > 0]
    if len(v):
        v = (v - t0) / 1e3
        print(f'  {n:28s} {v.min():8.2f} {v.median():8.2f} {v.max():8.2f}')
