"""Profiling aid: per-CTA timeline (globaltimer) of one hk_bilinear_pool_fwd launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hawkeye_b200 import _lib
for B in [int(a) for a in sys.argv[1:]] or [32]:
    xs = [torch.rand(B, 512, 14, 14, device='cuda') for _ in range(3)]
    ys = [torch.empty(B, 512 * 512, device='cuda') for _ in range(3)]
    nb = _lib.query('hk_bilinear_pool_fwd_workspace_bytes', B, 512, 196)
    ws = torch.empty(nb, dtype=torch.uint8, device='cuda')
    s = _lib.stream_ptr()
    tr = torch.zeros(148, 16, dtype=torch.int64, device='cuda')
    flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
    for i in range(2):
        _lib.call('hk_bilinear_pool_fwd', xs[i], ys[i], None, B, 512, 196, ws, nb, s)
    flush.zero_(); flush.sum()
    torch.cuda.synchronize()
    _lib.lib().hk_debug_gram_trace(tr.data_ptr())
    _lib.call('hk_bilinear_pool_fwd', xs[2], ys[2], None, B, 512, 196, ws, nb, s)
    torch.cuda.synchronize()
    _lib.lib().hk_debug_gram_trace(None)
    t = tr.cpu().double()
    t0 = t[:, 0][t[:, 0] > 0].min()
    names = ['start', 'first_full', 'acc0', 'acc1', 'acc2', 'acc3', 'norm0', 'norm1', 'norm2', 'norm3', 'done0', 'done1',
             'done2', 'done3', 'end']
    print(f'B={B}: ns since the first CTA start  (min / median / max over CTAs that reached the stamp)')
    for i, n in enumerate(names):
        col = t[:, i]
        col = col[col > 0] - t0
        if len(col):
            print(f'  {n:10s} n={len(col):3d}  {col.min():8.0f} {col.median():8.0f} {col.max():8.0f}')
