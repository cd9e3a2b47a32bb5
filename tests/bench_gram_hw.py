"""Profiling aid: bilinear-pool forward time vs H*W (row pitch alignment of the TMA operand loads)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hawkeye_b200 import _lib
B = 1024
for dbg in ('0', '3'):
    os.environ['HK_GRAM_DBG'] = dbg
    for HW in (192, 196, 200, 224, 256):
        xs = [torch.rand(B, 512, HW, device='cuda') for _ in range(2)]
        ys = [torch.empty(B, 512 * 512, device='cuda') for _ in range(2)]
        nb = _lib.query('hk_bilinear_pool_fwd_workspace_bytes', B, 512, HW)
        ws = torch.empty(nb, dtype=torch.uint8, device='cuda')
        s = _lib.stream_ptr()
        for i in range(2):
            _lib.call('hk_bilinear_pool_fwd', xs[i], ys[i], None, B, 512, HW, ws, nb, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(4):
            for i in range(2):
                _lib.call('hk_bilinear_pool_fwd', xs[i], ys[i], None, B, 512, HW, ws, nb, s)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 8 * 1e-3
        print(json.dumps(dict(dbg=dbg, HW=HW, us=round(t * 1e6, 1), x_gbs=round(B * 512 * HW * 4 / t / 1e9, 1),
                              alg_gbs=round(B * (512 * HW * 4 + 512 * 512 * 4) / t / 1e9, 1))), flush=True)
        del xs, ys
