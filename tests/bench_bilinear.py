"""Micro-benchmark of the fused bilinear-pool forward/backward (CUDA events, L2 flushed between launches)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hawkeye_b200 import _lib, ops
peak = bench.measured_peaks()[0]
out = {}
for B in (32, 256):
    t, tm = bench.time_bilinear_kernel(B, iters=20 if B == 32 else 8)
    out[f'fwd_B{B}'] = dict(us=t * 1e6, us_median=tm * 1e6, gbs=B * bench.K1_FWD_BYTES_PER_IMG / t / 1e9,
                            frac=B * bench.K1_FWD_BYTES_PER_IMG / t / 1e9 / peak)
    x = torch.rand(B, 512, 14, 14, device='cuda'); dy = torch.randn(B, 512 * 512, device='cuda'); dx = torch.empty_like(x)
    nb = _lib.query('hk_bilinear_pool_bwd_workspace_bytes', B, 512, 196); ws = torch.empty(nb, dtype=torch.uint8, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda'); s = _lib.stream_ptr(); evs = []
    for i in range(8):
        flush.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.call('hk_bilinear_pool_bwd', x, dy, dx, B, 512, 196, ws, nb, s); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize(); ts = sorted(a.elapsed_time(b) for a, b in evs)[1:-1]; tb = sum(ts) / len(ts) * 1e-3
    out[f'bwd_B{B}'] = dict(us=tb * 1e6, gbs=B * bench.K1_BWD_BYTES_PER_IMG / tb / 1e9, frac=B * bench.K1_BWD_BYTES_PER_IMG / tb / 1e9 / peak)
print(json.dumps(out))
