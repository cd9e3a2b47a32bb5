"""K1 (hk_bilinear_pool_fwd) correctness sweep + timing (not a pytest file; run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import detgen
from conftest import rel_l2
from oracle import hop_oracle as O
from hawkeye_b200 import ops
import bench

torch.set_num_threads(16)
ok = True
for (B, H, W, sparse) in [(1, 14, 14, 0), (2, 2, 2, 1), (3, 6, 6, 1), (5, 14, 14, 1), (33, 14, 14, 0), (34, 14, 14, 1), (70, 14, 14, 0), (133, 4, 4, 0)]:
    x = detgen.det_uniform((B, 512, H, W), 5)
    if sparse:
        x = torch.relu(x - 0.4)
    y = ops.bilinear_pool(x.cuda())
    torch.cuda.synchronize()
    y_ref = O.bilinear_pool_fwd(x.double())
    e = rel_l2(y.cpu(), y_ref)
    worst = max(rel_l2(y[b].cpu(), y_ref[b]) for b in range(B))
    print(f'K1 B={B} {H}x{W} sparse={sparse}: rel {e:.2e} worst image {worst:.2e}', flush=True)
    ok &= worst < 1e-3
print('K1 correctness', 'OK' if ok else 'FAILED', flush=True)
if ok and len(sys.argv) > 1:
    for B in (32, 256, 1024):
        t = bench.time_bilinear_kernel(B)
        gbs = B * bench.K1_FWD_BYTES_PER_IMG / t / 1e9
        print(f'K1 B={B}: {t * 1e6:.2f} us  {gbs:.0f} GB/s  frac {gbs / 6561.6:.3f}', flush=True)
