"""Timing-only sweep of hk_bilinear_pool_fwd (no correctness check: used with the HK_K1_DBG elimination flags)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
tag = ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('HK_'))
for B in [int(a) for a in sys.argv[1:]] or (32, 256):
    t = bench.time_bilinear_kernel(B)
    print(f'[{tag}] K1 B={B}: {t * 1e6:.2f} us  frac {B * bench.K1_FWD_BYTES_PER_IMG / t / 1e9 / 6561.6:.3f}', flush=True)
