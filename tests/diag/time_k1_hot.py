"""K1 timing with L2-resident buffers (2 buffer sets): separates the SM-side store limit from DRAM write-back."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
for B in [int(a) for a in sys.argv[1:]] or (32,):
    for fp in (1, 640 << 20):
        t = bench.time_bilinear_kernel(B, min_footprint=fp, reps=40 if fp == 1 else 4)
        print(f'K1 B={B} footprint>={fp >> 20} MB: {t * 1e6:.2f} us', flush=True)
