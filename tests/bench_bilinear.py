"""Micro-benchmark of the fused bilinear-pool forward/backward (CUDA events over a run of back-to-back launches that
rotate through buffer sets whose total footprint exceeds 4x the L2, so every launch sees cold inputs and the steady-state
write-back traffic of its predecessors).  `python tests/bench_bilinear.py [variants]` sweeps the kernel's env knobs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from hawkeye_b200 import _lib

peak = bench.measured_peaks()[0]


def check(B=4):
    x = torch.rand(B, 512, 14, 14, device='cuda', generator=torch.Generator('cuda').manual_seed(3))
    y = torch.empty(B, 512 * 512, device='cuda')
    nb = _lib.query('hk_bilinear_pool_fwd_workspace_bytes', B, 512, 196)
    ws = torch.empty(nb, dtype=torch.uint8, device='cuda')
    _lib.call('hk_bilinear_pool_fwd', x, y, None, B, 512, 196, ws, nb, _lib.stream_ptr())
    X = x.double().view(B, 512, 196)
    z = (X @ X.transpose(1, 2) / 196 + 1e-5).sqrt().view(B, -1)
    ref = z / z.norm(dim=1, keepdim=True)
    return ((y.double() - ref).norm() / ref.norm()).item()


def main():
    variants = [dict(HK_GRAM_FUSED='0'), dict(HK_GRAM_FUSED='2'), dict(HK_GRAM_FUSED='2', HK_GRAM_STORE='0'),
                dict(HK_GRAM_FUSED='2', HK_GRAM_STAGES='4'), dict(HK_GRAM_FUSED='2', HK_GRAM_PDL='0')]
    if len(sys.argv) > 1:
        variants = [dict(kv.split('=') for kv in a.split(',')) for a in sys.argv[1:]]
    out = []
    for v in variants:
        for k in ('HK_GRAM_FUSED', 'HK_GRAM_STORE', 'HK_GRAM_XHINT', 'HK_GRAM_DBG', 'HK_GRAM_STAGES', 'HK_GRAM_PDL', 'HK_GRAM_BALANCE'):
            os.environ.pop(k, None)
        os.environ.update(v)
        r = dict(variant=v, rel_err=check())
        for B in (32, 256, 1024):
            t = bench.time_bilinear_kernel(B)
            r[f'B{B}'] = dict(us=round(t * 1e6, 2), gbs=round(B * bench.K1_FWD_BYTES_PER_IMG / t / 1e9, 1),
                              frac=round(B * bench.K1_FWD_BYTES_PER_IMG / t / 1e9 / peak, 4))
        print(json.dumps(r), flush=True)
        out.append(r)
    for k in ('HK_GRAM_FUSED', 'HK_GRAM_STORE', 'HK_GRAM_XHINT', 'HK_GRAM_DBG', 'HK_GRAM_STAGES', 'HK_GRAM_PDL', 'HK_GRAM_BALANCE'):
        os.environ.pop(k, None)
    # backward (two kernels + GEMM), same protocol
    for B in (32, 256):
        t = bench.time_bilinear_kernel(B, bwd=True)
        print(json.dumps({f'bwd_B{B}': dict(us=round(t * 1e6, 2), gbs=round(B * bench.K1_BWD_BYTES_PER_IMG / t / 1e9, 1),
                                             frac=round(B * bench.K1_BWD_BYTES_PER_IMG / t / 1e9 / peak, 4))}), flush=True)


if __name__ == '__main__':
    main()
