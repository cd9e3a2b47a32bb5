"""Hardware probe (not a test): shared-memory image of an MN-major TMA tile (SWIZZLE_128B_ATOM_32B) and tcgen05.mma with the
A operand in tensor memory."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from hawkeye_b200 import _lib

lib = ctypes.CDLL(_lib.LIB_PATH)
f = lib.hk_debug_probe_tmem_a
f.argtypes = [ctypes.c_void_p] * 5
A = (torch.arange(64).view(64, 1) * 1000 + torch.arange(128).view(1, 128)).float().cuda()      # A[k][m] = 1000 k + m
Bm = torch.randn(32, 64, device='cuda')
Bm = (Bm.view(torch.int32) & ~0x1FFF).view(torch.float32)                                         # tf32-representable
dump = torch.zeros(8192, device='cuda')
D = torch.zeros(128, 32, device='cuda')
rc = f(A.data_ptr(), Bm.data_ptr(), dump.data_ptr(), D.data_ptr(), None)
torch.cuda.synchronize()
print('rc', rc)
# (1) swizzle: for each box j (m block), smem word index -> (k, m)
d = dump.cpu().view(4, 64, 32)          # [box][row r][word w] as laid out linearly: row = 128 B
ok = True
for j in range(4):
    for r in range(64):
        for w in range(32):
            val = int(d[j, r, w].item())
            k, m = val // 1000, val % 1000
            # hypothesis: row r holds k = r; 32-byte chunk c32 = w // 8 holds logical chunk c32 ^ (r % 4) of m-block j
            exp_m = j * 32 + (((w // 8) ^ (r % 4)) * 8) + (w % 8)
            if k != r or m != exp_m:
                ok = False
print('swizzle hypothesis chunk32 ^ (row % 4):', 'HOLDS' if ok else 'FAILS')
if not ok:
    for r in range(8):
        print('row', r, [int(d[0, r, w].item()) % 1000 for w in range(0, 32, 4)], 'k', int(d[0, r, 0].item()) // 1000)
# (2) TMEM-A MMA
ref = (A.double().t() @ Bm.double().t())                 # [128 m][32 n] = sum_k A[k][m] B[n][k]
err = ((D.double() - ref).norm() / ref.norm()).item()
print('tmem-A mma rel err', err)
