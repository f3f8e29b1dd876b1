import numpy as np, torch, sys, ctypes, itertools
sys.path.insert(0, '/root/repo')
import arecsys_b200
from arecsys_b200 import _lib
lib = _lib.load()
M, N, K = 128, 128, 32
rng = np.random.default_rng(0)
A = rng.integers(-4, 5, (M, K)).astype(np.float32)
B = rng.integers(-4, 5, (K, N)).astype(np.float32)
ref = A @ B
def run(ta, tb, dbg):
    arr = (ctypes.c_int * 8)(*dbg)
    lib.arx_gemm_tc_set_dbg(arr)
    dA = torch.tensor(np.ascontiguousarray(A.T if ta else A), device='cuda')
    dB = torch.tensor(np.ascontiguousarray(B.T if tb else B), device='cuda')
    C = torch.full((M, N), 777.0, device='cuda')
    _lib.call('arx_gemm_tc', dA.data_ptr(), dB.data_ptr(), C.data_ptr(), M, N, K, ta, tb, None, 1.0, 0.0)
    torch.cuda.synchronize()
    c = C.cpu().numpy()
    return (c == ref).mean(), float(np.abs(c).max())
print('baseline KK', run(0, 1, [0]*8))
for lbo, sbo, adv in itertools.product([16, 128, 512, 1024, 4096], [128, 512, 1024, 4096], [32, 128, 1024, 4096]):
    r = run(0, 0, [0, 0, 0, lbo, sbo, adv, 0, 0])
    if r[0] > 0.05 or (lbo, sbo, adv) == (4096, 1024, 1024):
        print('B_MN lbo %d sbo %d adv %d -> match %.3f max %.1f' % (lbo, sbo, adv, r[0], r[1]))
