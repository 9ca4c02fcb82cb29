import os, sys
sys.path.insert(0, "/root/repo")
import torch
from prego_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda:0"); s = torch.cuda.current_stream().cuda_stream; F16 = _lib.PRECISIONS["fp16"]
M, N, K = 262144, 3072, 2048
Y = (torch.randn(M, K, device=dev) * 0.6).half(); W2 = (torch.randn(N, K, device=dev) * 0.03).half(); b2 = torch.zeros(N, device=dev)
C = torch.empty(M, N, device=dev)
def timeit(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
t = timeit(lambda: lib.prego_gemm16_nt(Y.data_ptr(), W2.data_ptr(), b2.data_ptr(), C.data_ptr(), M, N, K, -256, F16, s)); print(f"plain {t:.3f} ms")
for dbg in (0, 4, 2, 1, 5):
    os.environ["PREGO_XF_DBG"] = str(dbg)
    t = timeit(lambda: lib.prego_gemm16_ln_nt(Y.data_ptr(), None, None, None, W2.data_ptr(), b2.data_ptr(), C.data_ptr(), M, N, K, F16, s)); print(f"identity dbg={dbg} {t:.3f} ms", flush=True)
