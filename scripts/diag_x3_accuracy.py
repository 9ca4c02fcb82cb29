"""Where does the 3xTF32 product lose accuracy?  C = A W^T (K = 2048, non-negative A like post-ReLU features) against fp64:
plain TF32, split product in one launch over 3K, and the same contraction cut into chunks whose partial results are added
in fp32 by the epilogue (accumulate = 1)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from prego_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
g = torch.Generator(device=dev).manual_seed(0)
M, N, K = 512, 1024, 2048
for name, A in (("abs-normal A", torch.randn(M, K, generator=g, device=dev).abs()), ("normal A", torch.randn(M, K, generator=g, device=dev))):
    W = torch.randn(N, K, generator=g, device=dev) * 0.02
    ref = A.double() @ W.double().T
    scale = ref.abs().max().item()

    def split(x):
        hi = ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        return hi, x - hi

    ah, al = split(A)
    wh, wl = split(W)
    A3 = torch.cat([ah, al, ah], 1).contiguous()
    W3 = torch.cat([wh, wh, wl], 1).contiguous()
    A3s = torch.cat([al, ah, ah], 1).contiguous()   # small terms first
    W3s = torch.cat([wh, wl, wh], 1).contiguous()

    def run(a, w, k, chunks=1):
        C = torch.zeros(M, N, device=dev)
        kc = k // chunks
        for c in range(chunks):
            ac, wc = a[:, c * kc:(c + 1) * kc].contiguous(), w[:, c * kc:(c + 1) * kc].contiguous()
            _lib.check(lib.prego_gemm_tf32_nt(ac.data_ptr(), wc.data_ptr(), None, C.data_ptr(), M, N, kc, 1 if c else 0, st), "gemm")
        torch.cuda.synchronize()
        d = C.double() - ref
        return d.abs().max().item() / scale, d.mean().item() / scale, (d.norm() / ref.norm()).item()

    print(name, "max|C| %.3f" % scale)
    for label, r in (("fp32 torch matmul (cuBLAS, no tf32)", None), ("tf32 one launch", run(A, W, K)), ("x3 one launch over 3K", run(A3, W3, 3 * K)),
                     ("x3 small terms first", run(A3s, W3s, 3 * K)),
                     ("x3 in 6 chunks", run(A3, W3, 3 * K, 6)), ("x3 in 24 chunks", run(A3, W3, 3 * K, 24)), ("x3 in 96 chunks", run(A3, W3, 3 * K, 96))):
        if r is None:
            torch.backends.cuda.matmul.allow_tf32 = False
            d = (A @ W.T).double() - ref
            r = (d.abs().max().item() / scale, d.mean().item() / scale, (d.norm() / ref.norm()).item())
        print(f"  {label:38s} max err / max|C| {r[0]:.2e}   mean signed {r[1]:+.2e}   fro {r[2]:.2e}")
