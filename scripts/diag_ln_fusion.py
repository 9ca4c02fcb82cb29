"""Fused-LayerNorm building blocks at the bench shape (M = 262 144 rows): correctness against torch and timing of
 (a) the plain CTA-pair GEMM2, (b) the same GEMM with the identity transform hop, (c) with the LayerNorm transform,
 (d) the separate layernorm pass it replaces (through MROAD.infer's phase profile), (e) GEMM1 with / without the statistics epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prego_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
F16 = _lib.PRECISIONS["fp16"]


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- correctness, small
for (M, N, K) in ((300, 512, 256), (4096, 3072, 2048)):
    g = torch.Generator(device=dev).manual_seed(M)
    A = (torch.randn(M, 4096, generator=g, device=dev).abs() * 0.5).half()
    W1 = (torch.randn(K, 4096, generator=g, device=dev) * 0.02).half()
    b1 = torch.randn(K, generator=g, device=dev) * 0.1
    Y = torch.empty(M, K, dtype=torch.float16, device=dev)
    stats = torch.zeros(K // 256, M, 2, device=dev)
    rowstat = torch.zeros(M, 2, device=dev)
    _lib.check(lib.prego_gemm16_stats_nt(A.data_ptr(), W1.data_ptr(), b1.data_ptr(), Y.data_ptr(), stats.data_ptr(), rowstat.data_ptr(), M, K, 4096, F16, 1e-5, s), "stats")
    torch.cuda.synchronize()
    yref = (A.float() @ W1.float().T + b1)
    assert (Y.float() - yref).abs().max() <= 2e-3 * yref.abs().max()
    y32 = Y.float()
    mu, var = y32.mean(-1), y32.var(-1, unbiased=False)
    rstd = 1 / torch.sqrt(var + 1e-5)
    print(f"M={M}: rowstat err rstd {((rowstat[:, 0] - rstd).abs() / rstd).max().item():.2e}  shift {(rowstat[:, 1] + mu * rstd).abs().max().item():.2e}")
    gamma = 1 + 0.1 * torch.randn(K, generator=g, device=dev)
    beta = 0.1 * torch.randn(K, generator=g, device=dev)
    W2 = (torch.randn(N, K, generator=g, device=dev) * 0.03).half()
    b2 = torch.randn(N, generator=g, device=dev) * 0.1
    C = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.prego_gemm16_ln_nt(Y.data_ptr(), rowstat.data_ptr(), gamma.data_ptr(), beta.data_ptr(), W2.data_ptr(), b2.data_ptr(), C.data_ptr(), M, N, K, F16, s), "ln gemm")
    torch.cuda.synchronize()
    e = torch.relu(torch.nn.functional.layer_norm(y32, (K,), gamma, beta, 1e-5)).half().float()
    ref = e @ W2.float().T + b2
    print(f"M={M}: fused LN+GEMM max err {(C - ref).abs().max().item():.3e} (scale {ref.abs().max().item():.2f}), finite {bool(torch.isfinite(C).all())}")
    Ci = torch.full((M, N), float("nan"), device=dev)
    _lib.check(lib.prego_gemm16_ln_nt(Y.data_ptr(), None, None, None, W2.data_ptr(), b2.data_ptr(), Ci.data_ptr(), M, N, K, F16, s), "identity")
    torch.cuda.synchronize()
    refi = y32 @ W2.float().T + b2
    print(f"M={M}: identity hop max err {(Ci - refi).abs().max().item():.3e}")

# ---- timing at the bench shape (fp32 output keeps the epilogue identical across the three GEMM2 variants)
M, N, K = 262144, 3072, 2048
Y = (torch.randn(M, K, device=dev) * 0.6).half()
W2 = (torch.randn(N, K, device=dev) * 0.03).half()
b2 = torch.zeros(N, device=dev)
gamma, beta = torch.ones(K, device=dev), torch.zeros(K, device=dev)
rowstat = torch.stack([torch.full((M,), 1.6, device=dev), torch.zeros(M, device=dev)], 1).contiguous()
C = torch.empty(M, N, device=dev)
fl = 2.0 * M * N * K
t = timeit(lambda: lib.prego_gemm16_nt(Y.data_ptr(), W2.data_ptr(), b2.data_ptr(), C.data_ptr(), M, N, K, -256, F16, s))
print(f"GEMM2 shape, plain CTA-pair kernel:   {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s")
t = timeit(lambda: lib.prego_gemm16_ln_nt(Y.data_ptr(), None, None, None, W2.data_ptr(), b2.data_ptr(), C.data_ptr(), M, N, K, F16, s))
print(f"GEMM2 shape, identity transform hop:  {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s")
t = timeit(lambda: lib.prego_gemm16_ln_nt(Y.data_ptr(), rowstat.data_ptr(), gamma.data_ptr(), beta.data_ptr(), W2.data_ptr(), b2.data_ptr(), C.data_ptr(), M, N, K, F16, s))
print(f"GEMM2 shape, LayerNorm transform:     {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s")
del C
M, N, K = 262144, 2048, 4096
A = (torch.randn(M, K, device=dev).abs() * 0.5).half()
W1 = (torch.randn(N, K, device=dev) * 0.02).half()
b1 = torch.zeros(N, device=dev)
Yo = torch.empty(M, N, dtype=torch.float16, device=dev)
stats = torch.zeros(N // 256, M, 2, device=dev)
rs = torch.zeros(M, 2, device=dev)
t = timeit(lambda: lib.prego_gemm16_stats_nt(A.data_ptr(), W1.data_ptr(), b1.data_ptr(), Yo.data_ptr(), stats.data_ptr(), rs.data_ptr(), M, N, K, F16, 1e-5, s))
print(f"GEMM1 shape with the statistics epilogue + finalize: {t:.3f} ms  {2.0 * M * N * K / t / 1e9:.0f} TFLOP/s")
