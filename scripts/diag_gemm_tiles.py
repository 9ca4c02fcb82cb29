"""Tile-shape experiment for the recurrence's GEMM (h W_hh'^T, K = 1024, N = 3072): the generic CTA-pair kernel with
192- vs 256-column tiles on the same problem, fp32 output (prego_gemm16_nt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prego_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
for (M, N, K) in ((32768, 3072, 1024), (32768, 3072, 2048), (65536, 2048, 4096)):
    A = (torch.randn(M, K, device=dev) * 0.5).half()
    W = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.zeros(N, device=dev)
    C = torch.empty(M, N, device=dev)
    for tile in (-256, -192):
        if N % (-tile):
            continue
        for _ in range(3):
            _lib.check(lib.prego_gemm16_nt(A.data_ptr(), W.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, tile, _lib.PRECISIONS["fp16"], s), "gemm")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            lib.prego_gemm16_nt(A.data_ptr(), W.data_ptr(), bias.data_ptr(), C.data_ptr(), M, N, K, tile, _lib.PRECISIONS["fp16"], s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"M={M} N={N} K={K} tile={-tile}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
