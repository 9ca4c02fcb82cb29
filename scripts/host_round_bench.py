"""Host-side rounding throughput on the GPU box's CPU (prego_host_round_features), alone and next to an H2D copy stream.
  python scripts/host_round_bench.py"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from prego_b200 import _lib

lib = _lib.load()
n = 1 << 28  # 1 GiB of fp32
src = torch.rand(n).pin_memory()
dst = torch.empty(n, dtype=torch.float16).pin_memory()
print("impl", lib.prego_host_round_impl(), "cpus", os.cpu_count())


def conv(threads, reps=3):
    lib.prego_host_round_features(src.data_ptr(), dst.data_ptr(), n, 2, threads)
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.prego_host_round_features(src.data_ptr(), dst.data_ptr(), n, 2, threads)
    return (time.perf_counter() - t0) / reps


for th in (1, 2, 4, 8, 12, 16, 24, 32):
    dt = conv(th)
    print(f"threads {th:2d}: {n * 4 / dt / 1e9:6.1f} GB/s fp32 read  ({dt * 1e3:6.1f} ms per GiB)")

if torch.cuda.is_available():
    dev = torch.device("cuda:0")
    d32 = torch.empty(n, dtype=torch.float32, device=dev)
    d16 = torch.empty(n, dtype=torch.float16, device=dev)
    for name, h, d in (("fp32", src, d32), ("fp16", dst, d16)):
        d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        print(f"H2D {name} alone: {h.numel() * h.element_size() / dt / 1e9:.1f} GB/s")
    import ctypes as C
    for slots, slot_bytes in ((4, 4 << 20), (3, 8 << 20), (8, 2 << 20), (4, 8 << 20), (2, 16 << 20)):
        h = C.c_void_p()
        assert lib.prego_host_stager_create(os.cpu_count(), slots, slot_bytes, C.byref(h)) == 0
        st = torch.cuda.Stream()
        lib.prego_host_stager_run(h, src.data_ptr(), d16.data_ptr(), n, 2, st.cuda_stream)
        st.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            lib.prego_host_stager_run(h, src.data_ptr(), d16.data_ptr(), n, 2, st.cuda_stream)
        st.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"ring stager {slots} x {slot_bytes >> 10} KiB, {os.cpu_count()} threads: round + H2D of 1 GiB fp32 in {dt * 1e3:.1f} ms = {n * 4 / dt / 1e9:.1f} GB/s fp32 read, {n * 2 / dt / 1e9:.1f} GB/s on the link")
        ref = torch.empty(n, dtype=torch.float16).pin_memory()
        lib.prego_host_round_features(src.data_ptr(), ref.data_ptr(), n, 2, 8)
        assert torch.equal(d16.cpu(), ref), "ring stager output differs"
        lib.prego_host_stager_destroy(h)
    stop = False

    def copier():
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            while not stop:
                d16.copy_(dst, non_blocking=True)
                s.synchronize()

    t = threading.Thread(target=copier)
    t.start()
    for th in (8, 16):
        dt = conv(th)
        print(f"threads {th:2d} next to an fp16 H2D loop: {n * 4 / dt / 1e9:6.1f} GB/s fp32 read")
    stop = True
    t.join()
