"""Where the host-rounded e2e step goes: per-submission staging time of HostRoundingStager with and without the device
step running next to it, for a few ring shapes.   python scripts/e2e_stage_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from prego_b200 import synthetic
from prego_b200.ingest import HostRoundingStager

dev = torch.device("cuda:0")
B, T = 4096, 64
model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20, device=dev)
rgb, flow = synthetic.device_features(B, T, dev, seed=1)
hr, hf = rgb.cpu().pin_memory(), flow.cpu().pin_memory()
del rgb, flow
h = torch.zeros(B, 1024, device=dev)
dl = torch.empty(B, T, dtype=torch.int32, device=dev)
hl = torch.empty(B, T, dtype=torch.int32).pin_memory()
for slots, sb, direct, compute, nth in ((3, 8 << 20, 0, False, 16), (3, 8 << 20, 0, True, 16), (3, 8 << 20, 0, True, 15), (3, 8 << 20, 0, True, 14),
                                        (3, 8 << 20, 0, True, 12), (3, 8 << 20, 256, True, 14), (3, 8 << 20, 256, True, 16), (3, 8 << 20, 0, True, 16)):
    st = HostRoundingStager(B, T, 2048, 2048, "fp16", dev, direct_streams=direct, ring_slots=slots, ring_slot_bytes=sb, threads=nth)
    orig = st._work
    times = []

    def timed(slot, srcs, _o=orig):
        t0 = time.perf_counter()
        _o(slot, srcs)
        times.append((time.perf_counter() - t0) * 1e3)

    st._work = timed
    n = 8
    for rep in range(2):
        times.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st.submit(0, hr, hf)
        for i in range(n):
            if i + 1 < n:
                st.submit(i + 1, hr, hf)
            if compute:
                st.infer(model, i, h_state=h, labels=dl, chunk_T=64)
                hl.copy_(dl, non_blocking=True)
            else:
                st.wait(i)
                st.release(i)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n * 1e3
    print(f"threads {nth}, ring {slots} x {sb >> 10} KiB, fp32 streams {direct:4d}, device step {'on ' if compute else 'off'}: {dt:6.1f} ms per step "
          f"({B * T / dt / 1e3:.2f} M frames/s); staging calls {', '.join(f'{t:.0f}' for t in times)} ms")
    st.close()
