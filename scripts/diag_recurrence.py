"""Where does the batched recurrence lose its time?  Runs the bench-shape step with the diagnostic knobs of
gru_seq_kernel (PREGO_GRU_DBG bits: 1 no gate math, 2 no dependency waits, 4 no gi loads, 8 no result stores /
publish; results are garbage with any bit set) and prints the recurrence phase time per step for each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from prego_b200 import synthetic

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20, device=dev)
rgb, flow = synthetic.device_features(B, T, dev, seed=1)
h = torch.zeros(B, 1024, device=dev)
for dbg in [int(x) for x in os.environ.get('DIAG_MODES', '0,1,2,4,10,5,15,0').split(',')]:  # 8 alone would starve the dependency waits
    os.environ["PREGO_GRU_DBG"] = str(dbg)
    for _ in range(2):
        model.infer(rgb, flow, h_state=h, want_probs=False, precision="fp16", chunk_T=T)
    torch.cuda.synchronize()
    model.profile_begin()
    n = 6
    for _ in range(n):
        model.infer(rgb, flow, h_state=h, want_probs=False, precision="fp16", chunk_T=T)
    p = model.profile_end()
    err = model.device_error()
    us = p["recurrence"]["ms"] / (n * T) * 1e3
    print(f"dbg={dbg:2d}  recurrence {us:6.2f} us/step  {B * 6.291456e6 / us / 1e6:7.1f} TFLOP/s   gemm1 {p['gemm1']['ms'] / n:6.3f} ms gemm2 {p['gemm2']['ms'] / n:6.3f} ms  err {err}", flush=True)
    h.zero_()

# wait-cycle counters of the kernel's roles (one launch, normal mode and handshake-only mode)
stats = torch.zeros(148, 16, dtype=torch.int64, device=dev)
os.environ["PREGO_GRU_STATS"] = str(stats.data_ptr())
for dbg in (0, 126):
    os.environ["PREGO_GRU_DBG"] = str(dbg)
    stats.zero_()
    model.infer(rgb, flow, h_state=h, want_probs=False, precision="fp16", chunk_T=T)
    torch.cuda.synchronize()
    s = stats.cpu().numpy().astype(float)
    lead, peer = s[0::2], s[1::2]
    items = lead[:, 6].mean()
    print(f"dbg={dbg}: items per pair {items:.1f}; cycles per item (leader CTA means): total {lead[:, 3].mean() / items:.0f}; "
          f"MMA thread waits: acc_empty {lead[:, 4].mean() / items:.0f}, full_bar {lead[:, 5].mean() / items:.0f}; "
          f"producer waits: dep {lead[:, 1].mean() / items:.0f}, empty {lead[:, 2].mean() / items:.0f} (peer: dep {peer[:, 1].mean() / items:.0f}, empty {peer[:, 2].mean() / items:.0f}); "
          f"epilogue warp waits: gi_full {lead[:, 9].mean() / items:.0f}, acc_full {lead[:, 10].mean() / items:.0f}, out_free {lead[:, 11].mean() / items:.0f} of {lead[:, 8].mean() / items:.0f}")
