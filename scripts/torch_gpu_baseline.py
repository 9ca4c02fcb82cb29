"""The reference's own modules (nn.Linear / LayerNorm / nn.GRU = cuBLAS + cuDNN / softmax, rnn.py:38-71) on the SAME B200,
bench shape (4096 streams x 64 frames, K = 86): the library baseline next to bench.py's number.
  python scripts/torch_gpu_baseline.py"""
import torch
import torch.nn as nn

dev = torch.device("cuda:0")
B, T, K = 4096, 64, 86


class Ref(nn.Module):
    def __init__(self):
        super().__init__()
        self.gru = nn.GRU(2048, 1024, 1, batch_first=True)
        self.layer1 = nn.Sequential(nn.Linear(4096, 2048), nn.LayerNorm(2048), nn.ReLU(), nn.Dropout(0.2))
        self.fc = nn.Linear(1024, K)

    def forward(self, rgb, flow):
        x = self.layer1(torch.cat((rgb, flow), 2))
        ht, _ = self.gru(x, torch.zeros(1, x.shape[0], 1024, device=x.device, dtype=x.dtype))
        return torch.softmax(self.fc(torch.relu(ht)), -1).argmax(-1)


torch.manual_seed(20)
m = Ref().to(dev).eval()
g = torch.Generator(device=dev).manual_seed(1)
rgb = torch.randn(B, T, 2048, generator=g, device=dev).abs_()
flow = torch.randn(B, T, 2048, generator=g, device=dev).abs_()
for mode in ("bf16 autocast", "tf32", "fp32"):
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode.startswith("bf16")):
            for _ in range(2):
                m(rgb, flow)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            n = 4
            for _ in range(n):
                m(rgb, flow)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"torch {torch.__version__} on B200, {mode}: {ms:.2f} ms per {B}x{T} step = {B * T / ms / 1e3:.2f} M frames/s", flush=True)
    except Exception as e:  # noqa
        print(mode, "failed:", repr(e)[:200], flush=True)
