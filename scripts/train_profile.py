"""Launch list of ONE training step (forward + BPTT + AdamW) for `ncu --profile-from-start off`:
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
        python scripts/train_profile.py 256 tf32
Without ncu it prints the CUDA-event time of the step."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

from prego_b200 import OadLoss, build_optimizer, synthetic, train_one_step

B, prec = int(sys.argv[1]), sys.argv[2]
dev = torch.device("cuda:0")
cfg = dict(synthetic.ASSEMBLY101_O)
model = synthetic.seeded_model(cfg, seed=20, device=dev)
model.train_precision = prec
crit = OadLoss(cfg)
opt = build_optimizer({"optimizer": "AdamW", "lr": 1e-4, "weight_decay": 0.05}, model)
T = 128
rgb, flow = synthetic.device_features(B, T, dev, seed=7, zero_flow=True)
target = torch.nn.functional.one_hot(torch.randint(0, 86, (B, T), device=dev), 86).float()
for _ in range(2):
    train_one_step(model, crit, opt, rgb, flow, target)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
train_one_step(model, crit, opt, rgb, flow, target)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"B {B} {prec}: {e0.elapsed_time(e1):.3f} ms per step")
