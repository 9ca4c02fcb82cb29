import sys; sys.path.insert(0, "/root/repo")
import torch, numpy as np
from prego_b200 import synthetic
dev = torch.device("cuda:0")
which = sys.argv[1]
if which == "a":   # fresh model, fp16 first
    model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20, device=dev)
    rgb, flow = synthetic.device_features(128, 1024, dev, seed=4242)
    out = model.infer(rgb, flow, want_probs=False, want_logits=True, precision="fp16", chunk_T=256)
    torch.cuda.synchronize(); print("a ok", model.device_error())
else:              # B = 1 x3 on one model, then a fresh model fp16 batched
    m1 = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20, device=dev)
    r1, f1 = synthetic.device_features(1, 500, dev, seed=1)
    m1.infer(r1, f1, want_probs=False, want_logits=True, precision="fp16x3"); torch.cuda.synchronize(); print("x3 B=1 ok")
    if which == "c":
        del m1
    model = synthetic.seeded_model(dict(synthetic.ASSEMBLY101_O), seed=20, device=dev)
    rgb, flow = synthetic.device_features(128, 1024, dev, seed=4242)
    out = model.infer(rgb, flow, want_probs=False, want_logits=True, precision="fp16", chunk_T=256)
    torch.cuda.synchronize(); print(which, "ok", model.device_error())
