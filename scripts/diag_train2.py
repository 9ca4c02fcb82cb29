import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from prego_b200 import synthetic
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
B, T, K = 16, 128, 86
H, E = 1024, 2048
M = B * T
def plan():
    off = 0; d = {}
    def take(name, floats):
        nonlocal off
        d[name] = off; off += (floats * 4 + 1023) // 1024 * 1024
    for n, f in [("yn", M*E), ("e", M*E), ("rstd", M), ("gi", M*3*H), ("gh", B*3*H), ("hall", (T+1)*B*H), ("hrelu", M*H),
                 ("r", M*H), ("z", M*H), ("n", M*H), ("ghn", M*H), ("logits_tm", M*K), ("dhrelu", M*H), ("dgh", M*3*H), ("de", M*E)]:
        take(n, f)
    return d
cfg = dict(synthetic.ASSEMBLY101_O, dropout=0.0, num_classes=K)
rgb, flow = synthetic.feature_batch(list(range(100, 100 + B)), T, dev, False)
wts = torch.randn(B, T, K, generator=torch.Generator().manual_seed(1)).to(dev)
model = synthetic.seeded_model(cfg, seed=20, device=dev).train()
logits = model(rgb, flow)["logits"]
(logits * wts).sum().backward()
torch.cuda.synchronize()
ws = model._train_ws; base = (-ws.data_ptr()) % 1024
p = plan()
def view(name, shape):
    n = 1
    for s in shape: n *= s
    return ws[base + p[name]: base + p[name] + n * 4].view(torch.float32).view(*shape)
dyhat = view("de", (T, B, E)).permute(1, 0, 2)      # [B,T,E] after bwd: dyhat
dy = view("e", (T, B, E)).permute(1, 0, 2)          # dy (aliases e)
yn = view("yn", (T, B, E)).permute(1, 0, 2)
# torch reference intermediates on GPU
sd = model.state_dict()
x = torch.cat((rgb, flow), 2)
y = F.linear(x, sd["layer1.0.weight"], sd["layer1.0.bias"]).requires_grad_(True)
ln = F.layer_norm(y, (E,), sd["layer1.1.weight"], sd["layer1.1.bias"], 1e-5); ln.retain_grad()
e = F.relu(ln); e.retain_grad()
gru = torch.nn.GRU(E, H, 1, batch_first=True).to(dev)
gru.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("gru.")})
ht, _ = gru(e, torch.zeros(1, B, H, device=dev))
lg = F.linear(F.relu(ht), sd["f_classification.0.weight"], sd["f_classification.0.bias"])
(lg * wts).sum().backward()
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
print("logits", rel(logits.detach(), lg.detach()))
print("de (grad wrt relu out) vs torch e.grad: n/a (overwritten); dyhat vs ln.grad", rel(dyhat, ln.grad))
print("dy vs y.grad", rel(dy, y.grad))
d = (dyhat - ln.grad).abs()
print("dyhat mismatches > 1e-4*max:", (d > 1e-4 * ln.grad.abs().max()).sum().item(), "of", d.numel())
idx = (d > 1e-4 * ln.grad.abs().max()).nonzero()[:10]
for i in idx.tolist():
    b, t, c = i
    print(i, "ours", dyhat[b, t, c].item(), "ref", ln.grad[b, t, c].item(), "yn", yn[b, t, c].item(), "ln", ln[b, t, c].item(), "e.grad", e.grad[b, t, c].item())
