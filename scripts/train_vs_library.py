import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench, torch, json, torch.nn as nn
dev=torch.device("cuda:0")
class Ref(nn.Module):
    def __init__(s):
        super().__init__(); s.gru=nn.GRU(2048,1024,1,batch_first=True); s.layer1=nn.Sequential(nn.Linear(4096,2048),nn.LayerNorm(2048),nn.ReLU(),nn.Dropout(0.2)); s.fc=nn.Linear(1024,86)
print(json.dumps(bench.library_train_step(dev, Ref), indent=0))
o=bench.training_leg(dev,1); o.pop("note"); print(json.dumps(o, indent=0))
