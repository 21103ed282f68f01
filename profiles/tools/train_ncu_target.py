import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200")): sys.path.insert(0, p)
import torch
from b200taco import synth
from b200taco.model import Tacotron2
B, S, Tn = 64, 40, 150
dev = torch.device("cuda:0")
torch.manual_seed(1234)
m = Tacotron2(); m._projection_bias_initialized = True; m = m.to(dev).train()
ids = synth.phoneme_ids(B, Tn, seed=0).to(dev)
mel = (torch.randn(B, 80, S) * 2 - 5).to(dev)
lens = torch.full((B,), Tn, dtype=torch.long, device=dev)
for _ in range(2):
    post, coarse, gate, aligns = m(ids, mel, lens)
    ((post - mel.transpose(1, 2)).abs().mean() + gate.mean()).backward()
torch.cuda.synchronize()
print("ok")
