"""Short run of the hot path for ncu captures: one C2-shaped batch, decoder capped at CAP frames."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
from b200taco import synth  # noqa: E402
from b200taco.model import Tacotron2  # noqa: E402

B, T, CAP = int(os.environ.get("B", 32)), int(os.environ.get("T", 150)), int(os.environ.get("CAP", 60))
torch.manual_seed(synth.WEIGHT_SEED)
m = Tacotron2().eval().to("cuda:0")
ids = synth.phoneme_ids(B, T, seed=0).cuda()
for _ in range(2):
    out = m.inference(ids, max_len_cap=CAP)
torch.cuda.synchronize()
print("frames", out[1].shape, "launches", m._engine.launch_count())
