"""One continuous-batching decoder launch for ncu captures: K C2-shaped batches (planted stop counter, stops in ~[300, 900]
frames) stream through the 32 rows of one persistent decoder kernel (Engine.decoder_infer_stream).

    TACO_TC_CLUSTERS=0 ncu --set full --clock-control none --import-source on -k regex:decoder_tc_kernel \
        -o gpurun_out/prof_stream python profiles/tools/ncu_stream_target.py
(ncu cannot replay the cooperative thread-block-cluster launch of the production variant; TACO_TC_CLUSTERS=0 selects the same
kernel template with the split-K / attention exchanges through L2.)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
from b200taco import synth  # noqa: E402
from b200taco.model import Tacotron2  # noqa: E402

K, CAP, ROWS = int(os.environ.get("K", 3)), int(os.environ.get("CAP", 1000)), int(os.environ.get("ROWS", 32))
dev = torch.device("cuda", 0)
torch.manual_seed(synth.WEIGHT_SEED)
m = Tacotron2().eval()
m._projection_bias_initialized = True
m = m.to(dev)
ids = synth.phoneme_ids(bench.B, bench.T, seed=0).to(dev)
m.load_state_dict(bench.planted_state_dict(m, ids, synth))
eng = m._get_engine()
mem = torch.cat([eng.encoder(ids) for _ in range(K)])
torch.cuda.synchronize()
mel, gate, align, len_h = eng.decoder_infer_stream(mem, ROWS, CAP, 0.5)
torch.cuda.synchronize()
print("utterances", mem.shape[0], "frames", int(len_h[:-1].sum()), "decode steps", int(len_h[-1]))
