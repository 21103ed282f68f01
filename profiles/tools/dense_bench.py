"""Device time of the encoder and the postnet on both dense paths (tcgen05 3xTF32 vs fp32 SIMT).

    python profiles/tools/dense_bench.py [B,T,S ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
from b200taco import synth  # noqa: E402
from b200taco.model import Tacotron2  # noqa: E402


def timed(fn, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def main():
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(32, 150, 842), (1, 200, 1000), (64, 150, 800)]
    outs = {}
    for path in ("simt", "auto"):
        torch.manual_seed(synth.WEIGHT_SEED)
        m = Tacotron2().eval()
        m.dense_path = path
        m = m.to("cuda:0")
        eng = m._get_engine()
        for B, T, S in shapes:
            ids = synth.phoneme_ids(B, T, seed=1).cuda()
            g = torch.Generator().manual_seed(5)
            mel = (torch.randn(B, S, 80, generator=g) * 2 - 5).cuda()
            t_enc, mem = timed(lambda: eng.encoder(ids))
            t_post, post = timed(lambda: eng.postnet(mel))
            key = (B, T, S)
            rec = {"dense_path": path, "B": B, "T": T, "S": S, "encoder_ms": round(t_enc, 4), "postnet_ms": round(t_post, 4),
                   "postnet_tflops_fp32_equiv": round(B * S * 8.68e6 / (t_post * 1e-3) / 1e12, 2)}
            if key in outs:
                rec["max_abs_diff_vs_simt"] = {"memory": float((mem - outs[key][0]).abs().max()),
                                               "postnet": float((post - outs[key][1]).abs().max())}
            else:
                outs[key] = (mem, post)
            print(json.dumps(rec))


if __name__ == "__main__":
    main()
