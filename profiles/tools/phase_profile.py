"""Per-phase device time of the persistent decoder (b200taco_decoder_profile) for several shapes.

    python profiles/tools/phase_profile.py [B,T,steps ...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
from b200taco import synth  # noqa: E402
from b200taco.engine import STOP_BATCH_WIDE  # noqa: E402
from b200taco.model import Tacotron2  # noqa: E402


def main():
    shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(1, 200, 300), (8, 150, 300), (32, 150, 300), (64, 150, 300)]
    dtypes = os.environ.get("WDT", "bf16").split(",")
    for wdt in dtypes:
        torch.manual_seed(synth.WEIGHT_SEED)
        m = Tacotron2().eval()
        m.lstm_weight_dtype = wdt
        m = m.to("cuda:0")
        eng = m._get_engine()
        for B, T, S in shapes:
            ids = synth.phoneme_ids(B, T, seed=1).cuda()
            mem = eng.encoder(ids)
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if os.environ.get("STREAM") and eng.stream_supported(B):     # utterance-queue instantiation, no refills
                    eng.decoder_infer_stream(mem, B, S, 0.5)
                else:
                    eng.decoder_infer(mem, S, 0.5, STOP_BATCH_WIDE, sync=False)
                e1.record()
                torch.cuda.synchronize()
            prof = eng.decoder_profile()
            tot = e0.elapsed_time(e1) * 1e3 / prof["steps"]
            print(json.dumps({"wdt": wdt, "B": B, "T": T, "steps": prof["steps"], "us_per_step_total": round(tot, 2),
                              "phases_us": {k: round(v, 2) for k, v in prof["us_per_step"].items()},
                              **({"fine": prof["fine_us_per_step"]} if "fine_us_per_step" in prof else {})}))


if __name__ == "__main__":
    main()
