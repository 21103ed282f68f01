"""Timeline of ONE decode step of the weight-resident small-batch decoder (decoder_res.cu): per-CTA globaltimer stamps
(TACO_RES_TRACE=<step>) reduced to role-level statistics, all times in ns relative to the moment the prenet output was
complete in the first gate CTA.

    TACO_RES_TRACE=200 python profiles/tools/res_trace.py [B,T]
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
from b200taco import synth  # noqa: E402
from b200taco._cabi import check  # noqa: E402
from b200taco.engine import STOP_BATCH_WIDE  # noqa: E402
from b200taco.model import Tacotron2  # noqa: E402

GATE = ["p in", "ah out", "ctx in", "dh out", "bg dh done", "bg ah done", "next p in"]
ATTN = ["pre done", "ah in (t0)", "pq done", "ep out", "ep in", "ew done", "ctx part", "ctx out", "w done"]
FRAME = ["ctx in (t0)", "dh in (t0)", "mel out", "mel in", "y1 done", "p out", "", "", "all in+reduced"]


def main():
    B, T = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,200").split(","))
    step = int(os.environ.get("TACO_RES_TRACE", "200"))
    os.environ["TACO_RES_TRACE"] = str(step)
    torch.manual_seed(synth.WEIGHT_SEED)
    m = Tacotron2().eval().to("cuda:0")
    eng = m._get_engine()
    mem = eng.encoder(synth.phoneme_ids(B, T, seed=1).cuda())
    for _ in range(2):
        eng.decoder_infer(mem, step + 50, 0.5, STOP_BATCH_WIDE, sync=True)
    EV = 64
    n = 160 * EV + 1024
    out = (C.c_int64 * n)()
    check(eng._lib.b200taco_decoder_trace(eng._h, out, n, eng._stream()), "b200taco_decoder_trace")
    ev = [[out[c * EV + i] for i in range(EV)] for c in range(148)]
    t0 = min(e[0] for e in ev[:128])
    natt = 148 - 128 - 8
    natt = natt // B * B

    def stats(rows, names):
        res = {}
        for i, nm in enumerate(names):
            if not nm:
                continue
            v = [r[i] - t0 for r in rows if r[i]]
            if v:
                res[nm] = {"min": min(v), "max": max(v), "mean": round(sum(v) / len(v))}
        return res
    for slot, nm in ((0, "attention cta 128: per-thread ah arrival"), (1, "frame cta 140: per-thread dh arrival")):
        v = [out[160 * EV + slot * 512 + i] - t0 for i in range(512)]
        print(nm, "by warp (min, max):", [(min(v[w * 32:w * 32 + 32]), max(v[w * 32:w * 32 + 32])) for w in range(16)])
    print("gate p-in  :", [ev[c][0] - t0 for c in range(128)])
    print("gate ah-out:", [ev[c][1] - t0 for c in range(128)])
    print("gate ctx-in:", [ev[c][2] - t0 for c in range(128)])
    print("gate dh-out:", [ev[c][3] - t0 for c in range(128)])
    for c in (128, 133, 139, 140, 147):
        print("cta", c, "per-warp ep poll rounds:", [ev[c][16 + w] for w in range(16)] if c < 140 else "", "per-warp words in:", [ev[c][16 + w] - t0 for w in range(16)], "reduced:", [ev[c][32 + w] - t0 for w in range(16)],
              "ep gathered:", [ev[c][48 + w] - t0 for w in range(16)])
    print(json.dumps({"B": B, "T": T, "step": step, "gate": stats(ev[:128], GATE), "attention": stats(ev[128:128 + natt], ATTN),
                      "frame": stats(ev[128 + natt:128 + natt + 8], FRAME)}, indent=1))


if __name__ == "__main__":
    main()
