"""Per-kernel device time of ONE C5 training step (B=64, T=150, 800-frame mels), collected with torch.profiler (CUPTI
activity records: real back-to-back durations with warm caches, unlike ncu's serialised cold-cache replays).

    python profiles/tools/c5_kernel_times.py [--out gpurun_out/c5_kernel_times.md]

Prints a markdown table: kernel, launches, total ms, mean us, share of the summed kernel time; plus the CUDA-event time of
the same step without the profiler attached (the number that counts).
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from b200taco import synth  # noqa: E402
from b200taco.loss import Tacotron2Loss, gate_targets  # noqa: E402
from b200taco.model import Tacotron2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=800)
    ap.add_argument("--text", type=int, default=150)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(synth.WEIGHT_SEED)
    model = Tacotron2()
    model._projection_bias_initialized = True
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    crit = Tacotron2Loss()
    B, S, T = args.batch, args.frames, args.text
    ids = synth.phoneme_ids(B, T, seed=1000).to(dev)
    g = torch.Generator().manual_seed(100)
    mel = (torch.randn(B, 80, S, generator=g) * 2 - 5).to(dev)
    tl = torch.full((B,), T, dtype=torch.long, device=dev)
    ml = torch.full((B,), S, dtype=torch.long, device=dev)
    gt = gate_targets(ml, S)

    def step():
        out = model(ids, mel, tl)
        total, _, _, _ = crit(out, (mel, gt, ml), text_lengths=tl)
        opt.zero_grad(set_to_none=True)
        total.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return total

    step()
    step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize(dev)
    plain_ms = e0.elapsed_time(e1)

    phase_text = ""
    if os.environ.get("TACO_ATTN_TIMES") == "1":
        import ctypes
        from b200taco import _cabi
        buf = (ctypes.c_uint64 * 8)()
        _cabi.lib().b200taco_attn_bwd_phase_ns(buf, 1)
        step()
        _cabi.lib().b200taco_attn_bwd_phase_ns(buf, 1)
        names = ["loads", "location features", "wait + dctx + dw/s", "reduction + cluster exchange 1", "per-position pass",
                 "reductions + halo exchange", "location-filter gradient", "location-input gradient"]
        phase_text = "\n\nattention_bwd_v3_kernel, CTA 0, us per launch (TIMED instantiation): " + ", ".join(
            f"{n} {buf[i] / S / 1e3:.2f}" for i, n in enumerate(names)) + f"; sum {sum(buf) / S / 1e3:.2f}"

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize(dev)
    agg = collections.OrderedDict()
    t_first, t_last = None, None
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = ev.name
        short = name.replace("b200taco::", "").replace("(anonymous namespace)::", "")
        short = short.split("(")[0]
        if short.startswith("void "):
            short = short[5:]
        d = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        n, tot = agg.get(short, (0, 0.0))
        agg[short] = (n + 1, tot + d)
        tr = ev.time_range
        t_first = tr.start if t_first is None else min(t_first, tr.start)
        t_last = tr.end if t_last is None else max(t_last, tr.end)
    total = sum(t for _, t in agg.values())
    lines = [f"C5 step B={B} T={T} S={S}: {plain_ms:.1f} ms by CUDA events without the profiler; under torch.profiler the kernels "
             f"sum to {total / 1e3:.1f} ms over a span of {(t_last - t_first) / 1e3:.1f} ms "
             f"({sum(n for n, _ in agg.values())} device activities)", "",
             "| kernel | launches | total ms | mean us | share |", "|---|---|---|---|---|"]
    for name, (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if tot / total < 0.001:
            continue
        lines.append(f"| `{name[:110]}` | {n} | {tot / 1e3:.2f} | {tot / n:.1f} | {100 * tot / total:.1f} % |")
    text = "\n".join(lines) + phase_text
    print(text)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
