#!/usr/bin/env python
"""Benchmark of the Tacotron 2 hot path (BASELINE.json metric: mel frames/sec at batch 32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one full free-running inference of one batch of BASELINE config 2 (C2):
32 synthetic utterances x 150 phonemes, cap 1000 decoder frames, per-utterance stop-gate
early exit (planted stop counter, SURVEY.md 8c/8d, stops land in ~[300, 900] frames):
encoder -> persistent decoder kernel -> postnet.  Frames counted = sum of the utterances'
own decoded lengths.

  value         frames/s with the phoneme ids already resident in HBM (device timed, CUDA events); the K
                batches go through the public corpus path (Tacotron2.inference_batches(continuous=True) ->
                Engine.stream_infer): ONE persistent decoder launch keeps 32 rows live and refills a row from
                the following batches as soon as its utterance stops (continuous batching; every utterance
                is bit-identical to its batch-by-batch decode, tests/test_gpu_parity.py)
  batch_by_batch  the same K batches with one decoder launch per batch (rows idle behind early stops),
                encoder(i+1) / postnet(i-1) overlapping decoder(i) on separate streams
  e2e           same metric through the same public API with HOST pinned ids, host->device copy and
                device->host read of the results (pinned) inside the timed region
  roofline      persistent decoder kernel: algorithmic bytes per decode step x
                decode steps per launch / CUDA-event duration, against the measured HBM copy
                bandwidth of MEASURED_PEAKS.json (the per-step working set is L2 resident, so the
                live-measured L2 read bandwidth is reported beside it)
  cpu_baseline  the CPU oracle port of the reference (torch CPU ops, all host threads) on a
                bounded sample of the same workload, rank 0, N=1 only
  b1            BASELINE config 3: B=1, T=200 decoder step latency in microseconds

--impl reference times the reference's own CPU implementation of the path: the unmodified reference model staged in
oracle/_ref by oracle/make_ref.py (the oracle port only if that was never staged), same ids / weights / cap.
N>1 (under torchrun): every rank decodes its own batches (weak scaling, no data-path collective).
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "mel frames/sec (batch 32)"
UNIT = "frames/s"
B, T, CAP = 32, 150, 1000
PLANT = dict(rate=8e-4, jitter=6e-3)                 # stops in ~[300, 900] frames (synth.plant_stop_counter_)
WORKLOAD = "C2: free-running inference, B=32 x T=150 phonemes, cap 1000 frames, per-utterance stop-gate early exit"
# algorithmic bytes per decoder step, bf16 LSTM weights + fp32 activations
BYTES_PER_STEP_C2 = 36_379_942 + B * T * (128 + 512) * 4 + B * T * 16 + B * 37_508
BYTES_PER_STEP_C3 = 36_379_942 + 1 * 200 * (128 + 512) * 4 + 200 * 16 + 37_508


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        if os.environ.get("BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
            t_end = time.time() + 3.0
            while not self.lines and time.time() < t_end:      # first sample = nvidia-smi is up
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx.append(int(f[1]))
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def planted_state_dict(model, ids_dev, synth):
    """Default init + planted stop counter centred on this batch's mean context."""
    with torch.no_grad():
        center = model._get_engine().encoder(ids_dev).mean((0, 1)).cpu()
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        synth.plant_stop_counter_(sd, ctx_center=center, **PLANT)
        synth.round_to_bf16_(sd)
    return sd


def run_b200(args, rank, world, local_rank):
    from b200taco import synth
    from b200taco.engine import STOP_PER_UTTERANCE
    from b200taco.model import Tacotron2

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(synth.WEIGHT_SEED)
    model = Tacotron2().eval()
    model._projection_bias_initialized = True
    model = model.to(dev)
    # weak scaling: per-GPU work is FIXED as N grows, so every rank decodes its own copy of the same synthetic
    # batch (corpus-level sharding with unequal utterances is BASELINE config 4, profiles/tools/c4_sharded.py)
    ids_host = synth.phoneme_ids(B, T, seed=0).pin_memory()
    ids_dev = ids_host.to(dev)
    model.load_state_dict(planted_state_dict(model, ids_dev, synth))
    eng = model._get_engine()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def step_resident():
        """value: inputs resident in HBM; returns (frames, decoder-kernel event pair)."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        flush.zero_()
        ev[1].record()
        mem = eng.encoder(ids_dev)
        ev[2].record()
        coarse, gate, align, lengths, S = eng.decoder_infer(mem, CAP, 0.5, STOP_PER_UTTERANCE, sync=False)
        ev[3].record()
        S = int(S.item())
        post = eng.postnet(coarse[:, :S].contiguous())
        ev[4].record()
        return lengths, S, ev, post

    def flush_l2():
        flush.zero_()

    def run_pipelined(n, host, continuous=True):
        """n batches through the public corpus path (Tacotron2.inference_batches -> Engine.stream_infer: continuous
        batching, 32 live rows refilled as utterances stop; continuous=False -> Engine.pipeline_infer: one decoder launch
        per batch).  Results of every batch are complete inside this call.
        Returns (frames, decoder event pairs, decode steps executed, last result)."""
        timings, lens, Ss, last = [], [], [], None
        feed = [ids_host if host else ids_dev] * n
        with contextlib.redirect_stdout(io.StringIO()):
            if continuous:
                it = eng.stream_infer(feed, CAP, 0.5, to_host=host, pre_hook=flush_l2, timings=timings)
            else:
                it = eng.pipeline_infer(feed, CAP, 0.5, STOP_PER_UTTERANCE, to_host=host, pre_hook=flush_l2, timings=timings)
            for res in it:
                lens.append(res[7] if host else res[4])
                Ss.append(res[5])
                last = res
        frames_ = int(torch.stack(lens).sum().item())
        steps_ = eng.last_stream_steps if continuous else sum(S + 1 for S in Ss)   # + the discarded first step (F6)
        return frames_, timings, steps_, last

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value (device resident ids, K batches pipelined) ----------------
    # the sampler starts before the warm-up: nvidia-smi's own start-up stalls CUDA calls of this process for
    # ~20 ms (measured), which must not land inside the timed region; its samples cover warm-up + timed steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    # warm-up: W batches, and never fewer than the timed K: one continuous-batching launch allocates its output buffers for all
    # its batches at once, and the caching allocator must have seen that size before the timed region
    run_pipelined(max(args.warmup, args.steps), False)
    barrier()
    eng.launch_count(reset=True)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    frames, dec_events, dec_steps, _ = run_pipelined(args.steps, False)
    t1.record()
    barrier()
    launches = eng.launch_count()
    clocks = sampler.stop()
    ms = t0.elapsed_time(t1)
    dec_ms = sum(a.elapsed_time(b) for a, b in dec_events)

    # ---------------- the same K batches, one decoder launch per batch (informational) ----------------
    run_pipelined(1, False, continuous=False)
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    bb_frames, bb_events, bb_steps, _ = run_pipelined(args.steps, False, continuous=False)
    b1.record()
    barrier()
    bb_ms = b0.elapsed_time(b1)
    bb_dec_ms = sum(a.elapsed_time(b) for a, b in bb_events)

    # ---------------- device time of the stages of ONE batch run one after the other (informational) ----------------
    stage_runs = [step_resident() for _ in range(3)]
    torch.cuda.synchronize(dev)
    stages_ms = {n: sum(r[2][i].elapsed_time(r[2][i + 1]) for r in stage_runs) / len(stage_runs)
                 for i, n in enumerate(("l2_flush", "encoder", "decoder", "postnet"))}

    # ---------------- e2e (pinned host ids in, pinned host results out, public API) ----------------
    run_pipelined(args.steps, True)            # warm-up incl. the pinned result buffers of all K batches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_frames, _, _, last = run_pipelined(args.steps, True)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    d2h = last[6].numel() * 4 + B * 4

    agg = torch.tensor([ms, e2e_ms, float(frames), float(e2e_frames), float(launches)], device=dev, dtype=torch.float64)
    if dist is not None:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(mx[0]), float(mx[1])
        frames, e2e_frames, launches = float(sm[2]), float(sm[3]), float(sm[4])

    out = None
    if rank == 0:
        hbm, hbm_src = _peaks()
        achieved = BYTES_PER_STEP_C2 * dec_steps / (dec_ms * 1e-3) / 1e9
        l2_gbps = eng.measure_l2_read_gbps(36 << 20, 40)
        out = {
            "metric": METRIC, "value": frames / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (decoder LSTM weights stored bf16, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "text_len": T, "cap": CAP,
                       "weights": "seed 1234 default init + planted stop counter (rate 8e-4, jitter 6e-3)",
                       "frames_counted": "sum of per-utterance decoded lengths",
                       "warmup_batches": max(args.warmup, args.steps),
                       "l2": "flushed before every decoder launch (256 MiB device memset on the decoder stream); the K batches' "
                             "encoder memory streams through the launch (123 MB at K=10)",
                       "pipeline": "continuous batching: the K batches' utterances stream through the 32 rows of ONE persistent "
                                   "decoder launch, a row is refilled from the next batch when its utterance stops; "
                                   "encoders before, postnets after the launch",
                       "sharding": f"dp{world}: one batch per rank per step, same synthetic batch on every rank, no collective"},
            "e2e": {"value": e2e_frames / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": ids_host.numel() * 8,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "batch_by_batch": {"value": bb_frames / (bb_ms * 1e-3), "unit": UNIT, "ms_per_step": bb_ms / args.steps,
                               "us_per_decode_step": bb_dec_ms * 1e3 / bb_steps, "decode_steps": bb_steps,
                               "what": "one persistent decoder launch per batch (rows idle behind early stops), rank 0"},
            "stages_ms": stages_ms,
            "clocks": clocks,
            "roofline": {"kernel": "decoder_tc_kernel<32> (persistent, one cooperative cluster launch for the K batches)", "bound": "hbm",
                         "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one continuous-batching launch (K=3 batches = 96
                         # utterances, 2002 decode steps), ncu --set full, profiles/r1_v15_stream_decoder_ncu_full.md: 66.9 KB
                         # of HBM traffic per decode step against 49.9 MB algorithmic - the working set is L2 resident
                         "traffic": 133_841_152, "traffic_unit": "bytes per launch (ncu, launch of 96 utterances / 2002 decode steps)",
                         "peak_source": hbm_src, "bytes_per_decode_step": BYTES_PER_STEP_C2,
                         "decode_steps_per_launch": dec_steps / max(1, len(dec_events)), "us_per_decode_step": dec_ms * 1e3 / dec_steps,
                         "row_occupancy": (frames / world + args.steps * B) / (dec_steps * B),
                         "kernel_share_of_step": dec_ms / ms,
                         "l2_read_peak_gbs_measured_live": l2_gbps, "frac_of_l2_read_peak": achieved / l2_gbps,
                         "note": "working set per decode step is L2 resident; HBM peak is the schema's denominator"},
        }
        if world == 1:
            out["b1"] = bench_b1(model, eng, synth, dev, l2_gbps, hbm)
            out["cpu_baseline"] = cpu_baseline(synth, steps=3)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out


def bench_b1(model, eng, synth, dev, l2_gbps, hbm):
    """BASELINE config 3: one 200-phoneme utterance, default init (gate never fires) -> 1000 frames."""
    from b200taco.engine import STOP_BATCH_WIDE
    from b200taco.model import Tacotron2
    torch.manual_seed(synth.WEIGHT_SEED)
    m = Tacotron2().eval()
    m._projection_bias_initialized = True
    with torch.no_grad():
        synth.round_to_bf16_(m.state_dict())
    m = m.to(dev)
    e = m._get_engine()
    ids = synth.phoneme_ids(1, 200, seed=3).to(dev)
    mem = e.encoder(ids)
    times = []
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, _, _, _, S = e.decoder_infer(mem, 1000, 0.5, STOP_BATCH_WIDE, sync=False)
        e1.record()
        S = int(S.item())
        if i >= 3:
            times.append(e0.elapsed_time(e1) * 1e3 / (S + 1))
    us = sum(times) / len(times)
    ach = BYTES_PER_STEP_C3 / (us * 1e-6) / 1e9
    return {"workload": "C3: B=1, T=200, 1000 frames, persistent decoder", "us_per_decode_step": us,
            "achieved_gbs": ach, "l2_read_peak_gbs_measured_live": l2_gbps, "l2_roofline_us": BYTES_PER_STEP_C3 / l2_gbps / 1e3,
            "x_of_l2_roofline": us / (BYTES_PER_STEP_C3 / l2_gbps / 1e3), "frac_of_hbm_peak": ach / hbm}


def _cpu_weights(synth, encoder_fn):
    """The GPU arm's weight set on the CPU: seed 1234 default init + planted stop counter centred on this batch's mean
    encoder output (computed by `encoder_fn` on the CPU; the GPU arm centres on its own encoder output, equal to ~1e-6),
    decoder LSTM matrices rounded to bf16-representable values."""
    from b200taco.model import Tacotron2
    sd = synth.seeded_state_dict(Tacotron2)
    ids = synth.phoneme_ids(B, T, seed=0)
    with torch.no_grad():
        center = encoder_fn(sd, ids).mean((0, 1))
        synth.plant_stop_counter_(sd, ctx_center=center, **PLANT)
        synth.round_to_bf16_(sd)
    return sd, ids


def cpu_baseline(synth, steps, cap=CAP):
    """The reference's own CPU path on the host cores, same ids / weights / cap as the GPU arm.  `kind: "reference"` = the
    unmodified reference model staged in oracle/_ref by oracle/make_ref.py (Tacotron2.inference, src/model.py:482);
    `kind: "port"` (fallback when oracle/_ref was never staged) = oracle/tacotron2_oracle.py, which issues the same ATen ops.
    The reference stops the whole batch when the FIRST row's gate fires (F7, src/model.py:327) and returns B x S frames;
    its per-step cost does not depend on the step index, so B x S / t is its frame rate at batch 32 whatever S is."""
    from oracle import tacotron2_oracle as O
    from oracle.make_ref import load_reference
    torch.set_num_threads(os.cpu_count() or 1)
    ref = load_reference()
    if ref is not None:
        model = ref.Tacotron2().eval()
        model._projection_bias_initialized = True

        def enc(sd, ids):
            model.load_state_dict(sd)
            return model.encoder(ids)
        sd, ids = _cpu_weights(synth, enc)
        model.load_state_dict(sd)
        run = lambda c: model.inference(ids, max_len_cap=c)           # noqa: E731
        kind = "reference"
    else:
        sd, ids = _cpu_weights(synth, O.encoder)
        run = lambda c: O.tacotron2_inference(sd, ids, max_len_cap=c)  # noqa: E731
        kind = "port"
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):      # F6 debug prints
        run(10)                                                           # warm-up
        t0 = time.perf_counter()
        for _ in range(steps):
            out = run(cap)
        dt = time.perf_counter() - t0
    S = out[1].shape[1]
    return {"value": steps * B * S / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} x Tacotron2.inference(ids(32,150), max_len_cap={cap}) with the GPU arm's ids and weights (planted "
                      f"stop counter): encoder + {S + 1} decode steps + postnet, batch-wide stop after {S} frames (F7), "
                      f"frames counted = 32 x {S}",
            "frames_per_call": B * S, "seconds": dt}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    from b200taco import synth
    t0 = time.perf_counter()
    r = cpu_baseline(synth, max(1, args.steps))
    ms = (time.perf_counter() - t0) * 1e3
    return {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3 / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "text_len": T, "cap": CAP,
                       "weights": "seed 1234 default init + planted stop counter (rate 8e-4, jitter 6e-3)",
                       "frames_counted": "B x S of the reference's own batch-wide stop (every row is still unfinished at S)",
                       "sample": r["sample"], "device": "host CPU", "arm": r["kind"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_ms": ms}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: anything a library prints meanwhile (NCCL's version banner, the
    # reference's debug prints) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        out = run_reference(args, rank, world)
    else:
        if world != args.gpus and rank == 0:
            print(f"[bench] WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
        out = run_b200(args, rank, world, local_rank)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
