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
                Engine.stream_infer): ONE persistent decoder launch keeps 64 rows live (two 32-utterance batches side by
                side; BENCH_LIVE_ROWS=32 for one) and refills a row from
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
  c4            BASELINE config 4 (every N): 8192 synthetic utterances, T ~ U{50..250}, sharded over the ranks (strong scaling,
                no data-path collective), ragged continuous batching (Tacotron2.inference_corpus, 64 live rows)
  c5            BASELINE config 5 (every N): teacher-forced training step fwd + loss + bwd + clip + Adam, B=64 / GPU, 800-frame
                mels, T=150, gradients all-reduced by DistributedDataParallel over NCCL (weak scaling); ms/step, TFLOP/s against
                the measured sustained bf16 peak, the all-reduce time left exposed, and at N=1 the reference's CPU step
--workload c4 | c5 makes that leg the JSON line's headline (metric / value / unit); --skip-extra drops the b1 / c4 / c5 legs.

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
LIVE_ROWS = int(os.environ.get("BENCH_LIVE_ROWS", "64"))      # decoder rows one continuous-batching launch keeps busy
PLANT = dict(rate=8e-4, jitter=6e-3)                 # stops in ~[300, 900] frames (synth.plant_stop_counter_)
WORKLOAD = "C2: free-running inference, B=32 x T=150 phonemes, cap 1000 frames, per-utterance stop-gate early exit"
# algorithmic bytes per decoder step, bf16 LSTM weights + fp32 activations
BYTES_PER_STEP_C2 = 36_379_942 + B * T * (128 + 512) * 4 + B * T * 16 + B * 37_508


def bytes_per_step(rows):
    """SURVEY.md 8(d) for `rows` utterances decoding side by side: the weights once, the activations per utterance."""
    return 36_379_942 + rows * T * (128 + 512) * 4 + rows * T * 16 + rows * 37_508
BYTES_PER_STEP_C3 = 36_379_942 + 1 * 200 * (128 + 512) * 4 + 200 * 16 + 37_508


# dram__bytes_read.sum + dram__bytes_write.sum of ONE continuous-batching launch, ncu --set full (profiles/r2_decoder_kernels_ncu_full.md)
NCU_TRAFFIC = {32: 102_501_120, 64: 1_463_833_400}
NCU_TRAFFIC_UNIT = {32: "bytes per launch (ncu, production cluster kernel, launch of 96 utterances / 2002 decode steps)",
                    64: "bytes per launch (ncu, production cluster kernel, launch of 128 utterances / 1475 decode steps = 0.99 MB of HBM "
                        "traffic per decode step against 63.5 MB algorithmic: the working set is L2 resident)"}


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

    def run_pipelined(n, host, continuous=True, everything=False):
        """n batches through the public corpus path (Tacotron2.inference_batches -> Engine.stream_infer: continuous
        batching, 32 live rows refilled as utterances stop; continuous=False -> Engine.pipeline_infer: one decoder launch
        per batch).  Results of every batch are complete inside this call.
        Returns (frames, decoder event pairs, decode steps executed, last result)."""
        timings, lens, Ss, last = [], [], [], None
        feed = [ids_host if host else ids_dev] * n
        with contextlib.redirect_stdout(io.StringIO()):
            if continuous:
                it = eng.stream_infer(feed, CAP, 0.5, to_host="all" if everything else host, pre_hook=flush_l2, timings=timings,
                                      live_rows=LIVE_ROWS)
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
    # the same with the reference's WHOLE 4-tuple copied to the host (mel_coarse, gate logits and alignments too)
    run_pipelined(args.steps, True, everything=True)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    full_frames, _, _, last_full = run_pipelined(args.steps, True, everything=True)
    f1.record()
    barrier()
    full_ms = f0.elapsed_time(f1)
    d2h_full = sum(t.numel() * t.element_size() for t in last_full[6:])

    agg = torch.tensor([ms, e2e_ms, float(frames), float(e2e_frames), float(launches), full_ms, float(full_frames)], device=dev,
                       dtype=torch.float64)
    if dist is not None:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(mx[0]), float(mx[1])
        frames, e2e_frames, launches = float(sm[2]), float(sm[3]), float(sm[4])
        full_ms, full_frames = float(mx[5]), float(sm[6])

    out = None
    if rank == 0:
        hbm, hbm_src = _peaks()
        live = B * max(m for m in range(max(1, LIVE_ROWS // B), 0, -1) if args.steps % m == 0)   # Engine.stream_infer's choice
        achieved = bytes_per_step(live) * dec_steps / (dec_ms * 1e-3) / 1e9
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
                       "pipeline": f"continuous batching: the K batches' utterances stream through the {live} rows of ONE persistent "
                                   f"decoder launch ({live // B} batch(es) of 32 side by side), a row is refilled from the next "
                                   "batch when its utterance stops; encoders before, postnets after the launch",
                       "live_rows": live,
                       "sharding": f"dp{world}: one batch per rank per step, same synthetic batch on every rank, no collective"},
            "e2e": {"value": e2e_frames / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": ids_host.numel() * 8,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "returns": "mel_postnet (B,S,80) + per-utterance lengths per batch, pinned host memory; mel_coarse, gate logits and "
                               "alignments of the reference's 4-tuple stay on the device (yielded as device tensors)"},
            "e2e_full": {"value": full_frames / (full_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": ids_host.numel() * 8,
                         "d2h_bytes_per_step": d2h_full, "ms_per_step": full_ms / args.steps,
                         "returns": "the reference's whole 4-tuple per batch in pinned host memory: mel_postnet, mel_coarse, gate logits, "
                                    "alignments (S,32,150), + lengths"},
            "gpu_launches": int(launches),
            "batch_by_batch": {"value": bb_frames / (bb_ms * 1e-3), "unit": UNIT, "ms_per_step": bb_ms / args.steps,
                               "us_per_decode_step": bb_dec_ms * 1e3 / bb_steps, "decode_steps": bb_steps,
                               "what": "one persistent decoder launch per batch (rows idle behind early stops), rank 0"},
            "stages_ms": stages_ms,
            "clocks": clocks,
            "roofline": {"kernel": f"decoder_tc_kernel<{live}> (persistent, one cooperative cluster launch for the K batches)", "bound": "hbm",
                         "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one continuous-batching launch (K=3 batches = 96
                         # utterances, 2002 decode steps), ncu --set full, profiles/r1_v15_stream_decoder_ncu_full.md: 66.9 KB
                         # of HBM traffic per decode step against 49.9 MB algorithmic - the working set is L2 resident
                         "traffic": NCU_TRAFFIC.get(live), "traffic_unit": NCU_TRAFFIC_UNIT.get(live),
                         "peak_source": hbm_src, "bytes_per_decode_step": bytes_per_step(live),
                         "decode_steps_per_launch": dec_steps / max(1, len(dec_events)), "us_per_decode_step": dec_ms * 1e3 / dec_steps,
                         "row_occupancy": (frames / world + args.steps * B) / (dec_steps * live),
                         "kernel_share_of_step": dec_ms / ms,
                         "l2_read_peak_gbs_measured_live": l2_gbps, "frac_of_l2_read_peak": achieved / l2_gbps,
                         "note": "working set per decode step is L2 resident; HBM peak is the schema's denominator"},
        }
        if world == 1:
            if not args.skip_extra:
                out["b1"] = bench_b1(model, eng, synth, dev, l2_gbps, hbm)
            out["cpu_baseline"] = cpu_baseline(synth, steps=3)
    if not args.skip_extra or args.workload != "c2":
        # the other two multi-GPU configurations of BASELINE.json, driver visible: every rank takes part
        c4 = c5 = None
        if args.workload in ("c2", "c4"):
            c4 = bench_c4(model, synth, dev, rank, world, dist)
        del eng, model, flush
        torch.cuda.empty_cache()
        if args.workload in ("c2", "c5"):
            c5 = bench_c5(synth, dev, rank, world, dist, local_rank)
        if rank == 0:
            if c4 is not None:
                out["c4"] = c4
            if c5 is not None:
                out["c5"] = c5
                if world == 1:
                    cb = cpu_baseline_c5(synth)
                    if cb is not None:
                        out["c5"]["cpu_baseline"] = cb
            head = {"c4": c4, "c5": c5}.get(args.workload)
            if head is not None:                       # --workload c4 | c5: that leg is the headline of the line
                out["c2"] = {k: out[k] for k in ("metric", "value", "unit", "ms_per_step", "scaling")}
                out.update(metric=head["metric"], value=head["value"], unit=head["unit"], scaling=head["scaling"],
                           ms_per_step=head.get("ms_per_step", head.get("seconds", 0.0) * 1e3))
                out["config"] = dict(out["config"], workload=head["workload"])
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out


C4_UTTS, C4_ROWS = 8192, 64
C5_B, C5_S, C5_T = 64, 800, 150
C5_TFLOP = 7.55                     # fwd + bwd of one C5 step


def bench_c4(model, synth, dev, rank, world, dist, n_utts=C4_UTTS, rows=C4_ROWS):
    """BASELINE config 4: the corpus is sorted by text length and dealt utterance by utterance to the ranks
    (every rank sees the same length distribution: strong scaling without a collective); each rank streams its share through persistent launches of 64
    live rows, utterances of different lengths side by side (ragged continuous batching), at least ~16 refills per row and
    launch (an utterance shorter than the launch's padded length only pays for its own positions).  Timed on the device, max over ranks; frames = sum of the decoded lengths of all utterances."""
    import random
    rng = random.Random(0)
    lens = [rng.randint(50, 250) for _ in range(n_utts)]
    # the length-sorted corpus dealt utterance by utterance in snake order (b200taco.shard.deal_sorted): the ranks' summed lengths
    # agree to 0.06 %.  (Before the encoder took ragged batches an utterance was encoded in a batch of its own exact length, and
    # whole exact-length groups had to go to one rank - shard.deal_length_groups, 4.2 % imbalance.)
    from b200taco import shard
    mine = shard.deal_sorted(lens, world, rank)
    corpus = {i: synth.phoneme_ids(1, lens[i], seed=i)[0] for i in mine}
    n_launch = max(1, min(4, len(mine) // (16 * rows)))           # >= ~16 refills per row and launch: the drain of a launch is fixed
    per_launch = (len(mine) + n_launch - 1) // n_launch

    def run(idx):
        frames = 0
        with contextlib.redirect_stdout(io.StringIO()):
            for _, _, coarse, _, _ in model.inference_corpus([corpus[i] for i in idx], rows=rows, launch_utterances=per_launch,
                                                             length_ratio=1e9):
                frames += coarse.shape[0]
        return frames

    run(mine[:: max(1, len(mine) // (2 * rows))][:2 * rows])          # warm-up: allocator, weight packing, both ends of the lengths
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    frames = run(mine)
    e1.record()
    torch.cuda.synchronize(dev)
    steps = model._get_engine().last_stream_steps
    agg = torch.tensor([e0.elapsed_time(e1), float(frames), float(steps)], device=dev, dtype=torch.float64)
    if dist is not None:
        mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        agg = torch.stack([mx[0], sm[1], sm[2]])
    ms, frames, steps = float(agg[0]), float(agg[1]), float(agg[2])
    return {"workload": f"C4: {n_utts} utterances, T~U{{50..250}}, length-sorted corpus dealt in snake order over {world} rank(s), ragged "
                        f"continuous batching with {rows} live rows, {n_launch} launch(es) per rank",
            "metric": "mel frames/sec (sharded corpus)", "value": frames / (ms * 1e-3), "unit": UNIT, "scaling": "strong",
            "seconds": ms * 1e-3, "frames": int(frames), "utterances": n_utts, "mean_frames_per_utterance": frames / n_utts,
            "row_occupancy": (frames + n_utts) / (steps * rows) if steps else None}


def bench_c5(synth, dev, rank, world, dist, local_rank, steps=2):
    """BASELINE config 5: one training step = Tacotron2.forward (train mode) + b200taco.loss.Tacotron2Loss (fused, no host sync)
    + backward (hand-written, csrc/train_*.cu) + gradient clipping + Adam, B=64 rows per GPU; under torchrun the gradients are
    all-reduced by DistributedDataParallel (NCCL over NVLink).  The all-reduce time left exposed = step time with DDP minus the
    same step inside ``no_sync()`` (no communication)."""
    from b200taco.loss import Tacotron2Loss, gate_targets
    from b200taco.model import Tacotron2
    torch.manual_seed(synth.WEIGHT_SEED)
    model = Tacotron2()
    model._projection_bias_initialized = True
    model = model.to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank]) if world > 1 else model
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    crit = Tacotron2Loss()
    Bc, S, Tn = C5_B, C5_S, C5_T
    ids = synth.phoneme_ids(Bc, Tn, seed=1000 + rank).to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    mel = (torch.randn(Bc, 80, S, generator=g) * 2 - 5).to(dev)
    tl = torch.full((Bc,), Tn, dtype=torch.long, device=dev)
    ml = torch.full((Bc,), S, dtype=torch.long, device=dev)
    gt = gate_targets(ml, S)

    def step(sync=True):
        ctx = contextlib.nullcontext() if (sync or world == 1) else net.no_sync()
        with ctx:
            out = net(ids, mel, tl)
            total, _, _, _ = crit(out, (mel, gt, ml), text_lengths=tl)
            opt.zero_grad(set_to_none=True)
            total.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return total

    def timed(n, sync=True):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = step(sync)
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), float(loss)

    step()                                                           # warm-up (allocator, DDP buckets)
    ms, loss = timed(steps)
    out = {"workload": f"C5: training step fwd + Tacotron2Loss + bwd + clip + Adam, B={Bc}/GPU, T={Tn}, {S}-frame mels"
                       + (", DistributedDataParallel (NCCL all-reduce of 112.6 MB fp32 gradients)" if world > 1 else ""),
           "metric": "training mel frames/sec", "value": world * Bc * S / (ms * 1e-3), "unit": UNIT, "scaling": "weak",
           "ms_per_step": ms, "tflops_per_gpu": C5_TFLOP / (ms * 1e-3), "loss": loss}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            peak = float(json.load(f).get("bf16_tflops_sustained", 0.0))
        if peak > 0:
            out["frac_of_sustained_bf16_peak"] = out["tflops_per_gpu"] / peak
            out["peak_tflops_sustained_measured"] = peak
    if world > 1:
        ms_nosync, _ = timed(steps, sync=False)
        out["ms_per_step_without_allreduce"] = ms_nosync
        out["exposed_allreduce_ms"] = ms - ms_nosync
    return out


def cpu_baseline_c5(synth, Bc=16, S=100, Tn=C5_T):
    """The reference's own training step on the host cores (oracle/_ref = the unmodified reference model, train mode, L1 + BCE
    loss, backward, Adam) on a bounded sample, scaled per frame (B=16, S=100)."""
    from oracle.make_ref import load_reference
    ref = load_reference()
    if ref is None:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(synth.WEIGHT_SEED)
    model = ref.Tacotron2()
    model._projection_bias_initialized = True
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    ids = synth.phoneme_ids(Bc, Tn, seed=1)
    mel = torch.randn(Bc, 80, S, generator=torch.Generator().manual_seed(2)) * 2 - 5
    tl = torch.full((Bc,), Tn, dtype=torch.long)

    def step():
        post, coarse, gate, _ = model(ids, mel, tl)
        tgt = mel.transpose(1, 2)
        loss = (post - tgt).abs().mean() + (coarse - tgt).abs().mean() + \
            torch.nn.functional.binary_cross_entropy_with_logits(gate, torch.zeros_like(gate))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
    with contextlib.redirect_stdout(io.StringIO()):
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
    return {"value": Bc * S / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "reference",
            "sample": f"one training step (fwd + L1/BCE loss + bwd + clip + Adam) of the unmodified reference at B={Bc}, T={Tn}, {S} frames",
            "seconds": dt}


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
    if args.workload == "c5":
        r = cpu_baseline_c5(synth)
        ms = (time.perf_counter() - t0) * 1e3
        return {"impl": "reference", "metric": "training mel frames/sec", "value": r["value"], "unit": UNIT, "n_gpus": world,
                "steps": 1, "warmup": 1, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C5 (CPU sample): " + r["sample"], "device": "host CPU", "arm": r["kind"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_ms": ms}
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
    ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--skip-extra", action="store_true", help="C2 only: no b1 / c4 / c5 legs")
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
