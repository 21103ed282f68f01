"""BASELINE config 4: length-bucketed sharded inference of N synthetic utterances (T ~ U{50..250}) across the
ranks of this job (one process per GPU, no data-path collective).

    python profiles/tools/c4_sharded.py [--utterances 8192]                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... profiles/tools/c4_sharded.py
"""
import argparse
import contextlib
import io
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from b200taco import shard, synth  # noqa: E402
from b200taco.model import Tacotron2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=64)      # rows per launch: a step costs 31 us at 32 rows, 42 us at 64
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    rng = random.Random(0)
    lens = [rng.randint(50, 250) for _ in range(args.utterances)]
    corpus = [synth.phoneme_ids(1, t, seed=i)[0] for i, t in enumerate(lens)]
    plan = shard.make_plan(lens, world, args.batch)

    torch.manual_seed(synth.WEIGHT_SEED)
    model = Tacotron2().eval()
    model._projection_bias_initialized = True
    model = model.to(dev)
    with torch.no_grad():
        ref_ids = synth.phoneme_ids(32, 150, seed=0).to(dev)
        center = model._get_engine().encoder(ref_ids).mean((0, 1)).cpu()
        sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        synth.plant_stop_counter_(sd, ctx_center=center, rate=8e-4, jitter=6e-3)
        synth.round_to_bf16_(sd)
    model.load_state_dict(sd)

    def decode(ids):
        with contextlib.redirect_stdout(io.StringIO()):
            out = model.inference_per_utterance(ids.to(dev))
        return out[4].tolist()

    decode(torch.stack([corpus[i] for i in plan.batches[0].indices]))          # warm-up (weight packing, allocator)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    local_out = shard.run_shard(plan, rank, corpus, decode)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)                              # after the data path: timings only
    total = shard.gather_lengths(local_out, len(lens), device=dev)
    if rank == 0:
        frames = int(total.sum())
        print(json.dumps({"workload": f"C4: {len(lens)} utterances, T~U{{50..250}}, exact-length batches of <= {args.batch}",
                          "n_gpus": world, "batches": len(plan.batches), "plan_imbalance": round(plan.imbalance(), 4),
                          "frames": frames, "seconds": float(ms) / 1e3, "frames_per_s": frames / (float(ms) / 1e3),
                          "mean_frames_per_utterance": frames / len(lens)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
