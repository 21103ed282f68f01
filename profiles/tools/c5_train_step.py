"""BASELINE config 5: teacher-forced training step (forward + backward, B rows per GPU, S-frame mels) with the gradient
all-reduce of torch DistributedDataParallel over NCCL when launched under torchrun.

    python profiles/tools/c5_train_step.py [--batch 64] [--frames 800] [--text 150] [--steps 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 ... profiles/tools/c5_train_step.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from b200taco import synth  # noqa: E402
from b200taco.model import Tacotron2  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=800)
    ap.add_argument("--text", type=int, default=150)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(synth.WEIGHT_SEED)
    model = Tacotron2()
    model._projection_bias_initialized = True
    model = model.to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    B, S, T = args.batch, args.frames, args.text
    ids = synth.phoneme_ids(B, T, seed=rank).to(dev)
    g = torch.Generator().manual_seed(100 + rank)
    mel = (torch.randn(B, 80, S, generator=g) * 2 - 5).to(dev)
    lens = torch.full((B,), T, dtype=torch.long, device=dev)

    def step():
        post, coarse, gate, aligns = net(ids, mel, lens)
        tgt = mel.transpose(1, 2)
        loss = (post - tgt).abs().mean() + (coarse - tgt).abs().mean() + torch.nn.functional.binary_cross_entropy_with_logits(
            gate, torch.zeros_like(gate))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return float(loss)

    step()                                              # warm-up
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = [step() for _ in range(args.steps)]
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # gradient all-reduce check: every rank must hold the same gradients after backward
        gsum = torch.stack([p.grad.double().sum() for p in model.parameters() if p.grad is not None]).sum().reshape(1)
        both = [torch.zeros_like(gsum) for _ in range(world)]
        dist.all_gather(both, gsum)
        same = all(bool(torch.allclose(both[0], b)) for b in both)
    if rank == 0:
        out = {"workload": f"C5: teacher-forced training step fwd+bwd+Adam, B={B}/GPU, T={T}, S={S} frames", "n_gpus": world,
               "ms_per_step": float(ms), "frames_per_s": world * B * S / (float(ms) * 1e-3), "losses": [round(x, 4) for x in losses]}
        if world > 1:
            out["grads_identical_across_ranks"] = same
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
