"""Length-bucketed utterance sharding for multi-GPU inference (BASELINE config 4).

Inference has no cross-utterance dependency: one process per GPU, weights replicated, every rank
decodes its own batches, **no collective on the data path**.  Only the per-utterance results
(lengths, timings) are gathered at the end.

Two facts of the reference shape the plan:

* ``Tacotron2.inference`` ignores ``text_lengths`` and attends padded positions (F8,
  src/model.py:482-494, :305), so a padded batch does not reproduce the solo result of its shorter
  rows.  Batches are therefore **exact-length**: every utterance of a batch has the same number of
  phonemes.
* A decoder step costs almost the same for 1 or 32 rows, so a batch costs ~ (frames of its longest
  row) x step_time(B); decoded length grows with the text length.  Batches are dealt to ranks
  longest-processing-time-first on that estimate, which keeps the tail imbalance small.

Everything here is host logic (pure Python + torch.distributed); it runs identically under ``gloo``
on CPU (tests/test_shard_cpu.py) and ``nccl`` on the GPU box.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch


@dataclass
class Batch:
    """Utterance indices (into the caller's corpus) that share one text length."""
    text_len: int
    indices: List[int]
    cost: float = 0.0


@dataclass
class Plan:
    world_size: int
    batches: List[Batch]
    assignment: List[List[int]] = field(default_factory=list)     # rank -> indices into ``batches``

    def load(self) -> List[float]:
        return [sum(self.batches[i].cost for i in a) for a in self.assignment]

    def imbalance(self) -> float:
        """max rank load / mean rank load (1.0 = perfect)."""
        load = self.load()
        mean = sum(load) / max(1, len(load))
        return max(load) / mean if mean > 0 else 1.0


def step_time_us(batch_rows: int) -> float:
    """Measured decoder step time on B200 (profiles/r1_v3_decoder_phase_times_fine.jsonl): nearly flat in B."""
    pts = [(1, 26.0), (8, 33.0), (32, 36.0), (64, 51.0)]
    if batch_rows <= pts[0][0]:
        return pts[0][1]
    for (b0, t0), (b1, t1) in zip(pts[:-1], pts[1:]):
        if batch_rows <= b1:
            return t0 + (t1 - t0) * (batch_rows - b0) / (b1 - b0)
    return pts[-1][1] * batch_rows / pts[-1][0]


def default_cost(text_len: int, rows: int, frames_per_phoneme: float = 5.5) -> float:
    """Estimated microseconds of one batch: decode steps of its longest row x step time (+ a small fixed part)."""
    return frames_per_phoneme * text_len * step_time_us(rows) + 3000.0


def plan_buckets(text_lengths: Sequence[int], batch_size: int = 32,
                 cost_fn: Callable[[int, int], float] = default_cost) -> List[Batch]:
    """Exact-length buckets cut into batches of at most ``batch_size`` rows (input order kept inside a bucket)."""
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    buckets: Dict[int, List[int]] = {}
    for i, n in enumerate(text_lengths):
        n = int(n)
        if n < 1:
            raise ValueError(f"utterance {i} has no phonemes")
        buckets.setdefault(n, []).append(i)
    batches = []
    for n in sorted(buckets, reverse=True):
        idx = buckets[n]
        for k in range(0, len(idx), batch_size):
            rows = idx[k:k + batch_size]
            batches.append(Batch(n, rows, float(cost_fn(n, len(rows)))))
    return batches


def assign(batches: List[Batch], world_size: int) -> Plan:
    """Longest-processing-time-first: heaviest batch to the least loaded rank.  Deterministic on every rank."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(batches)), key=lambda i: (-batches[i].cost, i))
    load = [0.0] * world_size
    assignment: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        assignment[r].append(i)
        load[r] += batches[i].cost
    return Plan(world_size, batches, assignment)


def make_plan(text_lengths: Sequence[int], world_size: int, batch_size: int = 32,
              cost_fn: Callable[[int, int], float] = default_cost) -> Plan:
    return assign(plan_buckets(text_lengths, batch_size, cost_fn), world_size)


def run_shard(plan: Plan, rank: int, corpus: Sequence[torch.Tensor],
              decode_fn: Callable[[torch.Tensor], Sequence[int]]) -> Dict[int, int]:
    """Decode this rank's batches.  ``corpus[i]`` is the 1-D int64 phoneme-id tensor of utterance i;
    ``decode_fn(ids (B,T)) -> per-row decoded lengths`` (e.g. a wrapper of
    ``Tacotron2.inference_per_utterance``).  Returns {utterance index: decoded length}."""
    out: Dict[int, int] = {}
    for bi in plan.assignment[rank]:
        b = plan.batches[bi]
        ids = torch.stack([corpus[i] for i in b.indices], dim=0)
        if ids.shape[1] != b.text_len:
            raise ValueError(f"batch {bi}: utterances are not all {b.text_len} phonemes long")
        lengths = [int(x) for x in decode_fn(ids)]
        if len(lengths) != len(b.indices):
            raise ValueError(f"decode_fn returned {len(lengths)} lengths for {len(b.indices)} rows")
        out.update(zip(b.indices, lengths))
    return out


def gather_lengths(local: Dict[int, int], n_utterances: int, device: Optional[torch.device] = None) -> torch.Tensor:
    """All ranks' decoded lengths as one (n_utterances,) int64 tensor on every rank.

    One small all-reduce after the data path has finished (each utterance is written by exactly one rank; the
    others contribute 0).  Without an initialised process group this is the identity.
    """
    import torch.distributed as dist
    t = torch.zeros(n_utterances, dtype=torch.int64, device=device)
    if local:
        idx = torch.tensor(list(local.keys()), dtype=torch.int64, device=device)
        val = torch.tensor(list(local.values()), dtype=torch.int64, device=device)
        t[idx] = val
    owners = torch.zeros(n_utterances, dtype=torch.int64, device=device)
    if local:
        owners[idx] = 1
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(owners, op=dist.ReduceOp.SUM)
    if not bool((owners == 1).all()):
        bad = (owners != 1).nonzero().flatten().tolist()[:8]
        raise RuntimeError(f"sharding error: utterances {bad} were decoded {owners[bad].tolist()} times")
    return t


def deal_length_groups(text_lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Sharding for the ragged continuous-batching path (``Tacotron2.inference_corpus``, bench.py's C4 leg): the utterance
    indices of ``rank``, sorted by text length.  WHOLE exact-length groups go to one rank - an utterance is encoded at its own
    length (F8), one encoder launch per length and rank, so splitting a length over the ranks would multiply the small
    encoder launches.  The distinct lengths are walked in ascending blocks of ``world_size``; inside a block every rank
    receives exactly one group, the largest group going to the rank with the fewest utterances so far: every rank spans the
    whole length range (the same length distribution, hence the same attention cost per decode step) with utterance counts
    within a few percent (the decode length of an utterance does not depend on its text length).  No communication: every
    rank derives its share from the same lengths."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    groups: Dict[int, List[int]] = {}
    for i, n in enumerate(text_lengths):
        groups.setdefault(int(n), []).append(i)
    distinct = sorted(groups)
    load = [0] * world_size
    mine: List[int] = []
    for lo in range(0, len(distinct), world_size):
        free = list(range(world_size))
        for n in sorted(distinct[lo:lo + world_size], key=lambda n: (-len(groups[n]), n)):
            r = min(free, key=lambda q: (load[q], q))
            free.remove(r)
            load[r] += len(groups[n])
            if r == rank:
                mine += groups[n]
    return sorted(mine, key=lambda i: (text_lengths[i], i))


def deal_sorted(text_lengths: Sequence[int], world_size: int, rank: int) -> List[int]:
    """Sharding for the ragged path once the ENCODER takes ragged batches too (``Engine.encoder(lengths=...)``): equal lengths no
    longer have to stay together, so the corpus is sorted by text length and every ``world_size``-th utterance goes to a rank,
    in snake order (0..W-1, W-1..0, ...) so that no rank always receives the longest of a round.  Utterance counts differ by at
    most one and the summed text lengths by a fraction of a percent (0.06 % for 8192 utterances on 8 ranks; whole exact-length
    groups: 4.2 %); every rank spans the whole length range.  Returns the indices of ``rank``, sorted by text length.  No
    communication: every rank derives its share from the same lengths."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    order = sorted(range(len(text_lengths)), key=lambda i: (int(text_lengths[i]), i))
    mine = []
    for pos, i in enumerate(order):
        rnd, k = divmod(pos, world_size)
        if (k if rnd % 2 == 0 else world_size - 1 - k) == rank:
            mine.append(i)
    return mine


def plan_ragged_launches(text_lengths: Sequence[int], rows: int, launch_utterances: int = 2048,
                         length_ratio: float = 1.35) -> List[tuple]:
    """Launch plan of the ragged continuous-batching path (``Engine.stream_infer_ragged``): the corpus is sorted by text length and
    cut into launches of at most ``launch_utterances`` utterances whose longest text is at most ``length_ratio`` x the shortest
    (attention work is spent on the padded width).  Inside a launch the queue runs LONGEST text first - a row takes the next
    utterance when its own stops, so the launch ends on the shortest utterances and the live utterances are of similar length at
    any time - and is filled up to a multiple of ``rows`` with repeats of its shortest utterances (results dropped).
    Returns ``[(queue, n_real), ...]``: ``queue`` = utterance indices in queue order, the first ``n_real`` of them real."""
    if rows < 1 or launch_utterances < 1:
        raise ValueError("rows and launch_utterances must be positive")
    lens = [int(n) for n in text_lengths]
    order = sorted(range(len(lens)), key=lambda i: (lens[i], i))
    launches, cur = [], []
    for i in order:
        if cur and (len(cur) >= launch_utterances or lens[i] > length_ratio * lens[cur[0]]):
            launches.append(cur)
            cur = []
        cur.append(i)
    if cur:
        launches.append(cur)
    plan = []
    for members in launches:
        queue = members[::-1]
        n_real = len(queue)
        queue = queue + [queue[n_real - 1 - k % n_real] for k in range((-n_real) % rows)]
        plan.append((queue, n_real))
    return plan
