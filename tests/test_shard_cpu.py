"""Host logic of the multi-GPU inference path (BASELINE config 4): exact-length bucketing, LPT
assignment, and the result gather over a real 2-rank ``gloo`` process group on CPU."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200taco import shard, synth


def corpus_lengths(n=8192, lo=50, hi=250, seed=0):
    rng = random.Random(seed)
    return [rng.randint(lo, hi) for _ in range(n)]


def test_buckets_are_exact_length_and_cover_everything():
    lens = corpus_lengths(1000)
    batches = shard.plan_buckets(lens, batch_size=32)
    seen = []
    for b in batches:
        assert 1 <= len(b.indices) <= 32
        assert all(lens[i] == b.text_len for i in b.indices)          # F8: no padding inside a batch
        seen += b.indices
    assert sorted(seen) == list(range(1000))


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_lpt_assignment_balances_c4(world):
    """BASELINE C4: 8192 utterances, T ~ U{50..250}: every batch on exactly one rank, tail imbalance < 2 %."""
    lens = corpus_lengths()
    plan = shard.make_plan(lens, world, batch_size=32)
    flat = sorted(i for a in plan.assignment for i in a)
    assert flat == list(range(len(plan.batches)))
    assert plan.imbalance() < 1.02
    # deterministic: every rank derives the same plan from the same lengths
    again = shard.make_plan(lens, world, batch_size=32)
    assert again.assignment == plan.assignment


def test_edge_cases():
    assert shard.plan_buckets([], 32) == []
    plan = shard.make_plan([7], 4)
    assert sum(len(a) for a in plan.assignment) == 1
    with pytest.raises(ValueError):
        shard.plan_buckets([5, 0], 32)
    with pytest.raises(ValueError):
        shard.make_plan([5], 0)
    # ragged leftovers: 33 utterances of one length -> a full batch and a single-row batch
    b = shard.plan_buckets([12] * 33, 32)
    assert [len(x.indices) for x in b] == [32, 1]


def _fake_decode(ids):
    """Stand-in for Tacotron2.inference_per_utterance on CPU: a deterministic function of the row."""
    return [int(row.sum().item()) % 700 + 2 for row in ids]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lens = corpus_lengths(n, 5, 40, seed=3)
        corpus = [synth.phoneme_ids(1, t, seed=i)[0] for i, t in enumerate(lens)]
        plan = shard.make_plan(lens, world, batch_size=8)
        local = shard.run_shard(plan, rank, corpus, _fake_decode)
        total = shard.gather_lengths(local, n)
        q.put((rank, len(local), total.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_run_and_gather():
    n, world = 200, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lens = corpus_lengths(n, 5, 40, seed=3)
    corpus = [synth.phoneme_ids(1, t, seed=i)[0] for i, t in enumerate(lens)]
    want = [_fake_decode(c[None])[0] for c in corpus]
    assert sum(g[1] for g in got) == n                     # every utterance decoded exactly once
    assert min(g[1] for g in got) > 0                      # both ranks worked
    for _, _, total in got:
        assert total == want                               # same gathered result on every rank, corpus order


def test_gather_detects_double_decoding():
    with pytest.raises(RuntimeError):
        shard.gather_lengths({0: 5}, 2)                    # utterance 1 was never decoded


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_length_groups_are_dealt_whole_and_balanced(world):
    """Sharding of the ragged continuous-batching path: every utterance on exactly one rank, no exact-length group split over
    the ranks, utterance counts and summed text lengths within 7 % of each other, every rank spanning the whole length range."""
    lens = corpus_lengths()
    shares = [shard.deal_length_groups(lens, world, r) for r in range(world)]
    assert sorted(i for s_ in shares for i in s_) == list(range(len(lens)))
    owner = {}
    for r, s_ in enumerate(shares):
        assert [lens[i] for i in s_] == sorted(lens[i] for i in s_)            # sorted by length: launches are length ranges
        for i in s_:
            assert owner.setdefault(lens[i], r) == r                           # a length never appears on two ranks
    counts = [len(s_) for s_ in shares]
    work = [sum(lens[i] for i in s_) for s_ in shares]
    assert max(counts) <= 1.07 * min(counts) and max(work) <= 1.07 * min(work)     # granularity: one group = 41 utterances of 1024
    if world > 1:
        for s_ in shares:
            assert min(lens[i] for i in s_) < 60 and max(lens[i] for i in s_) > 240
    with pytest.raises(ValueError):
        shard.deal_length_groups(lens, world, world)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sorted_snake_dealing_is_a_balanced_partition(world):
    """Sharding used by bench.py's C4 leg since the encoder takes ragged batches: every utterance on exactly one rank, counts
    within one, summed text lengths within 0.5 %, every share sorted by length and spanning the whole range, deterministic."""
    lens = corpus_lengths()
    shares = [shard.deal_sorted(lens, world, r) for r in range(world)]
    assert sorted(i for s_ in shares for i in s_) == list(range(len(lens)))
    counts = [len(s_) for s_ in shares]
    work = [sum(lens[i] for i in s_) for s_ in shares]
    assert max(counts) - min(counts) <= 1 and max(work) <= 1.005 * min(work)
    for s_ in shares:
        assert [lens[i] for i in s_] == sorted(lens[i] for i in s_)
        if world > 1:
            assert min(lens[i] for i in s_) < 60 and max(lens[i] for i in s_) > 240
    assert shares == [shard.deal_sorted(list(lens), world, r) for r in range(world)]
    with pytest.raises(ValueError):
        shard.deal_sorted(lens, world, -1)


def test_ragged_launch_plan():
    """Launch plan of the ragged continuous-batching path: every utterance queued exactly once, launches are length ranges within
    the ratio and the size cap, the queue of a launch runs longest text first and is filled up to a multiple of the rows with
    repeats of its shortest utterances."""
    lens = corpus_lengths()
    plan = shard.plan_ragged_launches(lens, rows=64, launch_utterances=2048, length_ratio=1.35)
    real = [i for q, n in plan for i in q[:n]]
    assert sorted(real) == list(range(len(lens)))
    prev_max = 0
    for q, n in plan:
        assert len(q) % 64 == 0 and 0 < n <= 2048 and len(q) - n < 64
        ql = [lens[i] for i in q[:n]]
        assert ql == sorted(ql, reverse=True) and max(ql) <= 1.35 * min(ql)
        assert min(ql) >= prev_max                                   # launches are ascending length ranges
        prev_max = max(ql)
        assert set(q[n:]) <= set(q[max(0, n - (len(q) - n)):n])      # the fill repeats the shortest utterances of the launch
    one = shard.plan_ragged_launches(lens[:100], rows=8, launch_utterances=10 ** 6, length_ratio=1e9)
    assert len(one) == 1 and one[0][1] == 100 and len(one[0][0]) == 104
    assert one[0][0][100:] == [one[0][0][99 - k] for k in range(4)]  # fill = the shortest ones, shortest first
    assert shard.plan_ragged_launches([], rows=8) == []
    with pytest.raises(ValueError):
        shard.plan_ragged_launches(lens, rows=0)
