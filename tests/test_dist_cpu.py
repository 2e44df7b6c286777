"""World-size-2 CPU (gloo) coverage of the multi-GPU host logic: stream sharding with the 5-frame halo reproduces
the single-process rows, and the bag-of-words count all-reduce reproduces the single-process histogram."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mofreak_b200
from mofreak_b200 import sharding
from tests import util


def test_shard_stream_covers_every_processed_frame_once():
    for n in (3, 5, 6, 7, 100, 101, 250):
        for world in (1, 2, 3, 8):
            parts = sharding.shard_stream(n, world)
            assert len(parts) == world
            covered = [c for (_, a, b) in parts for c in range(a, b)]
            assert covered == list(range(5, max(5, n)))
            for first, a, b in parts:
                assert first == a - 5 and first >= 0


def test_balance_streams_keeps_the_total_and_follows_the_rates():
    # the measured case: eight ranks, four links at 42.4 and four at 69.5 steps/s, 8 streams per group and rank
    rates = [42.38, 42.76, 42.66, 42.40, 69.96, 69.04, 68.81, 69.47]
    assert sharding.balance_streams(rates, 8) == [6, 6, 6, 6, 10, 10, 10, 10]
    assert sharding.balance_streams([97.8, 93.9], 8) == [8, 8]                 # equal links: nothing changes
    assert sharding.balance_streams([1.0], 8) == [8]
    assert sharding.balance_streams([0.0, 0.0], 8) == [8, 8]                   # no measurement: equal shares
    rng = np.random.default_rng(3)
    for _ in range(200):
        world = int(rng.integers(1, 9))
        per = int(rng.integers(1, 17))
        rates = rng.uniform(0.2, 3.0, world)
        got = sharding.balance_streams(rates, per)
        assert len(got) == world and all(min(2, per) <= g <= 2 * per for g in got)
        if all(per * world * r / rates.sum() >= 2 for r in rates) and all(per * world * r / rates.sum() <= 2 * per - 1 for r in rates):
            assert sum(got) == per * world                                     # nobody was clamped: the total is kept
            assert all(abs(g - per * world * r / rates.sum()) < 1.0 + 1e-9 for g, r in zip(got, rates))
        assert got == sharding.balance_streams(list(rates), per)               # deterministic


def test_shard_clips_balances_and_partitions():
    lengths = [400, 100, 250, 399, 20, 5, 3, 120]
    for world in (1, 2, 4, 8):
        parts = sharding.shard_clips(lengths, world)
        assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
        loads = [sum(max(0, lengths[i] - 5) for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests import oracle_binding
    orc = oracle_binding.load()
    clip = mofreak_b200.synth_clip(160, 120, 17, 3)
    first, a, b = sharding.shard_stream(clip.shape[0], world)[rank]
    sub = clip[first:b]                                     # halo + own range
    rows = orc.extract_clip(sub)
    rows["frame"] += first                                  # labels are relative to the uploaded range
    rng = np.random.default_rng(7)
    cb = rng.integers(0, 256, (64, 16), dtype=np.uint8)
    desc = np.concatenate([rows["appearance"], rows["motion"]], 1)
    counts = torch.from_numpy(np.bincount(orc.bow_assign(desc, cb), minlength=64).astype(np.int64))
    sharding.allreduce_counts(counts)
    np.save(os.path.join(out_dir, "rows%d.npy" % rank), rows)
    if rank == 0:
        np.save(os.path.join(out_dir, "counts.npy"), counts.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stream_sharding_and_count_allreduce(orc, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    clip = mofreak_b200.synth_clip(160, 120, 17, 3)
    want = orc.extract_clip(clip)
    got = np.concatenate([np.load(tmp_path / ("rows%d.npy" % r)) for r in range(world)])
    assert util.rows_equal(got, want), util.describe_diff(got, want)
    rng = np.random.default_rng(7)
    cb = rng.integers(0, 256, (64, 16), dtype=np.uint8)
    desc = np.concatenate([want["appearance"], want["motion"]], 1)
    assert np.array_equal(np.load(tmp_path / "counts.npy"), np.bincount(orc.bow_assign(desc, cb), minlength=64))
