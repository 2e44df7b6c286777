"""Two-GPU checks (skipped on a single-GPU box): streams sharded with the 5-frame halo and the one exchange step the path can
need, the NCCL sum of per-video word counts, through the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import mofreak_b200
from mofreak_b200 import sharding
from tests import util

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)            # only carries the 128-byte NCCL id
    torch.cuda.set_device(rank)
    clip = mofreak_b200.synth_clip(320, 240, 25, 5)
    first, a, b = sharding.shard_stream(clip.shape[0], world)[rank]
    rng = np.random.default_rng(3)
    cb = rng.integers(0, 256, (600, 16), dtype=np.uint8)
    with mofreak_b200.Context(320, 240, 32, 8, device=rank) as ctx:
        uid = torch.from_numpy(mofreak_b200.Context.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8))
        dist.broadcast(uid, 0)
        comm = ctx.nccl_comm_init(world, rank, uid.numpy())
        rows = ctx.extract_clip(clip[first:b])
        rows["frame"] += first
        ctx.bow_set_codebook(cb)
        desc = np.concatenate([rows["appearance"], rows["motion"]], 1)
        counts = torch.from_numpy(ctx.bow_histogram(desc).astype(np.int32)).cuda()
        ctx.allreduce_counts(counts.data_ptr(), 600, comm)
        ctx.sync()
        np.save(os.path.join(out_dir, "rows%d.npy" % rank), rows)
        np.save(os.path.join(out_dir, "counts%d.npy" % rank), counts.cpu().numpy())
        mofreak_b200.lib().mfk_nccl_comm_destroy(comm)
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharding_and_nccl_count_allreduce(orc, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    clip = mofreak_b200.synth_clip(320, 240, 25, 5)
    want = orc.extract_clip(clip)
    got = np.concatenate([np.load(tmp_path / ("rows%d.npy" % r)) for r in range(world)])
    assert util.rows_equal(got, want), util.describe_diff(got, want)
    rng = np.random.default_rng(3)
    cb = rng.integers(0, 256, (600, 16), dtype=np.uint8)
    desc = np.concatenate([want["appearance"], want["motion"]], 1)
    total = np.bincount(orc.bow_assign(desc, cb), minlength=600)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("counts%d.npy" % r)), total)
