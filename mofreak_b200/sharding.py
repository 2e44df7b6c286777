"""How the extraction path shards across the GPUs of one box: there is no data-path collective.

  * clip level   — whole clips are dealt round-robin, largest first (configs 2, 3, 5);
  * stream level — one long stream is cut into contiguous frame ranges, each with a read-only halo of the 5
                   frames before it, because processed frame c needs frame c-5 (reference
                   MoFREAKUtilities.cpp:391-401, 485-487); rows concatenate in (clip, frame, keypoint) order.
The only exchange step the path can need is the sum of per-video bag-of-words COUNT vectors when one video's rows
live on several GPUs: an integer all-reduce (exact), after which count/total is formed in float on the host as in
BagOfWordsRepresentation.cpp:125-135.
"""
FRAME_GAP = 5


def shard_clips(clip_lengths, world):
    """Greedy longest-first assignment; returns per-rank lists of clip indices (ascending within a rank)."""
    order = sorted(range(len(clip_lengths)), key=lambda i: (-clip_lengths[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += max(0, clip_lengths[i] - FRAME_GAP)
    return [sorted(x) for x in out]


def shard_stream(n_frames, world):
    """Splits processed frames FRAME_GAP .. n_frames-1 into `world` contiguous ranges.

    Returns per rank (first_frame_to_upload, first_processed, end_processed): the rank uploads frames
    [first_frame_to_upload, end_processed) and processes c in [first_processed, end_processed)."""
    total = max(0, n_frames - FRAME_GAP)
    base, extra = divmod(total, world)
    out, c = [], FRAME_GAP
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((c - FRAME_GAP, c, c + n))
        c += n
    return out


def allreduce_counts(counts, group=None):
    """Sums integer word counts over the ranks of a torch.distributed group (NCCL on GPU tensors, gloo on CPU
    tensors); in place, returns the tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
