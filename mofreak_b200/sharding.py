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


def balance_streams(rates, streams_per_rank, lo=2):
    """Streams for each rank in proportion to the rate it sustains (end to end every rank is bound by its host->device link,
    and the links of one box do not carry the same rate when all of them are busy: eight B200s, 18 GB/s on four of them and
    29.5 GB/s on the other four).  The box's total `streams_per_rank * len(rates)` is kept; a rank gets between `lo` and
    twice `streams_per_rank`; largest remainders take what flooring leaves (ties by rank), so every rank that holds the same
    `rates` computes the same assignment."""
    rates = [max(0.0, float(r)) for r in rates]
    world, total, hi = len(rates), streams_per_rank * len(rates), 2 * streams_per_rank
    if world == 0 or sum(rates) <= 0.0:
        return [streams_per_rank] * world
    share = [r / sum(rates) * total for r in rates]
    out = [min(hi, max(lo, int(x))) for x in share]
    order = sorted(range(world), key=lambda k: (-(share[k] - int(share[k])), k))
    i = 0
    while sum(out) < total and i < 4 * world:
        k = order[i % world]
        if out[k] < hi:
            out[k] += 1
        i += 1
    return out


def allreduce_counts(counts, group=None):
    """Sums integer word counts over the ranks of a torch.distributed group (NCCL on GPU tensors, gloo on CPU
    tensors); in place, returns the tensor."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
