#!/usr/bin/env python
"""Benchmark of the MoFREAK extraction hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path on the host cores (rank 0 only)
  python bench.py --config kth|ucf|trecvid                  BASELINE configs[1] / [2] / [3]; the headline (default) is trecvid

Workloads (BASELINE.json `configs`, synthetic frames of SURVEY §8d):
  trecvid  configs[3], the configuration the metric is quoted on: 16 continuous 720x576 gray streams per GPU, one step =
           64 new frames of every stream = 1024 processed frames (frame t against t-5);
  ucf      configs[2]: 256 clips of 200 frames, 320x240, extraction + per-clip 600-word Hamming BoW histograms;
  kth      configs[1]: 600 clips of 400 frames, 160x120, extraction only.
Videos shard across GPUs with no collective on the extraction path (weak scaling: every GPU gets its own videos).  The one
exchange step the path has - summing per-video word counts when a video's frame ranges live on several GPUs - is timed
as its own leg at N > 1 (`bow_allreduce`: mfk_bow_histogram_videos over an NCCL communicator).
  value   frames/s with the frames already resident in the HBM frame store (rows stay on the device)
  e2e     frames/s through the C ABI with HOST buffers: pinned-host -> device upload of every new frame of the step and
          device -> host fetch of the step's rows inside the timed region
The oracle (oracle/) is used only as the timed CPU baseline, never on the measured GPU path.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAP = 5
CONFIGS = {
    # videos per GPU, processed frames (pairs) per video per step, pairs per internal launch sequence
    "trecvid": dict(W=720, H=576, videos=16, pairs=64, max_pairs=256, mode="streams", bow=False, baseline="configs[3]",
                    what="TRECVID-shape 720x576 synthetic streams, extraction (diff-image BRISK + 8B FREAK + 8B MIP rows)"),
    "ucf": dict(W=320, H=240, videos=256, pairs=195, max_pairs=1024, mode="clips", bow=True, baseline="configs[2]",
                what="UCF101/HMDB51-shape 320x240 synthetic clips, extraction + per-clip 600-word Hamming BoW histograms"),
    "kth": dict(W=160, H=120, videos=600, pairs=395, max_pairs=4096, mode="clips", bow=False, baseline="configs[1]",
                what="KTH-shape 160x120 synthetic clips, extraction only"),
}
UNIT = "frames/s"


def metric_name(cfg):
    return "frames_per_sec_%dx%d" % (cfg["W"], cfg["H"])


def algorithmic_bytes_per_frame(cfg, rows_per_frame):
    # read I_t and I_{t-5} once each (u8), write the 32-byte rows
    return 2 * cfg["W"] * cfg["H"] + 32 * rows_per_frame


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profile_json(name):
    """A committed summary under profiles/ (ncu-derived figures the live run cannot measure itself), or None."""
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2]))
            except ValueError:
                continue
            try:
                power.append(float(c[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # under load = the upper half of the samples (the sampler also sees the idle edges)
            sm.sort()
            out["sm_mhz"] = statistics.median(sm[len(sm) // 2:])
            out["sm_mhz_min"] = sm[0]
            out["sm_max_mhz"] = max(smax)
            out["samples"] = len(sm)
            if power:
                out["power_w_max"] = max(power)
        out["reasons"] = sorted(reasons)
        return out


def workload_config(name, cfg):
    n_frames = cfg["videos"] * (cfg["pairs"] + GAP)
    pyr = 1.9 * cfg["W"] * cfg["H"]
    return {"workload": "BASELINE %s: %s; per GPU and step %d videos x %d processed frames (+ the 5-frame look-back)" %
                        (cfg["baseline"], cfg["what"], cfg["videos"], cfg["pairs"]),
            "name": name, "width": cfg["W"], "height": cfg["H"], "pairs_per_step_per_gpu": cfg["videos"] * cfg["pairs"],
            "pairs_per_launch": cfg["max_pairs"],
            "l2_policy": "inputs larger than L2: %.0f MB of frames + %.0f MB of per-batch intermediates per step" %
                         (n_frames * cfg["W"] * cfg["H"] / 1e6, min(cfg["max_pairs"], cfg["videos"] * cfg["pairs"]) * (4 * pyr + 4 * (cfg["W"] + 1) * (cfg["H"] + 1)) / 1e6),
            "parallelism": "videos sharded per GPU, no collective on the extraction path"}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def synth_videos(cfg, first_seed, n_videos, n_frames, out=None, distinct=None):
    """[n_videos, n_frames, H, W] synthetic clips (threads: the generator releases the GIL).  `distinct` < n_videos generates
    that many different clips and repeats them (stated in the bench line's `data`)."""
    import mofreak_b200
    W, H = cfg["W"], cfg["H"]
    if out is None:
        out = np.empty((n_videos, n_frames, H, W), np.uint8)
    distinct = n_videos if not distinct else min(distinct, n_videos)

    def gen(i):
        out[i] = mofreak_b200.synth_clip(W, H, n_frames, first_seed + i)
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(gen, range(distinct)))
    for i in range(distinct, n_videos):
        out[i] = out[i % distinct]
    return out


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference(cfg, steps, warmup, videos=None, as_written=False, pairs=None):
    """Times the oracle (reference detector object code + restated FREAK/MIP/loop) on all host cores over the SAME videos as
    the GPU arm's step (or the first `videos` of them: a bounded sample), each video cut into frame ranges with the 5-frame
    look-back so that every core has work - the fair multi-core form of the reference's serial per-video loop.  Imports
    nothing of the product: frames come from libmofreak_synth.so."""
    from tests import oracle_binding
    orc = oracle_binding.load()
    cores = os.cpu_count() or 1
    videos = cfg["videos"] if videos is None else min(videos, cfg["videos"])
    pairs = cfg["pairs"] if pairs is None else min(pairs, cfg["pairs"])
    clips = synth_videos(cfg, 1, videos, pairs + GAP)
    parts = 1
    while videos * parts < 2 * cores and pairs % (parts * 2) == 0 and pairs // (parts * 2) >= 4:
        parts *= 2
    seg = pairs // parts
    if parts > 1:       # frame ranges [k*seg, k*seg + seg + 5) of every video
        clips = np.ascontiguousarray(np.stack([clips[v, k * seg:k * seg + seg + GAP] for v in range(videos) for k in range(parts)]))
    for _ in range(max(0, warmup)):
        orc.time_extract(clips[:, :GAP + 1], threads=cores, as_written=as_written)
    t_total, frames, rows = 0.0, 0, 0
    for _ in range(max(1, steps)):
        sec, f, r = orc.time_extract(clips, threads=cores, as_written=as_written)
        t_total += sec; frames += f; rows += r
    return {
        "value": frames / t_total, "rows_per_sec": rows / t_total, "cores": cores,
        "kind": "reference" if orc.has_reference else "port",
        "sample": "%d of the step's %d videos x %d processed %dx%d frames (%d frames per step), cut into %d frame ranges of %d+5 frames, "
                  "one range per thread on %d threads, %d steps; detector = %s; FREAK/MIP/loop restated (OpenCV not available); "
                  "FREAK table %s" %
                  (videos, cfg["videos"], pairs, cfg["W"], cfg["H"], videos * pairs, videos * parts, seg, cores, max(1, steps),
                   "reference brisk.cpp object code" if orc.has_reference else "pinned port",
                   "rebuilt per frame as the reference does" if as_written else "built once (hoisted)"),
        "frames_per_step": videos * pairs,
        "ms_per_frame_per_core": 1e3 * t_total * cores / frames,
        "ms_per_step": 1e3 * t_total / max(1, steps),
    }


def cpu_sample_videos(name, cfg):
    """A bounded sample for the CPU legs: about 1-2 s of wall time per step on 16-24 cores (whole clips, a multiple of the
    core count so that the threads finish together; the trecvid streams are cut into frame ranges instead)."""
    cores = os.cpu_count() or 1
    return min(cfg["videos"], {"trecvid": cfg["videos"], "ucf": 2 * cores, "kth": 3 * cores}[name])


def run_reference(args, name, cfg):
    rank, _, world = dist_env()
    if rank != 0:
        return
    t0 = time.time()
    sample = cpu_sample_videos(name, cfg)
    r = cpu_reference(cfg, args.steps, args.warmup, videos=sample)
    aw = cpu_reference(cfg, 1, 0, videos=min(sample, os.cpu_count() or 1), as_written=True, pairs=4)
    config = workload_config(name, cfg)
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": config,
        "reference_step": "one GPU's share of the step (%d of %d videos, %d processed frames) on the host cores; the rate does not "
                          "depend on how many GPUs the other arm uses" % (sample, cfg["videos"], r["frames_per_step"]),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                         "as_written_value": aw["value"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "descriptors_per_sec": r["rows_per_sec"], "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, name, cfg):
    import torch
    import torch.distributed as dist
    import mofreak_b200
    from mofreak_b200 import sharding

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, H, V, P, MAX_PAIRS = cfg["W"], cfg["H"], cfg["videos"], cfg["pairs"], cfg["max_pairs"]
    streams_mode = cfg["mode"] == "streams"
    frame_bytes = W * H
    FPV = P + GAP                                   # frames per video in the resident leg
    n_frames, n_pairs = V * FPV, V * P
    # Page-locked host frames (--wc 1: write-combined).  Streams: per video [f0..f4 | F = f5..f5+P-1 | B = F played backwards],
    # so that a continuous stream can be fed P new frames per step for ever (forward, backward, forward, ...: every pair
    # compares frames 5 apart in that sequence).  Clips: per video its P + 5 frames.
    HPV = FPV + (P if streams_mode else 0)          # host frames per video
    host = mofreak_b200.HostBuffer(V * HPV * frame_bytes, write_combined=args.wc)
    hv = host.array.reshape(V, HPV, H, W)
    distinct = min(V, args.distinct) if args.distinct else V
    synth_videos(cfg, 1 + rank * V, V, FPV, out=hv[:, :FPV], distinct=distinct)
    if streams_mode:
        hv[:, FPV:] = hv[:, FPV - 2:GAP - 2:-1]     # B = f(P+3) .. f4

    cur = np.concatenate([np.arange(GAP, FPV) + v * FPV for v in range(V)]).astype(np.int32)
    prev = cur - GAP
    label = np.concatenate([np.arange(GAP, FPV) - 1 for _ in range(V)]).astype(np.int32)
    ppv = np.full(V, P, np.uint32)
    rng = np.random.default_rng(7)
    codebook = rng.integers(0, 256, (600, 16), dtype=np.uint8)

    ctx = mofreak_b200.Context(W, H, n_frames, MAX_PAIRS, device=local_rank)
    ctx.bow_set_codebook(codebook)

    # ---- resident: frames already in the HBM frame store ----
    for v in range(V):
        ctx.upload_raw(host.ptr + v * HPV * frame_bytes, FPV, v * FPV)
    ctx.sync()

    def step_resident():
        ctx.reset_rows()
        ctx.extract_pairs(cur, prev, label)
        if cfg["bow"]:
            ctx.bow_histogram_videos(ppv)            # per-clip word counts of the step's rows (configs[2]); blocks

    for _ in range(args.warmup):
        step_resident()
    ctx.sync()
    _, n_rows_step = ctx.rows_device()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        step_resident()
    ms_res = ctx.timer_stop()
    barrier()
    launches = ctx.launch_count() - launches0

    # ---- per-kernel accounting: same steps with the launch sequences serialised, so that every CUDA-event span is
    # one kernel family's own time (in the throughput run above four batches overlap on their own streams) ----
    acct_steps = max(1, min(args.steps, 20))
    ctx.set_serial(True)
    step_resident(); ctx.sync()
    ctx.profile_enable(True)
    ctx.timer_start()
    for _ in range(acct_steps):
        step_resident()
    ms_serial = ctx.timer_stop()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    ctx.set_serial(False)
    barrier()

    # ---- bag-of-words assignment of the step's rows against a 600-word codebook, on the device ----
    step_resident(); ctx.sync()
    ctx.bow_histogram_rows()
    bow_iters = max(3, min(args.steps, 20))
    ctx.timer_start()
    for _ in range(bow_iters):
        ctx.bow_histogram_rows()
    ms_bow = ctx.timer_stop() / bow_iters

    # ---- the path's one exchange step (N > 1): every video's frame ranges live on all ranks; per-video word counts are
    # summed over the ranks by the product's own NCCL all-reduce and checked against torch's ----
    bow_allreduce = None
    if world > 1:
        uid = torch.from_numpy(mofreak_b200.Context.nccl_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        dist.broadcast(uid, 0)
        comm = ctx.nccl_comm_init(world, rank, uid.cpu().numpy())
        local = ctx.bow_histogram_videos(ppv)
        want = torch.from_numpy(local.astype(np.int64)).cuda()
        dist.all_reduce(want)
        got = ctx.bow_histogram_videos(ppv, comm)
        ok = bool(np.array_equal(got.astype(np.int64), want.cpu().numpy()))
        barrier()
        t0 = time.perf_counter()
        for _ in range(bow_iters):
            ctx.bow_histogram_videos(ppv, comm)
        barrier()
        ms_red = (time.perf_counter() - t0) * 1e3 / bow_iters
        t0 = time.perf_counter()
        for _ in range(bow_iters):
            ctx.bow_histogram_videos(ppv)
        barrier()
        ms_loc = (time.perf_counter() - t0) * 1e3 / bow_iters
        t = torch.tensor([ms_red, ms_loc, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        bow_allreduce = {"call": "mfk_bow_histogram_videos(pairs_per_video, comm): k_bow_assign -> ncclAllReduce(uint32, sum) -> D2H, one host wait",
                         "videos": int(V), "words": 600, "allreduce_bytes": int(V * 600 * 4), "rows_assigned_per_rank": int(n_rows_step),
                         "ms_per_call": float(t[0]), "ms_per_call_without_exchange": float(t[1]),
                         "matches_torch_all_reduce_of_local_counts": float(t[2]) == 0.0, "n_ranks": world,
                         "timing": "host clock over %d calls between barriers, max over ranks" % bow_iters}
        mofreak_b200.lib().mfk_nccl_comm_destroy(comm)

    # ---- `.mofreak` text of the step's rows, formatted on the device and copied to pinned host memory ----
    text_host = torch.empty((int(n_rows_step) * 100 + 4096,), dtype=torch.uint8, pin_memory=True)
    step_resident(); ctx.sync()
    text_bytes = ctx.rows_text_raw(text_host.data_ptr(), text_host.numel())
    text_iters = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(text_iters):
        ctx.rows_text_raw(text_host.data_ptr(), text_host.numel())
    ms_text = (time.perf_counter() - t0) * 1e3 / text_iters
    del text_host
    ctx.close()

    # ---- end to end: host buffers in, host rows out, copies inside the timed region ----
    # Two contexts (each owns a frame store, row store and streams) take turns, so that the blocking row fetch of one
    # overlaps the uploads and kernels of the other.
    row_cap = int(n_rows_step * 1.5) + 65536
    if streams_mode:
        # The streams form two groups, one per context; a bench step advances ONE group by SEGS x P frames (= the step's
        # V x P processed frames) and the groups alternate, so that the blocking row fetch of one group overlaps the uploads
        # and kernels of the other.  Every stream is a frame RING of 2 x SEGS segments of P slots on the device: a turn writes
        # the stream's new frames into the segments the previous turn did not use; the 5 frames of look-back are still
        # there - nothing is uploaded twice.
        n_groups = 2 if V >= 2 and V % 2 == 0 else 1
        SEGS = n_groups                     # segments of P frames a turn appends to each stream of its group
        ring = 2 * SEGS * P
        chunk_v = max(1, MAX_PAIRS // P)
        e2e_ctx = []

        def build_stream_contexts(per_group):
            """Two contexts (groups) of `per_group` streams each; a stream's frames come from host stream (index % V)."""
            for e in e2e_ctx:
                e["ctx"].close()
            del e2e_ctx[:]
            cap = int(n_rows_step * 1.5 * max(1.0, per_group * n_groups / V)) + 65536
            for g in range(n_groups):
                vids = [(g * per_group + j) % V for j in range(per_group)]
                c = mofreak_b200.Context(W, H, len(vids) * ring, MAX_PAIRS, device=local_rank)
                rows_host = torch.empty((cap, 32), dtype=torch.uint8, pin_memory=True)
                # initial look-back f0..f4 at the end of the ring (what the first forward segment looks back to)
                for j, v in enumerate(vids):
                    c.upload_raw(host.ptr + v * HPV * frame_bytes, GAP, j * ring + ring - GAP)
                c.sync()
                # chunks of whole streams: a short first chunk, so that the first kernels start early
                sizes, left = [], len(vids)
                if len(vids) > chunk_v >= 2:
                    sizes += [1, chunk_v - 1]; left -= chunk_v
                while left > 0:
                    sizes.append(min(chunk_v, left)); left -= sizes[-1]
                e2e_ctx.append({"ctx": c, "rows": rows_host, "vids": vids, "sizes": sizes, "seg": 0, "cap": cap})

        per_group = V // n_groups
        build_stream_contexts(per_group)
        i_in = np.arange(P)

        def submit(e):
            """All uploads of the turn first (the copy stream never waits for the host to finish queuing kernels), then the
            extractions: every launch sequence waits only for the uploads that carry its own frames (per-upload events)."""
            c = e["ctx"]
            c.reset_rows()
            todo = []
            for _ in range(SEGS):
                k = e["seg"]
                q = k % (2 * SEGS)                                         # ring segment this pass writes
                src_off = (GAP if k % 2 == 0 else FPV) * frame_bytes       # forward segment F / backward segment B
                j0 = 0
                for n_s in e["sizes"]:
                    cu, pr, lb = [], [], []
                    for j in range(j0, j0 + n_s):
                        v = e["vids"][j]
                        base = j * ring
                        c.upload_raw(host.ptr + v * HPV * frame_bytes + src_off, P, base + q * P)
                        cu.append(base + q * P + i_in)
                        pr.append(base + (q * P + i_in - GAP) % ring)      # 5 frames earlier in the ring
                        lb.append(k * P + i_in + GAP - 1)
                    todo.append((np.concatenate(cu), np.concatenate(pr), np.concatenate(lb)))
                    j0 += n_s
                e["seg"] = k + 1
            for cu, pr, lb in todo:
                c.extract_pairs(cu, pr, lb)
    else:
        # clips: whole steps alternate between two contexts; a step uploads every frame of every clip
        chunk_v = max(1, MAX_PAIRS // P)
        e2e_ctx = []
        for _ in range(2):
            c = mofreak_b200.Context(W, H, n_frames, MAX_PAIRS, device=local_rank)
            c.bow_set_codebook(codebook)
            e2e_ctx.append({"ctx": c, "rows": torch.empty((row_cap, 32), dtype=torch.uint8, pin_memory=True), "cap": row_cap})

        def submit(e):
            c = e["ctx"]
            c.reset_rows()
            # the clips of a chunk are adjacent in the host buffer (HPV == FPV in clips mode): one copy per chunk, queued two
            # chunks ahead of the extraction that needs it (each launch sequence waits only for its own chunk's upload)
            starts = list(range(0, V, chunk_v))
            nxt = 0
            for i, v0 in enumerate(starts):
                while nxt < min(len(starts), i + 3):
                    u0 = starts[nxt]; u1 = min(V, u0 + chunk_v)
                    c.upload_raw(host.ptr + u0 * FPV * frame_bytes, (u1 - u0) * FPV, u0 * FPV)
                    nxt += 1
                v1 = min(V, v0 + chunk_v)
                ps = slice(v0 * P, v1 * P)
                c.extract_pairs(cur[ps], prev[ps], label[ps])

    def collect(e):
        if cfg["bow"]:
            e["ctx"].bow_histogram_videos(ppv)       # configs[2]: per-clip word counts belong to the step (blocks; the other context works)
        return e["ctx"].fetch_rows_raw(e["rows"].data_ptr(), e["cap"])

    def run_e2e(steps):
        """Returns the average number of rows fetched per step."""
        pending, acc = None, 0
        for k in range(steps):
            e = e2e_ctx[k % len(e2e_ctx)]
            submit(e)
            if pending is not None:
                acc += collect(pending)
            pending = e
        if pending is not None:
            acc += collect(pending)
        return acc / max(1, steps)

    run_e2e(max(2, args.warmup))
    # ---- N > 1, streams: the host->device links of one box do not carry the same rate when all of them are busy (eight ranks:
    # 23 GB/s on the slowest, 30 GB/s on average, profiles/h2d_ceiling.json), and end to end every rank is bound by its link.  So
    # streams are assigned to the GPUs in proportion to the rate each rank sustains while ALL ranks run (measured here over a
    # fixed wall time), keeping the box's total: a rank then advances `per_group` streams per step instead of V / 2. ----
    balance = None
    if world > 1 and streams_mode and not args.no_balance:
        barrier()
        t_cal, n_cal = time.perf_counter(), 0
        while n_cal < 4 or time.perf_counter() - t_cal < 0.6:
            run_e2e(2)
            n_cal += 2
        for e in e2e_ctx:
            e["ctx"].sync()
        rate = n_cal / (time.perf_counter() - t_cal)
        rates = torch.zeros(world, dtype=torch.float64, device="cuda")
        rates[rank] = rate
        dist.all_reduce(rates)
        rates = rates.cpu().numpy()
        assign = sharding.balance_streams(rates, per_group)      # streams per group over the box stay what they were
        balance = {"steps_per_s_per_rank_equal_shares": [float(x) for x in rates], "streams_per_group_per_rank": [int(x) for x in assign]}
        if int(assign[rank]) != per_group:
            per_group = int(assign[rank])
            build_stream_contexts(per_group)
        barrier()
        run_e2e(2)
    frames_per_step_e2e = (per_group * n_groups * P) if streams_mode else n_pairs
    h2d_per_step = (frames_per_step_e2e if streams_mode else n_frames) * frame_bytes
    launches_e2e0 = sum(e["ctx"].launch_count() for e in e2e_ctx)
    barrier()
    t0 = time.perf_counter()
    rows_e2e = run_e2e(args.steps)
    for e in e2e_ctx:
        e["ctx"].sync()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3   # every fetch blocks the host, so the host clock bounds the device work
    launches_e2e = sum(e["ctx"].launch_count() for e in e2e_ctx) - launches_e2e0
    for e in e2e_ctx:
        e["ctx"].close()
    clocks = sampler.stop() if sampler else None

    if world > 1:
        t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res, ms_e2e = float(t[0]), float(t[1])
        r = torch.tensor([float(n_rows_step), float(frames_per_step_e2e), float(h2d_per_step)], dtype=torch.float64, device="cuda")
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rows_all, frames_step_e2e_all, h2d_per_step = float(r[0]), float(r[1]), float(r[2]) / world      # bytes: mean over the ranks
    else:
        rows_all, frames_step_e2e_all = float(n_rows_step), float(frames_per_step_e2e)

    if rank == 0:
        frames_all = world * n_pairs * args.steps
        value = frames_all / (ms_res / 1e3)
        e2e = frames_step_e2e_all * args.steps / (ms_e2e / 1e3)
        rows_per_frame = n_rows_step / n_pairs
        # dominant kernel family by device time inside the timed region
        fam = {k: v for k, v in prof.items() if v[0] > 0 and k not in ("mip_only", "text") and (cfg["bow"] or k != "bow_assign")}
        total_ms = sum(v[1] for v in fam.values())
        top = max(fam, key=lambda k: fam[k][1])
        batches = acct_steps * ((n_pairs + MAX_PAIRS - 1) // MAX_PAIRS)
        pairs_per_launch = n_pairs / ((n_pairs + MAX_PAIRS - 1) // MAX_PAIRS)
        top_ms_per_launch = fam[top][1] / batches            # one launcher call per batch of <= MAX_PAIRS pairs
        alg_bytes = algorithmic_bytes_per_frame(cfg, rows_per_frame) * pairs_per_launch
        peak, peak_src = measured_peak_gbs()
        achieved = alg_bytes / (top_ms_per_launch * 1e-3) / 1e9
        traffic = profile_json("traffic.json") if name == "trecvid" else None
        issue = profile_json("issue.json") if name == "trecvid" else None
        ceiling = profile_json("h2d_ceiling.json")
        cpu = cpu_reference(cfg, 2, 1, videos=cpu_sample_videos(name, cfg)) if world == 1 else None
        sm_clock_ghz = ((clocks or {}).get("sm_mhz") or 1965.0) / 1e3
        line = {
            "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic" + ("" if distinct == V else " (%d distinct clips per GPU, repeated to %d)" % (distinct, V)),
            "config": workload_config(name, cfg),
            "timed_region_s": ms_res / 1e3,
            "descriptors_per_sec": rows_all * args.steps / (ms_res / 1e3),
            "rows_per_frame": rows_per_frame,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d_per_step), "d2h_bytes_per_step": int(rows_e2e) * 32,
                    "ms_per_step": ms_e2e / args.steps, "gpu_launches": int(launches_e2e), "contexts": 2,
                    "link_balance": balance,
                    "h2d_gb_per_s_per_gpu": h2d_per_step / (ms_e2e / args.steps * 1e-3) / 1e9,
                    "h2d_ceiling_gb_per_s_per_gpu": ((ceiling or {}).get(str(world)) or {}).get("gb_per_s_per_gpu_min"),
                    "h2d_ceiling_source": "profiles/h2d_ceiling.json (copies only, same buffers and chunk sizes, %d ranks)" % world if ceiling else None,
                    "frames": ("continuous streams in two groups of %d, one group per context: a step advances one group by %d new frames per "
                               "stream (= %d processed frames), written into device frame rings; the 5-frame look-back stays resident, "
                               "nothing is uploaded twice" % (V // 2, 2 * P, V * P) +
                               ("; with N > 1 the streams of the box are assigned to the ranks in proportion to the host->device rate each rank "
                                "sustains while all ranks run (link_balance): a rank's step advances streams_per_group_per_rank streams" if balance else ""))
                              if streams_mode else "clips: every step uploads every frame of every clip",
                    "timing": "host clock around the K steps, device synchronised on both sides (every fetch blocks, so "
                              "this bounds the device time from above); two contexts take turns so that the row fetch of one "
                              "overlaps the uploads and kernels of the other"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic or {}).get(top), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": top_ms_per_launch,
                         "share_of_step": fam[top][1] / total_ms if total_ms else None,
                         "whole_path_frac": (algorithmic_bytes_per_frame(cfg, rows_per_frame) * value / world) / 1e9 / peak,
                         "limiter": "the HBM figure is the judged one, but the counters say this kernel is bound by instruction issue and "
                                    "the shared-memory / integer pipes, not by DRAM: see roofline_issue"},
            "kernel_ms_per_step": {k: v[1] / acct_steps for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])},
            "kernel_accounting": "CUDA events per kernel family over %d serialised steps of the same run (%.2f ms/step serialised vs "
                                 "%.2f ms/step with four launch sequences in flight)" % (acct_steps, ms_serial / acct_steps, ms_res / args.steps),
            "clocks": clocks,
            "bow": {"codebook_words": 600, "descriptors_per_call": int(n_rows_step), "ms_per_call": ms_bow,
                    "descriptors_per_sec": n_rows_step / (ms_bow / 1e3), "popc64_per_sec": n_rows_step * 1200 / (ms_bow / 1e3),
                    "note": "mfk_bow_histogram_rows on the device-resident rows of one step incl. the 2.4 KB count read-back; "
                            "integer-pipe bound (2 XOR + 2 POPC.64 per codeword), HBM traffic 32 B per descriptor"},
            "text": {"rows_per_call": int(n_rows_step), "bytes_per_call": int(text_bytes), "ms_per_call": ms_text,
                     "rows_per_sec": n_rows_step / (ms_text / 1e3),
                     "note": "mfk_rows_text: `.mofreak` lines of one step's device-resident rows (exact %g in integer arithmetic), "
                             "incl. the D2H copy of the text to pinned memory; host clock"},
        }
        if issue and issue.get(top):
            wi = float(issue[top]["warp_instructions_per_launch"])
            bound_ms = wi / (148 * 4 * sm_clock_ghz * 1e9) * 1e3
            line["roofline_issue"] = {"kernel": top, "warp_instr_per_launch": wi, "issue_bound_ms": bound_ms,
                                      "frac": bound_ms / top_ms_per_launch, "ms_per_launch": top_ms_per_launch,
                                      "pipes": issue[top].get("pipes"),
                                      "source": "profiles/issue.json (ncu smsp__inst_executed.sum of the same launch, %s); bound = "
                                                "instructions / (148 SMs x 4 schedulers x SM clock)" % issue.get("capture", "")}
        if bow_allreduce:
            line["bow_allreduce"] = bow_allreduce
        if cpu:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"], "sample": cpu["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="trecvid", choices=sorted(CONFIGS), help="BASELINE configuration (headline: trecvid)")
    ap.add_argument("--videos", type=int, default=0, help="videos (streams / clips) per GPU per step; 0 = the configuration's")
    ap.add_argument("--max-pairs", type=int, default=0, help="frame pairs per internal launch sequence; 0 = the configuration's")
    ap.add_argument("--distinct", type=int, default=0, help="generate only this many distinct clips and repeat them (large clip sets)")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: every rank takes the same number of streams end to end (no link-rate balancing)")
    ap.add_argument("--wc", type=int, default=0, help="1: write-combined page-locked frame buffer for the end-to-end leg")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.videos:
        cfg["videos"] = args.videos
    if args.max_pairs:
        cfg["max_pairs"] = args.max_pairs
    if args.impl == "reference":
        run_reference(args, args.config, cfg)
    else:
        run_ours(args, args.config, cfg)


if __name__ == "__main__":
    main()
