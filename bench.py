#!/usr/bin/env python
"""Benchmark of the MoFREAK extraction hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's CPU path on the host cores (rank 0 only)

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): TRECVID-shape synthetic
720x576 gray streams.  One step = one pass of the hot path over STREAMS streams of 5 + PAIRS_PER_STREAM frames
per GPU, i.e. STREAMS * PAIRS_PER_STREAM processed frames (frame t against t-5).  Streams shard across GPUs with
no collective on the extraction path (weak scaling: every GPU gets its own streams).
  value   frames/s with the frames already resident in the HBM frame store (rows stay on the device)
  e2e     frames/s through the C ABI with HOST buffers: pinned-host -> device upload of every frame of the step
          and device -> host fetch of the step's rows inside the timed region
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 720, 576
STREAMS = 16                # per GPU
PAIRS_PER_STREAM = 64
FRAMES_PER_STREAM = PAIRS_PER_STREAM + 5
MAX_PAIRS = 256             # pairs per internal launch sequence
METRIC = "frames_per_sec_720x576"
UNIT = "frames/s"


def algorithmic_bytes_per_frame(rows_per_frame):
    # read I_t and I_{t-5} once each (u8), write the 32-byte rows
    return 2 * W * H + 32 * rows_per_frame


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_per_launch():
    """dram bytes per dominant-kernel launch from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
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
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2]))
            except ValueError:
                continue
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
            out["sm_max_mhz"] = max(smax)
        out["reasons"] = sorted(reasons)
        return out


def workload_config():
    n_frames = STREAMS * FRAMES_PER_STREAM
    return {"workload": "BASELINE configs[3]: TRECVID-shape 720x576 synthetic streams, extraction (diff-image BRISK + "
                        "8B FREAK + 8B MIP rows), %d streams x %d frames per GPU per step" % (STREAMS, FRAMES_PER_STREAM),
            "width": W, "height": H, "pairs_per_step_per_gpu": STREAMS * PAIRS_PER_STREAM, "pairs_per_launch": MAX_PAIRS,
            "l2_policy": "inputs larger than L2: %.0f MB of frames + %.0f MB of per-batch intermediates per step" %
                         (n_frames * W * H / 1e6, MAX_PAIRS * (4 * 786240 + 4 * 577 * 721) / 1e6),
            "parallelism": "streams sharded per GPU, no collective on the extraction path"}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation of the path on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, frames_per_clip_processed=6, as_written=False):
    """Times the oracle on all host cores, one clip per thread (the fair multi-core form of the reference's
    serial per-video loop).  Returns dict with frames/s etc."""
    from tests import oracle_binding
    import mofreak_b200
    orc = oracle_binding.load()
    cores = os.cpu_count() or 1
    n = 5 + frames_per_clip_processed
    clips = np.stack([mofreak_b200.synth_clip(W, H, n, 1000 + i) for i in range(cores)])
    for _ in range(max(0, warmup)):
        orc.time_extract(clips[:, :6], threads=cores, as_written=as_written)
    t_total, frames, rows = 0.0, 0, 0
    for _ in range(max(1, steps)):
        sec, f, r = orc.time_extract(clips, threads=cores, as_written=as_written)
        t_total += sec; frames += f; rows += r
    return {
        "value": frames / t_total, "rows_per_sec": rows / t_total, "cores": cores,
        "kind": "reference" if orc.has_reference else "port",
        "sample": "%d clips (one per core) x %d processed 720x576 frames per step, %d steps; detector = %s; FREAK/MIP/loop "
                  "restated (OpenCV not available); FREAK table %s" %
                  (cores, frames_per_clip_processed, max(1, steps),
                   "reference brisk.cpp object code" if orc.has_reference else "pinned port",
                   "rebuilt per frame as the reference does" if as_written else "built once (hoisted)"),
        "ms_per_frame_per_core": 1e3 * t_total * cores / frames,
        "ms_per_step": 1e3 * t_total / max(1, steps),
    }


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    t0 = time.time()
    r = cpu_reference(args.steps, args.warmup)
    aw = cpu_reference(1, 0, frames_per_clip_processed=3, as_written=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                         "as_written_value": aw["value"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "descriptors_per_sec": r["rows_per_sec"], "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import mofreak_b200

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

    n_frames = STREAMS * FRAMES_PER_STREAM
    n_pairs = STREAMS * PAIRS_PER_STREAM
    # synthetic streams, distinct per rank; pinned so that the e2e upload is a real async DMA
    # page-locked frames from the library's allocator; --wc asks for write-combined pages (the host only fills them)
    host = mofreak_b200.HostBuffer(n_frames * H * W, write_combined=args.wc)
    hv = host.array.reshape(n_frames, H, W)
    for s in range(STREAMS):
        hv[s * FRAMES_PER_STREAM:(s + 1) * FRAMES_PER_STREAM] = mofreak_b200.synth_clip(W, H, FRAMES_PER_STREAM, 1 + rank * STREAMS + s)
    cur = np.concatenate([np.arange(5, FRAMES_PER_STREAM) + s * FRAMES_PER_STREAM for s in range(STREAMS)]).astype(np.int32)
    prev = cur - 5
    label = np.concatenate([np.arange(5, FRAMES_PER_STREAM) - 1 for _ in range(STREAMS)]).astype(np.int32)

    ctx = mofreak_b200.Context(W, H, n_frames, MAX_PAIRS, device=local_rank)
    row_cap = n_pairs * 4096
    rows_host = torch.empty((row_cap, 32), dtype=torch.uint8, pin_memory=True)

    # ---- resident: frames already in the HBM frame store ----
    ctx.upload_raw(host.ptr, n_frames, 0)
    ctx.sync()

    def step_resident():
        ctx.reset_rows()
        ctx.extract_pairs(cur, prev, label)

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
    ctx.set_serial(True)
    step_resident(); ctx.sync()
    ctx.profile_enable(True)
    ctx.timer_start()
    for _ in range(args.steps):
        step_resident()
    ms_serial = ctx.timer_stop()
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    ctx.set_serial(False)
    barrier()

    # ---- bag-of-words assignment of the step's rows against a 600-word codebook (BASELINE configs[2] shape), on the device ----
    rng = np.random.default_rng(7)
    ctx.bow_set_codebook(rng.integers(0, 256, (600, 16), dtype=np.uint8))
    step_resident(); ctx.sync()
    ctx.bow_histogram_rows()
    ctx.timer_start()
    for _ in range(max(3, args.steps)):
        ctx.bow_histogram_rows()
    ms_bow = ctx.timer_stop() / max(3, args.steps)

    # ---- `.mofreak` text of the step's rows, formatted on the device and copied to pinned host memory ----
    text_host = torch.empty((int(n_rows_step) * 100 + 4096,), dtype=torch.uint8, pin_memory=True)
    step_resident(); ctx.sync()
    text_bytes = ctx.rows_text_raw(text_host.data_ptr(), text_host.numel())
    t0 = time.perf_counter()
    for _ in range(max(3, args.steps)):
        ctx.rows_text_raw(text_host.data_ptr(), text_host.numel())
    ms_text = (time.perf_counter() - t0) * 1e3 / max(3, args.steps)

    # ---- end to end: host buffers in, host rows out, copies inside the timed region ----
    # chunks of whole streams: the upload of chunk k+1 (copy stream) overlaps the kernels of chunk k
    chunk_streams = max(1, min(STREAMS, MAX_PAIRS // PAIRS_PER_STREAM))
    # a short first chunk keeps the one upload nothing can hide behind small
    sizes, left = [], STREAMS
    if STREAMS > chunk_streams >= 2:
        sizes += [1, chunk_streams - 1]; left -= chunk_streams
    while left > 0:
        sizes.append(min(chunk_streams, left)); left -= sizes[-1]
    chunks, s0 = [], 0
    for n_s in sizes:
        s1 = s0 + n_s
        chunks.append((s0 * FRAMES_PER_STREAM, (s1 - s0) * FRAMES_PER_STREAM, slice(s0 * PAIRS_PER_STREAM, s1 * PAIRS_PER_STREAM)))
        s0 = s1
    frame_bytes = W * H

    # two contexts alternate steps (each owns a frame store, row store and streams), so that the blocking row fetch of
    # step k overlaps the uploads and kernels of step k+1; every step still uploads all of its frames and fetches all
    # of its rows inside the timed region
    ctx_b = mofreak_b200.Context(W, H, n_frames, MAX_PAIRS, device=local_rank)
    rows_host_b = torch.empty((row_cap, 32), dtype=torch.uint8, pin_memory=True)
    e2e_ctx = [(ctx, rows_host), (ctx_b, rows_host_b)]

    def submit_e2e(c):
        c.reset_rows()
        for f0, nf, ps in chunks:
            c.upload_raw(host.ptr + f0 * frame_bytes, nf, f0)
            c.extract_pairs(cur[ps], prev[ps], label[ps])

    def run_e2e(steps):
        rows, pending = 0, None
        for k in range(steps):
            c, out = e2e_ctx[k % 2]
            submit_e2e(c)
            if pending is not None:
                rows = pending[0].fetch_rows_raw(pending[1].data_ptr(), row_cap)
            pending = (c, out)
        if pending is not None:
            rows = pending[0].fetch_rows_raw(pending[1].data_ptr(), row_cap)
        return rows

    run_e2e(max(2, args.warmup))
    launches_e2e0 = ctx.launch_count() + ctx_b.launch_count()
    barrier()
    t0 = time.perf_counter()
    rows_e2e = run_e2e(args.steps)
    ctx.sync(); ctx_b.sync()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3   # every fetch blocks the host, so the host clock bounds the device work
    launches_e2e = ctx.launch_count() + ctx_b.launch_count() - launches_e2e0
    del ctx_b
    clocks = sampler.stop() if sampler else None

    if world > 1:
        t = torch.tensor([ms_res, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_res, ms_e2e = float(t[0]), float(t[1])
        r = torch.tensor([float(n_rows_step)], dtype=torch.float64, device="cuda")
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        rows_all = float(r[0])
    else:
        rows_all = float(n_rows_step)

    if rank == 0:
        frames_all = world * n_pairs * args.steps
        value = frames_all / (ms_res / 1e3)
        e2e = frames_all / (ms_e2e / 1e3)
        rows_per_frame = n_rows_step / n_pairs
        # dominant kernel family by device time inside the timed region
        fam = {k: v for k, v in prof.items() if v[0] > 0 and k not in ("bow_assign", "mip_only")}
        total_ms = sum(v[1] for v in fam.values())
        top = max(fam, key=lambda k: fam[k][1])
        batches = args.steps * ((n_pairs + MAX_PAIRS - 1) // MAX_PAIRS)
        top_ms_per_launch = fam[top][1] / batches            # one launcher call per batch of MAX_PAIRS pairs
        alg_bytes = algorithmic_bytes_per_frame(rows_per_frame) * MAX_PAIRS
        peak, peak_src = measured_peak_gbs()
        achieved = alg_bytes / (top_ms_per_launch * 1e-3) / 1e9
        traffic = traffic_per_launch()
        cpu = cpu_reference(2, 1) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": workload_config(),
            "descriptors_per_sec": rows_all * args.steps / (ms_res / 1e3),
            "rows_per_frame": rows_per_frame,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n_frames * W * H, "d2h_bytes_per_step": int(rows_e2e) * 32,
                    "ms_per_step": ms_e2e / args.steps, "gpu_launches": int(launches_e2e), "contexts": 2,
                    "h2d_gb_per_s_per_gpu": n_frames * W * H / (ms_e2e / args.steps * 1e-3) / 1e9,
                    "timing": "host clock around the K steps, device synchronised on both sides (every fetch blocks, so "
                              "this bounds the device time from above); steps alternate between two contexts so the "
                              "fetch of step k overlaps the uploads and kernels of step k+1"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic or {}).get(top), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": top_ms_per_launch,
                         "share_of_step": fam[top][1] / total_ms if total_ms else None,
                         "whole_path_frac": (algorithmic_bytes_per_frame(rows_per_frame) * value / world) / 1e9 / peak},
            "kernel_ms_per_step": {k: v[1] / args.steps for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])},
            "kernel_accounting": "CUDA events per kernel family over %d serialised steps of the same run (%.2f ms/step serialised vs "
                                 "%.2f ms/step with four launch sequences in flight)" % (args.steps, ms_serial / args.steps, ms_res / args.steps),
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
        if cpu:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"], "sample": cpu["sample"]}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    global STREAMS, MAX_PAIRS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS, help="720x576 streams per GPU per step")
    ap.add_argument("--max-pairs", type=int, default=MAX_PAIRS, help="frame pairs per internal launch sequence")
    ap.add_argument("--wc", type=int, default=0, help="1: write-combined page-locked frame buffer for the end-to-end leg")
    args = ap.parse_args()
    STREAMS, MAX_PAIRS = args.streams, args.max_pairs
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
