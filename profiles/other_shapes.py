"""Resident throughput at the other BASELINE shapes (not bench lines: bench.py measures configs[3], 720x576).
    python profiles/other_shapes.py        (GPU box)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mofreak_b200

def run(w, h, clips, frames, max_pairs, steps=5):
    n = clips * frames
    ctx = mofreak_b200.Context(w, h, n, max_pairs)
    data = np.concatenate([mofreak_b200.synth_clip(w, h, frames, 1 + c) for c in range(clips)])
    ctx.upload_frames(data)
    cur = np.concatenate([np.arange(5, frames) + c * frames for c in range(clips)]).astype(np.int32)
    lab = np.concatenate([np.arange(5, frames) - 1 for _ in range(clips)]).astype(np.int32)
    def step():
        ctx.reset_rows(); ctx.extract_pairs(cur, cur - 5, lab)
    for _ in range(3): step()
    ctx.sync()
    _, rows = ctx.rows_device()
    ctx.timer_start()
    for _ in range(steps): step()
    ms = ctx.timer_stop()
    ctx.close()
    return {"shape": "%dx%d" % (w, h), "clips": clips, "frames_per_clip": frames, "pairs_per_launch": max_pairs,
            "frames_per_sec": len(cur) * steps / (ms / 1e3), "rows_per_frame": rows / len(cur), "ms_per_step": ms / steps}

if __name__ == "__main__":
    for args in [(160, 120, 256, 69, 4096), (160, 120, 256, 69, 1024), (320, 240, 64, 69, 1024), (320, 240, 64, 69, 512), (1280, 720, 8, 37, 64)]:
        print(json.dumps(run(*args)), flush=True)
