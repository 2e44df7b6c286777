"""Gantt data of the pipelined launch sequences: python profiles/timeline.py > gpurun_out/timeline.txt (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mofreak_b200

W, H, STREAMS, FPS = 720, 576, 16, 69
n_frames = STREAMS * FPS
ctx = mofreak_b200.Context(W, H, n_frames, 256)
clip = np.concatenate([mofreak_b200.synth_clip(W, H, FPS, 1 + s) for s in range(STREAMS)])
ctx.upload_frames(clip)
cur = np.concatenate([np.arange(5, FPS) + s * FPS for s in range(STREAMS)]).astype(np.int32)
prev = cur - 5
label = cur.copy()
def step():
    ctx.reset_rows(); ctx.extract_pairs(cur, prev, label)
for _ in range(3): step()
ctx.sync()
ctx.profile_enable(True)
ctx.timer_start()
for _ in range(3): step()
ms = ctx.timer_stop()
tl = ctx.profile_timeline()
print("total_ms", ms, "spans", len(tl))
for i, (name, a, b) in enumerate(tl):
    print("%4d batch %2d %-16s %8.3f %8.3f %7.3f" % (i, i // 11, name, a, b, b - a))
