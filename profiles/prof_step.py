"""The smallest program that launches every kernel of the path once per iteration - the target of the ncu captures.
One launch sequence of 256 pairs at 720x576 (4 streams x 69 frames), then bag of words, `.mofreak` text and MIP alone on
its rows.  `python profiles/prof_step.py [iterations]`; every iteration launches the same kernels in the same order
(printed as `launches per iteration`), so `ncu --launch-skip N*k -c N` selects the k-th iteration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mofreak_b200

W, H, STREAMS, FPS = int(os.environ.get("PROF_W", 720)), int(os.environ.get("PROF_H", 576)), int(os.environ.get("PROF_STREAMS", 4)), 69
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
clip = np.concatenate([mofreak_b200.synth_clip(W, H, FPS, 1 + s) for s in range(STREAMS)])
cur = np.concatenate([np.arange(5, FPS) + s * FPS for s in range(STREAMS)]).astype(np.int32)
lab = np.concatenate([np.arange(5, FPS) - 1 for _ in range(STREAMS)]).astype(np.int32)
ctx = mofreak_b200.Context(W, H, STREAMS * FPS, len(cur))
ctx.set_serial(True)
ctx.upload_frames(clip)
ctx.bow_set_codebook(np.random.default_rng(7).integers(0, 256, (600, 16), dtype=np.uint8))
margin = min(80, W // 3, H // 3)
kps = np.stack([np.random.default_rng(1).uniform(margin, W - margin, 4096), np.random.default_rng(2).uniform(margin, H - margin, 4096),
                np.random.default_rng(3).uniform(8.4, min(60, margin), 4096)], 1).astype(np.float32)
for it in range(iters):
    before = ctx.launch_count()
    ctx.reset_rows()
    ctx.extract_pairs(cur, cur - 5, lab)
    ctx.sync()
    counts = ctx.bow_histogram_rows()
    text = ctx.rows_text()
    ctx.mip_describe(5, 0, kps)
    print("iteration %d: launches per iteration %d, rows %d, text bytes %d" % (it, ctx.launch_count() - before, counts.sum(), len(text)), flush=True)
ctx.close()
