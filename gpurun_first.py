import sys, numpy as np, time
sys.path.insert(0,'/root/repo')
import mofreak_b200 as m
from tests import oracle_binding as ob
orc = ob.load()
print('oracle', orc.path, 'has_ref', orc.has_reference)
for (w,h,n) in [(160,120,20),(320,240,12),(720,576,9)]:
    clip = m.synth_clip(w,h,n,1)
    ctx = m.Context(w,h,32,8)
    ctx.upload_frames(clip,0)
    cur = np.arange(5,n); ctx.extract_pairs(cur, cur-5, cur-1)
    ctx.sync()
    # stage checks on the last batch
    lastF = (n-5) % 8 or 8
    first_of_last = n - lastF
    pair = 0; c = first_of_last
    d = ob.absdiff(clip[c], clip[c-5])
    ok = np.array_equal(ctx.debug_layer(0,0,pair), d); print(w,h,'L0', ok)
    L1 = orc.twothirdsample(d); L2 = orc.halfsample(d); L3=orc.halfsample(L1); L4=orc.halfsample(L2); L5=orc.halfsample(L3)
    for li,Lr in enumerate([d,L1,L2,L3,L4,L5]):
        g = ctx.debug_layer(0,li,pair)
        s = ctx.debug_layer(1,li,pair)
        print('  layer',li,g.shape,'img',np.array_equal(g,Lr),'S',np.array_equal(s,orc.score_map(Lr)))
    kg = ctx.debug_keypoints(pair); kr = orc.brisk(d)
    print('  kps', kg.shape, kr.shape, kg.shape==kr.shape and np.array_equal(kg.view(np.uint32),kr.view(np.uint32)))
    rows, per = ctx.fetch_rows(200000, n-5)
    ref = orc.extract_clip(clip)
    same = rows.shape==ref.shape and rows.tobytes()==ref.tobytes()
    print(w,h,'rows',rows.shape, ref.shape,'IDENTICAL' if same else 'DIFF', per[:6])
    if not same and rows.shape==ref.shape:
        for name in ['x','y','frame','scale','appearance','motion']:
            print('   field',name,'mismatches', int((rows[name]!=ref[name]).reshape(len(rows),-1).any(1).sum()))
    ctx.close()
# BoW
rng = np.random.default_rng(7)
cb = rng.integers(0,256,(600,16),dtype=np.uint8); desc = rng.integers(0,256,(20000,16),dtype=np.uint8)
desc[:600]=cb; cb[17]=cb[3]
ctx = m.Context(160,120,8,1); ctx.bow_set_codebook(cb)
idx = ctx.bow_assign(desc); ref = orc.bow_assign(desc, cb)
print('bow', np.array_equal(idx,ref))
print('hist', np.array_equal(ctx.bow_histogram(desc), np.bincount(ref,minlength=600)))
