"""Phase timestamps of the per-step tcgen05 kernels (debug aid, GPU only): python scripts/probe_steps.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import s2vt_b200

B, K, Tv = 64, 5, 80
m = s2vt_b200.Video_Caption_Generator(batch_size=B, n_video_lstm_step=Tv, max_videos=B, max_rows=K * B)
video = torch.rand(B, Tv, 1536, device='cuda')
m.rollout(video, K, 1); torch.cuda.synchronize()
buf = torch.zeros(8 * 4001, dtype=torch.int64, device='cuda')
m.lib.s2vt_debug_probe(C.c_void_p(buf.data_ptr()))
samp, gr = m.rollout(video, K, 2)
mask, _ = m.caption_masks(samp)
r = torch.rand(K * B, device='cuda'); b = torch.rand(K * B, device='cuda')
m.rl_backward(video, samp, mask, r, b, drop_seed=3)
torch.cuda.synchronize()
m.lib.s2vt_debug_probe(None)
a = buf.cpu().numpy().astype(np.uint64)
n = int(a[0])
rec = a[8:8 * (n + 1)].reshape(n, 8).astype(np.int64)
print('launches probed', n)
prev_end = None
rows = []
for i in range(n):
    t0, t1, t2, t3, t4, t5, meta, grid = rec[i]
    bn, k = int(meta >> 32), int(meta & 0xffffffff)
    rows.append((bn, k, int(grid), t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0, (t5 - prev_end) if prev_end else 0))
    prev_end = t5
import collections
agg = collections.OrderedDict()
for r_ in rows:
    agg.setdefault(r_[:3], []).append(r_[3:])
print('BN(+1000*KS) K grid | count | launch->dependency-wait-done | first-data | mma-loop | tmem-ready | epilogue | total(ns) | end-to-end cadence vs previous probed kernel')
for k_, v in agg.items():
    v = np.array(v, dtype=np.float64)
    print(k_, len(v), np.round(np.median(v, axis=0)).astype(int).tolist())
