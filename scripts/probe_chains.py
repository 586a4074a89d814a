"""Per-step phase timestamps of the persistent chain kernels (debug aid, GPU only): python scripts/probe_chains.py
For CTA (0,0,0) of every chain and every step s >= 1 (step 0 has no barrier):
  arrive -> pass   : grid barrier (release fence + atomic + spin until the last CTA arrives)
  pass -> first    : fence.proxy.async + TMA issue -> first K-block landed
  first -> last    : remaining K-blocks landed
  last -> mma      : MMAs of the last K-block issued + commit
  mma -> epi0      : tcgen05.commit -> epilogue warps released
  epi0 -> epi1     : TMEM load, [cluster exchange,] cell math, stores issued
  epi1 -> arrive'  : __syncthreads of the next step (slowest warp of this CTA)
"""
import collections
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
samp, gr = m.rollout(video, K, 1)
mask, _ = m.caption_masks(samp)
r = torch.rand(K * B, device='cuda'); b = torch.rand(K * B, device='cuda')
m.rl_backward(video, samp, mask, r, b, drop_seed=3); torch.cuda.synchronize()
buf = torch.zeros(8 * 4001, dtype=torch.int64, device='cuda')
m.lib.s2vt_debug_probe(C.c_void_p(buf.data_ptr()))
m.set_reuse_frontend(False)
samp, gr = m.rollout(video, K, 2)
m.rl_backward(video, samp, mask, r, b, drop_seed=3)
torch.cuda.synchronize()
m.lib.s2vt_debug_probe(None)
a = buf.cpu().numpy()
n = int(a[0])
rec = a[8:8 * (n + 1)].reshape(n, 8)
agg = collections.OrderedDict()
for i in range(1, n):
    t = rec[i]
    if t[6] == 0 or t[0] == 0 or rec[i - 1][5] == 0:
        continue
    meta = int(t[6])
    key = (meta >> 32, meta & 0xffffffff)
    prev_epi1 = rec[i - 1][5]
    nxt_arrive = rec[i + 1][0] if i + 1 < n and rec[i + 1][6] == t[6] else 0
    agg.setdefault(key, []).append((t[1] - t[0], t[2] - t[1], t[7] - t[2], t[3] - t[7], t[4] - t[3], t[5] - t[4], (nxt_arrive - t[5]) if nxt_arrive else -1,
                                    (nxt_arrive - t[0]) if nxt_arrive else -1))
print('BN+1000*KS(+100000 weights-stationary), K, CTAs | steps | barrier | first data | last data | mma issue tail | commit->epilogue | epilogue | to next arrive | step total (ns, medians)')
for k_, v in agg.items():
    v = np.array([x for x in v if x[6] >= 0], dtype=np.float64)
    if len(v):
        print(k_, len(v), np.round(np.median(v, axis=0)).astype(int).tolist())
