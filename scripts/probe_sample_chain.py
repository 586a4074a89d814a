"""Per-step phase timestamps of the overlapped persistent sampling chain (debug aid, GPU only; the library must be built with
S2VT_NVCC_EXTRA=-DS2VT_CHAIN_PROBE): python scripts/probe_sample_chain.py
CTA 0, every step s >= 1:  barrier A | cell epilogue | Wo prefetch + barrier B | pick loads + MMAs | TMEM -> smem staging |
pick epilogue (Philox / Gumbel arg-max) | cell MMAs of the next step issued (runs beside the pick epilogue) | step total"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import s2vt_b200

B, K, Tv = 64, 5, 80
vocab, by, order = bench.load_corpus()
w2i, bias = bench.peaked_bias(vocab, by)
m = s2vt_b200.Video_Caption_Generator(batch_size=B, n_video_lstm_step=Tv, bias_init_vector=bias, max_videos=B, max_rows=K * B)
m.variable('embed_word_W').mul_(3.0)
m.refresh()
m.lib.s2vt_set_overlap(m.h, 7 | 16)
video = torch.from_numpy(bench.features(B, Tv, 1)).cuda()
m.rollout(video, K, 1); torch.cuda.synchronize()
buf = torch.zeros(8 * 4001, dtype=torch.int64, device='cuda')
m.lib.s2vt_debug_probe(C.c_void_p(buf.data_ptr()))
m.rollout(video, K, 2)
torch.cuda.synchronize()
m.lib.s2vt_debug_probe(None)
a = buf.cpu().numpy()
n = int(a[0])
rec = a[8:8 * (n + 1)].reshape(n, 8)
rows = []
for s in range(1, n - 1):
    t = rec[s]
    # the LSTM chains of the encoder share the probe buffer: their slot 6 holds a shape code, the sampling chain's a timestamp
    if t[0] == 0 or t[5] == 0 or rec[s + 1][0] == 0 or t[0] < rec[s - 1][5] or t[6] < 10 ** 18 or rec[s + 1][6] < 10 ** 18 or rec[s - 1][6] < 10 ** 18:
        continue
    rows.append((t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[5], rec[s + 1][0] - t[0]))
v = np.array(rows, dtype=np.float64)
print('steps probed', len(v), 'of', n)
print('barrier A | cell epilogue | prefetch + barrier B | pick loads + MMAs | staging | pick epilogue | next cell MMAs issued | step total (ns, medians)')
print(np.round(np.median(v, axis=0)).astype(int).tolist())
