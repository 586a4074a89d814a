"""Per-step timeline of the pipelined weights-stationary chain (gemm_tcgen05_ws2.cuh), CTA (0,0); build with
S2VT_NVCC_EXTRA=-DS2VT_CHAIN_PROBE.  Per step and half: dependency satisfied (loads start) | all MMAs issued | accumulator ready
(epilogue starts) | half published.  Prints medians of the intervals (ns)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ge.build()
import s2vt_b200

B, K, Tv = 64, int(os.environ.get('PROBE_K', '5')), 80
m = s2vt_b200.Video_Caption_Generator(batch_size=B, n_video_lstm_step=Tv, max_videos=B, max_rows=K * B)
video = torch.rand(B, Tv, 1536, device='cuda')
samp, gr = m.rollout(video, K, 1)
m.teacher_forward(video, samp, drop_seed=3); torch.cuda.synchronize()
T = Tv + 35
for rep in range(2):
    buf = torch.zeros(8 * 4001, dtype=torch.int64, device='cuda')
    m.lib.s2vt_debug_probe(C.c_void_p(buf.data_ptr()))
    m.teacher_forward(video, samp, drop_seed=3)          # launches: LSTM1 chain (T steps), then the WS2 chain (T steps)
    torch.cuda.synchronize()
    m.lib.s2vt_debug_probe(None)
a = buf.cpu().numpy()
rec = a[8:8 * (int(a[0]) + 1)].reshape(-1, 8)[T:2 * T].astype(np.int64)      # second kernel of the call
d = lambda x: int(np.median(x))
s = slice(2, T - 1)
print('steps probed', rec.shape[0])
for hf in (0, 1):
    o = 4 * hf
    print('half %d: dep->mma issued %d | mma issued->acc ready %d | epilogue (acc ready->published) %d' %
          (hf, d(rec[s, o + 1] - rec[s, o + 0]), d(rec[s, o + 2] - rec[s, o + 1]), d(rec[s, o + 3] - rec[s, o + 2])))
print('half0 published -> half0 next dep satisfied (flag latency + slowest CTA of the row group): %d' % d(rec[3:T - 1, 0] - rec[2:T - 2, 3]))
print('half1 published -> half1 next dep satisfied: %d' % d(rec[3:T - 1, 4] - rec[2:T - 2, 7]))
print('half0 published -> half1 acc ready (idle gap of the epilogue warps if > 0): %d' % d(rec[s, 6] - rec[s, 3]))
print('half1 published -> next half0 acc ready: %d' % d(rec[3:T - 1, 2] - rec[2:T - 2, 7]))
print('step (half1 published to half1 published): %d' % d(rec[3:T - 1, 7] - rec[2:T - 2, 7]))
np.save('gpurun_out/probe_ws2.npy', rec)
