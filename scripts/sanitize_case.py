"""Small full iteration for compute-sanitizer (memcheck / synccheck): small dimensions, 160 caption rows so that the > 128-row kernels (pipelined
weights-stationary forward chain, split-K BPTT chain) run, plus beam search.  compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import s2vt_b200
from oracle import s2vt_numpy as M

dims = dict(D=256, E=96, H=200, V=501)
B, K, Tv, Tc = 32, 5, 4, 9
p = M.init_params(seed=4, dtype=np.float32, **dims)
m = s2vt_b200.Video_Caption_Generator(dim_image=dims['D'], n_words=dims['V'], word_dim=dims['E'], lstm_dim=dims['H'], batch_size=B, n_video_lstm_step=Tv,
                                      n_caption_lstm_step=Tc, dropout_rate=0.9, precision='bf16', beam_size=3, max_videos=B, max_rows=K * B)
m.load_variables(p)
video = M.synthetic_features(B, Tv, dims['D'])
for it in range(2):
    samp, greedy = m.rollout(video, K, seed=5 + it)
    mask, _ = m.caption_masks(samp)
    r = torch.rand(K * B, device='cuda'); b = torch.rand(K * B, device='cuda')
    loss = m.rl_backward(video, samp, mask, r, b, drop_seed=9 + it)
    out = m.optimizer_step(1e-3, 5.0)
m.xe_step(video, samp[:B], mask[:B], 1e-3)
m.beam_search(video, 3, 1.0)
torch.cuda.synchronize()
print('sanitize case ok: loss %.5f, grad norm %.5f' % (loss.item(), out[0].item()))
