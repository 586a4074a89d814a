"""Greedy-decode time of the temporal-attention decoder at the bench shape ([64, 32, 1536]): python scripts/att_time.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import s2vt_b200
B = 64
att = s2vt_b200.attention.Video_Caption_Generator(dim_image=1536, n_words=9972, dim_hidden=1000, batch_size=B, n_video_lstm_steps=32, n_caption_lstm_steps=35,
                                                  drop_out_rate=1.0, precision='bf16')
feats = torch.from_numpy(bench.features(B, 32, 4321)).cuda()
for _ in range(3):
    att.build_generator(feats)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    att.build_generator(feats)
e1.record(); torch.cuda.synchronize()
print('attention greedy: %.3f ms per %d-video batch' % (e0.elapsed_time(e1) / 10, B))
