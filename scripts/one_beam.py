"""A few beam-5 searches at the bench configuration (target for ncu captures): python scripts/one_beam.py [calls] [videos]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import s2vt_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
vocab, by, order = bench.load_corpus()
w2i, bias = bench.peaked_bias(vocab, by)
Tv = 80
model = s2vt_b200.Video_Caption_Generator(batch_size=B, n_video_lstm_step=Tv, bias_init_vector=bias, dropout_rate=1.0, beam_size=5, max_videos=B, max_rows=B)
model.variable('embed_word_W').mul_(3.0)
model.refresh()
feats = torch.from_numpy(bench.features(B, Tv, 1)).cuda()
for _ in range(n):
    out = model.beam_search(feats, 5, 1.0)
torch.cuda.synchronize()
print('done', out[1][:8].tolist())
